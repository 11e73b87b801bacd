// liloc_b200/csrc/s2m.cuh -- scan2MapOptimization as ONE persistent cooperative kernel.
//
// Replaces (reference, src/mapOptmization.cpp): surfOptimization :981-1048, combineOptimizationCoeffs
// :1050-1059, LMOptimization :1061-1181 and the iteration loop of scan2MapOptimization :1190-1200.
// The pose, the 6x6 normal equations, the degeneracy projector and the convergence test never leave
// the device: one grid barrier per Gauss-Newton iteration, zero host round trips.
//
// Per iteration, per batch of up to kBatchTiles tiles of kQPB scan points (all the tiles of a block on the usual launch):
//   A  per tile, eight lanes per query: transform by the current pose, exact 5-NN in the search grid (knn_grid.cuh), bounded from
//      the second iteration on by the distances to the previous iteration's neighbours; neighbour indices -> shared memory
//   B  per batch, ONE THREAD per query: 5x3 pivoted-Householder plane fit (Eigen colPivHouseholderQr semantics) or line fit,
//      validity gates, weight, Jacobian row (LOAM Euler form)            -> shared memory
//   C  per tile, in tile order: the 21 + 6 unique J^T J | J^T r sums and the correspondence count, accumulated in double from
//      exact float*float products by a fixed shuffle tree                -> one partial per block
//   -- grid barrier --
//   D  every block redundantly: fixed-order reduction of the block partials, float32 Householder QR
//      solve of the 6x6, (iteration 0) Jacobi eigen-decomposition + degeneracy projector, pose update,
//      convergence test.  Redundant evaluation is bit-identical in every block, so no second barrier.
// No intermediate (point, coefficient) array ever reaches HBM.
//
// Two feature classes share the loop: class 0 = surf, the reference's only class (plane residuals); class 1 =
// corner, the upstream LIO-SAM cornerOptimization the north_star names but the reference dropped (line
// residuals).  Corner tiles run first, then surf tiles (combineOptimizationCoeffs order).
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"
#include "knn_grid.cuh"

namespace liloc {

namespace cg = cooperative_groups;

constexpr int kS2mThreads = 256;
constexpr int kS2mWarps = kS2mThreads / 32;
constexpr int kQPB = kS2mThreads / kKnnLanes;   // queries per tile: one group of kKnnLanes lanes per query
constexpr int kNSums = 28;               // 21 (upper J^T J) + 6 (J^T r) + 1 (count)
constexpr int kRedGroups = kS2mThreads / kNSums;   // 9 groups of 28 threads for the cross-block reduction
constexpr int kSumsPerWarp = (kNSums + kS2mWarps - 1) / kS2mWarps;      // 4: warp w reduces sums 4w .. 4w+3 of a tile
static_assert(kQPB == 32, "the tile sums map one query row to one lane");
constexpr int kBatchTiles = kS2mThreads / kQPB;   // tiles whose plane / line fits run together: one fit per thread
constexpr int kBatchQ = kBatchTiles * kQPB;

struct S2mState {                        // members isDegenerate / matP (src/mapOptmization.cpp:90-91)
    int is_degenerate;
    float matP[36];
};

struct S2mResult {                       // device mirror of liloc_s2m_result (same field order)
    int iters, converged, n_sel, is_degenerate, too_few, skipped, q_all, map_points;
    float matP[36];
    float AtA0[36];
    float pose[6];
    int extra[8];                        // q_corner, map_points_corner, tiles_corner, tiles_surf, 0...
    float trace[32][16];                 // per iteration: n_sel, X[6], pose_after[6], AtB[0..2]
    int clk[32][8];                      // LILOC_S2M_PROFILE: SM-clock deltas of the iteration phases (block 0)
};

struct FeatView {                        // one feature class
    const float4* scan;                  // down-sampled scan, body frame
    const int* q_all;                    // device count
    const float4* map;                   // down-sampled local map, original (voxel) order
    const int* map_count;
    GridView grid;
    int stride;
};

struct S2mArgs {
    FeatView f[2];                       // 0 surf, 1 corner
    int enable_corner;
    int max_iters, min_sel, min_points, min_corner;
    double* partial;                     // [2][gridDim.x][kNSums]
    float pose_in[6];                    // initial guess (travels with the launch: no copy in front of the kernel)
    unsigned* done_flag;                 // optional (mapped host memory): receives done_value after everything else is written
    unsigned done_value;
    float* pose_io;                      // 6 floats: out = result (also in S2mResult::pose)
    const S2mState* state;               // members isDegenerate / matP before this call ...
    S2mState* state_out;                 // ... and after it (double-buffered by the host, see liloc_frame_features)
    S2mResult* result;
    unsigned long long* marks;           // optional latency breakdown: %globaltimer at [0] start and, per iteration i,
                                         // [1+3i] tiles done, [2+3i] past the grid barrier, [3+3i] pose updated (block 0)
};
constexpr int kS2mMarks = 2 + 3 * 32;      // start, three per iteration, [kS2mMarks - 1] = end of block 0

LILOC_DEV unsigned long long s2m_global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// ---- 5x3 least squares  A x = -1  (Eigen::colPivHouseholderQr().solve, src/mapOptmization.cpp:1009) --
// Same operation sequence as oracle plane_lsq_5x3.  Register-only: every array index is static.
LILOC_DEV void plane_lsq_5x3(const float (&P)[5][3], float (&x)[3]) {
    float q[3][5];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 5; ++i) q[j][i] = P[i][j];
    int perm[3] = {0, 1, 2};
    float hco[3] = {0.f, 0.f, 0.f};
    float nUpd[3], nDir[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 5; ++i) s += q[j][i] * q[j][i];
        nDir[j] = nUpd[j] = sqrtf(s);
    }
    const float eps = 1.1920929e-07f;
    float mxn = nUpd[0]; if (nUpd[1] > mxn) mxn = nUpd[1]; if (nUpd[2] > mxn) mxn = nUpd[2];
    const float thr_helper = ((mxn * eps) * (mxn * eps)) / 5.0f;
    const float downdate_thr = sqrtf(eps);
    int nonzero = 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int big = k; float bign = nUpd[k];
#pragma unroll
        for (int j = k + 1; j < 3; ++j) if (nUpd[j] > bign) { bign = nUpd[j]; big = j; }
        if (nonzero == 3 && bign * bign < thr_helper * (float)(5 - k)) nonzero = k;
#pragma unroll
        for (int j = k + 1; j < 3; ++j) {
            if (big == j) {
#pragma unroll
                for (int i = 0; i < 5; ++i) { const float t = q[k][i]; q[k][i] = q[j][i]; q[j][i] = t; }
                { const float t = nUpd[k]; nUpd[k] = nUpd[j]; nUpd[j] = t; }
                { const float t = nDir[k]; nDir[k] = nDir[j]; nDir[j] = t; }
                { const int t = perm[k]; perm[k] = perm[j]; perm[j] = t; }
            }
        }
        float tailSq = 0.f;
#pragma unroll
        for (int i = k + 1; i < 5; ++i) tailSq += q[k][i] * q[k][i];
        const float c0 = q[k][k];
        float beta, tau;
        if (tailSq <= 1.17549435e-38f) {
            tau = 0.f; beta = c0;
#pragma unroll
            for (int i = k + 1; i < 5; ++i) q[k][i] = 0.f;
        } else {
            beta = sqrtf(c0 * c0 + tailSq);
            if (c0 >= 0.f) beta = -beta;
            const float den = c0 - beta;
#pragma unroll
            for (int i = k + 1; i < 5; ++i) q[k][i] = q[k][i] / den;
            tau = (beta - c0) / beta;
        }
        q[k][k] = beta; hco[k] = tau;
#pragma unroll
        for (int j = k + 1; j < 3; ++j) {
            float tmp = 0.f;
#pragma unroll
            for (int i = k + 1; i < 5; ++i) tmp += q[k][i] * q[j][i];
            tmp += q[j][k];
            q[j][k] -= tau * tmp;
#pragma unroll
            for (int i = k + 1; i < 5; ++i) q[j][i] -= (tau * q[k][i]) * tmp;
        }
#pragma unroll
        for (int j = k + 1; j < 3; ++j) {
            if (nUpd[j] != 0.f) {
                float temp = fabsf(q[j][k]) / nUpd[j];
                temp = (1.0f + temp) * (1.0f - temp);
                if (temp < 0.f) temp = 0.f;
                const float r = nUpd[j] / nDir[j];
                const float temp2 = temp * (r * r);
                if (temp2 <= downdate_thr) {
                    float s = 0.f;
#pragma unroll
                    for (int i = k + 1; i < 5; ++i) s += q[j][i] * q[j][i];
                    nDir[j] = sqrtf(s); nUpd[j] = nDir[j];
                } else {
                    nUpd[j] *= sqrtf(temp);
                }
            }
        }
    }
    float c[5] = {-1.f, -1.f, -1.f, -1.f, -1.f};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (k < nonzero) {
            float tmp = 0.f;
#pragma unroll
            for (int i = k + 1; i < 5; ++i) tmp += q[k][i] * c[i];
            tmp += c[k];
            c[k] -= hco[k] * tmp;
#pragma unroll
            for (int i = k + 1; i < 5; ++i) c[i] -= (hco[k] * q[k][i]) * tmp;
        }
    }
#pragma unroll
    for (int i = 2; i >= 0; --i) {
        if (i < nonzero) {
            c[i] = c[i] / q[i][i];
#pragma unroll
            for (int r = 0; r < i; ++r) c[r] -= q[i][r] * c[i];
        }
    }
    x[0] = x[1] = x[2] = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (i < nonzero) {
            if (perm[i] == 0) x[0] = c[i];
            else if (perm[i] == 1) x[1] = c[i];
            else x[2] = c[i];
        }
    }
}

// Loop body of surfOptimization after the kNN (src/mapOptmization.cpp:1002-1045).
// nb[j] = j-th neighbour (map frame).  Returns true when the correspondence is kept.
LILOC_DEV bool surf_residual(const float (&nb)[5][3], const float d2_4, const float4 ori, const float4 sel, float4* coeff) {
    if (!(d2_4 < 1.0f)) return false;
    float X[3];
    plane_lsq_5x3(nb, X);
    float pa = X[0], pb = X[1], pc = X[2], pd = 1.f;
    const float ps = sqrtf(pa * pa + pb * pb + pc * pc);
    pa /= ps; pb /= ps; pc /= ps; pd /= ps;
    bool valid = true;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const float r = pa * nb[j][0] + pb * nb[j][1] + pc * nb[j][2] + pd;
        if ((double)fabsf(r) > 0.2) valid = false;
    }
    if (!valid) return false;
    const float pd2 = pa * sel.x + pb * sel.y + pc * sel.z + pd;
    const float rng = sqrtf(sqrtf(ori.x * ori.x + ori.y * ori.y + ori.z * ori.z));
    const float s = (float)(1 - 0.9 * (double)fabsf(pd2) / (double)rng);
    *coeff = make_float4(s * pa, s * pb, s * pc, s * pd2);
    return (double)s > 0.1;
}

// cv::eigen(sym 3x3) stand-in for cornerOptimization: the same cyclic Jacobi as eigen6_jacobi below (float32,
// eigenvalues descending, eigenvectors as rows), indices compile-time so everything stays in registers.
template <int P, int Q>
__host__ __device__ __forceinline__ void jacobi_rotate3(float (&A)[3][3], float (&U)[3][3]) {
    const float apq = A[P][Q];
    if (apq == 0.f) return;
    const float theta = (A[Q][Q] - A[P][P]) / (2.0f * apq);
    const float t = (theta >= 0.f ? 1.0f : -1.0f) / (fabsf(theta) + sqrtf(theta * theta + 1.0f));
    const float c = 1.0f / sqrtf(t * t + 1.0f);
    const float s = t * c;
    const float h = t * apq;
    A[P][P] -= h; A[Q][Q] += h; A[P][Q] = 0.f; A[Q][P] = 0.f;
    constexpr int K = 3 - P - Q;         // the one remaining index
    {
        const float akp = A[K][P], akq = A[K][Q];
        const float np_ = c * akp - s * akq, nq_ = s * akp + c * akq;
        A[K][P] = np_; A[P][K] = np_; A[K][Q] = nq_; A[Q][K] = nq_;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float upk = U[P][k], uqk = U[Q][k];
        U[P][k] = c * upk - s * uqk; U[Q][k] = s * upk + c * uqk;
    }
}

__host__ __device__ inline void eigen3_jacobi(const float* A_in, float* w, float* V) {
    float A[3][3], U[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) { A[i][j] = A_in[i * 3 + j]; U[i][j] = (i == j) ? 1.f : 0.f; }
    const float eps = 1.1920929e-07f;
#pragma unroll 1
    for (int sweep = 0; sweep < 30; ++sweep) {
        float off = 0.f, diag = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            diag += A[i][i] * A[i][i];
#pragma unroll
            for (int j = i + 1; j < 3; ++j) off += A[i][j] * A[i][j];
        }
        if (off <= (eps * eps) * diag || off == 0.f) break;
        jacobi_rotate3<0, 1>(A, U); jacobi_rotate3<0, 2>(A, U); jacobi_rotate3<1, 2>(A, U);
    }
    float d[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) d[i] = A[i][i];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        int r = 0;
#pragma unroll
        for (int j = 0; j < 3; ++j) r += (d[j] > d[i] || (d[j] == d[i] && j < i)) ? 1 : 0;
#pragma unroll
        for (int o = 0; o < 3; ++o) {
            if (r == o) {
                w[o] = d[i];
#pragma unroll
                for (int k = 0; k < 3; ++k) V[o * 3 + k] = U[i][k];
            }
        }
    }
}

// Loop body of upstream LIO-SAM cornerOptimization after the kNN (the reference has no
// corner path).  Same operation sequence and type promotions as the oracle's corner_residual.
LILOC_DEV bool corner_residual(const float (&nb)[5][3], const float d2_4, const float4 sel, float4* coeff) {
    if (!(d2_4 < 1.0f)) return false;
    float cx = 0.f, cy = 0.f, cz = 0.f;
#pragma unroll
    for (int j = 0; j < 5; ++j) { cx += nb[j][0]; cy += nb[j][1]; cz += nb[j][2]; }
    cx /= 5.f; cy /= 5.f; cz /= 5.f;
    float a11 = 0.f, a12 = 0.f, a13 = 0.f, a22 = 0.f, a23 = 0.f, a33 = 0.f;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const float ax = nb[j][0] - cx, ay = nb[j][1] - cy, az = nb[j][2] - cz;
        a11 += ax * ax; a12 += ax * ay; a13 += ax * az;
        a22 += ay * ay; a23 += ay * az;
        a33 += az * az;
    }
    a11 /= 5.f; a12 /= 5.f; a13 /= 5.f; a22 /= 5.f; a23 /= 5.f; a33 /= 5.f;
    const float A[9] = {a11, a12, a13, a12, a22, a23, a13, a23, a33};
    float w[3], V[9];
    eigen3_jacobi(A, w, V);
    if (!(w[0] > 3.f * w[1])) return false;
    const float x0 = sel.x, y0 = sel.y, z0 = sel.z;
    const float x1 = (float)((double)cx + 0.1 * (double)V[0]);
    const float y1 = (float)((double)cy + 0.1 * (double)V[1]);
    const float z1 = (float)((double)cz + 0.1 * (double)V[2]);
    const float x2 = (float)((double)cx - 0.1 * (double)V[0]);
    const float y2 = (float)((double)cy - 0.1 * (double)V[1]);
    const float z2 = (float)((double)cz - 0.1 * (double)V[2]);
    const float m11 = (x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1);
    const float m12 = (x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1);
    const float m13 = (y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1);
    const float a012 = sqrtf(m11 * m11 + m12 * m12 + m13 * m13);
    const float l12 = sqrtf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
    const float la = ((y1 - y2) * m11 + (z1 - z2) * m12) / a012 / l12;
    const float lb = -((x1 - x2) * m11 - (z1 - z2) * m13) / a012 / l12;
    const float lc = -((x1 - x2) * m12 + (y1 - y2) * m13) / a012 / l12;
    const float ld2 = a012 / l12;
    const float s = (float)(1 - 0.9 * (double)fabsf(ld2));
    *coeff = make_float4(s * la, s * lb, s * lc, s * ld2);
    return (double)s > 0.1;
}

struct Trig { float srx, crx, sry, cry, srz, crz; };
LILOC_DEV Trig pose_trig(const float* pose6) {
    Trig t;
    t.srx = (float)sin((double)pose6[2]); t.crx = (float)cos((double)pose6[2]);
    t.sry = (float)sin((double)pose6[1]); t.cry = (float)cos((double)pose6[1]);
    t.srz = (float)sin((double)pose6[0]); t.crz = (float)cos((double)pose6[0]);
    return t;
}

// Jacobian row of LMOptimization (src/mapOptmization.cpp:1093-1125): [d/droll, d/dpitch, d/dyaw, cx, cy, cz | -c.w]
LILOC_DEV void jacobian_row(const Trig& g, const float4 p, const float4 c, float (&row)[7]) {
    const float srx = g.srx, crx = g.crx, sry = g.sry, cry = g.cry, srz = g.srz, crz = g.crz;
    const float arx = (-srx * cry * p.x - (srx * sry * srz + crx * crz) * p.y + (crx * srz - srx * sry * crz) * p.z) * c.x
                    + (crx * cry * p.x - (srx * crz - crx * sry * srz) * p.y + (crx * sry * crz + srx * srz) * p.z) * c.y;
    const float ary = (-crx * sry * p.x + crx * cry * srz * p.y + crx * cry * crz * p.z) * c.x
                    + (-srx * sry * p.x + srx * sry * srz * p.y + srx * cry * crz * p.z) * c.y
                    + (-cry * p.x - sry * srz * p.y - sry * crz * p.z) * c.z;
    const float arz = ((crx * sry * crz + srx * srz) * p.y + (srx * crz - crx * sry * srz) * p.z) * c.x
                    + ((-crx * srz + srx * sry * crz) * p.y + (-srx * sry * srz - crx * crz) * p.z) * c.y
                    + (cry * crz * p.y - cry * srz * p.z) * c.z;
    row[0] = arz; row[1] = ary; row[2] = arx; row[3] = c.x; row[4] = c.y; row[5] = c.z; row[6] = -c.w;
}

// ---- 6x6 float32 linear algebra (OpenCV stand-ins); same operation order as the oracle ----
__host__ __device__ inline void solve6_qr(const float* A_in, const float* b_in, float* x) {
    float A[6][6], b[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        b[i] = b_in[i];
#pragma unroll
        for (int j = 0; j < 6; ++j) A[i][j] = A_in[i * 6 + j];
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        float tailSq = 0.f;
#pragma unroll
        for (int i = k + 1; i < 6; ++i) tailSq += A[i][k] * A[i][k];
        const float c0 = A[k][k];
        if (tailSq > 1.17549435e-38f) {
            float beta = sqrtf(c0 * c0 + tailSq);
            if (c0 >= 0.f) beta = -beta;
            const float den = c0 - beta;
            float v[6];
#pragma unroll
            for (int i = k + 1; i < 6; ++i) v[i] = A[i][k] / den;
            const float tau = (beta - c0) / beta;
            A[k][k] = beta;
#pragma unroll
            for (int i = k + 1; i < 6; ++i) A[i][k] = 0.f;
#pragma unroll
            for (int j = k + 1; j < 6; ++j) {
                float tmp = 0.f;
#pragma unroll
                for (int i = k + 1; i < 6; ++i) tmp += v[i] * A[i][j];
                tmp += A[k][j];
                A[k][j] -= tau * tmp;
#pragma unroll
                for (int i = k + 1; i < 6; ++i) A[i][j] -= (tau * v[i]) * tmp;
            }
            float tmp = 0.f;
#pragma unroll
            for (int i = k + 1; i < 6; ++i) tmp += v[i] * b[i];
            tmp += b[k];
            b[k] -= tau * tmp;
#pragma unroll
            for (int i = k + 1; i < 6; ++i) b[i] -= (tau * v[i]) * tmp;
        }
    }
    float xr[6];
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        float s = b[i];
#pragma unroll
        for (int j = i + 1; j < 6; ++j) s -= A[i][j] * xr[j];
        xr[i] = s / A[i][i];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) x[i] = xr[i];
}

// cv::eigen stand-in: cyclic Jacobi, float32, eigenvalues descending, eigenvectors as rows of V.
// Every matrix index is a compile-time constant (template parameters + full unrolling) so A and U
// live in registers; the loop form of the same operation sequence is the oracle's eigen6_jacobi.
template <int P, int Q>
__host__ __device__ __forceinline__ void jacobi_rotate(float (&A)[6][6], float (&U)[6][6]) {
    const float apq = A[P][Q];
    if (apq == 0.f) return;
    const float theta = (A[Q][Q] - A[P][P]) / (2.0f * apq);
    const float t = (theta >= 0.f ? 1.0f : -1.0f) / (fabsf(theta) + sqrtf(theta * theta + 1.0f));
    const float c = 1.0f / sqrtf(t * t + 1.0f);
    const float s = t * c;
    const float h = t * apq;
    A[P][P] -= h; A[Q][Q] += h; A[P][Q] = 0.f; A[Q][P] = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        if (k == P || k == Q) continue;
        const float akp = A[k][P], akq = A[k][Q];
        const float np_ = c * akp - s * akq, nq_ = s * akp + c * akq;
        A[k][P] = np_; A[P][k] = np_; A[k][Q] = nq_; A[Q][k] = nq_;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const float upk = U[P][k], uqk = U[Q][k];
        U[P][k] = c * upk - s * uqk; U[Q][k] = s * upk + c * uqk;
    }
}

__host__ __device__ inline void eigen6_jacobi(const float* A_in, float* w, float* V) {
    float A[6][6], U[6][6];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) { A[i][j] = A_in[i * 6 + j]; U[i][j] = (i == j) ? 1.f : 0.f; }
    const float eps = 1.1920929e-07f;
#pragma unroll 1
    for (int sweep = 0; sweep < 30; ++sweep) {
        float off = 0.f, diag = 0.f;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            diag += A[i][i] * A[i][i];
#pragma unroll
            for (int j = i + 1; j < 6; ++j) off += A[i][j] * A[i][j];
        }
        if (off <= (eps * eps) * diag || off == 0.f) break;
        jacobi_rotate<0, 1>(A, U); jacobi_rotate<0, 2>(A, U); jacobi_rotate<0, 3>(A, U); jacobi_rotate<0, 4>(A, U); jacobi_rotate<0, 5>(A, U);
        jacobi_rotate<1, 2>(A, U); jacobi_rotate<1, 3>(A, U); jacobi_rotate<1, 4>(A, U); jacobi_rotate<1, 5>(A, U);
        jacobi_rotate<2, 3>(A, U); jacobi_rotate<2, 4>(A, U); jacobi_rotate<2, 5>(A, U);
        jacobi_rotate<3, 4>(A, U); jacobi_rotate<3, 5>(A, U);
        jacobi_rotate<4, 5>(A, U);
    }
    // stable descending selection on the diagonal; rank[i] = output position of eigenpair i
    float d[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) d[i] = A[i][i];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        int r = 0;
#pragma unroll
        for (int j = 0; j < 6; ++j) r += (d[j] > d[i] || (d[j] == d[i] && j < i)) ? 1 : 0;
        w[r] = d[i];
#pragma unroll
        for (int k = 0; k < 6; ++k) V[r * 6 + k] = U[i][k];
    }
}

// src/mapOptmization.cpp:1132-1154
__host__ __device__ inline int degeneracy_projector(const float* AtA, float* matP) {
    float w[6], V[36], V2[36];
    eigen6_jacobi(AtA, w, V);
    for (int i = 0; i < 36; ++i) V2[i] = V[i];
    int degenerate = 0;
    for (int i = 5; i >= 0; --i) {
        if (w[i] < 100.f) { for (int j = 0; j < 6; ++j) V2[i * 6 + j] = 0.f; degenerate = 1; }
        else break;
    }
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) {
        double s = 0.0;
        for (int k = 0; k < 6; ++k) s += (double)V[k * 6 + i] * (double)V2[k * 6 + j];
        matP[i * 6 + j] = (float)s;
    }
    return degenerate;
}

// True when every eigenvalue of the symmetric 6x6 A is provably above `bound`: Cholesky of (A - bound*I)
// in double succeeds.  Used to skip the Jacobi sweep in the common non-degenerate case: with
// bound = 100 + 1e-4 * max|diag| the float32 Jacobi (error ~1e-6 * |A|) would reach the same decision.
__host__ __device__ inline bool eigenvalues_above(const float* A, double bound) {
    double L[6][6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double d = (double)A[j * 6 + j] - bound;
#pragma unroll
        for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
        if (!(d > 0.0)) return false;
        const double inv = 1.0 / sqrt(d);
        L[j][j] = d;
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
            double v = (double)A[i * 6 + j];
#pragma unroll
            for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k];
            L[i][j] = v * inv;
        }
    }
    return true;
}

// T (3x4) and the LOAM trig set from sc = {sin roll, sin pitch, sin yaw, cos roll, cos pitch, cos yaw}
// (each evaluated in double and rounded to float; identical to pose_to_T / pose_trig).
LILOC_DEV void T_from_sincos(const float* pose6, const float* sc, float* T) {
    const float A = sc[5], B = sc[2], C = sc[4], D = sc[1], E = sc[3], F = sc[0];
    const float DE = D * E, DF = D * F;
    T[0] = A * C;  T[1] = A * DF - B * E;  T[2]  = B * F + A * DE;  T[3]  = pose6[3];
    T[4] = B * C;  T[5] = A * E + B * DF;  T[6]  = B * DE - A * F;  T[7]  = pose6[4];
    T[8] = -D;     T[9] = C * F;           T[10] = C * E;           T[11] = pose6[5];
}

// which (a, c) pair of the row feeds sum t: t < 21 upper triangle row-major, 21..26 (a, rhs), 27 count
__constant__ unsigned char c_sum_a[kNSums] = {0,0,0,0,0,0, 1,1,1,1,1, 2,2,2,2, 3,3,3, 4,4, 5, 0,1,2,3,4,5, 7};
__constant__ unsigned char c_sum_c[kNSums] = {0,1,2,3,4,5, 1,2,3,4,5, 2,3,4,5, 3,4,5, 4,5, 5, 6,6,6,6,6,6, 7};

struct S2mShared {
    FeatView fv[2];
    GridParams gp[2];
    float T[12];
    float pose[6];
    float sc[6];
    Trig trig;
    float row[kBatchQ][9];     // 6 Jacobian entries, rhs, valid (1.0 / 0.0); padded to 9 words (one row per lane, no bank conflicts)
    float4 ori[kBatchQ];       // the batch's queries, body frame
    int nn[kBatchQ][5];        // ... their five neighbours (indices into the class's local map)
    int qk[kBatchQ];           // ... their index among the class's queries, -1: slot not in use
    unsigned char qcls[kBatchQ], qok[kBatchQ];   // ... class; 1 when five neighbours lie within 1 m
    double sums[kNSums];
    double red[kRedGroups][kNSums];
    float X[6];
    long long dbg[4];
    float matP[36];
    int is_degenerate;
    int done;                  // 0 continue, 1 converged, 2 too few
    int n_sel;
};

// Phase A for one tile of queries of class `cls`: a group of kKnnLanes lanes owns one query (cooperative exact 5-NN); the
// neighbour indices go to batch slots [slot0, slot0 + kQPB) in shared memory.
// `seeded`: the slots hold this tile's queries and their neighbours of the previous iteration (a block whose tiles fit one batch
// sees the same queries in every iteration): the distances to those five points bound the search (knn5_group).
LILOC_DEV void s2m_tile_knn(S2mShared& sh, const int cls, const int tile, const int nq, const int slot0, const bool seeded = false) {
    const FeatView& f = sh.fv[cls];
    const int s = threadIdx.x / kKnnLanes, sub = threadIdx.x % kKnnLanes;
    const int qk = tile * kQPB + s;
    const bool active = qk < nq;
    float4 ori = make_float4(0.f, 0.f, 0.f, 0.f), sel = ori;
    if (active) { ori = __ldg(&f.scan[qk * f.stride]); sel = apply_T(sh.T, ori); }
    float bound = 3.4e38f;
    if (seeded) {                                     // uniform over the block
        float d = 0.f;
        const bool have = active && sh.qok[slot0 + s] != 0;
        if (have && sub < 5) d = dist2(sel, __ldg(&f.map[sh.nn[slot0 + s][sub]]));
#pragma unroll
        for (int o = kKnnLanes / 2; o > 0; o >>= 1) d = fmaxf(d, __shfl_xor_sync(0xffffffffu, d, o));
        if (have) bound = d;
        __syncwarp();                                 // the slot is rewritten below by lane 0 of the group
    }
    const Knn5 r = knn5_group<kKnnLanes>(f.grid, sh.gp[cls], sel, sub, active, bound);
    if (sub == 0) {
        const int slot = slot0 + s;
        sh.qk[slot] = active ? qk : -1;
        sh.qcls[slot] = (unsigned char)cls;
        sh.qok[slot] = (active && r.i[4] != 0x7fffffff && r.d[4] < 1.0f) ? 1 : 0;
        sh.ori[slot] = ori;
#pragma unroll
        for (int j = 0; j < 5; ++j) sh.nn[slot][j] = r.i[j];
    }
}

// Phase B for a batch of `nslots` queries: ONE THREAD PER QUERY does the plane (or line) fit, the gates and the Jacobian row.
// The fit is a ~3 us chain of dependent IEEE divisions and square roots whatever the number of lanes that walk it; run behind
// every tile's 5-NN on one lane of eight it was paid once per tile (two tiles per block on cfg 1: 2 x 2.9 us per iteration),
// here once per batch of up to kBatchTiles tiles, with all lanes of the fitting warps busy.
// Writes coeff_out / flag_out as well when they are non-null (k_feature_pass).
LILOC_DEV void s2m_batch_rows(S2mShared& sh, const int nslots, float4* __restrict__ coeff_out, unsigned char* __restrict__ flag_out) {
    const int slot = threadIdx.x;
    if (slot >= nslots) return;
    float row[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float valid = 0.f;
    float4 coeff = make_float4(0.f, 0.f, 0.f, 0.f);
    const int qk = sh.qk[slot];
    const int cls = sh.qcls[slot];
    const FeatView& f = sh.fv[cls];
    if (qk >= 0 && sh.qok[slot]) {
        const float4 ori = sh.ori[slot];
        const float4 sel = apply_T(sh.T, ori);
        float nb[5][3];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const float4 p = __ldg(&f.map[sh.nn[slot][j]]);
            nb[j][0] = p.x; nb[j][1] = p.y; nb[j][2] = p.z;
        }
        float4 c;
        const bool keep = cls ? corner_residual(nb, 0.f, sel, &c) : surf_residual(nb, 0.f, ori, sel, &c);      // the 1 m gate is in qok
        if (keep) { coeff = c; jacobian_row(sh.trig, ori, c, row); valid = 1.f; }
    }
    if (valid == 0.f) {
#pragma unroll
        for (int j = 0; j < 7; ++j) row[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 7; ++j) sh.row[slot][j] = row[j];
    sh.row[slot][7] = valid;
    if (coeff_out && qk >= 0) { coeff_out[qk * f.stride] = coeff; flag_out[qk * f.stride] = valid != 0.f ? 1 : 0; }
}

// Stage the class views and the device-resident grid geometry in shared memory (once per kernel).
LILOC_DEV void s2m_stage_views(const S2mArgs& a, S2mShared& sh) {
    if (threadIdx.x == 0) {
        sh.fv[0] = a.f[0]; sh.fv[1] = a.f[1];
        sh.gp[0] = *a.f[0].grid.p;
        if (a.enable_corner) sh.gp[1] = *a.f[1].grid.p;
    }
}

// sin/cos of the three angles on six lanes in parallel, then T and the trig set by one thread.
LILOC_DEV void s2m_refresh_transform(S2mShared& sh) {
    if (threadIdx.x < 6) {
        const double ang = (double)sh.pose[threadIdx.x % 3];
        sh.sc[threadIdx.x] = (float)(threadIdx.x < 3 ? sin(ang) : cos(ang));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        T_from_sincos(sh.pose, sh.sc, sh.T);
        sh.trig.srx = sh.sc[2]; sh.trig.crx = sh.sc[5]; sh.trig.sry = sh.sc[1]; sh.trig.cry = sh.sc[4];
        sh.trig.srz = sh.sc[0]; sh.trig.crz = sh.sc[3];
    }
    __syncthreads();
}

#ifndef LILOC_S2M_MINBLOCKS
#define LILOC_S2M_MINBLOCKS 2
#endif
__global__ void __launch_bounds__(kS2mThreads, LILOC_S2M_MINBLOCKS) k_scan2map(S2mArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ S2mShared sh;
    const int q_all = *a.f[0].q_all;
    const int M = *a.f[0].map_count;
    const int nq = (q_all + a.f[0].stride - 1) / a.f[0].stride;
    const int q_corner = a.enable_corner ? *a.f[1].q_all : 0;
    const int nq_c = a.enable_corner ? (q_corner + a.f[1].stride - 1) / a.f[1].stride : 0;
    const int n_tiles_c = (nq_c + kQPB - 1) / kQPB;
    const int n_tiles = n_tiles_c + (nq + kQPB - 1) / kQPB;

    s2m_stage_views(a, sh);
    if (threadIdx.x < 6) sh.pose[threadIdx.x] = a.pose_in[threadIdx.x];
    if (threadIdx.x >= 32 && threadIdx.x < 68) sh.matP[threadIdx.x - 32] = a.state->matP[threadIdx.x - 32];
    if (threadIdx.x == 0) { sh.is_degenerate = a.state->is_degenerate; sh.done = 0; sh.n_sel = 0; }
    __syncthreads();
    s2m_refresh_transform(sh);

    // :1187; with corners the upstream gate (cornerDS > edgeFeatureMinValidNum && surfDS > surfFeatureMinValidNum)
    const bool skipped = !(q_all > a.min_points) || (a.enable_corner && !(q_corner > a.min_corner));
    const bool marking = a.marks != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    if (marking) a.marks[0] = s2m_global_ns();
    int iters = 0, converged = 0, too_few = 0;
#ifdef LILOC_S2M_PROFILE
    long long clk[8];
#define S2M_CLK(i) if (blockIdx.x == 0 && threadIdx.x == 0) clk[i] = clock64();
#else
#define S2M_CLK(i)
#endif
    const bool one_batch = (long long)blockIdx.x + (long long)kBatchTiles * gridDim.x >= n_tiles;      // all tiles of this block in one batch
    if (!skipped) {
        for (int iter = 0; iter < a.max_iters; ++iter) {
            S2M_CLK(0)
            // C: lane = query row of the tile, warp w owns sums 4w .. 4w+3.  Products of two floats are exact in double;
            // the 32 rows are added by a fixed xor-shuffle tree (5 levels instead of a 32-long chain of dependent
            // FP64 adds, which cost ~290 cycles each on this part and used to be the longest piece of the tile).
            double acc[kSumsPerWarp];                        // lane 0 of every warp: running sums over this block's tiles
#pragma unroll
            for (int u = 0; u < kSumsPerWarp; ++u) acc[u] = 0.0;
            for (int tile0 = blockIdx.x; tile0 < n_tiles; tile0 += kBatchTiles * gridDim.x) {
                int nb = 0;                                    // tiles of this batch
#ifdef LILOC_S2M_PROFILE
                if (threadIdx.x == 0) sh.dbg[0] = clock64();
#endif
                for (int tile = tile0; tile < n_tiles && nb < kBatchTiles; tile += gridDim.x, ++nb) {
                    const int cls = tile < n_tiles_c ? 1 : 0;  // corner tiles first (combineOptimizationCoeffs order)
                    s2m_tile_knn(sh, cls, cls ? tile : tile - n_tiles_c, cls ? nq_c : nq, nb * kQPB, iter > 0 && one_batch);
                }
#ifdef LILOC_S2M_PROFILE
                if (threadIdx.x == 0) sh.dbg[1] = clock64();
#endif
                __syncthreads();
                s2m_batch_rows(sh, nb * kQPB, nullptr, nullptr);
#ifdef LILOC_S2M_PROFILE
                if (threadIdx.x == 0) sh.dbg[2] = clock64();
#endif
                __syncthreads();
#ifdef LILOC_S2M_PROFILE
                if (threadIdx.x == 0) sh.dbg[3] = clock64();
#endif
                for (int i = 0; i < nb; ++i) {                 // tile sums, tiles in ascending order as before
                    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
                    const float* r = sh.row[i * kQPB + lane];
                    double v[kSumsPerWarp];
#pragma unroll
                    for (int u = 0; u < kSumsPerWarp; ++u) {
                        const int k = warp * kSumsPerWarp + u;
                        v[u] = 0.0;
                        if (k < kNSums) v[u] = (double)r[c_sum_a[k]] * (double)r[c_sum_c[k]];
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                        for (int u = 0; u < kSumsPerWarp; ++u) v[u] += shfl_xor_d(v[u], o);
                    }
#pragma unroll
                    for (int u = 0; u < kSumsPerWarp; ++u) acc[u] += v[u];
                }
                __syncthreads();
            }
            S2M_CLK(1)
            if (marking && iter < 32) a.marks[1 + 3 * iter] = s2m_global_ns();
            double* part = a.partial + (size_t)(iter & 1) * gridDim.x * kNSums;
            if ((threadIdx.x & 31) == 0) {
#pragma unroll
                for (int u = 0; u < kSumsPerWarp; ++u) {
                    const int k = (threadIdx.x >> 5) * kSumsPerWarp + u;
                    if (k < kNSums) part[(size_t)blockIdx.x * kNSums + k] = acc[u];
                }
            }
            grid.sync();
            S2M_CLK(2)
            if (marking && iter < 32) a.marks[2 + 3 * iter] = s2m_global_ns();
            // D: fixed-order reduction of the block partials, identical in every block.  Thread (j, c) sums
            // blocks j, j + kRedGroups, ... of column c (coalesced 224-byte rows; an L2 round trip is ~0.5 us here, so
            // all of a thread's loads are issued before the first add), then column c adds its kRedGroups group sums.
            if (threadIdx.x < kRedGroups * kNSums) {
                const int c = threadIdx.x % kNSums, j = threadIdx.x / kNSums;
                const int nb = min((int)gridDim.x, n_tiles);         // blocks beyond the tile count wrote zeros: skip them
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;      // four independent chains
#pragma unroll 1
                for (int b = j; b < nb; b += 16 * kRedGroups) {     // 16 (predicated) L2 loads in flight per thread
                    double v[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const int bb = b + u * kRedGroups;
                        v[u] = bb < nb ? __ldcg(&part[(size_t)bb * kNSums + c]) : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 16; u += 4) { s0 += v[u]; s1 += v[u + 1]; s2 += v[u + 2]; s3 += v[u + 3]; }
                }
                sh.red[j][c] = (s0 + s1) + (s2 + s3);
            }
            __syncthreads();
            if (threadIdx.x < kNSums) {
                const double* r = &sh.red[0][threadIdx.x];
                static_assert(kRedGroups == 9, "fixed reduction tree below");
                sh.sums[threadIdx.x] = (((r[0 * kNSums] + r[1 * kNSums]) + (r[2 * kNSums] + r[3 * kNSums])) +
                                        ((r[4 * kNSums] + r[5 * kNSums]) + (r[6 * kNSums] + r[7 * kNSums]))) + r[8 * kNSums];
            }
            __syncthreads();
            S2M_CLK(3)
            const int n_sel = (int)sh.sums[27];
            const bool few = n_sel < a.min_sel;             // :1079-1082 (pose unchanged: later iterations identical)
            if (!few && (threadIdx.x == 0 || (threadIdx.x == 32 && iter == 0))) {
                float AtA[36];
                int k = 0;
#pragma unroll
                for (int r = 0; r < 6; ++r)
#pragma unroll
                    for (int c = r; c < 6; ++c) { AtA[r * 6 + c] = AtA[c * 6 + r] = (float)sh.sums[k++]; }
                if (threadIdx.x == 0) {                     // cv::solve(DECOMP_QR), :1130
                    float AtB[6];
#pragma unroll
                    for (int r = 0; r < 6; ++r) AtB[r] = (float)sh.sums[21 + r];
                    solve6_qr(AtA, AtB, sh.X);
                    if (iter == 0 && blockIdx.x == 0) for (int i = 0; i < 36; ++i) a.result->AtA0[i] = AtA[i];
                } else {                                    // iteration 0, second warp: degeneracy check :1132-1154
                    float mxd = AtA[0];
#pragma unroll
                    for (int i = 1; i < 6; ++i) mxd = fmaxf(mxd, AtA[i * 7]);
                    if (eigenvalues_above(AtA, 100.0 + 1e-4 * (double)mxd)) {
                        sh.is_degenerate = 0;               // all eigenvalues >= 100: matP is never read
                        for (int i = 0; i < 36; ++i) sh.matP[i] = (i % 7 == 0) ? 1.f : 0.f;
                    } else {
                        sh.is_degenerate = degeneracy_projector(AtA, sh.matP);
                    }
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                sh.n_sel = n_sel;
                if (few) {
                    sh.done = 2;
                } else {
                    float X[6];
#pragma unroll
                    for (int i = 0; i < 6; ++i) X[i] = sh.X[i];
                    if (sh.is_degenerate) {                                             // :1156-1160
                        float X2[6];
                        for (int i = 0; i < 6; ++i) {
                            double sd = 0.0;
                            for (int j = 0; j < 6; ++j) sd += (double)sh.matP[i * 6 + j] * (double)X[j];
                            X2[i] = (float)sd;
                        }
                        for (int i = 0; i < 6; ++i) X[i] = X2[i];
                    }
                    for (int i = 0; i < 6; ++i) sh.pose[i] += X[i];                     // :1162-1167
                    if (blockIdx.x == 0 && iter < 32) {
                        float* tr = a.result->trace[iter];
                        tr[0] = (float)n_sel;
                        for (int i = 0; i < 6; ++i) { tr[1 + i] = X[i]; tr[7 + i] = sh.pose[i]; }
                        tr[13] = (float)sh.sums[21]; tr[14] = (float)sh.sums[22]; tr[15] = (float)sh.sums[23];
                    }
                    const float r0 = X[0] * 57.29578f, r1 = X[1] * 57.29578f, r2 = X[2] * 57.29578f;
                    const float t0 = X[3] * 100, t1 = X[4] * 100, t2 = X[5] * 100;
                    const float deltaR = (float)sqrt((double)r0 * (double)r0 + (double)r1 * (double)r1 + (double)r2 * (double)r2);
                    const float deltaT = (float)sqrt((double)t0 * (double)t0 + (double)t1 * (double)t1 + (double)t2 * (double)t2);
                    if ((double)deltaR < 0.05 && (double)deltaT < 0.05) sh.done = 1;    // :1177
                }
            }
            __syncthreads();
            S2M_CLK(4)
            if (sh.done == 0) s2m_refresh_transform(sh);
            S2M_CLK(5)
            if (marking && iter < 32) a.marks[3 + 3 * iter] = s2m_global_ns();
#ifdef LILOC_S2M_PROFILE
            if (blockIdx.x == 0 && threadIdx.x == 0 && iter < 32)
            {
                for (int i = 0; i < 5; ++i) a.result->clk[iter][i] = (int)(clk[i + 1] - clk[i]);
                a.result->clk[iter][5] = (int)(sh.dbg[1] - sh.dbg[0]);     // kNN
                a.result->clk[iter][6] = (int)(sh.dbg[2] - sh.dbg[1]);     // fit + row (thread 0)
                a.result->clk[iter][7] = (int)(sh.dbg[3] - sh.dbg[2]);     // wait for the slowest group
            }
#endif
            iters = iter + 1;
            if (sh.done == 1) { converged = 1; break; }
            if (sh.done == 2) { too_few = 1; break; }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        S2mResult* r = a.result;
        r->iters = iters; r->converged = converged; r->n_sel = sh.n_sel; r->is_degenerate = sh.is_degenerate;
        r->too_few = too_few; r->skipped = skipped ? 1 : 0; r->q_all = q_all; r->map_points = M;
        r->extra[0] = q_corner; r->extra[1] = a.enable_corner ? *a.f[1].map_count : 0;
        r->extra[2] = n_tiles_c; r->extra[3] = n_tiles - n_tiles_c;
        for (int i = 0; i < 36; ++i) r->matP[i] = sh.matP[i];
        for (int i = 0; i < 6; ++i) { r->pose[i] = sh.pose[i]; a.pose_io[i] = sh.pose[i]; }
        a.state_out->is_degenerate = sh.is_degenerate;
        for (int i = 0; i < 36; ++i) a.state_out->matP[i] = sh.matP[i];
        if (a.marks) a.marks[kS2mMarks - 1] = s2m_global_ns();
        if (a.done_flag) {                   // the host polls this word instead of waiting for the kernel to retire
            __threadfence_system();
            *(volatile unsigned*)a.done_flag = a.done_value;
        }
    }
}

// surfOptimization / cornerOptimization at a fixed pose for the residual-level parity tests (liloc_surf_pass,
// liloc_feature_pass): the same tile function as the registration kernel, with the coefficients written out
// instead of reduced.
__global__ void __launch_bounds__(kS2mThreads)
k_feature_pass(S2mArgs a, int cls, const float* __restrict__ pose6, float4* __restrict__ coeff_out, unsigned char* __restrict__ flag_out) {
    __shared__ S2mShared sh;
    if (threadIdx.x == 0) { sh.fv[cls] = a.f[cls]; sh.gp[cls] = *a.f[cls].grid.p; }
    if (threadIdx.x < 6) sh.pose[threadIdx.x] = pose6[threadIdx.x];
    __syncthreads();
    const int q_all = *sh.fv[cls].q_all;
    const int nq = (q_all + sh.fv[cls].stride - 1) / sh.fv[cls].stride;
    const int n_tiles = (nq + kQPB - 1) / kQPB;
    s2m_refresh_transform(sh);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        s2m_tile_knn(sh, cls, tile, nq, 0);
        __syncthreads();
        s2m_batch_rows(sh, kQPB, coeff_out, flag_out);
        __syncthreads();
    }
}

}  // namespace liloc
