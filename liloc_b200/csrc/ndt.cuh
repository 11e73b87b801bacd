// liloc_b200/csrc/ndt.cuh -- relo-mode NDT registration on the device, batched over independent jobs.
//
// Replaces (reference): registration->setInputTarget / setInputSource / align / getFinalTransformation
//   src/mapOptmization.cpp:273-287 (relocalisation init, 2 x align) and
//   include/egoOptimization/anchorOptimize.hpp:394-407 (per-frame scan-matching factor),
// which the reference delegates to ndt_omp (pclomp::NormalDistributionsTransform, un-vendored; algorithm
// restated in SURVEY.md A.6 and oracle/ndt_oracle.cpp, whose operation order this file follows).
//
// Layout in HBM per target (prior submap): a dense cell -> leaf table (int32) over the target's voxel
// bounding box and one 64-byte leaf record {mean double3, inverse covariance float9, count}.  A job is one
// (target, source scan, initial guess) alignment; jobs never interact, which is what the multi-hypothesis
// relocalisation (cfg 4) and the multi-submap JFGO matching (cfg 5) batch over.
//
// Kernels:  k_ndt_leaves   voxel runs -> Gaussian leaf records (sequential double sums in source order)
//           k_ndt_align    the whole batch as one persistent cooperative kernel: derivative sweeps (score, gradient 6,
//                          Hessian 21; a tile = 1024 points of one job, float per-point terms, double
//                          accumulation, one shuffle + shared-memory reduction per tile), then per job a
//                          fixed-order reduction of the tile partials and ONE step of the Newton / More-Thuente
//                          state machine (double), then compaction of the unfinished-job list; no host polling
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"
#include "voxelgrid.cuh"
#include "pipeline.cuh"

namespace liloc {

namespace cg = cooperative_groups;

struct NdtLeaf { double mean[3]; float icov[9]; int n; int pad; };   // 64 bytes

struct NdtTargetView {
    const int* table;          // dense cell -> leaf index, -1 none
    const NdtLeaf* leaves;
    float inv_leaf;
    int minb[3], maxb[3], divb[3];
};

struct NdtSourceView { const float4* pts; int n; };

struct NdtOptsDev {
    double d1, d2;             // gauss_d1_, gauss_d2_
    double step_max, step_min, eps;
    int max_iterations, search, n_align;      // search: 0 DIRECT1, 1 DIRECT7, 2 KDTREE
    float radius2;                            // KDTREE: resolution^2 (float, like pcl's radius * radius)
};

enum { NDT_AWAIT_FIRST = 0, NDT_LS_FIRST = 1, NDT_LS_LOOP = 2, NDT_DONE = 3 };

struct NdtJob {
    int target, source, phase, align_pass;
    int nr_iterations, converged, sweeps, step_iterations, open_interval, interval_converged;
    double p[6], x_t[6], dir[6];
    double score, phi_0, d_phi_0, a_l, f_l, g_l, a_u, f_u, g_u, a_t;
    float Tfinal[12];
    float guess[12];
    int total_iterations, total_sweeps;
    int n_tiles, pad_;            // ceil(source points / kNdtTile), set by the host
};

constexpr int kNdtThreads = 256;
constexpr int kNdtPPT = 4;                            // source points per thread and tile
constexpr int kNdtTile = kNdtThreads * kNdtPPT;       // 1024 source points per tile
constexpr int kNdtSums = 28;      // score, gradient[6], Hessian upper triangle[21]

// ---- small dense algebra in double, static indices only -------------------------------------------------
template <int N, int P, int Q>
__host__ __device__ __forceinline__ void jacobi_rotate_d(double (&A)[N][N], double (&U)[N][N]) {
    const double apq = A[P][Q];
    if (apq == 0.0) return;
    const double theta = (A[Q][Q] - A[P][P]) / (2.0 * apq);
    const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
    const double c = 1.0 / sqrt(t * t + 1.0), s = t * c, h = t * apq;
    A[P][P] -= h; A[Q][Q] += h; A[P][Q] = 0.0; A[Q][P] = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        if (k == P || k == Q) continue;
        const double akp = A[k][P], akq = A[k][Q];
        const double np_ = c * akp - s * akq, nq_ = s * akp + c * akq;
        A[k][P] = np_; A[P][k] = np_; A[k][Q] = nq_; A[Q][k] = nq_;
    }
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double ukp = U[k][P], ukq = U[k][Q];
        U[k][P] = c * ukp - s * ukq; U[k][Q] = s * ukp + c * ukq;
    }
}

template <int N>
__host__ __device__ __forceinline__ bool jacobi_converged_d(const double (&A)[N][N]) {
    double off = 0.0, diag = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        diag += A[i][i] * A[i][i];
#pragma unroll
        for (int j = i + 1; j < N; ++j) off += A[i][j] * A[i][j];
    }
    return off <= 1e-300 || off <= (2.220446049250313e-16 * 2.220446049250313e-16) * diag;
}

// eigenvalues w, eigenvectors = columns of U
__host__ __device__ inline void jacobi3_d(double (&A)[3][3], double (&U)[3][3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) U[i][j] = (i == j) ? 1.0 : 0.0;
#pragma unroll 1
    for (int sweep = 0; sweep < 60; ++sweep) {
        if (jacobi_converged_d<3>(A)) break;
        jacobi_rotate_d<3, 0, 1>(A, U); jacobi_rotate_d<3, 0, 2>(A, U); jacobi_rotate_d<3, 1, 2>(A, U);
    }
}

__host__ __device__ inline void jacobi6_d(double (&A)[6][6], double (&U)[6][6]) {
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) U[i][j] = (i == j) ? 1.0 : 0.0;
#pragma unroll 1
    for (int sweep = 0; sweep < 60; ++sweep) {
        if (jacobi_converged_d<6>(A)) break;
        jacobi_rotate_d<6, 0, 1>(A, U); jacobi_rotate_d<6, 0, 2>(A, U); jacobi_rotate_d<6, 0, 3>(A, U); jacobi_rotate_d<6, 0, 4>(A, U); jacobi_rotate_d<6, 0, 5>(A, U);
        jacobi_rotate_d<6, 1, 2>(A, U); jacobi_rotate_d<6, 1, 3>(A, U); jacobi_rotate_d<6, 1, 4>(A, U); jacobi_rotate_d<6, 1, 5>(A, U);
        jacobi_rotate_d<6, 2, 3>(A, U); jacobi_rotate_d<6, 2, 4>(A, U); jacobi_rotate_d<6, 2, 5>(A, U);
        jacobi_rotate_d<6, 3, 4>(A, U); jacobi_rotate_d<6, 3, 5>(A, U);
        jacobi_rotate_d<6, 4, 5>(A, U);
    }
}

// ---- pose algebra at the 4x4 boundary (same operation order as the oracle) ------------------------------
__host__ __device__ inline void ndt_matrix_from_sincos(const double* p, float ca, float sa, float cb, float sb, float cc, float sc, float* T);
__host__ __device__ inline void ndt_pose_to_matrix(const double* p, float* T) {
    const float a = (float)p[3], b = (float)p[4], c = (float)p[5];
    const float ca = (float)cos((double)a), sa = (float)sin((double)a);
    const float cb = (float)cos((double)b), sb = (float)sin((double)b);
    const float cc = (float)cos((double)c), sc = (float)sin((double)c);
    ndt_matrix_from_sincos(p, ca, sa, cb, sb, cc, sc, T);
}
__host__ __device__ inline void ndt_matrix_from_sincos(const double* p, float ca, float sa, float cb, float sb, float cc, float sc, float* T) {
    const float Rx[9] = {1, 0, 0, 0, ca, -sa, 0, sa, ca}, Ry[9] = {cb, 0, sb, 0, 1, 0, -sb, 0, cb}, Rz[9] = {cc, -sc, 0, sc, cc, 0, 0, 0, 1};
    float M[9], R[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) M[i * 3 + j] = Rx[i * 3] * Ry[j] + Rx[i * 3 + 1] * Ry[3 + j] + Rx[i * 3 + 2] * Ry[6 + j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) R[i * 3 + j] = M[i * 3] * Rz[j] + M[i * 3 + 1] * Rz[3 + j] + M[i * 3 + 2] * Rz[6 + j];
#pragma unroll
    for (int i = 0; i < 3; ++i) { T[i * 4] = R[i * 3]; T[i * 4 + 1] = R[i * 3 + 1]; T[i * 4 + 2] = R[i * 3 + 2]; T[i * 4 + 3] = (float)p[i]; }
}

__host__ __device__ inline float ndt_atan2f(float y, float x) { return (float)atan2((double)y, (double)x); }

__host__ __device__ inline void ndt_matrix_to_pose(const float* T, double* p) {
    const float m00 = T[0], m01 = T[1], m02 = T[2], m10 = T[4], m11 = T[5], m12 = T[6], m20 = T[8], m21 = T[9], m22 = T[10];
    float r0 = ndt_atan2f(m12, m22);
    const float c2 = sqrtf(m00 * m00 + m01 * m01);
    float r1;
    if (r0 > 0.f) { r0 -= 3.14159265358979323846f; r1 = ndt_atan2f(-m02, -c2); }
    else r1 = ndt_atan2f(-m02, c2);
    const float s1 = (float)sin((double)r0), c1 = (float)cos((double)r0);
    const float r2 = ndt_atan2f(s1 * m20 - c1 * m10, c1 * m11 - s1 * m21);
    p[0] = T[3]; p[1] = T[7]; p[2] = T[11];
    p[3] = -r0; p[4] = -r1; p[5] = -r2;
}

struct NdtAng { float j[8][3]; float h[15][3]; };

__host__ __device__ inline void ndt_angles_from_sincos(double cx, double sx, double cy, double sy, double cz, double sz, NdtAng* A);
__host__ __device__ inline void ndt_angle_derivatives(const double* p, NdtAng* A) {
    double cx, cy, cz, sx, sy, sz;
    if (fabs(p[3]) < 10e-5) { cx = 1.0; sx = 0.0; } else { cx = cos(p[3]); sx = sin(p[3]); }
    if (fabs(p[4]) < 10e-5) { cy = 1.0; sy = 0.0; } else { cy = cos(p[4]); sy = sin(p[4]); }
    if (fabs(p[5]) < 10e-5) { cz = 1.0; sz = 0.0; } else { cz = cos(p[5]); sz = sin(p[5]); }
    ndt_angles_from_sincos(cx, sx, cy, sy, cz, sz, A);
}
__host__ __device__ inline void ndt_angles_from_sincos(double cx, double sx, double cy, double sy, double cz, double sz, NdtAng* A) {
    const double J[8][3] = {
        {-sx * sz + cx * sy * cz, -sx * cz - cx * sy * sz, -cx * cy}, {cx * sz + sx * sy * cz, cx * cz - sx * sy * sz, -sx * cy},
        {-sy * cz, sy * sz, cy}, {sx * cy * cz, -sx * cy * sz, sx * sy}, {-cx * cy * cz, cx * cy * sz, -cx * sy},
        {-cy * sz, -cy * cz, 0}, {cx * cz - sx * sy * sz, -cx * sz - sx * sy * cz, 0}, {sx * cz + cx * sy * sz, cx * sy * cz - sx * sz, 0}};
    const double H[15][3] = {
        {-cx * sz - sx * sy * cz, -cx * cz + sx * sy * sz, sx * cy}, {-sx * sz + cx * sy * cz, -cx * sy * sz - sx * cz, -cx * cy},
        {cx * cy * cz, -cx * cy * sz, cx * sy}, {sx * cy * cz, -sx * cy * sz, sx * sy},
        {-sx * cz - cx * sy * sz, sx * sz - cx * sy * cz, 0}, {cx * cz - sx * sy * sz, -sx * sy * cz - cx * sz, 0},
        {-cy * cz, cy * sz, sy}, {-sx * sy * cz, sx * sy * sz, sx * cy}, {cx * sy * cz, -cx * sy * sz, -cx * cy},
        {sy * sz, sy * cz, 0}, {-sx * cy * sz, -sx * cy * cz, 0}, {cx * cy * sz, cx * cy * cz, 0},
        {-cy * cz, cy * sz, 0}, {-cx * sz - sx * sy * cz, -cx * cz + sx * sy * sz, 0}, {-sx * sz + cx * sy * cz, -cx * sy * sz - sx * cz, 0}};
    for (int r = 0; r < 8; ++r) for (int c = 0; c < 3; ++c) A->j[r][c] = (float)J[r][c];
    for (int r = 0; r < 15; ++r) for (int c = 0; c < 3; ++c) A->h[r][c] = (float)H[r][c];
}

// ---- target build: one Gaussian leaf per voxel run ---------------------------------------------------------
// skeys / svals: voxel keys sorted (stable) with their source indices.  One thread per run head walks its
// run (double sums of p and p p^T in ascending source index), builds the leaf (ndt_omp VoxelGridCovariance,
// SURVEY A.6) and publishes it in the dense table when it is usable (>= 6 points, finite inverse).
__global__ void __launch_bounds__(256)
k_ndt_leaves(const unsigned* __restrict__ skeys, const unsigned* __restrict__ svals, int n, const float4* __restrict__ pts,
             NdtLeaf* __restrict__ leaves, int* __restrict__ table, int* __restrict__ n_leaves) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned key = skeys[i];
        if (i > 0 && skeys[i - 1] == key) continue;
        double s[3] = {0, 0, 0}, ss[6] = {0, 0, 0, 0, 0, 0};      // xx xy xz yy yz zz
        int e = i;
        do {
            const float4 p = __ldg(&pts[svals[e]]);
            const double x = (double)p.x, y = (double)p.y, z = (double)p.z;
            s[0] += x; s[1] += y; s[2] += z;
            ss[0] += x * x; ss[1] += x * y; ss[2] += x * z; ss[3] += y * y; ss[4] += y * z; ss[5] += z * z;
            ++e;
        } while (e < n && skeys[e] == key);
        const int cnt = e - i;
        if (cnt < 6) continue;
        const double nn = (double)cnt;
        const double mean[3] = {s[0] / nn, s[1] / nn, s[2] / nn};
        const double S[3][3] = {{ss[0], ss[1], ss[2]}, {ss[1], ss[3], ss[4]}, {ss[2], ss[4], ss[5]}};
        double cov[3][3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) cov[a][b] = ((S[a][b] - 2.0 * (s[a] * mean[b])) / nn + mean[a] * mean[b]) * ((nn - 1.0) / nn);
        double A[3][3], U[3][3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) A[a][b] = 0.5 * (cov[a][b] + cov[b][a]);
        jacobi3_d(A, U);
        // ascending eigenvalues, ties by index
        double w[3] = {A[0][0], A[1][1], A[2][2]};
        int o0 = 0, o1 = 1, o2 = 2;
        if (w[o1] < w[o0]) { const int t = o0; o0 = o1; o1 = t; }
        if (w[o2] < w[o1]) { const int t = o1; o1 = o2; o2 = t; }
        if (w[o1] < w[o0]) { const int t = o0; o0 = o1; o1 = t; }
        double ev0 = (o0 == 0 ? w[0] : o0 == 1 ? w[1] : w[2]), ev1 = (o1 == 0 ? w[0] : o1 == 1 ? w[1] : w[2]), ev2 = (o2 == 0 ? w[0] : o2 == 1 ? w[1] : w[2]);
        if (ev0 < 0 || ev1 < 0 || ev2 <= 0) continue;
        const double min_ev = 0.01 * ev2;
        if (ev0 < min_ev) {
            ev0 = min_ev;
            if (ev1 < min_ev) ev1 = min_ev;
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    const double va0 = (o0 == 0 ? U[a][0] : o0 == 1 ? U[a][1] : U[a][2]), vb0 = (o0 == 0 ? U[b][0] : o0 == 1 ? U[b][1] : U[b][2]);
                    const double va1 = (o1 == 0 ? U[a][0] : o1 == 1 ? U[a][1] : U[a][2]), vb1 = (o1 == 0 ? U[b][0] : o1 == 1 ? U[b][1] : U[b][2]);
                    const double va2 = (o2 == 0 ? U[a][0] : o2 == 1 ? U[a][1] : U[a][2]), vb2 = (o2 == 0 ? U[b][0] : o2 == 1 ? U[b][1] : U[b][2]);
                    double acc = 0.0;
                    acc += va0 * ev0 * vb0; acc += va1 * ev1 * vb1; acc += va2 * ev2 * vb2;
                    cov[a][b] = acc;
                }
        }
        const double a = cov[0][0], b = cov[0][1], c = cov[0][2], d = cov[1][0], e2 = cov[1][1], f = cov[1][2], g = cov[2][0], h = cov[2][1], k = cov[2][2];
        const double det = a * (e2 * k - f * h) - b * (d * k - f * g) + c * (d * h - e2 * g);
        const double id = 1.0 / det;
        const double inv[9] = {(e2 * k - f * h) * id, (c * h - b * k) * id, (b * f - c * e2) * id,
                               (f * g - d * k) * id, (a * k - c * g) * id, (c * d - a * f) * id,
                               (d * h - e2 * g) * id, (b * g - a * h) * id, (a * e2 - b * d) * id};
        bool finite = true;
#pragma unroll
        for (int q = 0; q < 9; ++q) finite = finite && isfinite(inv[q]);
        if (!finite) continue;
        const int slot = atomicAdd(n_leaves, 1);
        NdtLeaf L;
        L.mean[0] = mean[0]; L.mean[1] = mean[1]; L.mean[2] = mean[2];
#pragma unroll
        for (int q = 0; q < 9; ++q) L.icov[q] = (float)inv[q];
        L.n = cnt; L.pad = (int)key;
        leaves[slot] = L;
        table[key] = slot;
    }
}

__global__ void __launch_bounds__(256) k_fill_int(int* __restrict__ a, long long n, int v) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) a[i] = v;
}

// ---- derivative sweep ------------------------------------------------------------------------------------------
LILOC_DEV int ndt_lookup(const NdtTargetView& t, float x, float y, float z, int dx, int dy, int dz) {
    const int i = (int)floorf(x * t.inv_leaf) + dx, j = (int)floorf(y * t.inv_leaf) + dy, k = (int)floorf(z * t.inv_leaf) + dz;
    if (i < t.minb[0] || i > t.maxb[0] || j < t.minb[1] || j > t.maxb[1] || k < t.minb[2] || k > t.maxb[2]) return -1;
    return __ldg(&t.table[(size_t)(i - t.minb[0]) + (size_t)(j - t.minb[1]) * t.divb[0] + (size_t)(k - t.minb[2]) * t.divb[0] * t.divb[1]]);
}

// updateDerivatives (compute_hessian = true) for one (point, cell) pair, added into acc[28] (double).
LILOC_DEV void ndt_point_cell(const float (&x)[3], const float (&xt)[3], const NdtLeaf* __restrict__ L, const NdtAng& A,
                              const double d1, const float d2, double (&acc)[kNdtSums]) {
    const double m0 = L->mean[0], m1 = L->mean[1], m2 = L->mean[2];
    float C[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) C[q] = L->icov[q];
    const float q0 = (float)((double)xt[0] - m0), q1 = (float)((double)xt[1] - m1), q2 = (float)((double)xt[2] - m2);
    const float qC[3] = {q0 * C[0] + q1 * C[3] + q2 * C[6], q0 * C[1] + q1 * C[4] + q2 * C[7], q0 * C[2] + q1 * C[5] + q2 * C[8]};
    const float qCq = q0 * qC[0] + q1 * qC[1] + q2 * qC[2];
    float e = (float)exp((double)(-d2 * qCq * 0.5f));
    const float score_inc = (float)(-d1 * (double)e);
    e = d2 * e;
    if (e > 1.f || e < 0.f || e != e) return;
    e = (float)((double)e * d1);
    acc[0] += (double)score_inc;
    // point gradient g (3 x 6): columns 0-2 are the identity, column 3 is (0, xj0, xj1), columns 4 and 5 are dense.
    // ndt_omp multiplies the zeros and ones out; they are skipped here: x * 1 = x, x * 0 = +-0 and s + (+-0) = s leave every
    // finite value unchanged (only the sign of an exact zero can differ, which the double sums do not see), so the
    // result is the same float the full expression produces -- tests/test_gpu_ndt.py::test_derivatives_match_oracle
    // compares against the oracle, which evaluates the full expressions.
    float xj[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) xj[r] = A.j[r][0] * x[0] + A.j[r][1] * x[1] + A.j[r][2] * x[2];
    float xh[15];
#pragma unroll
    for (int r = 0; r < 15; ++r) xh[r] = A.h[r][0] * x[0] + A.h[r][1] * x[1] + A.h[r][2] * x[2];
    float Cg[3][6], qCg[6];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        Cg[r][0] = C[r * 3]; Cg[r][1] = C[r * 3 + 1]; Cg[r][2] = C[r * 3 + 2];
        Cg[r][3] = C[r * 3 + 1] * xj[0] + C[r * 3 + 2] * xj[1];
        Cg[r][4] = C[r * 3] * xj[2] + C[r * 3 + 1] * xj[3] + C[r * 3 + 2] * xj[4];
        Cg[r][5] = C[r * 3] * xj[5] + C[r * 3 + 1] * xj[6] + C[r * 3 + 2] * xj[7];
    }
    qCg[0] = qC[0]; qCg[1] = qC[1]; qCg[2] = qC[2];          // q^T C e_c: the same expression as qC[c]
#pragma unroll
    for (int c = 3; c < 6; ++c) qCg[c] = q0 * Cg[0][c] + q1 * Cg[1][c] + q2 * Cg[2][c];
#pragma unroll
    for (int c = 0; c < 6; ++c) acc[1 + c] += (double)(e * qCg[c]);
    // H_E blocks (i, j) for i, j in {3, 4, 5}: a b c / b d e / c e f, each a 3-vector; a, b, c have a zero first component
    const float hterm[6] = {qC[1] * xh[0] + qC[2] * xh[1], qC[1] * xh[2] + qC[2] * xh[3], qC[1] * xh[4] + qC[2] * xh[5],
                            qC[0] * xh[6] + qC[1] * xh[7] + qC[2] * xh[8], qC[0] * xh[9] + qC[1] * xh[10] + qC[2] * xh[11],
                            qC[0] * xh[12] + qC[1] * xh[13] + qC[2] * xh[14]};
    int k = 7;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = i; j < 6; ++j) {
            float ht = 0.f;
            if (i >= 3) {
                // block index in {a,b,c,d,e,f}: (3,3)=0 (3,4)=1 (3,5)=2 (4,4)=3 (4,5)=4 (5,5)=5
                const int bi = (i == 3) ? (j - 3) : (i == 4) ? (j - 1) : 5;
                ht = hterm[bi];
            }
            float gCg;                                        // g_j^T C g_i
            if (j < 3) gCg = Cg[j][i];
            else if (j == 3) gCg = xj[0] * Cg[1][i] + xj[1] * Cg[2][i];
            else if (j == 4) gCg = xj[2] * Cg[0][i] + xj[3] * Cg[1][i] + xj[4] * Cg[2][i];
            else gCg = xj[5] * Cg[0][i] + xj[6] * Cg[1][i] + xj[7] * Cg[2][i];
            acc[k] += (double)(e * (-d2 * qCg[i] * qCg[j] + ht + gCg));
            ++k;
        }
}

// What a sweep needs of a job's trial pose x_t: the 3x4 transform and the angle-derivative tables.  Evaluated once
// per job and round by the thread that moves x_t (double sin/cos: ~thousands of cycles), not once per tile.
struct NdtPoseCache { float T[12]; NdtAng ang; };
constexpr int kNdtPoseFloats = sizeof(NdtPoseCache) / sizeof(float);
LILOC_DEV void ndt_pose_cache_fill(const double* x_t, NdtPoseCache* pc) { ndt_pose_to_matrix(x_t, pc->T); ndt_angle_derivatives(x_t, &pc->ang); }

// The same by a whole warp: the twelve double sin / cos evaluations (six for the float transform, six for the
// angle-derivative tables) are what the serial version spends its time on; lanes 0..5 do one angle each -- lanes 0..2 the
// float-rounded angles of ndt_pose_to_matrix, lanes 3..5 the double angles of ndt_angle_derivatives with its small-angle
// rule -- and lane 0 assembles the tables from the shuffled results.  Same values as ndt_pose_cache_fill.
LILOC_DEV void ndt_pose_cache_fill_warp(const double* x_t_lane0, bool go, NdtPoseCache* pc) {
    const int lane = threadIdx.x & 31;
    double x[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) x[i] = __shfl_sync(0xffffffffu, go ? x_t_lane0[i] : 0.0, 0);
    double c = 1.0, s = 0.0;
    if (lane < 3) {
        const double ang = (double)(float)(lane == 0 ? x[3] : lane == 1 ? x[4] : x[5]);
        c = cos(ang); s = sin(ang);
    } else if (lane < 6) {
        const double ang = lane == 3 ? x[3] : lane == 4 ? x[4] : x[5];
        if (!(fabs(ang) < 10e-5)) { c = cos(ang); s = sin(ang); }
    }
    double cs[6], sn[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) { cs[i] = __shfl_sync(0xffffffffu, c, i); sn[i] = __shfl_sync(0xffffffffu, s, i); }
    if (lane == 0 && go) {
        ndt_matrix_from_sincos(x, (float)cs[0], (float)sn[0], (float)cs[1], (float)sn[1], (float)cs[2], (float)sn[2], pc->T);
        ndt_angles_from_sincos(cs[3], sn[3], cs[4], sn[4], cs[5], sn[5], &pc->ang);
    }
}

constexpr int kNdtWarpCap = 256;          // per-warp ring of gathered (point, leaf) hits: < 32 left over + 32 * 7 new per round
struct NdtSweepShared {
    NdtPoseCache pc;
    double red[kNdtThreads / 32][kNdtSums];
    int2 hits[kNdtThreads / 32][kNdtWarpCap];      // {source point index, leaf index}
};

// One tile (kNdtTile source points) of one derivative sweep of a job at its trial pose: cell look-ups, warp-level
// compaction of the (point, leaf) hits, float per-hit terms with double accumulation in registers, then ONE
// warp-shuffle + shared-memory reduction per tile -> 28 doubles at `out`.  All threads of the block call this together.
// kPoseStaged: the caller already put the pose tables into sh.pc (and synchronised); otherwise they are copied from `pc`.
template <bool kPoseStaged>
LILOC_DEV void ndt_sweep_tile(const NdtPoseCache* __restrict__ pc, const NdtTargetView& tg, const NdtSourceView& src, const NdtOptsDev& o,
                              const int tile, NdtSweepShared& sh, double* __restrict__ out) {
    __syncthreads();                                      // sh.red / sh.hits may still be read by the previous tile's tail
    if (!kPoseStaged) {
        if (threadIdx.x < kNdtPoseFloats) reinterpret_cast<float*>(&sh.pc)[threadIdx.x] = reinterpret_cast<const float*>(pc)[threadIdx.x];
        __syncthreads();
    }
    double acc[kNdtSums];
#pragma unroll
    for (int k = 0; k < kNdtSums; ++k) acc[k] = 0.0;
    const int base = tile * kNdtTile;
    const int ncell = o.search == 2 ? 27 : (o.search == 1 ? 7 : 1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* T = sh.pc.T;
    // Only about half of the points fall into a voxel that holds a Gaussian; evaluated in place, half of every warp
    // idles through the ~800 instructions of ndt_point_cell (ncu: 17 of 32 lanes active on average).  Each warp
    // therefore gathers its (point, leaf) hits into a small ring in shared memory with ballots -- no block barrier --
    // and evaluates them 32 at a time with every lane busy.
    int2* ring = sh.hits[warp];
    int head = 0, count = 0;                              // warp-uniform
    auto evaluate = [&](const int2 h) {
        const float4 p = __ldg(&src.pts[h.x]);
        const float x[3] = {p.x, p.y, p.z};
        const float xt[3] = {T[0] * p.x + T[1] * p.y + T[2] * p.z + T[3],
                             T[4] * p.x + T[5] * p.y + T[6] * p.z + T[7],
                             T[8] * p.x + T[9] * p.y + T[10] * p.z + T[11]};
        ndt_point_cell(x, xt, &tg.leaves[h.y], sh.pc.ang, o.d1, (float)o.d2, acc);
    };
    if (ncell == 1) {
        // DIRECT1: the four cell look-ups of a thread (dependent L2 gathers) are issued together, then gathered
        int li[kNdtPPT];
#pragma unroll
        for (int r = 0; r < kNdtPPT; ++r) {
            const int i = base + r * kNdtThreads + threadIdx.x;
            li[r] = -1;
            if (i < src.n) {
                const float4 p = __ldg(&src.pts[i]);
                li[r] = ndt_lookup(tg, T[0] * p.x + T[1] * p.y + T[2] * p.z + T[3], T[4] * p.x + T[5] * p.y + T[6] * p.z + T[7],
                                   T[8] * p.x + T[9] * p.y + T[10] * p.z + T[11], 0, 0, 0);
            }
        }
#pragma unroll
        for (int r = 0; r < kNdtPPT; ++r) {
            const unsigned m = __ballot_sync(0xffffffffu, li[r] >= 0);
            if (li[r] >= 0) ring[count + __popc(m & ((1u << lane) - 1u))] = make_int2(base + r * kNdtThreads + threadIdx.x, li[r]);
            count += __popc(m);
        }
        __syncwarp();
        while (count >= 32) { evaluate(ring[head + lane]); head += 32; count -= 32; }
    } else {
#pragma unroll 1
        for (int r = 0; r < kNdtPPT; ++r) {
            const int i = base + r * kNdtThreads + threadIdx.x;
            const bool valid = i < src.n;
            float xt0 = 0.f, xt1 = 0.f, xt2 = 0.f;
            if (valid) {
                const float4 p = __ldg(&src.pts[i]);
                xt0 = T[0] * p.x + T[1] * p.y + T[2] * p.z + T[3];
                xt1 = T[4] * p.x + T[5] * p.y + T[6] * p.z + T[7];
                xt2 = T[8] * p.x + T[9] * p.y + T[10] * p.z + T[11];
            }
            for (int c = 0; c < ncell; ++c) {
                int dx, dy, dz;
                if (ncell == 7) {       // DIRECT7 displacement order: centre, +x, -x, +y, -y, +z, -z
                    dx = (c == 1) - (c == 2); dy = (c == 3) - (c == 4); dz = (c == 5) - (c == 6);
                } else {                // KDTREE: radiusSearch(x_trans, resolution) over the leaf centroids.  A centroid closer
                    dx = c % 3 - 1; dy = (c / 3) % 3 - 1; dz = c / 9 - 1;       // than one voxel edge lies in the 3x3x3 block around the point's voxel
                }
                int li = valid ? ndt_lookup(tg, xt0, xt1, xt2, dx, dy, dz) : -1;
                if (ncell == 27 && li >= 0) {
                    const NdtLeaf* L = &tg.leaves[li];
                    const float ex = (float)__ldg(&L->mean[0]) - xt0, ey = (float)__ldg(&L->mean[1]) - xt1, ez = (float)__ldg(&L->mean[2]) - xt2;
                    if (!(ex * ex + ey * ey + ez * ez < o.radius2)) li = -1;     // FLANN radius search: squared distance < radius^2
                }
                const unsigned m = __ballot_sync(0xffffffffu, li >= 0);
                if (li >= 0) ring[(head + count + __popc(m & ((1u << lane) - 1u))) % kNdtWarpCap] = make_int2(i, li);
                count += __popc(m);
                if (c % 7 == 6 || c == ncell - 1) {     // drain: at most 31 left over + 7 x 32 new entries fit the ring
                    __syncwarp();
                    while (count >= 32) {
                        evaluate(ring[(head + lane) % kNdtWarpCap]);
                        head = (head + 32) % kNdtWarpCap; count -= 32;
                    }
                    __syncwarp();
                }
            }
        }
    }
    if (lane < count) evaluate(ring[(head + lane) % kNdtWarpCap]);
    // Warp reduction of the 28 sums as a reduce-scatter butterfly: at every level a lane keeps one half of its values
    // and sends the other half to its partner (62 shuffles and 31 adds instead of 280 and 140 for 28 separate trees);
    // afterwards lane k holds the warp total of sum k.
    {
        double v16[16];
        const bool b16 = (lane & 16) != 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const double lo = acc[j], hi = (j + 16 < kNdtSums) ? acc[j + 16] : 0.0;
            v16[j] = (b16 ? hi : lo) + shfl_xor_d(b16 ? lo : hi, 16);
        }
        double v8[8];
        const bool b8 = (lane & 8) != 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) v8[j] = (b8 ? v16[j + 8] : v16[j]) + shfl_xor_d(b8 ? v16[j] : v16[j + 8], 8);
        double v4[4];
        const bool b4 = (lane & 4) != 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) v4[j] = (b4 ? v8[j + 4] : v8[j]) + shfl_xor_d(b4 ? v8[j] : v8[j + 4], 4);
        const bool b2 = (lane & 2) != 0, b1 = (lane & 1) != 0;
        double v2[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) v2[j] = (b2 ? v4[j + 2] : v4[j]) + shfl_xor_d(b2 ? v4[j] : v4[j + 2], 2);
        const double v1 = (b1 ? v2[1] : v2[0]) + shfl_xor_d(b1 ? v2[0] : v2[1], 1);
        if (lane < kNdtSums) sh.red[warp][lane] = v1;
    }
    __syncthreads();
    if (threadIdx.x < kNdtSums) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kNdtThreads / 32; ++w) v += sh.red[w][threadIdx.x];
        out[threadIdx.x] = v;
    }
}

// Stand-alone sweep of ONE job (liloc_ndt_derivatives, parity tests): grid = point tiles.
__global__ void __launch_bounds__(kNdtThreads)
k_ndt_sweep(const NdtJob* __restrict__ jobs, const NdtTargetView* __restrict__ targets, const NdtSourceView* __restrict__ sources,
            NdtOptsDev o, double* __restrict__ partial) {
    __shared__ NdtSweepShared sh;
    __shared__ NdtPoseCache s_pc;
    const NdtJob& J = jobs[0];
    if (threadIdx.x == 0) ndt_pose_cache_fill(J.x_t, &s_pc);
    __syncthreads();
    ndt_sweep_tile<false>(&s_pc, targets[J.target], sources[J.source], o, blockIdx.x, sh, partial + (size_t)blockIdx.x * kNdtSums);
}

// ---- More-Thuente helpers ------------------------------------------------------------------------------------
LILOC_DEV double ndt_psi(double a, double f_a, double f_0, double g_0, double mu) { return f_a - f_0 - mu * g_0 * a; }
LILOC_DEV double ndt_dpsi(double g_a, double g_0, double mu) { return g_a - mu * g_0; }

LILOC_DEV bool ndt_update_interval(double& a_l, double& f_l, double& g_l, double& a_u, double& f_u, double& g_u, double a_t, double f_t, double g_t) {
    if (f_t > f_l) { a_u = a_t; f_u = f_t; g_u = g_t; return false; }
    else if (g_t * (a_l - a_t) > 0) { a_l = a_t; f_l = f_t; g_l = g_t; return false; }
    else if (g_t * (a_l - a_t) < 0) { a_u = a_l; f_u = f_l; g_u = g_l; a_l = a_t; f_l = f_t; g_l = g_t; return false; }
    return true;
}

LILOC_DEV double ndt_trial_value(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u, double a_t, double f_t, double g_t) {
    if (f_t > f_l) {
        const double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l, w = sqrt(z * z - g_t * g_l);
        const double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
        const double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
        return (fabs(a_c - a_l) < fabs(a_q - a_l)) ? a_c : 0.5 * (a_q + a_c);
    } else if (g_t * g_l < 0) {
        const double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l, w = sqrt(z * z - g_t * g_l);
        const double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
        const double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
        return (fabs(a_c - a_t) >= fabs(a_s - a_t)) ? a_c : a_s;
    } else if (fabs(g_t) <= fabs(g_l)) {
        const double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l, w = sqrt(z * z - g_t * g_l);
        const double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
        const double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
        const double a_n = (fabs(a_c - a_t) < fabs(a_s - a_t)) ? a_c : a_s;
        return (a_t > a_l) ? fmin(a_t + 0.66 * (a_u - a_t), a_n) : fmax(a_t + 0.66 * (a_u - a_t), a_n);
    }
    const double z = 3 * (f_t - f_u) / (a_t - a_u) - g_t - g_u, w = sqrt(z * z - g_t * g_u);
    return a_u + (a_t - a_u) * (w - g_u - z) / (g_t - g_u + 2 * w);
}

// JacobiSVD(H).solve(-g) for the symmetrised Hessian (upper triangle Hup[21]).
// Fast path: when H is well conditioned the minimum-norm solution IS H^-1 (-g), so it is computed by Gaussian
// elimination with partial pivoting in double (pivot ratio > 1e-10; agrees with the SVD solution to ~1e-6 of the
// tolerance the line search works at).  Otherwise the Jacobi eigen-decomposition gives the pseudo-inverse exactly
// as JacobiSVD would: directions whose singular value is below 6 eps smax are dropped.
LILOC_DEV void ndt_newton_solve(const double* Hup, const double* g, double* dp) {
    {
        // every index below is a compile-time constant (full unrolling, row swaps as selects) so that the 6x7 system
        // can stay in registers instead of being indexed in local memory
        double M[6][7];
        {
            int k = 0;
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = i; j < 6; ++j) { M[i][j] = Hup[k]; M[j][i] = Hup[k]; ++k; }
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) M[i][6] = -g[i];
        double pmin = 1e300, pmax = 0.0;
        bool ok = true;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            int pr = c; double best = fabs(M[c][c]);
#pragma unroll
            for (int r = c + 1; r < 6; ++r) { const double v = fabs(M[r][c]); if (v > best) { best = v; pr = r; } }
            if (!(best > 0.0) || best != best) ok = false;
#pragma unroll
            for (int r = c + 1; r < 6; ++r) {
                const bool sw = pr == r;
#pragma unroll
                for (int j = c; j < 7; ++j) { const double a0 = M[c][j], a1 = M[r][j]; M[c][j] = sw ? a1 : a0; M[r][j] = sw ? a0 : a1; }
            }
            pmin = fmin(pmin, best); pmax = fmax(pmax, best);
            const double inv = 1.0 / M[c][c];
#pragma unroll
            for (int r = c + 1; r < 6; ++r) {
                const double f = M[r][c] * inv;
#pragma unroll
                for (int j = c + 1; j < 7; ++j) M[r][j] -= f * M[c][j];
            }
        }
        if (ok && pmin > 1e-10 * pmax) {
            double x[6];
#pragma unroll
            for (int i = 5; i >= 0; --i) {
                double v = M[i][6];
#pragma unroll
                for (int j = i + 1; j < 6; ++j) v -= M[i][j] * x[j];
                x[i] = v / M[i][i];
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) dp[i] = x[i];
            return;
        }
    }
    double A[6][6], U[6][6];
    int k = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = i; j < 6; ++j) { A[i][j] = Hup[k]; A[j][i] = Hup[k]; ++k; }
    jacobi6_d(A, U);
    double smax = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) smax = fmax(smax, fabs(A[i][i]));
    const double thr = fmax(smax * 6.0 * 2.220446049250313e-16, 2.2250738585072014e-308);
#pragma unroll
    for (int i = 0; i < 6; ++i) dp[i] = 0.0;
#pragma unroll
    for (int m = 0; m < 6; ++m) {
        if (!(fabs(A[m][m]) > thr)) continue;
        double proj = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) proj += U[i][m] * (-g[i]);
        proj /= A[m][m];
#pragma unroll
        for (int i = 0; i < 6; ++i) dp[i] += U[i][m] * proj;
    }
}

// One step of computeTransformation / computeStepLengthMT given the sums of the sweep at J.x_t.  Returns true when a NEW
// trial pose x_t was set (the caller then evaluates T(x_t) -- final_transformation_ of the line search -- and the angle
// tables for the next sweep); the 3x4 transform itself is not computed here (six double sin / cos on one lane).
__device__ __noinline__ bool ndt_job_step(NdtJob& J, const double* sums, const NdtOptsDev& o) {
    const double mu = 1.e-4, nu = 0.9;
    ++J.sweeps;
    J.score = sums[0];
    bool need_check = false;
    double phi_t = 0, d_phi_t = 0, psi_t = 0, d_psi_t = 0;
    if (J.phase == NDT_LS_FIRST || J.phase == NDT_LS_LOOP) {
        phi_t = -J.score;
        d_phi_t = 0.0;
        for (int i = 0; i < 6; ++i) d_phi_t -= sums[1 + i] * J.dir[i];
        psi_t = ndt_psi(J.a_t, phi_t, J.phi_0, J.d_phi_0, mu);
        d_psi_t = ndt_dpsi(d_phi_t, J.d_phi_0, mu);
        if (J.phase == NDT_LS_LOOP) {
            if (J.open_interval && (psi_t <= 0 && d_psi_t >= 0)) {
                J.open_interval = 0;
                J.f_l = J.f_l + J.phi_0 - mu * J.d_phi_0 * J.a_l; J.g_l = J.g_l + mu * J.d_phi_0;
                J.f_u = J.f_u + J.phi_0 - mu * J.d_phi_0 * J.a_u; J.g_u = J.g_u + mu * J.d_phi_0;
            }
            if (J.open_interval) J.interval_converged = ndt_update_interval(J.a_l, J.f_l, J.g_l, J.a_u, J.f_u, J.g_u, J.a_t, psi_t, d_psi_t) ? 1 : 0;
            else J.interval_converged = ndt_update_interval(J.a_l, J.f_l, J.g_l, J.a_u, J.f_u, J.g_u, J.a_t, phi_t, d_phi_t) ? 1 : 0;
            J.step_iterations++;
        } else {
            J.step_iterations = 0;
        }
        need_check = true;
    }
    for (int guard = 0; guard < 64; ++guard) {
        if (need_check) {
            if (!J.interval_converged && J.step_iterations < 10 && !(psi_t <= 0 && d_phi_t <= -nu * J.d_phi_0)) {
                if (J.open_interval) J.a_t = ndt_trial_value(J.a_l, J.f_l, J.g_l, J.a_u, J.f_u, J.g_u, J.a_t, psi_t, d_psi_t);
                else J.a_t = ndt_trial_value(J.a_l, J.f_l, J.g_l, J.a_u, J.f_u, J.g_u, J.a_t, phi_t, d_phi_t);
                J.a_t = fmin(J.a_t, o.step_max); J.a_t = fmax(J.a_t, o.step_min);
                for (int i = 0; i < 6; ++i) J.x_t[i] = J.p[i] + J.dir[i] * J.a_t;
                J.phase = NDT_LS_LOOP;
                return true;                              // next sweep at the new trial point
            }
            need_check = false;
            // line search finished: accept a_t
            for (int i = 0; i < 6; ++i) J.p[i] += J.dir[i] * J.a_t;
            const bool conv = J.nr_iterations > o.max_iterations || (J.nr_iterations && (fabs(J.a_t) < o.eps));
            J.nr_iterations++;
            if (conv) { J.converged = 1; J.phase = NDT_DONE; return false; }
        }
        // Newton direction from the latest sums (they belong to the current p)
        double dp[6];
        ndt_newton_solve(sums + 7, sums + 1, dp);
        double norm = 0.0;
        for (int i = 0; i < 6; ++i) norm += dp[i] * dp[i];
        norm = sqrt(norm);
        if (norm == 0 || norm != norm) { J.converged = (norm == norm) ? 1 : 0; J.phase = NDT_DONE; return false; }
        for (int i = 0; i < 6; ++i) J.dir[i] = dp[i] / norm;
        J.phi_0 = -J.score;
        J.d_phi_0 = 0.0;
        for (int i = 0; i < 6; ++i) J.d_phi_0 -= sums[1 + i] * J.dir[i];
        if (J.d_phi_0 >= 0) {
            if (J.d_phi_0 == 0) {                          // zero step: p unchanged, iteration counted
                J.a_t = 0.0;
                const bool conv = J.nr_iterations > o.max_iterations || (J.nr_iterations && (0.0 < o.eps));
                J.nr_iterations++;
                if (conv) { J.converged = 1; J.phase = NDT_DONE; return false; }
                continue;
            }
            J.d_phi_0 *= -1;
            for (int i = 0; i < 6; ++i) J.dir[i] *= -1;
        }
        J.a_l = 0; J.a_u = 0;
        J.f_l = ndt_psi(J.a_l, J.phi_0, J.phi_0, J.d_phi_0, mu); J.g_l = ndt_dpsi(J.d_phi_0, J.d_phi_0, mu);
        J.f_u = ndt_psi(J.a_u, J.phi_0, J.phi_0, J.d_phi_0, mu); J.g_u = ndt_dpsi(J.d_phi_0, J.d_phi_0, mu);
        J.interval_converged = ((o.step_max - o.step_min) < 0) ? 1 : 0;
        J.open_interval = 1;
        J.a_t = norm;
        J.a_t = fmin(J.a_t, o.step_max); J.a_t = fmax(J.a_t, o.step_min);
        for (int i = 0; i < 6; ++i) J.x_t[i] = J.p[i] + J.dir[i] * J.a_t;
        J.phase = NDT_LS_FIRST;
        return true;
    }
    J.phase = NDT_DONE;                                   // unreachable in practice (guard)
    return false;
}

// start (or restart, for the second of the reference's two back-to-back align calls) from a 3x4 guess
LILOC_DEV void ndt_job_begin(NdtJob& J, const float* guess12) {
    for (int i = 0; i < 12; ++i) J.Tfinal[i] = guess12[i];
    ndt_matrix_to_pose(guess12, J.p);
    for (int i = 0; i < 6; ++i) J.x_t[i] = J.p[i];
    J.phase = NDT_AWAIT_FIRST;
    J.nr_iterations = 0; J.converged = 0; J.sweeps = 0; J.step_iterations = 0; J.open_interval = 1; J.interval_converged = 0;
}

// ---- the whole batch of alignments as ONE persistent kernel driven by a ticket queue ----------------------------------------
// Round 1 ran the batch in lock-step (sweep phase / grid barrier / update phase / grid barrier / compaction): every round
// waited for its slowest job, and a batch lasted as many rounds as its longest job needs whatever the number of GPUs
// (profiles/frame_latency_r01.md).  Now jobs advance INDEPENDENTLY:
//   * a ticket = (job, first tile, tile count); tickets sit in a ring in global memory; a block claims the next ticket
//     with ONE atomicAdd on `head` and waits for that slot to be published (entries are tagged with their ticket number, so
//     a block may hold a ticket that does not exist yet -- that is how idle blocks wait for work);
//   * the block sweeps the ticket's tiles (ndt_sweep_tile: 28 doubles per tile into `partial`), then adds the tile count to
//     the job's arrival counter; the block that completes the sweep (all tiles arrived) reduces the tile partials in tile
//     order (the canonical summation order shared with the oracle), runs ONE step of the Newton / More-Thuente state
//     machine on a shared-memory copy of the job and either publishes the tickets of the job's next sweep or retires the job;
//   * no grid barrier, no cooperative launch, no compaction: a job's sweep starts as soon as its previous one is
//     reduced, on whatever blocks are free.  Ticket granularity adapts: whole jobs while there are several unfinished
//     jobs per block (no atomics inside a sweep), single tiles when blocks would otherwise idle (a job's sweep then
//     spreads over up to n_tiles blocks).
// Results are independent of the schedule: every sum has a fixed order.
struct NdtSync { int done; int pad; };        // per job: tiles of the current sweep already in `partial`

struct NdtQueue {
    unsigned long long* ring;      // [cap] (ticket + 1) << 32 | 1 << 31 | job << 10 | first tile
    unsigned mask;                 // cap - 1, cap a power of two >= the tiles of all jobs + the blocks of the launch
    unsigned* head;                // next ticket to hand out
    unsigned* tail;                // tickets published so far
    int* finished;                 // jobs retired
    int* error;                    // a wait gave up (1) -- every block leaves
};
constexpr int kNdtMaxTilesPerJob = 1024;      // 10 bits of a ring entry (1 M source points)
constexpr int kNdtMaxJobs = 1 << 21;

struct NdtJobView {                // the constant part of what a tile needs of its job
    NdtTargetView tg;
    const float4* pts; int n; int n_tiles;
};
// Everything a ticket needs, contiguous: ONE L2 round trip per ticket (104 threads, one word each) instead of the dependent
// chain job -> source / target views -> pose tables.  view: written once in the prologue; grain, pc: by every update.
struct NdtHot { NdtJobView view; int grain; int pad; NdtPoseCache pc; };
constexpr int kNdtHotWords = sizeof(NdtHot) / sizeof(int);
static_assert(sizeof(NdtHot) % sizeof(int) == 0 && kNdtHotWords <= kNdtThreads, "NdtHot is staged by one word per thread");

struct NdtAlignArgs {
    NdtJob* jobs;
    int n_jobs;
    const NdtTargetView* targets;
    const NdtSourceView* sources;
    NdtOptsDev o;
    int tiles_max;                 // ceil(max source size / kNdtTile)
    double* partial;               // [n_jobs][tiles_max][kNdtSums]
    NdtHot* hot;                   // [n_jobs] view + ticket granularity + transform / angle tables of the job's current trial pose
    NdtSync* sync;                 // [n_jobs]
    NdtQueue q;
    float* rows;                   // [n_jobs][8] packed results {roll, pitch, yaw, x, y, z, score / n, iterations} or nullptr
    int grain_div;                 // tickets aimed at per block over the unfinished jobs (ndt_grain_for)
    unsigned long long* prof;      // optional (zeroed by the host): [0] ~(first block start) (complemented: a maximum of a zeroed word), [1] last block end (globaltimer ns), [2] sum over blocks of
                                   // sweep time, [3] update time, [4] time waiting for a ticket, [5] tickets, [6] updates,
                                   // [8..11] SM cycles inside the updates: reduce + load, state machine, pose tables, store + publish,
                                   // [12] sweeps of all retired jobs, [13] sweeps of the longest job, [14] sum of sweeps x source points
};

LILOC_DEV unsigned long long ndt_ld_entry(const unsigned long long* p) { return *((const volatile unsigned long long*)p); }

// Publish the tickets of one sweep of `job` (whole warp).  The ticket numbers are reserved first (ndt_reserve, lane 0): the
// atomic's round trip then overlaps whatever the warp still has to write; consumers wait on the tags, not on `tail`.
LILOC_DEV unsigned ndt_reserve(const NdtAlignArgs& a, int n_tiles, int grain) {
    return atomicAdd(a.q.tail, (unsigned)((n_tiles + grain - 1) / grain));
}
LILOC_DEV void ndt_publish(const NdtAlignArgs& a, int job, int n_tiles, int grain, unsigned base_lane0) {
    const int lane = threadIdx.x & 31;
    const int n_tick = (n_tiles + grain - 1) / grain;
    __threadfence();                                      // job state, pose tables, arrival counter before the tickets
    __syncwarp();
    const unsigned base = __shfl_sync(0xffffffffu, base_lane0, 0);
    for (int i = lane; i < n_tick; i += 32) {
        const unsigned t = base + (unsigned)i;
        const unsigned long long e = ((unsigned long long)(t + 1u) << 32) | 0x80000000ull | ((unsigned long long)job << 10) | (unsigned long long)(i * grain);
        *((volatile unsigned long long*)&a.q.ring[t & a.q.mask]) = e;
    }
}

// Tiles per ticket.  A ticket costs a claim, a 416-byte descriptor load and an arrival (~2-3 us) on top of its tiles
// (~3.7 us each), so tickets should be long -- but a sweep cut into few tickets cannot spread over idle blocks, and the
// longest job's chain of sweeps is what a small batch waits for.  Aim at ~`div` tickets in flight per block over the
// unfinished jobs: whole sweeps while jobs are plentiful, single tiles once fewer jobs than blocks remain.
LILOC_DEV int ndt_grain_for(int active_jobs, int n_tiles, int grid, int div) {
    long long g = ((long long)active_jobs * n_tiles) / ((long long)div * grid);
    if (g < 1) g = 1;
    if (g > n_tiles) g = n_tiles;
    return (int)g;
}


// RotMtoEuler (include/math_tools.h:92-111) of the rotation block + translation -> [roll, pitch, yaw, x, y, z], evaluated in
// double and rounded to float (what liloc_b200/relo.py::matrices_to_pose6 does on the host)
LILOC_DEV void ndt_pack_row(const NdtJob& J, int n_source, float* row) {
    const double r00 = J.Tfinal[0], r10 = J.Tfinal[4], r20 = J.Tfinal[8], r21 = J.Tfinal[9], r22 = J.Tfinal[10], r11 = J.Tfinal[5], r12 = J.Tfinal[6];
    const double sy = sqrt(r00 * r00 + r10 * r10);
    double roll, pitch, yaw;
    if (sy >= 1e-6) { roll = atan2(r21, r22); pitch = atan2(-r20, sy); yaw = atan2(r10, r00); }
    else { roll = atan2(-r12, r11); pitch = atan2(-r20, sy); yaw = 0.0; }
    row[0] = (float)roll; row[1] = (float)pitch; row[2] = (float)yaw;
    row[3] = J.Tfinal[3]; row[4] = J.Tfinal[7]; row[5] = J.Tfinal[11];
    row[6] = (float)(J.score / (double)(n_source > 0 ? n_source : 1));
    row[7] = (float)J.total_iterations;
}

struct NdtUpdateShared { NdtJob job; double sums[kNdtSums]; };

// The block that completed a job's sweep: fixed-order reduction + one state-machine step + publish / retire (warp 0).
LILOC_DEV void ndt_job_update(const NdtAlignArgs& a, int job, NdtUpdateShared& us) {
    const int lane = threadIdx.x & 31;
    const bool prof = a.prof != nullptr && lane == 0;
    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    if (prof) c0 = clock64();
    NdtHot* hot = &a.hot[job];
    const NdtJobView* vw = &hot->view;
    const int n_tiles = __ldcg(&vw->n_tiles);
    int finished = 0;                                     // read now: in flight together with everything below
    if (lane == 0) finished = *((volatile int*)a.q.finished);
    int jw[(sizeof(NdtJob) / sizeof(int) + 31) / 32];     // the job state, 32 lanes wide: in flight together with the partial sums
    {
        const int* src = reinterpret_cast<const int*>(&a.jobs[job]);
#pragma unroll
        for (int u = 0; u < (int)(sizeof(jw) / sizeof(int)); ++u) { const int i = lane + 32 * u; jw[u] = i < (int)(sizeof(NdtJob) / sizeof(int)) ? __ldcg(src + i) : 0; }
    }
    if (lane < kNdtSums) {
        const double* pj = a.partial + (size_t)job * a.tiles_max * kNdtSums + lane;
        double v = 0.0;
        // 32 (predicated) loads in flight per trip -- an L2 round trip is ~0.5 us and a VLP-16 scan is 29 tiles --, summed in tile order
#pragma unroll 1
        for (int t = 0; t < n_tiles; t += 32) {
            double w[32];
#pragma unroll
            for (int u = 0; u < 32; ++u) w[u] = (t + u < n_tiles) ? __ldcg(&pj[(size_t)(t + u) * kNdtSums]) : 0.0;
#pragma unroll
            for (int u = 0; u < 32; ++u) if (t + u < n_tiles) v += w[u];
        }
        us.sums[lane] = v;
    }
    {
        int* dst = reinterpret_cast<int*>(&us.job);
#pragma unroll
        for (int u = 0; u < (int)(sizeof(jw) / sizeof(int)); ++u) { const int i = lane + 32 * u; if (i < (int)(sizeof(NdtJob) / sizeof(int))) dst[i] = jw[u]; }
    }
    __syncwarp();
    if (prof) c1 = clock64();
    bool new_trial = false, go = false;
    int grain = 1;
    unsigned base = 0;
    double xt[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (lane == 0) {
        NdtJob& L = us.job;
        new_trial = ndt_job_step(L, us.sums, a.o);
        if (L.phase == NDT_DONE) {
            L.total_iterations += L.nr_iterations; L.total_sweeps += L.sweeps;
            if (L.align_pass + 1 < a.o.n_align) {         // registration->align() again from the previous result (:282-290)
                const int pass = L.align_pass + 1;
                float g[12];
                for (int i = 0; i < 12; ++i) g[i] = L.Tfinal[i];
                ndt_job_begin(L, g);
                L.align_pass = pass;
            }
        }
        go = L.phase != NDT_DONE;
#pragma unroll
        for (int i = 0; i < 6; ++i) xt[i] = L.x_t[i];
        if (go) {                                         // ticket numbers and granularity of the next sweep: reserved now, the
            grain = ndt_grain_for(a.n_jobs - finished, n_tiles, (int)gridDim.x, a.grain_div);      // round trip of the atomic overlaps the pose tables and the stores
            base = ndt_reserve(a, n_tiles, grain);
        }
    }
    go = __shfl_sync(0xffffffffu, go ? 1 : 0, 0) != 0;
    new_trial = __shfl_sync(0xffffffffu, new_trial ? 1 : 0, 0) != 0;
    if (prof) c2 = clock64();
    NdtPoseCache* pc = &hot->pc;
    ndt_pose_cache_fill_warp(xt, go, pc);                 // the next sweep's transform and angle tables, six lanes for the sin / cos
    __syncwarp();
    if (prof) c3 = clock64();
    if (lane == 0 && new_trial) {                         // final_transformation_ = T(x_t) of the line search's latest trial
#pragma unroll
        for (int i = 0; i < 12; ++i) us.job.Tfinal[i] = pc->T[i];
    }
    __syncwarp();
    {
        int* dst = reinterpret_cast<int*>(&a.jobs[job]);
        const int* src = reinterpret_cast<const int*>(&us.job);
        for (int i = lane; i < (int)(sizeof(NdtJob) / sizeof(int)); i += 32) dst[i] = src[i];
    }
    if (go) {
        if (lane == 0) {
            *((volatile int*)&a.sync[job].done) = 0;
            *((volatile int*)&hot->grain) = grain;
        }
        grain = __shfl_sync(0xffffffffu, grain, 0);
        ndt_publish(a, job, n_tiles, grain, base);
    } else {
        if (lane == 0 && a.rows) ndt_pack_row(us.job, __ldcg(&vw->n), a.rows + (size_t)job * 8);
        if (lane == 0 && a.prof) {                        // what the host's counters need of a retired job: no job records travel back for them
            const unsigned long long sw = (unsigned long long)us.job.total_sweeps;
            atomicAdd(&a.prof[12], sw); atomicMax(&a.prof[13], sw); atomicAdd(&a.prof[14], sw * (unsigned long long)__ldcg(&vw->n));
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(a.q.finished, 1);
    }
    if (prof) {         // SM clock cycles of the four parts of an update: reduce + load, state machine, pose tables, store + publish
        const long long c4 = clock64();
        atomicAdd(&a.prof[8], (unsigned long long)(c1 - c0)); atomicAdd(&a.prof[9], (unsigned long long)(c2 - c1));
        atomicAdd(&a.prof[10], (unsigned long long)(c3 - c2)); atomicAdd(&a.prof[11], (unsigned long long)(c4 - c3));
    }
}

__global__ void __launch_bounds__(kNdtThreads, 2) k_ndt_align(NdtAlignArgs a) {
    __shared__ NdtSweepShared sh;
    __shared__ NdtUpdateShared us;
    __shared__ NdtJobView s_view;
    __shared__ int s_ticket, s_last, s_grain;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long t_sweep = 0, t_update = 0, t_wait = 0, n_tick = 0, n_upd = 0, t_begin = 0;
    const bool prof = a.prof != nullptr && threadIdx.x == 0;
    if (prof) { t_begin = global_ns(); atomicMax(&a.prof[0], ~t_begin); }
    // ---- prologue: one warp per job -- state, pose tables, view, first tickets.  No barrier: consumers wait on ticket tags.
    {
        const int gwarp = warp * gridDim.x + blockIdx.x, nwarps = (kNdtThreads / 32) * gridDim.x;
        for (int j = gwarp; j < a.n_jobs; j += nwarps) {
            NdtJob& J = a.jobs[j];
            double xt[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            int n_tiles = 0;
            if (lane == 0) {
                J.align_pass = 0; J.total_iterations = 0; J.total_sweeps = 0;
                ndt_job_begin(J, J.guess);
#pragma unroll
                for (int i = 0; i < 6; ++i) xt[i] = J.x_t[i];
                const NdtSourceView sv = a.sources[J.source];
                NdtJobView v;
                v.tg = a.targets[J.target]; v.pts = sv.pts; v.n = sv.n; v.n_tiles = (sv.n + kNdtTile - 1) / kNdtTile;
                a.hot[j].view = v;
                n_tiles = v.n_tiles;
            }
            n_tiles = __shfl_sync(0xffffffffu, n_tiles, 0);
            ndt_pose_cache_fill_warp(xt, true, &a.hot[j].pc);
            int grain = 1;
            unsigned base = 0;
            if (lane == 0) {
                grain = ndt_grain_for(a.n_jobs, n_tiles, (int)gridDim.x, a.grain_div);
                base = ndt_reserve(a, n_tiles, grain);
                a.sync[j].done = 0; a.hot[j].grain = grain; a.hot[j].pad = 0;
            }
            grain = __shfl_sync(0xffffffffu, grain, 0);
            ndt_publish(a, j, n_tiles, grain, base);
        }
    }
    // ---- consumer loop
    for (;;) {
        if (threadIdx.x == 0) {
            unsigned long long t0 = 0;
            if (prof) t0 = global_ns();
            const unsigned my = atomicAdd(a.q.head, 1u);
            const unsigned long long* slot = &a.q.ring[my & a.q.mask];
            const unsigned long long want = ((unsigned long long)(my + 1u) << 32) | 0x80000000ull;
            int ticket = -1;
            for (long long spin = 0;; ++spin) {
                const unsigned long long e = ndt_ld_entry(slot);
                if ((e & 0xFFFFFFFF80000000ull) == want) { ticket = (int)(e & 0x7FFFFFFFull); break; }
                if ((spin & 15) == 15) {
                    if (*((volatile int*)a.q.finished) >= a.n_jobs || *((volatile int*)a.q.error)) break;
                    if (spin > (1ll << 26)) { atomicExch(a.q.error, 1); break; }        // ~ seconds: something is wrong, leave
                    __nanosleep(spin < 256 ? 20 : 200);
                }
            }
            s_ticket = ticket;
            if (prof) t_wait += global_ns() - t0;
        }
        __syncthreads();
        const int ticket = s_ticket;
        if (ticket < 0) break;
        const int job = ticket >> 10, tile0 = ticket & 1023;
        // the job's view, ticket granularity and pose tables (written by other blocks: through L2), one word per thread
        if (threadIdx.x < kNdtHotWords) {
            const int w = __ldcg(reinterpret_cast<const int*>(&a.hot[job]) + threadIdx.x);
            constexpr int kViewWords = sizeof(NdtJobView) / sizeof(int);
            if (threadIdx.x < kViewWords) reinterpret_cast<int*>(&s_view)[threadIdx.x] = w;
            else if (threadIdx.x == kViewWords) s_grain = w;
            else if (threadIdx.x > kViewWords + 1) reinterpret_cast<int*>(&sh.pc)[threadIdx.x - kViewWords - 2] = w;
        }
        __syncthreads();
        unsigned long long t1 = 0;
        if (prof) t1 = global_ns();
        const NdtSourceView src{s_view.pts, s_view.n};
        const int n_tiles = s_view.n_tiles;
        int t_end = tile0 + s_grain;
        if (t_end > n_tiles) t_end = n_tiles;
        for (int t = tile0; t < t_end; ++t) {
            double* out = a.partial + ((size_t)job * a.tiles_max + t) * kNdtSums;
            ndt_sweep_tile<true>(nullptr, s_view.tg, src, a.o, t, sh, out);
        }
        __syncthreads();                                  // every thread's partial sums are written ...
        if (threadIdx.x == 0) {
            const int cnt = t_end - tile0;
            __threadfence();                              // ... and ordered before the arrival (cumulative over the barrier above)
            const int before = atomicAdd(&a.sync[job].done, cnt);
            s_last = (before + cnt == n_tiles) ? 1 : 0;
            if (prof) { const unsigned long long t2 = global_ns(); t_sweep += t2 - t1; t1 = t2; ++n_tick; }
        }
        __syncthreads();
        if (s_last) {
            if (warp == 0) {
                __threadfence();
                ndt_job_update(a, job, us);
            }
            if (prof) { t_update += global_ns() - t1; ++n_upd; }
        }
    }
    if (prof) {
        atomicMax(&a.prof[1], global_ns());
        atomicAdd(&a.prof[2], t_sweep); atomicAdd(&a.prof[3], t_update); atomicAdd(&a.prof[4], t_wait);
        atomicAdd(&a.prof[5], n_tick); atomicAdd(&a.prof[6], n_upd);
    }
}

}  // namespace liloc
