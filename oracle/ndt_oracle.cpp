// oracle/ndt_oracle.cpp -- CPU restatement of the relo-mode NDT registration that LiLoc delegates to
// ndt_omp (pclomp::NormalDistributionsTransform + pclomp::VoxelGridCovariance).
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (see liloc_oracle.cpp header).
//
// PARITY UNPINNED.  ndt_omp is a third-party dependency of the reference that is NOT vendored under
// /root/reference and is unpinned there (README.md:39,47 `git clone` of koide3/ndt_omp default branch;
// package.xml:44-45).  Nothing in the reference tests or pins its output.  This file restates the
// published algorithm (Magnusson 2009, "The Three-Dimensional Normal-Distributions Transform";
// More & Thuente 1994 line search; the PCL/ndt_omp lineage summarised in SURVEY.md A.6) and anchors on
// the reference's own call sites:
//   setTransformationEpsilon(ndtEpsilon) / setResolution(ndtResolution)      src/mapOptmization.cpp:178-179
//   setNeighborhoodSearchMethod(DIRECT1 | DIRECT7)                           src/mapOptmization.cpp:181-190
//   setInputTarget(submap) ; setInputSource(full scan) ; align(guess)        src/mapOptmization.cpp:273-287,
//                                                                            include/egoOptimization/anchorOptimize.hpp:394-407
//   guess = [eulerToRotation(r,p,y) | t] ; result decoded with RotMtoEuler   include/math_tools.h:229-237, :92-111
// ndt_omp defaults kept: step_size 0.1, outlier_ratio 0.55, max_iterations 35, min_points_per_voxel 6,
// min_covar_eigvalue_mult 0.01.
//
// Canonical choices shared with the CUDA path (DESIGN.md "NDT arithmetic"):
//   * per-point terms in float32 exactly as ndt_omp's updateDerivatives (float 4x6 / 24x6 blocks), exp()
//     evaluated in double and rounded to float, accumulation of score / gradient / Hessian in double;
//   * the Hessian is evaluated in every derivative sweep (ndt_omp recomputes it after the line search with
//     a double-precision twin of the same formula);
//   * ONE summation order for the 28 sums of a sweep, independent of thread count (see derivative_sweep);
//   * the Newton step solves the symmetrised 6x6 Hessian: Gaussian elimination with partial pivoting when it is well
//     conditioned, Jacobi eigen-decomposition pseudo-inverse (== JacobiSVD::solve for a symmetric matrix) otherwise;
//   * voxel covariance eigen-decomposition: cyclic Jacobi in double; 3x3 inverse by cofactors.

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "oracle_types.h"

namespace {

// ---- 3x3 / 6x6 symmetric Jacobi in double ----------------------------------------------------------
template <int N>
void jacobi_sym(const double* A_in, double* w, double* V /* columns = eigenvectors, row-major NxN */) {
    double A[N][N], U[N][N];
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) { A[i][j] = A_in[i * N + j]; U[i][j] = (i == j) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < N; ++i) { diag += A[i][i] * A[i][i]; for (int j = i + 1; j < N; ++j) off += A[i][j] * A[i][j]; }
        if (off <= 1e-300 || off <= (DBL_EPSILON * DBL_EPSILON) * diag) break;
        for (int p = 0; p < N - 1; ++p) for (int q = p + 1; q < N; ++q) {
            const double apq = A[p][q];
            if (apq == 0.0) continue;
            const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
            const double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
            const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c, h = t * apq;
            A[p][p] -= h; A[q][q] += h; A[p][q] = 0.0; A[q][p] = 0.0;
            for (int k = 0; k < N; ++k) {
                if (k == p || k == q) continue;
                const double akp = A[k][p], akq = A[k][q];
                const double np_ = c * akp - s * akq, nq_ = s * akp + c * akq;
                A[k][p] = np_; A[p][k] = np_; A[k][q] = nq_; A[q][k] = nq_;
            }
            for (int k = 0; k < N; ++k) {          // columns p, q of U
                const double ukp = U[k][p], ukq = U[k][q];
                U[k][p] = c * ukp - s * ukq; U[k][q] = s * ukp + c * ukq;
            }
        }
    }
    for (int i = 0; i < N; ++i) w[i] = A[i][i];
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) V[i * N + j] = U[i][j];
}

bool inverse3(const double* m, double* inv) {
    const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    const double id = 1.0 / det;
    inv[0] = (e * i - f * h) * id; inv[1] = (c * h - b * i) * id; inv[2] = (b * f - c * e) * id;
    inv[3] = (f * g - d * i) * id; inv[4] = (a * i - c * g) * id; inv[5] = (c * d - a * f) * id;
    inv[6] = (d * h - e * g) * id; inv[7] = (b * g - a * h) * id; inv[8] = (a * e - b * d) * id;
    for (int k = 0; k < 9; ++k) if (!std::isfinite(inv[k])) return false;
    return true;
}

// ---- target grid: pclomp::VoxelGridCovariance --------------------------------------------
struct NdtLeaf { double mean[3]; float icov[9]; int n; int key; };

struct NdtTarget {
    float inv_leaf;
    float resolution = 0.f;
    int minb[3], maxb[3], divb[3];
    std::vector<int> table;          // dense cell -> leaf index, -1 = no usable leaf
    std::vector<NdtLeaf> leaves;
    bool ok = false;

    void build(const orc_point* pts, int n, float resolution) {
        ok = false; leaves.clear(); table.clear();
        if (n <= 0) return;
        inv_leaf = 1.0f / resolution;
        this->resolution = resolution;
        float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (int i = 0; i < n; ++i) {
            const float v[3] = {pts[i].x, pts[i].y, pts[i].z};
            for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], v[a]); mx[a] = std::max(mx[a], v[a]); }
        }
        const int64_t dx = (int64_t)((mx[0] - mn[0]) * inv_leaf) + 1, dy = (int64_t)((mx[1] - mn[1]) * inv_leaf) + 1, dz = (int64_t)((mx[2] - mn[2]) * inv_leaf) + 1;
        if (dx * dy * dz > (int64_t)INT32_MAX) return;     // "Leaf size is too small": no grid
        for (int a = 0; a < 3; ++a) {
            minb[a] = (int)std::floor(mn[a] * inv_leaf);
            maxb[a] = (int)std::floor(mx[a] * inv_leaf);
            divb[a] = maxb[a] - minb[a] + 1;
        }
        const int64_t ncell = (int64_t)divb[0] * divb[1] * divb[2];
        std::vector<std::pair<int, int>> keys((size_t)n);
        for (int i = 0; i < n; ++i) {
            const int i0 = (int)(std::floor(pts[i].x * inv_leaf) - (float)minb[0]);
            const int i1 = (int)(std::floor(pts[i].y * inv_leaf) - (float)minb[1]);
            const int i2 = (int)(std::floor(pts[i].z * inv_leaf) - (float)minb[2]);
            keys[i] = {i0 + i1 * divb[0] + i2 * divb[0] * divb[1], i};
        }
        std::sort(keys.begin(), keys.end());
        table.assign((size_t)ncell, -1);
        for (size_t i = 0; i < keys.size();) {
            size_t j = i;
            double s[3] = {0, 0, 0}, ss[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (; j < keys.size() && keys[j].first == keys[i].first; ++j) {
                const orc_point& p = pts[keys[j].second];
                const double v[3] = {(double)p.x, (double)p.y, (double)p.z};
                for (int a = 0; a < 3; ++a) { s[a] += v[a]; for (int b = 0; b < 3; ++b) ss[a * 3 + b] += v[a] * v[b]; }
            }
            const int cnt = (int)(j - i);
            const int key = keys[i].first;
            i = j;
            if (cnt < 6) continue;                          // min_points_per_voxel_
            const double nn = (double)cnt;
            double mean[3] = {s[0] / nn, s[1] / nn, s[2] / nn}, cov[9];
            for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b)
                cov[a * 3 + b] = ((ss[a * 3 + b] - 2.0 * (s[a] * mean[b])) / nn + mean[a] * mean[b]) * ((nn - 1.0) / nn);
            double sym[9];
            for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) sym[a * 3 + b] = 0.5 * (cov[a * 3 + b] + cov[b * 3 + a]);
            double w[3], V[9];
            jacobi_sym<3>(sym, w, V);
            int o[3] = {0, 1, 2};                           // ascending eigenvalues (SelfAdjointEigenSolver order)
            std::sort(o, o + 3, [&](int a, int b) { return w[a] < w[b] || (w[a] == w[b] && a < b); });
            double ev[3] = {w[o[0]], w[o[1]], w[o[2]]};
            if (ev[0] < 0 || ev[1] < 0 || ev[2] <= 0) continue;
            const double min_ev = 0.01 * ev[2];             // min_covar_eigvalue_mult_
            if (ev[0] < min_ev) {
                ev[0] = min_ev;
                if (ev[1] < min_ev) ev[1] = min_ev;
                for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) {      // cov = V diag(ev) V^T
                    double acc = 0.0;
                    for (int k = 0; k < 3; ++k) acc += V[a * 3 + o[k]] * ev[k] * V[b * 3 + o[k]];
                    cov[a * 3 + b] = acc;
                }
            }
            double icov[9];
            if (!inverse3(cov, icov)) continue;
            NdtLeaf L;
            for (int a = 0; a < 3; ++a) L.mean[a] = mean[a];
            for (int k = 0; k < 9; ++k) L.icov[k] = (float)icov[k];
            L.n = cnt; L.key = key;
            table[(size_t)key] = (int)leaves.size();
            leaves.push_back(L);
        }
        ok = true;
    }

    // getNeighborhoodAtPoint1 / 7: leaf index of the cell displaced by (dx,dy,dz) from the cell of p, or -1
    inline int lookup(float x, float y, float z, int ddx, int ddy, int ddz) const {
        const int i = (int)std::floor(x * inv_leaf) + ddx, j = (int)std::floor(y * inv_leaf) + ddy, k = (int)std::floor(z * inv_leaf) + ddz;
        if (i < minb[0] || i > maxb[0] || j < minb[1] || j > maxb[1] || k < minb[2] || k > maxb[2]) return -1;
        return table[(size_t)(i - minb[0]) + (size_t)(j - minb[1]) * divb[0] + (size_t)(k - minb[2]) * divb[0] * divb[1]];
    }
};

const int kDisp7[7][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};

// ---- pose algebra at the 4x4 boundary ---------------------------------------------------------------
// T(p) = Translation(p0..2) * Rx(p3) * Ry(p4) * Rz(p5), float (ndt_omp computeStepLengthMT)
void pose_to_matrix(const double* p, float* T /* 3x4 row-major */) {
    const float a = (float)p[3], b = (float)p[4], c = (float)p[5];
    const float ca = (float)std::cos((double)a), sa = (float)std::sin((double)a);
    const float cb = (float)std::cos((double)b), sb = (float)std::sin((double)b);
    const float cc = (float)std::cos((double)c), sc = (float)std::sin((double)c);
    const float Rx[9] = {1, 0, 0, 0, ca, -sa, 0, sa, ca}, Ry[9] = {cb, 0, sb, 0, 1, 0, -sb, 0, cb}, Rz[9] = {cc, -sc, 0, sc, cc, 0, 0, 0, 1};
    float M[9], R[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M[i * 3 + j] = Rx[i * 3] * Ry[j] + Rx[i * 3 + 1] * Ry[3 + j] + Rx[i * 3 + 2] * Ry[6 + j];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R[i * 3 + j] = M[i * 3] * Rz[j] + M[i * 3 + 1] * Rz[3 + j] + M[i * 3 + 2] * Rz[6 + j];
    for (int i = 0; i < 3; ++i) { T[i * 4] = R[i * 3]; T[i * 4 + 1] = R[i * 3 + 1]; T[i * 4 + 2] = R[i * 3 + 2]; T[i * 4 + 3] = (float)p[i]; }
}

// Eigen eulerAngles(0,1,2) of the rotation block (float), translation -> p (double)
void matrix_to_pose(const float* T /* 3x4 */, double* p) {
    const float m00 = T[0], m01 = T[1], m02 = T[2], m10 = T[4], m11 = T[5], m12 = T[6], m20 = T[8], m21 = T[9], m22 = T[10];
    // transcendental functions evaluated in double and rounded to float (canonical, shared with the device)
    auto at2 = [](float y, float x) { return (float)std::atan2((double)y, (double)x); };
    float r0 = at2(m12, m22);
    const float c2 = std::sqrt(m00 * m00 + m01 * m01);
    float r1;
    if (r0 > 0.f) { r0 -= 3.14159265358979323846f; r1 = at2(-m02, -c2); }
    else r1 = at2(-m02, c2);
    const float s1 = (float)std::sin((double)r0), c1 = (float)std::cos((double)r0);
    const float r2 = at2(s1 * m20 - c1 * m10, c1 * m11 - s1 * m21);
    p[0] = T[3]; p[1] = T[7]; p[2] = T[11];
    p[3] = -r0; p[4] = -r1; p[5] = -r2;
}

// ---- derivative sweep ---------------------------------------------------------------------------------
struct AngDeriv { float j[8][3]; float h[15][3]; };

void angle_derivatives(const double* p, AngDeriv* A) {      // computeAngleDerivatives
    double cx, cy, cz, sx, sy, sz;
    if (std::fabs(p[3]) < 10e-5) { cx = 1.0; sx = 0.0; } else { cx = std::cos(p[3]); sx = std::sin(p[3]); }
    if (std::fabs(p[4]) < 10e-5) { cy = 1.0; sy = 0.0; } else { cy = std::cos(p[4]); sy = std::sin(p[4]); }
    if (std::fabs(p[5]) < 10e-5) { cz = 1.0; sz = 0.0; } else { cz = std::cos(p[5]); sz = std::sin(p[5]); }
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

struct NdtConsts { double d1, d2; };
NdtConsts gauss_consts(double resolution, double outlier_ratio) {
    const double c1 = 10.0 * (1.0 - outlier_ratio), c2 = outlier_ratio / std::pow(resolution, 3);
    const double d3 = -std::log(c2);
    NdtConsts k;
    k.d1 = -std::log(c1 + c2) - d3;
    k.d2 = -2.0 * std::log((-std::log(c1 * std::exp(-0.5) + c2) - d3) / k.d1);
    return k;
}

// Contribution of one (point, cell) pair: updateDerivatives with compute_hessian = true.
// out[0] score, out[1..6] gradient, out[7..27] upper triangle of the Hessian (row-major), all float terms.
inline bool point_cell_terms(const float x[3], const float xt[3], const NdtLeaf& L, const AngDeriv& A, double d1, float d2, float* out) {
    const float q[3] = {(float)((double)xt[0] - L.mean[0]), (float)((double)xt[1] - L.mean[1]), (float)((double)xt[2] - L.mean[2])};
    const float* C = L.icov;
    // x_trans4 * c_inv4 (row vector times matrix), then dot with x_trans4
    const float qC[3] = {q[0] * C[0] + q[1] * C[3] + q[2] * C[6], q[0] * C[1] + q[1] * C[4] + q[2] * C[7], q[0] * C[2] + q[1] * C[5] + q[2] * C[8]};
    const float qCq = q[0] * qC[0] + q[1] * qC[1] + q[2] * qC[2];
    float e = (float)std::exp((double)(-d2 * qCq * 0.5f));
    out[0] = (float)(-d1 * (double)e);                 // gauss_d1_ is a double member, score_inc a float
    e = d2 * e;
    if (e > 1.f || e < 0.f || e != e) { for (int k = 0; k < 28; ++k) out[k] = 0.f; return false; }
    e = (float)((double)e * d1);
    // point gradient columns 3..5 (columns 0..2 are the identity)
    float g[3][6];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 6; ++c) g[r][c] = (r == c) ? 1.f : 0.f;
    float xj[8];
    for (int r = 0; r < 8; ++r) xj[r] = A.j[r][0] * x[0] + A.j[r][1] * x[1] + A.j[r][2] * x[2];
    g[1][3] = xj[0]; g[2][3] = xj[1];
    g[0][4] = xj[2]; g[1][4] = xj[3]; g[2][4] = xj[4];
    g[0][5] = xj[5]; g[1][5] = xj[6]; g[2][5] = xj[7];
    float xh[15];
    for (int r = 0; r < 15; ++r) xh[r] = A.h[r][0] * x[0] + A.h[r][1] * x[1] + A.h[r][2] * x[2];
    const float ha[3] = {0.f, xh[0], xh[1]}, hb[3] = {0.f, xh[2], xh[3]}, hc[3] = {0.f, xh[4], xh[5]};
    const float hd[3] = {xh[6], xh[7], xh[8]}, he[3] = {xh[9], xh[10], xh[11]}, hf[3] = {xh[12], xh[13], xh[14]};
    // c_inv4 * point_gradient4 and x_trans4 . (that)
    float Cg[3][6], qCg[6];
    for (int c = 0; c < 6; ++c) {
        for (int r = 0; r < 3; ++r) Cg[r][c] = C[r * 3] * g[0][c] + C[r * 3 + 1] * g[1][c] + C[r * 3 + 2] * g[2][c];
        qCg[c] = q[0] * Cg[0][c] + q[1] * Cg[1][c] + q[2] * Cg[2][c];
    }
    for (int c = 0; c < 6; ++c) out[1 + c] = e * qCg[c];
    // Hessian: e * (-d2 * qCg_i * qCg_j + qC . H_ij + g_j . C g_i)
    const float* Hblk[3][3] = {{ha, hb, hc}, {hb, hd, he}, {hc, he, hf}};
    int k = 7;
    for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) {
        float hterm = 0.f;
        if (i >= 3 && j >= 3) { const float* v = Hblk[i - 3][j - 3]; hterm = qC[0] * v[0] + qC[1] * v[1] + qC[2] * v[2]; }
        const float gCg = g[0][j] * Cg[0][i] + g[1][j] * Cg[1][i] + g[2][j] * Cg[2][i];
        out[k++] = e * (-d2 * qCg[i] * qCg[j] + hterm + gCg);
    }
    return true;
}

struct Sums { double v[28]; };

// CANONICAL SUMMATION ORDER of the 28 sums of a derivative sweep (shared with the device, liloc_b200/csrc/ndt.cuh
// ndt_sweep_tile + the job update).  ndt_omp itself sums per OpenMP thread and then over threads, so its result
// depends on the thread count in the last bits -- and the More-Thuente line search branches on such bits.  Both
// sides of the parity tests therefore use ONE fixed order, independent of thread count / grid size:
//   * the source cloud is cut into tiles of 1024 consecutive points; a tile is 8 "warps"; warp w owns the points
//     base + r * 256 + w * 32 + lane   (r = 0..3 rounds, lane = 0..31);
//   * a warp's (point, cell) hits -- cells that hold a Gaussian (and, for KDTREE, pass the radius test) -- are
//     enumerated in (round, cell, lane) order and dealt round-robin to 32 accumulators: hit number s goes to
//     accumulator s % 32, each accumulator adds its hits' float terms in that order in double, starting from 0;
//   * the 32 accumulators are combined by a butterfly: a[l] += a[l + s] for l < s, s = 16, 8, 4, 2, 1;
//   * the 8 warp totals are added in warp order to 0.0, the tile totals in tile order to 0.0.
constexpr int kTile = 1024, kWarps = 8, kLanes = 32, kRounds = 4;

void sweep_tile(const NdtTarget& tg, const orc_point* src, int n, int tile, const float* T, const AngDeriv& A, double d1, float d2,
                int search, double* out28) {
    const int ncell = search == 2 ? 27 : (search == 1 ? 7 : 1);
    const float radius2 = tg.resolution * tg.resolution;
    const int base = tile * kTile;
    double tile_tot[28];
    for (int k = 0; k < 28; ++k) tile_tot[k] = 0.0;
    for (int w = 0; w < kWarps; ++w) {
        double acc[kLanes][28];
        for (int l = 0; l < kLanes; ++l) for (int k = 0; k < 28; ++k) acc[l][k] = 0.0;
        long long seq = 0;
        for (int r = 0; r < kRounds; ++r) {
            float xs[kLanes][3], xts[kLanes][3];
            bool valid[kLanes];
            for (int l = 0; l < kLanes; ++l) {
                const int i = base + r * (kWarps * kLanes) + w * kLanes + l;
                valid[l] = i < n;
                if (!valid[l]) continue;
                const float x[3] = {src[i].x, src[i].y, src[i].z};
                for (int a = 0; a < 3; ++a) xs[l][a] = x[a];
                xts[l][0] = T[0] * x[0] + T[1] * x[1] + T[2] * x[2] + T[3];
                xts[l][1] = T[4] * x[0] + T[5] * x[1] + T[6] * x[2] + T[7];
                xts[l][2] = T[8] * x[0] + T[9] * x[1] + T[10] * x[2] + T[11];
            }
            // search: 0 DIRECT1 (the point's voxel), 1 DIRECT7 (+ the six face neighbours), 2 KDTREE (radiusSearch over the
            // leaf centroids with radius = resolution: squared float distance < radius^2, FLANN RadiusResultSet; such a
            // centroid can only belong to one of the 27 voxels around the point's voxel).  ndt_omp sums the neighbours in
            // FLANN's distance order; here in voxel order (canonical).
            for (int c = 0; c < ncell; ++c) {
                const int dx = ncell == 27 ? c % 3 - 1 : kDisp7[c][0], dy = ncell == 27 ? (c / 3) % 3 - 1 : kDisp7[c][1], dz = ncell == 27 ? c / 9 - 1 : kDisp7[c][2];
                for (int l = 0; l < kLanes; ++l) {
                    if (!valid[l]) continue;
                    const float* xt = xts[l];
                    const int li = tg.lookup(xt[0], xt[1], xt[2], dx, dy, dz);
                    if (li < 0) continue;
                    const NdtLeaf& L = tg.leaves[(size_t)li];
                    if (ncell == 27) {
                        const float ex = (float)L.mean[0] - xt[0], ey = (float)L.mean[1] - xt[1], ez = (float)L.mean[2] - xt[2];
                        if (!(ex * ex + ey * ey + ez * ez < radius2)) continue;
                    }
                    float t[28];
                    const bool used = point_cell_terms(xs[l], xt, L, A, d1, d2, t);
                    double* a = acc[seq % kLanes];
                    ++seq;                                  // a rejected pair (d2 * e outside [0, 1]) still takes its slot
                    if (used) for (int k = 0; k < 28; ++k) a[k] += (double)t[k];
                }
            }
        }
        for (int s = kLanes / 2; s >= 1; s >>= 1)
            for (int l = 0; l < s; ++l) for (int k = 0; k < 28; ++k) acc[l][k] = acc[l][k] + acc[l + s][k];
        for (int k = 0; k < 28; ++k) tile_tot[k] += acc[0][k];
    }
    for (int k = 0; k < 28; ++k) out28[k] = tile_tot[k];
}

Sums derivative_sweep(const NdtTarget& tg, const orc_point* src, int n, const double* p, const NdtConsts& K, int search, int threads) {
    float T[12]; pose_to_matrix(p, T);
    AngDeriv A; angle_derivatives(p, &A);
    const double d1 = K.d1; const float d2 = (float)K.d2;
    const int n_tiles = (n + kTile - 1) / kTile;
    std::vector<Sums> per((size_t)std::max(n_tiles, 1));
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(dynamic, 1)
    for (int t = 0; t < n_tiles; ++t) sweep_tile(tg, src, n, t, T, A, d1, d2, search, per[(size_t)t].v);
    Sums tot; for (int k = 0; k < 28; ++k) tot.v[k] = 0.0;
    for (int t = 0; t < n_tiles; ++t) for (int k = 0; k < 28; ++k) tot.v[k] += per[(size_t)t].v[k];
    return tot;
}

// Plain per-point summation in ascending point order (what a single-threaded ndt_omp does): kept as the cross-check
// of the canonical order above (tests/test_ndt_oracle.py: the two agree to ~1e-13 relative).
Sums derivative_sweep_sequential(const NdtTarget& tg, const orc_point* src, int n, const double* p, const NdtConsts& K, int search) {
    float T[12]; pose_to_matrix(p, T);
    AngDeriv A; angle_derivatives(p, &A);
    const double d1 = K.d1; const float d2 = (float)K.d2;
    const int ncell = search == 2 ? 27 : (search == 1 ? 7 : 1);
    const float radius2 = tg.resolution * tg.resolution;
    Sums tot; for (int k = 0; k < 28; ++k) tot.v[k] = 0.0;
    for (int i = 0; i < n; ++i) {
        const float x[3] = {src[i].x, src[i].y, src[i].z};
        const float xt[3] = {T[0] * x[0] + T[1] * x[1] + T[2] * x[2] + T[3], T[4] * x[0] + T[5] * x[1] + T[6] * x[2] + T[7],
                             T[8] * x[0] + T[9] * x[1] + T[10] * x[2] + T[11]};
        for (int c = 0; c < ncell; ++c) {
            const int dx = ncell == 27 ? c % 3 - 1 : kDisp7[c][0], dy = ncell == 27 ? (c / 3) % 3 - 1 : kDisp7[c][1], dz = ncell == 27 ? c / 9 - 1 : kDisp7[c][2];
            const int li = tg.lookup(xt[0], xt[1], xt[2], dx, dy, dz);
            if (li < 0) continue;
            if (ncell == 27) {
                const auto& L = tg.leaves[(size_t)li];
                const float ex = (float)L.mean[0] - xt[0], ey = (float)L.mean[1] - xt[1], ez = (float)L.mean[2] - xt[2];
                if (!(ex * ex + ey * ey + ez * ez < radius2)) continue;
            }
            float t[28];
            if (point_cell_terms(x, xt, tg.leaves[(size_t)li], A, d1, d2, t)) for (int k = 0; k < 28; ++k) tot.v[k] += (double)t[k];
        }
    }
    return tot;
}

// Newton direction: JacobiSVD(H).solve(-g) of the symmetrised 6x6 Hessian.
// newton_solve_pinv: the pseudo-inverse through a Jacobi eigen-decomposition (== JacobiSVD::solve for a symmetric matrix:
// directions whose singular value is below 6 eps smax are dropped).
void newton_solve_pinv(const double* Hup /*21*/, const double* g, double* dp) {
    double H[36], w[6], V[36];
    int k = 0;
    for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { H[i * 6 + j] = Hup[k]; H[j * 6 + i] = Hup[k]; ++k; }
    jacobi_sym<6>(H, w, V);
    double smax = 0.0;
    for (int i = 0; i < 6; ++i) smax = std::max(smax, std::fabs(w[i]));
    const double thr = std::max(smax * 6.0 * DBL_EPSILON, DBL_MIN);
    for (int i = 0; i < 6; ++i) dp[i] = 0.0;
    for (int m = 0; m < 6; ++m) {
        if (!(std::fabs(w[m]) > thr)) continue;
        double proj = 0.0;
        for (int i = 0; i < 6; ++i) proj += V[i * 6 + m] * (-g[i]);
        proj /= w[m];
        for (int i = 0; i < 6; ++i) dp[i] += V[i * 6 + m] * proj;
    }
}

// CANONICAL Newton solve (shared with the device, ndt.cuh ndt_newton_solve): when H is well conditioned the
// minimum-norm solution IS H^-1 (-g); it is then computed by Gaussian elimination with partial pivoting in double
// (first row of maximal |pivot|; accepted when min |pivot| > 1e-10 max |pivot|), which costs a fraction of the Jacobi
// sweeps; otherwise the pseudo-inverse above.  The two agree to ~1e-12 relative where the fast path applies
// (tests/test_ndt_oracle.py::test_newton_paths_agree); the choice only fixes WHICH last bits both sides share.
void newton_solve(const double* Hup /*21*/, const double* g, double* dp) {
    double M[6][7];
    {
        int k = 0;
        for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { M[i][j] = Hup[k]; M[j][i] = Hup[k]; ++k; }
    }
    for (int i = 0; i < 6; ++i) M[i][6] = -g[i];
    double pmin = 1e300, pmax = 0.0;
    bool ok = true;
    for (int c = 0; c < 6; ++c) {
        int pr = c; double best = std::fabs(M[c][c]);
        for (int r = c + 1; r < 6; ++r) { const double v = std::fabs(M[r][c]); if (v > best) { best = v; pr = r; } }
        if (!(best > 0.0) || best != best) ok = false;
        if (pr != c) for (int j = c; j < 7; ++j) std::swap(M[c][j], M[pr][j]);
        pmin = std::min(pmin, best); pmax = std::max(pmax, best);
        const double inv = 1.0 / M[c][c];
        for (int r = c + 1; r < 6; ++r) {
            const double f = M[r][c] * inv;
            for (int j = c + 1; j < 7; ++j) M[r][j] -= f * M[c][j];
        }
    }
    if (ok && pmin > 1e-10 * pmax) {
        double x[6];
        for (int i = 5; i >= 0; --i) {
            double v = M[i][6];
            for (int j = i + 1; j < 6; ++j) v -= M[i][j] * x[j];
            x[i] = v / M[i][i];
        }
        for (int i = 0; i < 6; ++i) dp[i] = x[i];
        return;
    }
    newton_solve_pinv(Hup, g, dp);
}

// ---- More-Thuente helpers (computeStepLengthMT and friends) --------------------------------------------
inline double psi_mt(double a, double f_a, double f_0, double g_0, double mu) { return f_a - f_0 - mu * g_0 * a; }
inline double dpsi_mt(double g_a, double g_0, double mu) { return g_a - mu * g_0; }

bool update_interval_mt(double& a_l, double& f_l, double& g_l, double& a_u, double& f_u, double& g_u, double a_t, double f_t, double g_t) {
    if (f_t > f_l) { a_u = a_t; f_u = f_t; g_u = g_t; return false; }
    else if (g_t * (a_l - a_t) > 0) { a_l = a_t; f_l = f_t; g_l = g_t; return false; }
    else if (g_t * (a_l - a_t) < 0) { a_u = a_l; f_u = f_l; g_u = g_l; a_l = a_t; f_l = f_t; g_l = g_t; return false; }
    return true;
}

double trial_value_mt(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u, double a_t, double f_t, double g_t) {
    if (f_t > f_l) {
        const double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l, w = std::sqrt(z * z - g_t * g_l);
        const double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
        const double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
        return (std::fabs(a_c - a_l) < std::fabs(a_q - a_l)) ? a_c : 0.5 * (a_q + a_c);
    } else if (g_t * g_l < 0) {
        const double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l, w = std::sqrt(z * z - g_t * g_l);
        const double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
        const double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
        return (std::fabs(a_c - a_t) >= std::fabs(a_s - a_t)) ? a_c : a_s;
    } else if (std::fabs(g_t) <= std::fabs(g_l)) {
        const double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l, w = std::sqrt(z * z - g_t * g_l);
        const double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
        const double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
        const double a_n = (std::fabs(a_c - a_t) < std::fabs(a_s - a_t)) ? a_c : a_s;
        return (a_t > a_l) ? std::min(a_t + 0.66 * (a_u - a_t), a_n) : std::max(a_t + 0.66 * (a_u - a_t), a_n);
    }
    const double z = 3 * (f_t - f_u) / (a_t - a_u) - g_t - g_u, w = std::sqrt(z * z - g_t * g_u);
    return a_u + (a_t - a_u) * (w - g_u - z) / (g_t - g_u + 2 * w);
}

}  // namespace

extern "C" {

typedef struct {
    double step_size;        // 0.1   (pclomp default)
    double outlier_ratio;    // 0.55  (pclomp default)
    double trans_eps;        // (double) ndtEpsilon, a float parameter in the reference (include/utility.h:285)
    float resolution;        // ndtResolution
    int max_iterations;      // 35
    int search;              // 0 DIRECT1, 1 DIRECT7, 2 KDTREE
    int threads;
} orc_ndt_opts;

typedef struct {
    int iterations;          // nr_iterations_
    int converged;
    int sweeps;              // derivative sweeps executed (incl. line-search evaluations)
    double score;            // final score (sum), trans_probability = score / n
} orc_ndt_result;

void* orc_ndt_target_create(const orc_point* pts, int n, float resolution) {
    NdtTarget* t = new NdtTarget();
    t->build(pts, n, resolution);
    return t;
}
void orc_ndt_target_destroy(void* t) { delete static_cast<NdtTarget*>(t); }
int orc_ndt_target_leaves(void* t) { return (int)static_cast<NdtTarget*>(t)->leaves.size(); }
// leaf records in ascending voxel-index order: mean[3] (double), icov[9] (float), n
int orc_ndt_target_keys(void* tp, int* keys, int cap) {
    NdtTarget* t = static_cast<NdtTarget*>(tp);
    const int n = (int)t->leaves.size();
    for (int i = 0; i < n && i < cap; ++i) keys[i] = t->leaves[i].key;
    return n;
}
int orc_ndt_target_get(void* tp, double* means, float* icovs, int* counts, int cap) {
    NdtTarget* t = static_cast<NdtTarget*>(tp);
    const int n = (int)t->leaves.size();
    for (int i = 0; i < n && i < cap; ++i) {
        for (int a = 0; a < 3; ++a) means[3 * i + a] = t->leaves[i].mean[a];
        for (int k = 0; k < 9; ++k) icovs[9 * i + k] = t->leaves[i].icov[k];
        counts[i] = t->leaves[i].n;
    }
    return n;
}

void orc_ndt_pose_to_matrix(const double* p6, float* T12) { pose_to_matrix(p6, T12); }
void orc_ndt_matrix_to_pose(const float* T12, double* p6) { matrix_to_pose(T12, p6); }

// the Newton direction of computeTransformation (JacobiSVD(H).solve(-g)) and the cubic / quadratic trial value of the
// More-Thuente step (trialValueSelectionMT), exposed for the known-answer tests
void orc_ndt_newton(const double* Hup21, const double* g6, double* dp6) { newton_solve(Hup21, g6, dp6); }
void orc_ndt_newton_pinv(const double* Hup21, const double* g6, double* dp6) { newton_solve_pinv(Hup21, g6, dp6); }
double orc_ndt_trial_value(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u, double a_t, double f_t, double g_t) {
    return trial_value_mt(a_l, f_l, g_l, a_u, f_u, g_u, a_t, f_t, g_t);
}

// one derivative sweep at pose p: out28 = {score, grad[6], hessian upper[21]}
void orc_ndt_derivatives(void* tp, const orc_point* src, int n, const double* p6, const orc_ndt_opts* o, double* out28) {
    const NdtConsts K = gauss_consts(o->resolution, o->outlier_ratio);
    const Sums s = derivative_sweep(*static_cast<NdtTarget*>(tp), src, n, p6, K, o->search, o->threads);
    std::memcpy(out28, s.v, sizeof(s.v));
}

// the same sums in plain ascending point order (cross-check of the canonical order)
void orc_ndt_derivatives_sequential(void* tp, const orc_point* src, int n, const double* p6, const orc_ndt_opts* o, double* out28) {
    const NdtConsts K = gauss_consts(o->resolution, o->outlier_ratio);
    const Sums s = derivative_sweep_sequential(*static_cast<NdtTarget*>(tp), src, n, p6, K, o->search);
    std::memcpy(out28, s.v, sizeof(s.v));
}

// pclomp::NormalDistributionsTransform::computeTransformation: guess / result are 3x4 row-major float.
int orc_ndt_align(void* tp, const orc_point* src, int n, const float* guess12, const orc_ndt_opts* o, float* final12, orc_ndt_result* res) {
    const NdtTarget& tg = *static_cast<NdtTarget*>(tp);
    const NdtConsts K = gauss_consts(o->resolution, o->outlier_ratio);
    int nr_iterations = 0, sweeps = 0;
    bool converged = false;
    float Tfinal[12];
    std::memcpy(Tfinal, guess12, sizeof(Tfinal));
    double p[6];
    matrix_to_pose(guess12, p);
    Sums cur = derivative_sweep(tg, src, n, p, K, o->search, o->threads); ++sweeps;
    double score = cur.v[0];
    const double step_max = o->step_size, step_min = o->trans_eps / 2.0, eps = o->trans_eps;
    while (!converged) {
        double dp[6];
        newton_solve(cur.v + 7, cur.v + 1, dp);
        double norm = 0.0;
        for (int i = 0; i < 6; ++i) norm += dp[i] * dp[i];
        norm = std::sqrt(norm);
        if (norm == 0 || norm != norm) { converged = (norm == norm); break; }
        double dir[6];
        for (int i = 0; i < 6; ++i) dir[i] = dp[i] / norm;
        // ---- computeStepLengthMT ----
        double a_t;
        {
            const double phi_0 = -score;
            double d_phi_0 = 0.0;
            for (int i = 0; i < 6; ++i) d_phi_0 -= cur.v[1 + i] * dir[i];
            bool zero_step = false;
            if (d_phi_0 >= 0) {
                if (d_phi_0 == 0) zero_step = true;
                else { d_phi_0 *= -1; for (int i = 0; i < 6; ++i) dir[i] *= -1; }
            }
            if (zero_step) a_t = 0.0;
            else {
                const int max_step_iterations = 10;
                int step_iterations = 0;
                const double mu = 1.e-4, nu = 0.9;
                double a_l = 0, a_u = 0;
                double f_l = psi_mt(a_l, phi_0, phi_0, d_phi_0, mu), g_l = dpsi_mt(d_phi_0, d_phi_0, mu);
                double f_u = psi_mt(a_u, phi_0, phi_0, d_phi_0, mu), g_u = dpsi_mt(d_phi_0, d_phi_0, mu);
                bool interval_converged = (step_max - step_min) < 0, open_interval = true;
                a_t = norm;
                a_t = std::min(a_t, step_max); a_t = std::max(a_t, step_min);
                double x_t[6];
                for (int i = 0; i < 6; ++i) x_t[i] = p[i] + dir[i] * a_t;
                pose_to_matrix(x_t, Tfinal);
                cur = derivative_sweep(tg, src, n, x_t, K, o->search, o->threads); ++sweeps;
                score = cur.v[0];
                double phi_t = -score, d_phi_t = 0.0;
                for (int i = 0; i < 6; ++i) d_phi_t -= cur.v[1 + i] * dir[i];
                double psi_t = psi_mt(a_t, phi_t, phi_0, d_phi_0, mu), d_psi_t = dpsi_mt(d_phi_t, d_phi_0, mu);
                while (!interval_converged && step_iterations < max_step_iterations && !(psi_t <= 0 && d_phi_t <= -nu * d_phi_0)) {
                    if (open_interval) a_t = trial_value_mt(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
                    else a_t = trial_value_mt(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
                    a_t = std::min(a_t, step_max); a_t = std::max(a_t, step_min);
                    for (int i = 0; i < 6; ++i) x_t[i] = p[i] + dir[i] * a_t;
                    pose_to_matrix(x_t, Tfinal);
                    cur = derivative_sweep(tg, src, n, x_t, K, o->search, o->threads); ++sweeps;
                    score = cur.v[0];
                    phi_t = -score; d_phi_t = 0.0;
                    for (int i = 0; i < 6; ++i) d_phi_t -= cur.v[1 + i] * dir[i];
                    psi_t = psi_mt(a_t, phi_t, phi_0, d_phi_0, mu); d_psi_t = dpsi_mt(d_phi_t, d_phi_0, mu);
                    if (open_interval && (psi_t <= 0 && d_psi_t >= 0)) {
                        open_interval = false;
                        f_l = f_l + phi_0 - mu * d_phi_0 * a_l; g_l = g_l + mu * d_phi_0;
                        f_u = f_u + phi_0 - mu * d_phi_0 * a_u; g_u = g_u + mu * d_phi_0;
                    }
                    if (open_interval) interval_converged = update_interval_mt(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
                    else interval_converged = update_interval_mt(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
                    step_iterations++;
                }
            }
        }
        for (int i = 0; i < 6; ++i) p[i] += dir[i] * a_t;
        if (nr_iterations > o->max_iterations || (nr_iterations && (std::fabs(a_t) < eps))) converged = true;
        nr_iterations++;
    }
    std::memcpy(final12, Tfinal, sizeof(Tfinal));
    if (res) { res->iterations = nr_iterations; res->converged = converged ? 1 : 0; res->sweeps = sweeps; res->score = score; }
    return 0;
}

}  // extern "C"
