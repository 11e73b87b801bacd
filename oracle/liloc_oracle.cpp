// oracle/liloc_oracle.cpp -- CPU restatement of LiLoc's scan-to-map registration hot path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference leg may load it.  The product (liloc_b200/) never
// links, includes or calls anything in oracle/.
//
// PARITY UNPINNED: the reference (/root/reference) ships no tests, golden vectors or fixtures for
// this path, and its own implementation cannot be compiled here (ROS, PCL, FLANN, Eigen, OpenCV C++,
// GTSAM, ndt_omp are absent).  This file restates, function by function, the reference lines cited
// below plus the published third-party algorithms they delegate to; it is
// cross-checked in tests/ against scipy.cKDTree, numpy.linalg and cv2 (the Python OpenCV build).
//
// Arithmetic contract shared with the CUDA path (so that per-point results are bit-identical):
//   * float32 IEEE ops, NO fused multiply-add (build with -ffp-contract=off), left-to-right as written
//   * sin/cos of pose angles are evaluated in double and rounded to float (canonical choice; the
//     reference calls sinf/cosf, whose glibc and CUDA versions differ by <=1 ulp)
//   * kNN total order is (d2, index) ascending; d2 = ((dx*dx)+(dy*dy))+(dz*dz)
//   * J^T J / J^T r are accumulated in double from exact float*float products, rounded once to float
//
// Build: see oracle/Makefile  (g++ -O3 -fopenmp -ffp-contract=off -shared -fPIC)

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "oracle_types.h"

namespace {

inline float sin_c(float a) { return (float)std::sin((double)a); }
inline float cos_c(float a) { return (float)std::cos((double)a); }

// ---------------------------------------------------------------------------------------------
// pcl::getTransformation(x,y,z,roll,pitch,yaw) float instantiation, called by
// trans2Affine3f  src/mapOptmization.cpp:404-406 and transformPointCloud include/utility.h:394.
// pose6 = [roll, pitch, yaw, x, y, z] (transformTobeMapped layout, src/mapOptmization.cpp:85).
// T is the top 3x4 block, row-major.
void pose_to_T(const float* pose6, float* T) {
    const float roll = pose6[0], pitch = pose6[1], yaw = pose6[2];
    const float A = cos_c(yaw), B = sin_c(yaw), C = cos_c(pitch), D = sin_c(pitch);
    const float E = cos_c(roll), F = sin_c(roll), DE = D * E, DF = D * F;
    T[0] = A * C;  T[1] = A * DF - B * E;  T[2]  = B * F + A * DE;  T[3]  = pose6[3];
    T[4] = B * C;  T[5] = A * E + B * DF;  T[6]  = B * DE - A * F;  T[7]  = pose6[4];
    T[8] = -D;     T[9] = C * F;           T[10] = C * E;           T[11] = pose6[5];
}

// pointAssociateToMap src/mapOptmization.cpp:388-393 == loop body of transformPointCloud
// include/utility.h:400-405: m0*x + m1*y + m2*z + t, left to right, intensity copied.
inline orc_point apply_T(const float* T, const orc_point& p) {
    orc_point o;
    o.x = T[0] * p.x + T[1] * p.y + T[2]  * p.z + T[3];
    o.y = T[4] * p.x + T[5] * p.y + T[6]  * p.z + T[7];
    o.z = T[8] * p.x + T[9] * p.y + T[10] * p.z + T[11];
    o.w = p.w;
    return o;
}

// ---------------------------------------------------------------------------------------------
// pcl::VoxelGrid<PointXYZI>::applyFilter (PCL 1.10 defaults), used at
// src/mapOptmization.cpp:954-955 (local map) and :972-973 (current scan).
// Canonical within-voxel order: ascending source index.  Returns number of output points.
struct VoxKey { int idx; int src; };

int voxelgrid(const orc_point* in, int n, float leaf, orc_point* out) {
    if (n <= 0) return 0;
    const float inv = 1.0f / leaf;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = 0; i < n; ++i) {
        const float v[3] = {in[i].x, in[i].y, in[i].z};
        if (!std::isfinite(v[0]) || !std::isfinite(v[1]) || !std::isfinite(v[2])) continue;
        for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], v[a]); mx[a] = std::max(mx[a], v[a]); }
    }
    const int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1;
    const int64_t dy = (int64_t)((mx[1] - mn[1]) * inv) + 1;
    const int64_t dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
    if (dx * dy * dz > (int64_t)INT32_MAX) {   // "Leaf size is too small": output = input unchanged
        std::memcpy(out, in, sizeof(orc_point) * (size_t)n);
        return n;
    }
    int minb[3], maxb[3], divb[3];
    for (int a = 0; a < 3; ++a) {
        minb[a] = (int)std::floor(mn[a] * inv);
        maxb[a] = (int)std::floor(mx[a] * inv);
        divb[a] = maxb[a] - minb[a] + 1;
    }
    const int mul[3] = {1, divb[0], divb[0] * divb[1]};
    std::vector<VoxKey> keys((size_t)n);
    for (int i = 0; i < n; ++i) {
        const int i0 = (int)(std::floor(in[i].x * inv) - (float)minb[0]);
        const int i1 = (int)(std::floor(in[i].y * inv) - (float)minb[1]);
        const int i2 = (int)(std::floor(in[i].z * inv) - (float)minb[2]);
        keys[i].idx = i0 * mul[0] + i1 * mul[1] + i2 * mul[2];
        keys[i].src = i;
    }
    std::sort(keys.begin(), keys.end(), [](const VoxKey& a, const VoxKey& b) {
        return a.idx < b.idx || (a.idx == b.idx && a.src < b.src);
    });
    int m = 0;
    size_t i = 0;
    while (i < keys.size()) {
        size_t j = i;
        float sx = 0.f, sy = 0.f, sz = 0.f, sw = 0.f;
        while (j < keys.size() && keys[j].idx == keys[i].idx) {
            const orc_point& p = in[keys[j].src];
            sx += p.x; sy += p.y; sz += p.z; sw += p.w;
            ++j;
        }
        const float cnt = (float)(j - i);
        out[m].x = sx / cnt; out[m].y = sy / cnt; out[m].z = sz / cnt; out[m].w = sw / cnt;
        ++m;
        i = j;
    }
    return m;
}

// ---------------------------------------------------------------------------------------------
// Exact 5-NN.  pcl::KdTreeFLANN::nearestKSearch at src/mapOptmization.cpp:992:
// exact, sorted ascending; canonical tie-break by ascending index.
struct Top5 {
    float d[5]; int id[5]; int n;
    void init() { n = 0; for (int k = 0; k < 5; ++k) { d[k] = FLT_MAX; id[k] = -1; } }
    static inline bool less(float da, int ia, float db, int ib) { return da < db || (da == db && ia < ib); }
    inline bool accepts(float dd, int ii) const { return n < 5 || less(dd, ii, d[4], id[4]); }
    inline void push(float dd, int ii) {
        if (!accepts(dd, ii)) return;
        int k = (n < 5) ? n : 4;
        while (k > 0 && less(dd, ii, d[k - 1], id[k - 1])) { d[k] = d[k - 1]; id[k] = id[k - 1]; --k; }
        d[k] = dd; id[k] = ii;
        if (n < 5) ++n;
    }
};

inline float dist2(const orc_point& q, const orc_point& p) {
    const float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
    return ((dx * dx) + (dy * dy)) + (dz * dz);
}

void knn5_brute_one(const orc_point* map, int M, const orc_point& q, int* idx, float* d2) {
    Top5 t; t.init();
    for (int j = 0; j < M; ++j) t.push(dist2(q, map[j]), j);
    for (int k = 0; k < 5; ++k) { idx[k] = t.id[k]; d2[k] = t.d[k]; }
}

// kd-tree of FLANN KDTreeSingleIndex shape (leaf_max_size 15, points reordered) so that the CPU
// baseline has the reference's asymptotics (kdtreeSurfFromMap->setInputCloud src/mapOptmization.cpp:1188
// rebuilds it every frame).  Results are exact and use the canonical (d2, index) order.
struct KdTree {
    struct Node { int left, right; int lo, hi; int dim; float split_lo, split_hi; };
    std::vector<Node> nodes;
    std::vector<orc_point> pts;     // reordered copy
    std::vector<int> ids;           // original indices
    float bbmin[3], bbmax[3];
    static constexpr int kLeaf = 15;

    static inline float coord(const orc_point& p, int d) { return d == 0 ? p.x : (d == 1 ? p.y : p.z); }

    void build(const orc_point* map, int M) {
        nodes.clear(); pts.assign(map, map + M); ids.resize((size_t)M);
        for (int i = 0; i < M; ++i) ids[i] = i;
        for (int a = 0; a < 3; ++a) { bbmin[a] = FLT_MAX; bbmax[a] = -FLT_MAX; }
        for (int i = 0; i < M; ++i) for (int a = 0; a < 3; ++a) {
            const float v = coord(pts[i], a);
            bbmin[a] = std::min(bbmin[a], v); bbmax[a] = std::max(bbmax[a], v);
        }
        if (M > 0) { nodes.reserve((size_t)(2 * M / kLeaf + 8)); divide(0, M); }
    }

    int divide(int lo, int hi) {
        const int me = (int)nodes.size();
        nodes.push_back(Node{-1, -1, lo, hi, -1, 0.f, 0.f});
        if (hi - lo <= kLeaf) return me;
        float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (int i = lo; i < hi; ++i) for (int a = 0; a < 3; ++a) {
            const float v = coord(pts[i], a);
            mn[a] = std::min(mn[a], v); mx[a] = std::max(mx[a], v);
        }
        int dim = 0; float span = mx[0] - mn[0];
        for (int a = 1; a < 3; ++a) if (mx[a] - mn[a] > span) { span = mx[a] - mn[a]; dim = a; }
        if (span <= 0.f) return me;   // all coincident: keep as an oversized leaf
        const int mid = lo + (hi - lo) / 2;
        // order by (coord, original id) so the split is deterministic
        std::vector<int> ord((size_t)(hi - lo));
        for (int i = lo; i < hi; ++i) ord[i - lo] = i;
        std::nth_element(ord.begin(), ord.begin() + (mid - lo), ord.end(), [&](int a, int b) {
            const float ca = coord(pts[a], dim), cb = coord(pts[b], dim);
            return ca < cb || (ca == cb && ids[a] < ids[b]);
        });
        std::vector<orc_point> tp((size_t)(hi - lo)); std::vector<int> ti((size_t)(hi - lo));
        for (int k = 0; k < hi - lo; ++k) { tp[k] = pts[ord[k]]; ti[k] = ids[ord[k]]; }
        for (int k = 0; k < hi - lo; ++k) { pts[lo + k] = tp[k]; ids[lo + k] = ti[k]; }
        float lmax = -FLT_MAX, rmin = FLT_MAX;
        for (int i = lo; i < mid; ++i) lmax = std::max(lmax, coord(pts[i], dim));
        for (int i = mid; i < hi; ++i) rmin = std::min(rmin, coord(pts[i], dim));
        nodes[me].dim = dim; nodes[me].split_lo = lmax; nodes[me].split_hi = rmin;
        const int l = divide(lo, mid);
        const int r = divide(mid, hi);
        nodes[me].left = l; nodes[me].right = r;
        return me;
    }

    void search(int ni, const orc_point& q, Top5& t) const {
        const Node& nd = nodes[ni];
        if (nd.dim < 0) {
            for (int i = nd.lo; i < nd.hi; ++i) t.push(dist2(q, pts[i]), ids[i]);
            return;
        }
        const float v = coord(q, nd.dim);
        const float dl = v - nd.split_lo, dh = nd.split_hi - v;   // >0: outside that child's slab
        int first, second; float gap;
        if (dl <= dh) { first = nd.left;  second = nd.right; gap = dh; }
        else          { first = nd.right; second = nd.left;  gap = dl; }
        search(first, q, t);
        // the far child can only hold points with d2 >= gap*gap; keep it unless that is strictly worse
        // than the current 5th (ties at equal d2 may still win on index)
        if (t.n < 5 || gap <= 0.f || gap * gap <= t.d[4]) search(second, q, t);
    }

    void knn5(const orc_point& q, int* idx, float* d2) const {
        Top5 t; t.init();
        if (!nodes.empty()) search(0, q, t);
        for (int k = 0; k < 5; ++k) { idx[k] = t.id[k]; d2[k] = t.d[k]; }
    }
};

// ---------------------------------------------------------------------------------------------
// Eigen::Matrix<float,5,3>::colPivHouseholderQr().solve(b), b = -1 (src/mapOptmization.cpp:994-1009),
// restated from the Eigen 3.3 algorithm with sequential float sums.
void plane_lsq_5x3(const float P[5][3], float x[3]) {
    float q[3][5];                      // column-major working copy
    for (int j = 0; j < 3; ++j) for (int i = 0; i < 5; ++i) q[j][i] = P[i][j];
    int perm[3] = {0, 1, 2};
    float hco[3] = {0.f, 0.f, 0.f};
    float nUpd[3], nDir[3];
    for (int j = 0; j < 3; ++j) {
        float s = 0.f;
        for (int i = 0; i < 5; ++i) s += q[j][i] * q[j][i];
        nDir[j] = nUpd[j] = std::sqrt(s);
    }
    const float eps = FLT_EPSILON;
    float mxn = nUpd[0]; if (nUpd[1] > mxn) mxn = nUpd[1]; if (nUpd[2] > mxn) mxn = nUpd[2];
    const float thr_helper = ((mxn * eps) * (mxn * eps)) / 5.0f;
    const float downdate_thr = std::sqrt(eps);
    int nonzero = 3;
    for (int k = 0; k < 3; ++k) {
        int big = k; float bign = nUpd[k];
        for (int j = k + 1; j < 3; ++j) if (nUpd[j] > bign) { bign = nUpd[j]; big = j; }
        if (nonzero == 3 && bign * bign < thr_helper * (float)(5 - k)) nonzero = k;
        if (big != k) {
            for (int i = 0; i < 5; ++i) std::swap(q[k][i], q[big][i]);
            std::swap(nUpd[k], nUpd[big]); std::swap(nDir[k], nDir[big]); std::swap(perm[k], perm[big]);
        }
        // Householder reflector for rows k..4 of column k
        float tailSq = 0.f;
        for (int i = k + 1; i < 5; ++i) tailSq += q[k][i] * q[k][i];
        const float c0 = q[k][k];
        float beta, tau;
        if (tailSq <= FLT_MIN) {
            tau = 0.f; beta = c0;
            for (int i = k + 1; i < 5; ++i) q[k][i] = 0.f;
        } else {
            beta = std::sqrt(c0 * c0 + tailSq);
            if (c0 >= 0.f) beta = -beta;
            const float den = c0 - beta;
            for (int i = k + 1; i < 5; ++i) q[k][i] = q[k][i] / den;
            tau = (beta - c0) / beta;
        }
        q[k][k] = beta; hco[k] = tau;
        for (int j = k + 1; j < 3; ++j) {
            float tmp = 0.f;
            for (int i = k + 1; i < 5; ++i) tmp += q[k][i] * q[j][i];
            tmp += q[j][k];
            q[j][k] -= tau * tmp;
            for (int i = k + 1; i < 5; ++i) q[j][i] -= (tau * q[k][i]) * tmp;
        }
        for (int j = k + 1; j < 3; ++j) {
            if (nUpd[j] != 0.f) {
                float temp = std::fabs(q[j][k]) / nUpd[j];
                temp = (1.0f + temp) * (1.0f - temp);
                if (temp < 0.f) temp = 0.f;
                const float r = nUpd[j] / nDir[j];
                const float temp2 = temp * (r * r);
                if (temp2 <= downdate_thr) {
                    float s = 0.f;
                    for (int i = k + 1; i < 5; ++i) s += q[j][i] * q[j][i];
                    nDir[j] = std::sqrt(s); nUpd[j] = nDir[j];
                } else {
                    nUpd[j] *= std::sqrt(temp);
                }
            }
        }
    }
    float c[5] = {-1.f, -1.f, -1.f, -1.f, -1.f};
    for (int k = 0; k < nonzero; ++k) {
        float tmp = 0.f;
        for (int i = k + 1; i < 5; ++i) tmp += q[k][i] * c[i];
        tmp += c[k];
        c[k] -= hco[k] * tmp;
        for (int i = k + 1; i < 5; ++i) c[i] -= (hco[k] * q[k][i]) * tmp;
    }
    for (int i = nonzero - 1; i >= 0; --i) {          // column-oriented back substitution
        c[i] = c[i] / q[i][i];
        for (int r = 0; r < i; ++r) c[r] -= q[i][r] * c[i];
    }
    x[0] = x[1] = x[2] = 0.f;
    for (int i = 0; i < nonzero; ++i) x[perm[i]] = c[i];
}

// ---------------------------------------------------------------------------------------------
// Residual of one scan point: loop body of surfOptimization src/mapOptmization.cpp:986-1046.
// Returns true when the correspondence is kept (s > 0.1); coeff = (s*a, s*b, s*c, s*pd2).
bool surf_residual(const orc_point* map, const int* idx5, const float* d2, const orc_point& ori,
                   const orc_point& sel, orc_point* coeff) {
    if (idx5[4] < 0) return false;                        // fewer than 5 map points (SURVEY A.2 edge)
    if (!(d2[4] < 1.0f)) return false;                    // :1002
    float A[5][3];
    for (int j = 0; j < 5; ++j) { A[j][0] = map[idx5[j]].x; A[j][1] = map[idx5[j]].y; A[j][2] = map[idx5[j]].z; }
    float X[3];
    plane_lsq_5x3(A, X);                                  // :1009
    float pa = X[0], pb = X[1], pc = X[2], pd = 1.f;
    const float ps = std::sqrt(pa * pa + pb * pb + pc * pc);   // :1016
    pa /= ps; pb /= ps; pc /= ps; pd /= ps;
    for (int j = 0; j < 5; ++j) {                         // :1020-1027
        const float r = pa * A[j][0] + pb * A[j][1] + pc * A[j][2] + pd;
        if ((double)std::fabs(r) > 0.2) return false;
    }
    const float pd2 = pa * sel.x + pb * sel.y + pc * sel.z + pd;                      // :1030
    const float rng = std::sqrt(std::sqrt(ori.x * ori.x + ori.y * ori.y + ori.z * ori.z));
    const float s = (float)(1 - 0.9 * (double)std::fabs(pd2) / (double)rng);          // :1032 (double expr)
    coeff->x = s * pa; coeff->y = s * pb; coeff->z = s * pc; coeff->w = s * pd2;
    return (double)s > 0.1;                               // :1040
}

// ---------------------------------------------------------------------------------------------
// 6x6 float32 linear algebra standing in for OpenCV.

// cv::solve(AtA, AtB, X, DECOMP_QR) src/mapOptmization.cpp:1130: Householder QR, float32.
void solve6_qr(const float A_in[36], const float b_in[6], float x[6]) {
    float A[6][6], b[6];
    for (int i = 0; i < 6; ++i) { b[i] = b_in[i]; for (int j = 0; j < 6; ++j) A[i][j] = A_in[i * 6 + j]; }
    for (int k = 0; k < 6; ++k) {
        float tailSq = 0.f;
        for (int i = k + 1; i < 6; ++i) tailSq += A[i][k] * A[i][k];
        const float c0 = A[k][k];
        if (tailSq <= FLT_MIN) continue;                  // already upper-triangular in this column
        float beta = std::sqrt(c0 * c0 + tailSq);
        if (c0 >= 0.f) beta = -beta;
        const float den = c0 - beta;
        float v[6];
        for (int i = k + 1; i < 6; ++i) v[i] = A[i][k] / den;
        const float tau = (beta - c0) / beta;
        A[k][k] = beta;
        for (int i = k + 1; i < 6; ++i) A[i][k] = 0.f;
        for (int j = k + 1; j < 6; ++j) {
            float tmp = 0.f;
            for (int i = k + 1; i < 6; ++i) tmp += v[i] * A[i][j];
            tmp += A[k][j];
            A[k][j] -= tau * tmp;
            for (int i = k + 1; i < 6; ++i) A[i][j] -= (tau * v[i]) * tmp;
        }
        float tmp = 0.f;
        for (int i = k + 1; i < 6; ++i) tmp += v[i] * b[i];
        tmp += b[k];
        b[k] -= tau * tmp;
        for (int i = k + 1; i < 6; ++i) b[i] -= (tau * v[i]) * tmp;
    }
    for (int i = 5; i >= 0; --i) {
        float s = b[i];
        for (int j = i + 1; j < 6; ++j) s -= A[i][j] * x[j];
        x[i] = s / A[i][i];
    }
}

// cv::eigen(sym 6x6) src/mapOptmization.cpp:1138: Jacobi rotations in float32; eigenvalues sorted
// descending, eigenvectors returned as ROWS of V.  Cyclic-by-row sweeps.
void eigen6_jacobi(const float A_in[36], float w[6], float V[36]) {
    float A[6][6], U[6][6];
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { A[i][j] = A_in[i * 6 + j]; U[i][j] = (i == j) ? 1.f : 0.f; }
    for (int sweep = 0; sweep < 30; ++sweep) {
        float off = 0.f, diag = 0.f;
        for (int i = 0; i < 6; ++i) { diag += A[i][i] * A[i][i]; for (int j = i + 1; j < 6; ++j) off += A[i][j] * A[i][j]; }
        if (off <= (FLT_EPSILON * FLT_EPSILON) * diag || off == 0.f) break;
        for (int p = 0; p < 5; ++p) for (int q = p + 1; q < 6; ++q) {
            const float apq = A[p][q];
            if (apq == 0.f) continue;
            const float theta = (A[q][q] - A[p][p]) / (2.0f * apq);
            const float t = (theta >= 0.f ? 1.0f : -1.0f) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0f));
            const float c = 1.0f / std::sqrt(t * t + 1.0f);
            const float s = t * c;
            const float h = t * apq;
            A[p][p] -= h; A[q][q] += h; A[p][q] = 0.f; A[q][p] = 0.f;
            for (int k = 0; k < 6; ++k) {                 // symmetric update of the off-diagonal rows/columns
                if (k == p || k == q) continue;
                const float akp = A[k][p], akq = A[k][q];
                const float np_ = c * akp - s * akq, nq_ = s * akp + c * akq;
                A[k][p] = np_; A[p][k] = np_; A[k][q] = nq_; A[q][k] = nq_;
            }
            for (int k = 0; k < 6; ++k) {                 // rows of U are the eigenvectors
                const float upk = U[p][k], uqk = U[q][k];
                U[p][k] = c * upk - s * uqk; U[q][k] = s * upk + c * uqk;
            }
        }
    }
    int ord[6] = {0, 1, 2, 3, 4, 5};
    for (int i = 0; i < 5; ++i) {                         // selection sort, descending, stable on ties
        int m = i;
        for (int j = i + 1; j < 6; ++j) if (A[ord[j]][ord[j]] > A[ord[m]][ord[m]]) m = j;
        const int t = ord[m]; for (int j = m; j > i; --j) ord[j] = ord[j - 1]; ord[i] = t;
    }
    for (int i = 0; i < 6; ++i) { w[i] = A[ord[i]][ord[i]]; for (int k = 0; k < 6; ++k) V[i * 6 + k] = U[ord[i]][k]; }
}

// Degeneracy projector src/mapOptmization.cpp:1132-1154: zero the rows of V2 whose eigenvalue is
// below 100 walking from the smallest; matP = V^-1 * V2 with V orthonormal => V^T * V2.
bool degeneracy_projector(const float AtA[36], float matP[36]) {
    float w[6], V[36], V2[36];
    eigen6_jacobi(AtA, w, V);
    std::memcpy(V2, V, sizeof(V2));
    bool degenerate = false;
    for (int i = 5; i >= 0; --i) {
        if (w[i] < 100.f) { for (int j = 0; j < 6; ++j) V2[i * 6 + j] = 0.f; degenerate = true; }
        else break;
    }
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) {
        double s = 0.0;
        for (int k = 0; k < 6; ++k) s += (double)V[k * 6 + i] * (double)V2[k * 6 + j];
        matP[i * 6 + j] = (float)s;
    }
    return degenerate;
}

// Jacobian row of LMOptimization src/mapOptmization.cpp:1093-1125 (LOAM camera-frame relabelling:
// srx=sin(yaw) sry=sin(pitch) srz=sin(roll)).  row = [d/droll, d/dpitch, d/dyaw, cx, cy, cz], rhs = -c.w
struct Trig { float srx, crx, sry, cry, srz, crz; };
inline Trig pose_trig(const float* pose6) {
    Trig t;
    t.srx = sin_c(pose6[2]); t.crx = cos_c(pose6[2]);
    t.sry = sin_c(pose6[1]); t.cry = cos_c(pose6[1]);
    t.srz = sin_c(pose6[0]); t.crz = cos_c(pose6[0]);
    return t;
}
inline void jacobian_row(const Trig& g, const orc_point& p, const orc_point& c, float row[6], float* rhs) {
    const float srx = g.srx, crx = g.crx, sry = g.sry, cry = g.cry, srz = g.srz, crz = g.crz;
    const float arx = (-srx * cry * p.x - (srx * sry * srz + crx * crz) * p.y + (crx * srz - srx * sry * crz) * p.z) * c.x
                    + (crx * cry * p.x - (srx * crz - crx * sry * srz) * p.y + (crx * sry * crz + srx * srz) * p.z) * c.y;
    const float ary = (-crx * sry * p.x + crx * cry * srz * p.y + crx * cry * crz * p.z) * c.x
                    + (-srx * sry * p.x + srx * sry * srz * p.y + srx * cry * crz * p.z) * c.y
                    + (-cry * p.x - sry * srz * p.y - sry * crz * p.z) * c.z;
    const float arz = ((crx * sry * crz + srx * srz) * p.y + (srx * crz - crx * sry * srz) * p.z) * c.x
                    + ((-crx * srz + srx * sry * crz) * p.y + (-srx * sry * srz - crx * crz) * p.z) * c.y
                    + (cry * crz * p.y - cry * srz * p.z) * c.z;
    row[0] = arz; row[1] = ary; row[2] = arx; row[3] = c.x; row[4] = c.y; row[5] = c.z;
    *rhs = -c.w;
}

// LMOptimization src/mapOptmization.cpp:1061-1181.  state->is_degenerate / matP persist (members :90-91).
// Returns 1 converged, 0 not converged, -1 skipped (<min_sel correspondences).
int lm_step(const orc_point* ori, const orc_point* coef, int n_sel, int iter, int min_sel,
            float* pose6, orc_lm_state* st, float* AtA_out, float* AtB_out) {
    if (n_sel < min_sel) return -1;                                   // :1079-1082
    const Trig g = pose_trig(pose6);
    double acc[27];
    for (int k = 0; k < 27; ++k) acc[k] = 0.0;
    for (int i = 0; i < n_sel; ++i) {
        float r[6], b;
        jacobian_row(g, ori[i], coef[i], r, &b);
        int k = 0;
        for (int a = 0; a < 6; ++a) for (int c = a; c < 6; ++c) acc[k++] += (double)r[a] * (double)r[c];
        for (int a = 0; a < 6; ++a) acc[21 + a] += (double)r[a] * (double)b;
    }
    float AtA[36], AtB[6], X[6];
    { int k = 0; for (int a = 0; a < 6; ++a) for (int c = a; c < 6; ++c) { AtA[a * 6 + c] = AtA[c * 6 + a] = (float)acc[k++]; } }
    for (int a = 0; a < 6; ++a) AtB[a] = (float)acc[21 + a];
    if (AtA_out) std::memcpy(AtA_out, AtA, sizeof(AtA));
    if (AtB_out) std::memcpy(AtB_out, AtB, sizeof(AtB));
    solve6_qr(AtA, AtB, X);                                            // :1130
    if (iter == 0) st->is_degenerate = degeneracy_projector(AtA, st->matP) ? 1 : 0;   // :1132-1154
    if (st->is_degenerate) {                                           // :1156-1160
        float X2[6];
        for (int i = 0; i < 6; ++i) {
            double s = 0.0;
            for (int k = 0; k < 6; ++k) s += (double)st->matP[i * 6 + k] * (double)X[k];
            X2[i] = (float)s;
        }
        std::memcpy(X, X2, sizeof(X));
    }
    for (int k = 0; k < 6; ++k) pose6[k] += X[k];                      // :1162-1167
    const float r0 = X[0] * 57.29578f, r1 = X[1] * 57.29578f, r2 = X[2] * 57.29578f;   // pcl::rad2deg(float)
    const float t0 = X[3] * 100, t1 = X[4] * 100, t2 = X[5] * 100;
    const float deltaR = (float)std::sqrt((double)r0 * (double)r0 + (double)r1 * (double)r1 + (double)r2 * (double)r2);
    const float deltaT = (float)std::sqrt((double)t0 * (double)t0 + (double)t1 * (double)t1 + (double)t2 * (double)t2);
    return ((double)deltaR < 0.05 && (double)deltaT < 0.05) ? 1 : 0;   // :1177
}

// ---------------------------------------------------------------------------------------------
// Upstream LIO-SAM corner path.  NOT in the reference (it dropped feature extraction and
// cornerOptimization; remnants: laserCloudCornerTemp src/mapOptmization.cpp:947,950 and the unused
// msg/cloud_info.msg:28 field); the north_star names it, so it is restated from upstream LIO-SAM
// (TixiaoShan/LIO-SAM mapOptmization.cpp cornerOptimization) as an optional
// superset.  Parity for this class is GPU-vs-this-file only.

// cv::eigen(sym 3x3): same cyclic Jacobi as eigen6_jacobi, eigenvalues descending, eigenvectors as rows.
void eigen3_jacobi(const float A_in[9], float w[3], float V[9]) {
    float A[3][3], U[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { A[i][j] = A_in[i * 3 + j]; U[i][j] = (i == j) ? 1.f : 0.f; }
    for (int sweep = 0; sweep < 30; ++sweep) {
        float off = 0.f, diag = 0.f;
        for (int i = 0; i < 3; ++i) { diag += A[i][i] * A[i][i]; for (int j = i + 1; j < 3; ++j) off += A[i][j] * A[i][j]; }
        if (off <= (FLT_EPSILON * FLT_EPSILON) * diag || off == 0.f) break;
        for (int p = 0; p < 2; ++p) for (int q = p + 1; q < 3; ++q) {
            const float apq = A[p][q];
            if (apq == 0.f) continue;
            const float theta = (A[q][q] - A[p][p]) / (2.0f * apq);
            const float t = (theta >= 0.f ? 1.0f : -1.0f) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0f));
            const float c = 1.0f / std::sqrt(t * t + 1.0f);
            const float s = t * c;
            const float h = t * apq;
            A[p][p] -= h; A[q][q] += h; A[p][q] = 0.f; A[q][p] = 0.f;
            for (int k = 0; k < 3; ++k) {
                if (k == p || k == q) continue;
                const float akp = A[k][p], akq = A[k][q];
                const float np_ = c * akp - s * akq, nq_ = s * akp + c * akq;
                A[k][p] = np_; A[p][k] = np_; A[k][q] = nq_; A[q][k] = nq_;
            }
            for (int k = 0; k < 3; ++k) {
                const float upk = U[p][k], uqk = U[q][k];
                U[p][k] = c * upk - s * uqk; U[q][k] = s * upk + c * uqk;
            }
        }
    }
    int ord[3] = {0, 1, 2};
    for (int i = 0; i < 2; ++i) {                         // selection sort, descending, stable on ties
        int m = i;
        for (int j = i + 1; j < 3; ++j) if (A[ord[j]][ord[j]] > A[ord[m]][ord[m]]) m = j;
        const int t = ord[m]; for (int j = m; j > i; --j) ord[j] = ord[j - 1]; ord[i] = t;
    }
    for (int i = 0; i < 3; ++i) { w[i] = A[ord[i]][ord[i]]; for (int k = 0; k < 3; ++k) V[i * 3 + k] = U[ord[i]][k]; }
}

// Loop body of upstream cornerOptimization after the kNN: centroid and covariance of the 5 neighbours,
// line test (largest eigenvalue > 3 x second), point-to-line distance and weight.  Type promotions as
// upstream: `0.1 * v` and `1 - 0.9 * fabs(ld2)` are double expressions narrowed to float.
bool line_residual(const float P[5][3], const orc_point& sel, orc_point* coeff) {
    float cx = 0.f, cy = 0.f, cz = 0.f;
    for (int j = 0; j < 5; ++j) { cx += P[j][0]; cy += P[j][1]; cz += P[j][2]; }
    cx /= 5.f; cy /= 5.f; cz /= 5.f;
    float a11 = 0.f, a12 = 0.f, a13 = 0.f, a22 = 0.f, a23 = 0.f, a33 = 0.f;
    for (int j = 0; j < 5; ++j) {
        const float ax = P[j][0] - cx, ay = P[j][1] - cy, az = P[j][2] - cz;
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
    const float a012 = std::sqrt(m11 * m11 + m12 * m12 + m13 * m13);
    const float l12 = std::sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
    const float la = ((y1 - y2) * m11 + (z1 - z2) * m12) / a012 / l12;
    const float lb = -((x1 - x2) * m11 - (z1 - z2) * m13) / a012 / l12;
    const float lc = -((x1 - x2) * m12 + (y1 - y2) * m13) / a012 / l12;
    const float ld2 = a012 / l12;
    const float s = (float)(1 - 0.9 * (double)std::fabs(ld2));
    coeff->x = s * la; coeff->y = s * lb; coeff->z = s * lc; coeff->w = s * ld2;
    return (double)s > 0.1;
}

bool corner_residual(const orc_point* map, const int* idx5, const float* d2, const orc_point& sel, orc_point* coeff) {
    if (idx5[4] < 0) return false;
    if (!(d2[4] < 1.0f)) return false;
    float P[5][3];
    for (int j = 0; j < 5; ++j) { P[j][0] = map[idx5[j]].x; P[j][1] = map[idx5[j]].y; P[j][2] = map[idx5[j]].z; }
    return line_residual(P, sel, coeff);
}

double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

}  // namespace

// =============================================================================================
// extern "C" surface (ctypes)
// =============================================================================================
extern "C" {

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_pose_to_T(const float* pose6, float* T12) { pose_to_T(pose6, T12); }

void orc_transform_cloud(const orc_point* in, int n, const float* pose6, orc_point* out, int threads) {
    float T[12]; pose_to_T(pose6, T);
    (void)threads;
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(static)
    for (int i = 0; i < n; ++i) out[i] = apply_T(T, in[i]);
}

void orc_apply_T(const orc_point* in, int n, const float* T12, orc_point* out) {
    for (int i = 0; i < n; ++i) out[i] = apply_T(T12, in[i]);
}

int orc_voxelgrid(const orc_point* in, int n, float leaf, orc_point* out) { return voxelgrid(in, n, leaf, out); }

void orc_knn5_brute(const orc_point* map, int M, const orc_point* q, int nq, int* idx5, float* d2, int threads) {
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(static)
    for (int i = 0; i < nq; ++i) knn5_brute_one(map, M, q[i], idx5 + 5 * (size_t)i, d2 + 5 * (size_t)i);
}

void orc_knn5_kdtree(const orc_point* map, int M, const orc_point* q, int nq, int* idx5, float* d2, int threads) {
    KdTree t; t.build(map, M);
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(static)
    for (int i = 0; i < nq; ++i) t.knn5(q[i], idx5 + 5 * (size_t)i, d2 + 5 * (size_t)i);
}

void orc_plane_lsq(const float* pts15, float* x3) {
    float P[5][3];
    for (int i = 0; i < 5; ++i) for (int j = 0; j < 3; ++j) P[i][j] = pts15[i * 3 + j];
    plane_lsq_5x3(P, x3);
}

void orc_solve6(const float* A36, const float* b6, float* x6) { solve6_qr(A36, b6, x6); }
void orc_eigen6(const float* A36, float* w6, float* V36) { eigen6_jacobi(A36, w6, V36); }
int  orc_degeneracy(const float* A36, float* P36) { return degeneracy_projector(A36, P36) ? 1 : 0; }

void orc_jacobian_row(const float* pose6, const orc_point* ori, const orc_point* coef, float* row6, float* rhs) {
    const Trig g = pose_trig(pose6);
    jacobian_row(g, *ori, *coef, row6, rhs);
}

// One pass of surfOptimization + combineOptimizationCoeffs (src/mapOptmization.cpp:981-1059) at a
// fixed pose.  knn: 0 brute force, 1 kd-tree.  Outputs are compacted in ascending i; sel_index gets
// the scan index of every kept correspondence.  Returns n_sel.
int orc_surf_pass(const orc_point* map, int M, const orc_point* scan, int Q_all, int stride,
                  const float* pose6, int knn, int threads, orc_point* ori_out, orc_point* coef_out,
                  int* sel_index, int* idx5_out /*Q_all*5 or NULL*/, float* d2_out /*Q_all*5 or NULL*/) {
    float T[12]; pose_to_T(pose6, T);
    KdTree tree; if (knn == 1) tree.build(map, M);
    std::vector<uint8_t> flag((size_t)Q_all, 0);
    std::vector<orc_point> coef((size_t)Q_all);
    const int nq = (Q_all + stride - 1) / stride;
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(static)
    for (int k = 0; k < nq; ++k) {
        const int i = k * stride;
        const orc_point sel = apply_T(T, scan[i]);
        int idx[5]; float d2[5];
        if (knn == 1) tree.knn5(sel, idx, d2); else knn5_brute_one(map, M, sel, idx, d2);
        if (idx5_out) for (int j = 0; j < 5; ++j) { idx5_out[5 * (size_t)i + j] = idx[j]; d2_out[5 * (size_t)i + j] = d2[j]; }
        orc_point c;
        if (surf_residual(map, idx, d2, scan[i], sel, &c)) { coef[i] = c; flag[i] = 1; }
    }
    int n = 0;
    for (int i = 0; i < Q_all; ++i) if (flag[i]) { ori_out[n] = scan[i]; coef_out[n] = coef[i]; if (sel_index) sel_index[n] = i; ++n; }
    return n;
}

int orc_lm_step(const orc_point* ori, const orc_point* coef, int n_sel, int iter, int min_sel,
                float* pose6, orc_lm_state* st, float* AtA36, float* AtB6) {
    return lm_step(ori, coef, n_sel, iter, min_sel, pose6, st, AtA36, AtB6);
}

// scan2MapOptimization src/mapOptmization.cpp:1183-1201 minus transformUpdate (:1202, host side).
// Returns 0 ok, 1 skipped (Q_all <= min_points, :1187).
int orc_scan2map(const orc_point* map, int M, const orc_point* scan, int Q_all, float* pose6,
                 const orc_s2m_opts* o, orc_lm_state* st, orc_s2m_result* res) {
    std::memset(res, 0, sizeof(*res));
    if (!(Q_all > o->min_points)) { res->skipped = 1; return 1; }
    KdTree tree;
    if (o->knn == 1) tree.build(map, M);                               // :1188
    std::vector<uint8_t> flag((size_t)Q_all, 0);
    std::vector<orc_point> coefv((size_t)Q_all), ori, coef;
    ori.reserve((size_t)Q_all); coef.reserve((size_t)Q_all);
    const int stride = o->stride, nq = (Q_all + stride - 1) / stride;
    const int threads = o->threads > 0 ? o->threads : 1;
    for (int iter = 0; iter < o->max_iters; ++iter) {                  // :1190
        ori.clear(); coef.clear();
        float T[12]; pose_to_T(pose6, T);                              // :982
#pragma omp parallel for num_threads(threads) schedule(static)
        for (int k = 0; k < nq; ++k) {                                 // :984-985
            const int i = k * stride;
            const orc_point sel = apply_T(T, scan[i]);
            int idx[5]; float d2[5];
            if (o->knn == 1) tree.knn5(sel, idx, d2); else knn5_brute_one(map, M, sel, idx, d2);
            orc_point c;
            if (surf_residual(map, idx, d2, scan[i], sel, &c)) { coefv[i] = c; flag[i] = 1; }
        }
        for (int i = 0; i < Q_all; ++i) if (flag[i]) { ori.push_back(scan[i]); coef.push_back(coefv[i]); flag[i] = 0; }   // :1050-1059
        res->iters = iter + 1;
        res->n_sel = (int)ori.size();
        const int r = lm_step(ori.data(), coef.data(), (int)ori.size(), iter, o->min_sel, pose6, st, nullptr, nullptr);
        if (r == 1) { res->converged = 1; break; }                     // :1198-1199
        if (r == -1) { res->too_few = 1; break; }   // pose unchanged => every later iteration is identical (:1080)
    }
    res->is_degenerate = st->is_degenerate;
    return 0;
}

// extractCloud src/mapOptmization.cpp:936-961 for an already-selected keyframe list: transform each
// keyframe by its pose (include/utility.h:388-407), concatenate in the given order, VoxelGrid.
// kf_offsets has K+1 entries into kf_points.  raw_out may be NULL.  Returns M.
int orc_localmap_build(const orc_point* kf_points, const int64_t* kf_offsets, const float* poses6, int K,
                       float leaf, int threads, orc_point* map_out, int64_t* n_raw_out) {
    const int64_t n_raw = kf_offsets[K] - kf_offsets[0];
    std::vector<orc_point> raw((size_t)n_raw);
    int64_t w = 0;
    for (int k = 0; k < K; ++k) {
        const int64_t n = kf_offsets[k + 1] - kf_offsets[k];
        orc_transform_cloud(kf_points + kf_offsets[k], (int)n, poses6 + 6 * (size_t)k, raw.data() + w, threads);
        w += n;
    }
    if (n_raw_out) *n_raw_out = n_raw;
    return voxelgrid(raw.data(), (int)n_raw, leaf, map_out);
}

// One whole reference frame for CPU timing: assembly + map VoxelGrid + scan VoxelGrid + scan2map,
// keeping the reference's redundancy (kd-tree rebuilt per frame, single-threaded VoxelGrid, OpenMP
// only over the residual loop and the keyframe transform).  ms_out[4] = {assemble+mapDS, scanDS, s2m, total}.
int orc_frame(const orc_point* kf_points, const int64_t* kf_offsets, const float* poses6, int K, float map_leaf,
              const orc_point* scan_raw, int n_scan, float scan_leaf, float* pose6, const orc_s2m_opts* o,
              orc_lm_state* st, orc_s2m_result* res, int* M_out, int* Q_out, double* ms_out) {
    const double t0 = now_ms();
    const int64_t n_raw = kf_offsets[K] - kf_offsets[0];
    std::vector<orc_point> map((size_t)std::max<int64_t>(n_raw, 1));
    const int M = orc_localmap_build(kf_points, kf_offsets, poses6, K, map_leaf, o->threads, map.data(), nullptr);
    const double t1 = now_ms();
    std::vector<orc_point> ds((size_t)std::max(n_scan, 1));
    const int Q_all = voxelgrid(scan_raw, n_scan, scan_leaf, ds.data());
    const double t2 = now_ms();
    const int rc = orc_scan2map(map.data(), M, ds.data(), Q_all, pose6, o, st, res);
    const double t3 = now_ms();
    if (M_out) *M_out = M;
    if (Q_out) *Q_out = Q_all;
    if (ms_out) { ms_out[0] = t1 - t0; ms_out[1] = t2 - t1; ms_out[2] = t3 - t2; ms_out[3] = t3 - t0; }
    return rc;
}

// ---- upstream LIO-SAM corner + surf superset (not in the reference; see the corner section above) -------
void orc_eigen3(const float* A9, float* w3, float* V9) { eigen3_jacobi(A9, w3, V9); }

int orc_line_residual(const float* pts15, const float* sel3, orc_point* coeff) {
    float P[5][3];
    for (int i = 0; i < 5; ++i) for (int j = 0; j < 3; ++j) P[i][j] = pts15[i * 3 + j];
    orc_point sel{sel3[0], sel3[1], sel3[2], 0.f};
    coeff->x = coeff->y = coeff->z = coeff->w = 0.f;
    return line_residual(P, sel, coeff) ? 1 : 0;
}

// One pass of cornerOptimization at a fixed pose; same output convention as orc_surf_pass.
int orc_corner_pass(const orc_point* map, int M, const orc_point* scan, int Q_all, int stride,
                    const float* pose6, int knn, int threads, orc_point* ori_out, orc_point* coef_out, int* sel_index) {
    float T[12]; pose_to_T(pose6, T);
    KdTree tree; if (knn == 1) tree.build(map, M);
    std::vector<uint8_t> flag((size_t)Q_all, 0);
    std::vector<orc_point> coef((size_t)Q_all);
    const int nq = (Q_all + stride - 1) / stride;
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(static)
    for (int k = 0; k < nq; ++k) {
        const int i = k * stride;
        const orc_point sel = apply_T(T, scan[i]);
        int idx[5]; float d2[5];
        if (knn == 1) tree.knn5(sel, idx, d2); else knn5_brute_one(map, M, sel, idx, d2);
        orc_point c;
        if (corner_residual(map, idx, d2, sel, &c)) { coef[i] = c; flag[i] = 1; }
    }
    int n = 0;
    for (int i = 0; i < Q_all; ++i) if (flag[i]) { ori_out[n] = scan[i]; coef_out[n] = coef[i]; if (sel_index) sel_index[n] = i; ++n; }
    return n;
}

// Upstream scan2MapOptimization: gate cornerDS > min_corner && surfDS > min_points; per iteration
// cornerOptimization, surfOptimization, combineOptimizationCoeffs (corners first), LMOptimization.
// o->stride applies to the surf loop (upstream 1), stride_corner to the corner loop (upstream 1).
int orc_scan2map_features(const orc_point* map_s, int Ms, const orc_point* scan_s, int Qs,
                          const orc_point* map_c, int Mc, const orc_point* scan_c, int Qc,
                          int min_corner, int stride_corner, float* pose6,
                          const orc_s2m_opts* o, orc_lm_state* st, orc_s2m_result* res) {
    std::memset(res, 0, sizeof(*res));
    if (!(Qc > min_corner && Qs > o->min_points)) { res->skipped = 1; return 1; }
    KdTree tree_s, tree_c;
    if (o->knn == 1) { tree_c.build(map_c, Mc); tree_s.build(map_s, Ms); }
    std::vector<uint8_t> flag_s((size_t)Qs, 0), flag_c((size_t)Qc, 0);
    std::vector<orc_point> coef_s((size_t)Qs), coef_c((size_t)Qc), ori, coef;
    const int ss = o->stride > 0 ? o->stride : 1, sc = stride_corner > 0 ? stride_corner : 1;
    const int nqs = (Qs + ss - 1) / ss, nqc = (Qc + sc - 1) / sc;
    const int threads = o->threads > 0 ? o->threads : 1;
    for (int iter = 0; iter < o->max_iters; ++iter) {
        ori.clear(); coef.clear();
        float T[12]; pose_to_T(pose6, T);
#pragma omp parallel for num_threads(threads) schedule(static)
        for (int k = 0; k < nqc; ++k) {
            const int i = k * sc;
            const orc_point sel = apply_T(T, scan_c[i]);
            int idx[5]; float d2[5];
            if (o->knn == 1) tree_c.knn5(sel, idx, d2); else knn5_brute_one(map_c, Mc, sel, idx, d2);
            orc_point c;
            if (corner_residual(map_c, idx, d2, sel, &c)) { coef_c[i] = c; flag_c[i] = 1; }
        }
#pragma omp parallel for num_threads(threads) schedule(static)
        for (int k = 0; k < nqs; ++k) {
            const int i = k * ss;
            const orc_point sel = apply_T(T, scan_s[i]);
            int idx[5]; float d2[5];
            if (o->knn == 1) tree_s.knn5(sel, idx, d2); else knn5_brute_one(map_s, Ms, sel, idx, d2);
            orc_point c;
            if (surf_residual(map_s, idx, d2, scan_s[i], sel, &c)) { coef_s[i] = c; flag_s[i] = 1; }
        }
        for (int i = 0; i < Qc; ++i) if (flag_c[i]) { ori.push_back(scan_c[i]); coef.push_back(coef_c[i]); flag_c[i] = 0; }
        for (int i = 0; i < Qs; ++i) if (flag_s[i]) { ori.push_back(scan_s[i]); coef.push_back(coef_s[i]); flag_s[i] = 0; }
        res->iters = iter + 1;
        res->n_sel = (int)ori.size();
        const int r = lm_step(ori.data(), coef.data(), (int)ori.size(), iter, o->min_sel, pose6, st, nullptr, nullptr);
        if (r == 1) { res->converged = 1; break; }
        if (r == -1) { res->too_few = 1; break; }
    }
    res->is_degenerate = st->is_degenerate;
    return 0;
}

// AnchorOptimization::getICPFitnessScore (include/egoOptimization/anchorOptimize.hpp:465-481) on the two clouds the
// caller built with transformPointCloud (:425, :439): pcl CorrespondenceEstimation::determineCorrespondences with the
// default (unbounded) max distance = exact 1-NN of every cloud-1 point in cloud 2, Correspondence::distance = SQUARED
// distance; `fitness_score += corr.distance` sequentially in float, divided by the number of correspondences.
float orc_icp_fitness(const orc_point* c1, int n1, const float* pose6_1, const orc_point* c2, int n2, const float* pose6_2,
                      int knn, int threads) {
    std::vector<orc_point> a((size_t)n1), b((size_t)n2);
    float T1[12], T2[12];
    pose_to_T(pose6_1, T1); pose_to_T(pose6_2, T2);
    for (int i = 0; i < n1; ++i) a[i] = apply_T(T1, c1[i]);
    for (int i = 0; i < n2; ++i) b[i] = apply_T(T2, c2[i]);
    KdTree tree; if (knn == 1) tree.build(b.data(), n2);
    std::vector<float> d((size_t)n1);
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(static)
    for (int i = 0; i < n1; ++i) {
        int idx[5]; float d2[5];
        if (knn == 1) tree.knn5(a[i], idx, d2); else knn5_brute_one(b.data(), n2, a[i], idx, d2);
        d[i] = d2[0];
    }
    float s = 0.f;
    for (int i = 0; i < n1; ++i) s += d[i];
    return s / (float)n1;
}

}  // extern "C"
