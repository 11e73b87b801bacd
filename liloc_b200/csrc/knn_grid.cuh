// liloc_b200/csrc/knn_grid.cuh -- uniform search grid over the down-sampled local map and the exact
// warp-cooperative 5-NN query that replaces pcl::KdTreeFLANN.
//
// Replaces (reference): kdtreeSurfFromMap->setInputCloud(laserCloudSurfFromMapDS)  src/mapOptmization.cpp:1188
//                       kdtreeSurfFromMap->nearestKSearch(pointSel, 5, ...)         src/mapOptmization.cpp:992
//
// Exactness argument.  The reference keeps a correspondence only if d2[4] < 1.0 (:1002), i.e. only
// if all five neighbours lie strictly within 1 m.  Cells are cubes of edge c = 2^k >= 1 m and cell
// coordinates are floor(x * (1/c)) with 1/c a power of two, so the coordinate is computed without
// rounding; two points closer than 1 m therefore sit in cells that differ by at most 1 per axis, and
// scanning the 27-cell neighbourhood with the (d2, index) total order returns exactly the reference's
// neighbour set whenever that set can pass the gate.  Otherwise fewer than 5 candidates below 1.0
// are found and the query is rejected -- the same decision the reference takes.
//
// Layout in HBM: map points sorted by cell (x fastest) as float4 {x, y, z, bitcast(original index)};
// cell_start[ncells + 1] exclusive offsets (built by the G phases of pipeline.cuh).  The three x-adjacent cells of a (y, z) row are one
// contiguous range, so a query reads 9 ranges with coalesced 16-byte loads, one lane per point.
#pragma once
#include "common.cuh"
#include "voxelgrid.cuh"

namespace liloc {

struct GridParams {
    float inv_cell;        // 1 / cell edge (power of two)
    int g0[3];             // cell coordinate of the grid origin
    int dim[3];
    int ncells;
};

// The geometry is derived ON THE DEVICE from the local map's bounding box (pipeline.cuh, grid_params_compute), so building the
// search structure needs no host round trip; consumers read it through this pointer.
struct GridView {
    const int* __restrict__ cell_start;
    const float4* __restrict__ cpts;
    const GridParams* __restrict__ p;
};

LILOC_DEV int cell_coord(float v, float inv_cell, int g0) { return (int)floorf(v * inv_cell) - g0; }

constexpr int kScanThreads = 256;
constexpr int kScanPerThread = 16;   // four int4 loads in flight per thread
constexpr int kScanTile = kScanThreads * kScanPerThread;     // cells per block and step of the cell-count scan (pipeline.cuh)

// ---- exact 5-NN, one warp per query ---------------------------------------------------------------
struct Knn5 { float d[5]; int i[5]; };     // ascending (d2, index); i == INT_MAX marks "none"

LILOC_DEV void top5_insert(Knn5& t, float d, int i) {
    if (!nn_less(d, i, t.d[4], t.i[4])) return;
    t.d[4] = d; t.i[4] = i;
#pragma unroll
    for (int k = 4; k > 0; --k) {
        if (nn_less(t.d[k], t.i[k], t.d[k - 1], t.i[k - 1])) {
            const float td = t.d[k]; t.d[k] = t.d[k - 1]; t.d[k - 1] = td;
            const int ti = t.i[k]; t.i[k] = t.i[k - 1]; t.i[k - 1] = ti;
        }
    }
}

// Exact 5-NN by a group of L lanes (L = 1, 2, 4, 8 ...; groups are aligned inside a warp, and ALL 32 lanes of
// the warp must call this together -- lanes without a query pass active = false).  Lane `sub` of the group
// scans every L-th candidate of the 9 cell rows (taken as one sequence) with its own sorted top-5 in registers; the L lists are then
// merged with five rounds of (d2, index) arg-min over xor-shuffles.  On return every lane of the group
// holds the same result.  Far fewer instructions than one warp per query: the map is a surface, so a
// 27-cell neighbourhood holds only ~100-250 candidates.
// `bound`: an upper bound of the fifth neighbour's squared distance, when the caller has one (the registration kernel: the
// largest distance to the five neighbours of the previous iteration, which are five distinct map points); 3.4e38 otherwise.
// Cells that cannot hold a point with d2 <= bound and d2 < 1 are not read: of a cell the query sees its nearest face, and the
// lower bound lb = ((ex^2) + (ey^2)) + (ez^2) of the face distances is evaluated in the operation order of dist2, so -- float
// subtraction, multiplication and addition being monotone -- lb <= dist2(q, p) holds in float for every p of the cell.  A point
// with d2 > bound is not among the five smallest (d2, index) and one with d2 >= 1 is never inserted: the result is unchanged.
template <int L>
LILOC_DEV Knn5 knn5_group(const GridView& g, const GridParams& gp, const float4 q, const int sub, const bool active, const float bound = 3.4e38f) {
    const int cx = cell_coord(q.x, gp.inv_cell, gp.g0[0]);
    const int cy = cell_coord(q.y, gp.inv_cell, gp.g0[1]);
    const int cz = cell_coord(q.z, gp.inv_cell, gp.g0[2]);
    // squared distance to the nearest face of the cells at offset -1 / 0 / +1 (cell edges are powers of two: exact coordinates)
    const float edge = __int_as_float(0x7f000000 - __float_as_int(gp.inv_cell));      // 1 / inv_cell for a power of two, without a division
    float sx[3], sy[3], sz[3];
    {
        const float ox = (float)(cx + gp.g0[0]) * edge, oy = (float)(cy + gp.g0[1]) * edge, oz = (float)(cz + gp.g0[2]) * edge;
        const float xl = q.x - ox, xr = (ox + edge) - q.x, yl = q.y - oy, yr = (oy + edge) - q.y, zl = q.z - oz, zr = (oz + edge) - q.z;
        sx[0] = xl * xl; sx[1] = 0.f; sx[2] = xr * xr;
        sy[0] = yl * yl; sy[1] = 0.f; sy[2] = yr * yr;
        sz[0] = zl * zl; sz[1] = 0.f; sz[2] = zr * zr;
    }
    int rs[9], re[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        const int yy = cy + (r % 3) - 1, zz = cz + (r / 3) - 1;
        rs[r] = 0; re[r] = 0;
        const float lb_l = (sx[0] + sy[r % 3]) + sz[r / 3], lb_c = (sx[1] + sy[r % 3]) + sz[r / 3], lb_r = (sx[2] + sy[r % 3]) + sz[r / 3];
        const bool need_c = lb_c <= bound && lb_c < 1.0f;
        const int xa = (lb_l <= bound && lb_l < 1.0f) ? cx - 1 : cx, xb = (lb_r <= bound && lb_r < 1.0f) ? cx + 1 : cx;
        const int x0 = max(xa, 0), x1 = min(xb, gp.dim[0] - 1);
        if (active && need_c && yy >= 0 && yy < gp.dim[1] && zz >= 0 && zz < gp.dim[2] && x0 <= x1) {
            const int base = (zz * gp.dim[1] + yy) * gp.dim[0];
            rs[r] = __ldg(&g.cell_start[base + x0]);
            re[r] = __ldg(&g.cell_start[base + x1 + 1]);
        }
    }
    // The nine rows as ONE candidate sequence: candidate k of the query lives at cpts[k + delta[r]] for the row r with
    // pre[r] <= k < pre[r + 1].  A lane takes candidates sub, sub + L, ... and issues nine loads per pass whatever row they
    // fall in, so a query with ~70 candidates is done in one or two memory latencies; scanning row by row took as many
    // passes as the longest row needs (3-4) with most loads of a pass idle.  The result does not depend on the order in
    // which candidates are seen: the top-5 is kept in the total order (d2, index).
    int pre[10], delta[9];
    pre[0] = 0;
#pragma unroll
    for (int r = 0; r < 9; ++r) { pre[r + 1] = pre[r] + (re[r] - rs[r]); delta[r] = rs[r] - pre[r]; }
    const int total = pre[9];
    Knn5 t;
#pragma unroll
    for (int k = 0; k < 5; ++k) { t.d[k] = 3.4e38f; t.i[k] = 0x7fffffff; }
    for (int k0 = sub; k0 < total; k0 += 9 * L) {
        float4 p[9];
#pragma unroll
        for (int u = 0; u < 9; ++u) {
            const int k = k0 + u * L;
            if (k < total) {
                int off = delta[0];
#pragma unroll
                for (int r = 1; r < 9; ++r) if (k >= pre[r]) off = delta[r];
                p[u] = __ldg(&g.cpts[k + off]);
            }
        }
#pragma unroll
        for (int u = 0; u < 9; ++u) {
            if (k0 + u * L < total) {
                const float d = dist2(q, p[u]);
                if (d < 1.0f) top5_insert(t, d, __float_as_int(p[u].w));
            }
        }
    }
    __syncwarp();
    Knn5 out;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        float md = t.d[0]; int mi = t.i[0];
#pragma unroll
        for (int o = L / 2; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, md, o);
            const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
            if (nn_less(od, oi, md, mi)) { md = od; mi = oi; }
        }
        out.d[k] = md; out.i[k] = mi;
        if (mi != 0x7fffffff && t.i[0] == mi) {          // the winning lane pops its head
#pragma unroll
            for (int j = 0; j < 4; ++j) { t.d[j] = t.d[j + 1]; t.i[j] = t.i[j + 1]; }
            t.d[4] = 3.4e38f; t.i[4] = 0x7fffffff;
        }
    }
    return out;
}

constexpr int kKnnLanes = 8;     // lanes per query in the registration kernels

// Standalone kNN kernel behind liloc_knn5 (kernel-level parity tests): the same group search the
// registration kernel uses.
__global__ void __launch_bounds__(256)
k_knn5(const float4* __restrict__ queries, int nq, GridView g, int* __restrict__ idx5, float* __restrict__ d2, int* __restrict__ found,
       const float* __restrict__ bounds = nullptr) {
    const int sub = threadIdx.x % kKnnLanes;
    const GridParams gp = *g.p;
    const int groups_per_grid = (gridDim.x * blockDim.x) / kKnnLanes;
    const int nq_pad = ((nq + groups_per_grid - 1) / groups_per_grid) * groups_per_grid;   // keep warps converged
    for (int qi = (blockIdx.x * blockDim.x + threadIdx.x) / kKnnLanes; qi < nq_pad; qi += groups_per_grid) {
        const bool active = qi < nq;
        const float4 q = active ? __ldg(&queries[qi]) : make_float4(0.f, 0.f, 0.f, 0.f);
        const Knn5 r = knn5_group<kKnnLanes>(g, gp, q, sub, active, (bounds && active) ? __ldg(&bounds[qi]) : 3.4e38f);
        if (active && sub == 0) {
            int f = 0;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const bool ok = r.i[k] != 0x7fffffff;
                idx5[5 * qi + k] = ok ? r.i[k] : -1;
                d2[5 * qi + k] = ok ? r.d[k] : 3.4e38f;
                f += ok ? 1 : 0;
            }
            found[qi] = f;
        }
    }
}

// ---- exact UNBOUNDED 1-NN (pcl::registration::CorrespondenceEstimation, max distance = infinity) -------------------
// Replaces the kd-tree query of AnchorOptimization::getICPFitnessScore (include/egoOptimization/anchorOptimize.hpp:465-481).
// A group of kKnnLanes lanes owns one query and searches cubic shells of cells around the query's cell, ring by ring.
// After rings 0..r every unseen point lies in a cell at Chebyshev distance >= r + 1, i.e. farther than r cell edges from
// the query, so the search stops as soon as best <= (r * edge)^2 (or the shells have left the grid).  Distance as
// FLANN's L2_Simple: ((dx^2) + (dy^2)) + (dz^2) in float.
__global__ void __launch_bounds__(256)
k_nn1_dist2(const float4* __restrict__ queries, int nq, const float* __restrict__ T12, GridView g, float* __restrict__ d2_out) {
    const GridParams gp = *g.p;
    __shared__ float sT[12];
    if (threadIdx.x < 12) sT[threadIdx.x] = T12[threadIdx.x];
    __syncthreads();
    const int sub = threadIdx.x % kKnnLanes;
    const int groups_per_grid = (gridDim.x * blockDim.x) / kKnnLanes;
    const int nq_pad = ((nq + groups_per_grid - 1) / groups_per_grid) * groups_per_grid;     // keep warps converged
    const float edge = 1.0f / gp.inv_cell;
    for (int qi = (blockIdx.x * blockDim.x + threadIdx.x) / kKnnLanes; qi < nq_pad; qi += groups_per_grid) {
        const bool active = qi < nq;
        const float4 q = active ? apply_T(sT, __ldg(&queries[qi])) : make_float4(0.f, 0.f, 0.f, 0.f);
        const int cx = cell_coord(q.x, gp.inv_cell, gp.g0[0]);
        const int cy = cell_coord(q.y, gp.inv_cell, gp.g0[1]);
        const int cz = cell_coord(q.z, gp.inv_cell, gp.g0[2]);
        // the last ring that still touches the grid
        int rmax = max(max(max(cx, gp.dim[0] - 1 - cx), max(cy, gp.dim[1] - 1 - cy)), max(cz, gp.dim[2] - 1 - cz));
        if (!active) rmax = -1;
        float best = 3.4e38f;
        for (int r = 0;; ++r) {
            // group-uniform loop control: every lane of the warp takes part in the shuffles below
            const bool mine = active && r <= rmax && !(r >= 1 && best <= ((float)(r - 1) * edge) * ((float)(r - 1) * edge));
            if (__ballot_sync(0xffffffffu, mine) == 0u) break;
            if (mine) {
                for (int dz = -r; dz <= r; ++dz) {
                    const int zz = cz + dz;
                    if (zz < 0 || zz >= gp.dim[2]) continue;
                    for (int dy = -r; dy <= r; ++dy) {
                        const int yy = cy + dy;
                        if (yy < 0 || yy >= gp.dim[1]) continue;
                        const bool face = (dz == -r || dz == r || dy == -r || dy == r);
                        const int base = (zz * gp.dim[1] + yy) * gp.dim[0];
                        // on a face of the shell the whole x-range belongs to ring r; elsewhere only its two end cells
                        for (int part = 0; part < (face ? 1 : 2); ++part) {
                            int x0, x1;
                            if (face) { x0 = cx - r; x1 = cx + r; }
                            else { x0 = x1 = (part == 0 ? cx - r : cx + r); if (r == 0 && part == 1) continue; }
                            x0 = max(x0, 0); x1 = min(x1, gp.dim[0] - 1);
                            if (x0 > x1) continue;
                            const int js = __ldg(&g.cell_start[base + x0]), je = __ldg(&g.cell_start[base + x1 + 1]);
                            for (int j = js + sub; j < je; j += kKnnLanes) best = fminf(best, dist2(q, __ldg(&g.cpts[j])));
                        }
                    }
                }
            }
#pragma unroll
            for (int o = kKnnLanes / 2; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
        }
        if (active && sub == 0) d2_out[qi] = best;
    }
}

// fitness_score += corr.distance, sequentially in source order in float, then / n (anchorOptimize.hpp:473-478):
// one thread, eight loads in flight.
__global__ void k_seq_mean(const float* __restrict__ v, int n, float* __restrict__ out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    float s = 0.f;
    int i = 0;
    for (; i + 8 <= n; i += 8) {
        float w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) w[u] = v[i + u];
#pragma unroll
        for (int u = 0; u < 8; ++u) s += w[u];
    }
    for (; i < n; ++i) s += v[i];
    *out = s / (float)n;
}

}  // namespace liloc
