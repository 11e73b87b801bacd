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
// cell_start[ncells + 1] exclusive offsets.  The three x-adjacent cells of a (y, z) row are one
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

// The geometry is derived ON THE DEVICE from the local map's bounding box (k_grid_params), so building the
// search structure needs no host round trip; consumers read it through this pointer.
struct GridView {
    const int* __restrict__ cell_start;
    const float4* __restrict__ cpts;
    const GridParams* __restrict__ p;
};

LILOC_DEV int cell_coord(float v, float inv_cell, int g0) { return (int)floorf(v * inv_cell) - g0; }

__global__ void __launch_bounds__(256) k_zero_int(int* __restrict__ a, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) a[i] = 0;
}

// Search-grid geometry from the bounding box of the cloud the map was down-sampled from (a superset of the
// centroids' box): cell edge 2^k >= 1 m, one cell of padding, at most max_cells cells.  One thread; also
// zeroes nothing -- the counters are cleared by k_grid_zero once ncells is known.
__global__ void k_grid_params(const float* __restrict__ mn, const float* __restrict__ mx, const int* __restrict__ M,
                              int max_cells, GridParams* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float lo[3], hi[3];
    const bool empty = *M <= 0;
    for (int a = 0; a < 3; ++a) { lo[a] = empty ? 0.f : mn[a]; hi[a] = empty ? 0.f : mx[a]; }
    GridParams g;
    g.ncells = -1;
    for (int k = 0; k < 24; ++k) {
        const float inv = 1.0f / (float)(1 << k);
        long long cells = 1;
        for (int a = 0; a < 3; ++a) {
            g.g0[a] = (int)floorf(lo[a] * inv) - 1;
            const int g1 = (int)floorf(hi[a] * inv) + 1;
            g.dim[a] = g1 - g.g0[a] + 1;
            cells *= g.dim[a];
        }
        g.inv_cell = inv;
        if (cells <= (long long)max_cells) { g.ncells = (int)cells; break; }
    }
    if (g.ncells < 0) { g.ncells = 1; g.dim[0] = g.dim[1] = g.dim[2] = 1; g.inv_cell = 0.f; g.g0[0] = g.g0[1] = g.g0[2] = 0; }   // unreachable for finite input
    *out = g;
}

__global__ void __launch_bounds__(256) k_grid_zero(const GridParams* __restrict__ gp, int* __restrict__ counts) {
    const int n = gp->ncells;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) counts[i] = 0;
}

__global__ void __launch_bounds__(256)
k_grid_count(const float4* __restrict__ map, const int* __restrict__ M_dev, const GridParams* __restrict__ gpp, int* __restrict__ counts) {
    const GridParams gp = *gpp;
    const int M = *M_dev;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
        const float4 p = __ldg(&map[i]);
        const int cx = cell_coord(p.x, gp.inv_cell, gp.g0[0]);
        const int cy = cell_coord(p.y, gp.inv_cell, gp.g0[1]);
        const int cz = cell_coord(p.z, gp.inv_cell, gp.g0[2]);
        atomicAdd(&counts[(cz * gp.dim[1] + cy) * gp.dim[0] + cx], 1);
    }
}

// cursor[] starts as a copy of cell_start[]; order inside a cell is arbitrary, which is harmless
// because the query uses the (d2, index) total order.
__global__ void __launch_bounds__(256)
k_grid_scatter(const float4* __restrict__ map, const int* __restrict__ M_dev, const GridParams* __restrict__ gpp,
               int* __restrict__ cursor, float4* __restrict__ cpts) {
    const GridParams gp = *gpp;
    const int M = *M_dev;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
        const float4 p = __ldg(&map[i]);
        const int cx = cell_coord(p.x, gp.inv_cell, gp.g0[0]);
        const int cy = cell_coord(p.y, gp.inv_cell, gp.g0[1]);
        const int cz = cell_coord(p.z, gp.inv_cell, gp.g0[2]);
        const int pos = atomicAdd(&cursor[(cz * gp.dim[1] + cy) * gp.dim[0] + cx], 1);
        cpts[pos] = make_float4(p.x, p.y, p.z, __int_as_float(i));
    }
}

// ---- multi-tile exclusive scan of the cell counts (-> cell_start); n = gp->ncells lives on the device,
// so the kernels run on a fixed launch grid and loop over tiles -------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanPerThread = 8;
constexpr int kScanTile = kScanThreads * kScanPerThread;

__global__ void __launch_bounds__(kScanThreads)
k_scan_reduce(const int* __restrict__ in, const GridParams* __restrict__ gp, int* __restrict__ tile_sums) {
    __shared__ int s_w[kScanThreads / 32];
    const int n = gp->ncells;
    const int n_tiles = (n + kScanTile - 1) / kScanTile;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int base = tile * kScanTile;
        int c = 0;
#pragma unroll
        for (int j = 0; j < kScanPerThread; ++j) {
            const int i = base + j * kScanThreads + threadIdx.x;
            if (i < n) c += in[i];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
#pragma unroll
            for (int w = 0; w < kScanThreads / 32; ++w) t += s_w[w];
            tile_sums[tile] = t;
        }
    }
}

__global__ void __launch_bounds__(1024)
k_scan_tiles_grid(int* __restrict__ tile_counts, const GridParams* __restrict__ gp, int* __restrict__ total_out) {
    scan_tiles_body(tile_counts, (gp->ncells + kScanTile - 1) / kScanTile, total_out);
}

// out[i] = exclusive prefix, out[n] = grand total; `cursor` (which may alias `in`) receives the same values
// (the scatter cursor).
__global__ void __launch_bounds__(kScanThreads)
k_scan_down(const int* in, const GridParams* __restrict__ gp, const int* __restrict__ tile_offsets,
            int* __restrict__ out, int* cursor) {
    __shared__ int s_w[kScanThreads / 32];
    const int n = gp->ncells;
    const int n_tiles = (n + kScanTile - 1) / kScanTile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int base = tile * kScanTile + threadIdx.x * kScanPerThread;
        int v[kScanPerThread];
        int c = 0;
#pragma unroll
        for (int j = 0; j < kScanPerThread; ++j) { v[j] = (base + j < n) ? in[base + j] : 0; c += v[j]; }
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        __syncthreads();
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        int run = tile_offsets[tile] + inc - c;
#pragma unroll
        for (int w = 0; w < kScanThreads / 32; ++w) if (w < warp) run += s_w[w];
#pragma unroll
        for (int j = 0; j < kScanPerThread; ++j) {
            if (base + j < n) { out[base + j] = run; cursor[base + j] = run; }
            run += v[j];
            if (base + j == n - 1) { out[n] = run; }
        }
    }
}

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
// scans every L-th point of the 9 cell rows with its own sorted top-5 in registers; the L lists are then
// merged with five rounds of (d2, index) arg-min over xor-shuffles.  On return every lane of the group
// holds the same result.  Far fewer instructions than one warp per query: the map is a surface, so a
// 27-cell neighbourhood holds only ~100-250 candidates.
template <int L>
LILOC_DEV Knn5 knn5_group(const GridView& g, const GridParams& gp, const float4 q, const int sub, const bool active) {
    const int cx = cell_coord(q.x, gp.inv_cell, gp.g0[0]);
    const int cy = cell_coord(q.y, gp.inv_cell, gp.g0[1]);
    const int cz = cell_coord(q.z, gp.inv_cell, gp.g0[2]);
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, gp.dim[0] - 1);
    int rs[9], re[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        const int yy = cy + (r % 3) - 1, zz = cz + (r / 3) - 1;
        rs[r] = 0; re[r] = 0;
        if (active && yy >= 0 && yy < gp.dim[1] && zz >= 0 && zz < gp.dim[2] && x0 <= x1) {
            const int base = (zz * gp.dim[1] + yy) * gp.dim[0];
            rs[r] = __ldg(&g.cell_start[base + x0]);
            re[r] = __ldg(&g.cell_start[base + x1 + 1]);
        }
    }
    Knn5 t;
#pragma unroll
    for (int k = 0; k < 5; ++k) { t.d[k] = 3.4e38f; t.i[k] = 0x7fffffff; }
    // Row-parallel scan: every pass issues one load per row (up to 9 independent 16-byte loads in
    // flight per lane) before any distance is evaluated, so a pass costs one memory latency, not nine.
    for (int k = sub;; k += L) {
        float4 p[9];
        bool any = false;
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            const int j = rs[r] + k;
            if (j < re[r]) { p[r] = __ldg(&g.cpts[j]); any = true; }
        }
        if (!any) break;
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            if (rs[r] + k < re[r]) {
                const float d = dist2(q, p[r]);
                if (d < 1.0f) top5_insert(t, d, __float_as_int(p[r].w));
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
k_knn5(const float4* __restrict__ queries, int nq, GridView g, int* __restrict__ idx5, float* __restrict__ d2, int* __restrict__ found) {
    const int sub = threadIdx.x % kKnnLanes;
    const GridParams gp = *g.p;
    const int groups_per_grid = (gridDim.x * blockDim.x) / kKnnLanes;
    const int nq_pad = ((nq + groups_per_grid - 1) / groups_per_grid) * groups_per_grid;   // keep warps converged
    for (int qi = (blockIdx.x * blockDim.x + threadIdx.x) / kKnnLanes; qi < nq_pad; qi += groups_per_grid) {
        const bool active = qi < nq;
        const float4 q = active ? __ldg(&queries[qi]) : make_float4(0.f, 0.f, 0.f, 0.f);
        const Knn5 r = knn5_group<kKnnLanes>(g, gp, q, sub, active);
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

}  // namespace liloc
