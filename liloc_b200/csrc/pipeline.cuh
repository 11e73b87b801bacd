// liloc_b200/csrc/pipeline.cuh -- the whole point-cloud side of a frame as ONE persistent cooperative kernel.
//
// Replaces (reference, per frame): extractCloud's transform + concatenation (src/mapOptmization.cpp:936-961,
// include/utility.h:388-407), the two pcl::VoxelGrid filters (:954-955, :972-973; semantics SURVEY.md A.1) and
// kdtreeSurfFromMap->setInputCloud (:1188).  Up to four independent "problems" (surf map, surf scan, corner map,
// corner scan) run through the same phases between grid barriers, so a frame costs one launch here plus one
// launch of the registration kernel (s2m.cuh) instead of ~60 small launches:
//
//   P0  transform + concatenate the selected keyframes (map problems) and the bounding box of every cloud
//   P1  VoxelGrid parameters; int32 voxel key per point; digit histogram of the first radix pass
//   Sx  stable LSD radix sort of (key, source index), 8 bits per pass, ceil(key_bits / 8) passes: per-block
//       digit offsets from the block histograms, warp-match ranking inside a tile, scatter; the histogram of the
//       NEXT pass is accumulated while scattering, so a pass is a single phase
//   R2  one centroid per voxel run (float sums in ascending source index = the canonical order), written in
//       ascending voxel index; the output offset of a tile is the number of run heads in all earlier tiles, obtained
//       by a decoupled look-back over per-tile words (no separate counting phase, no barrier); the centroid's
//       search-grid cell is counted in the same pass
//   G2, G3  exclusive scan of the cell counts (one pass, the same look-back), then the points sorted by cell (the kNN
//       search structure)
//
// Every size that only exists on the device (M, Q_all, search-grid geometry) stays there.  No library calls:
// the radix sort is part of this kernel (the first version of this path used cub::DeviceRadixSort).
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"
#include "voxelgrid.cuh"
#include "knn_grid.cuh"

namespace liloc {

namespace cg = cooperative_groups;

constexpr int kPipeThreads = 256;
constexpr int kSortBlocksMax = 296;                       // blocks that can take part in the sort (every block of a two-per-SM launch)
constexpr int kSortBlocksSmall = 148;                     // ... below kSortTwoLevelMin items: one block per SM
constexpr int kSortTwoLevelMin = 1200000;                 // items from which the sort runs on every block with two-level offsets
constexpr int kSortRounds = 4;
constexpr int kSortTile = kPipeThreads * kSortRounds;     // 1024 items
constexpr int kHistStride = kSortBlocksMax * 256;         // ints per histogram buffer (three rotate)
constexpr int kHistInts = 3 * kHistStride + 256;          // + the digit totals of the pass being scattered
constexpr int kMaxProbs = 4;

struct VgProb {
    const float4* pts;         // cloud to down-sample (== raw when assembling)
    int n;                     // points (known on the host), or their upper bound when n_dev is set
    const int* n_dev;          // optional: the count lives on the device (written by an earlier launch: the de-skewed scan)
    float leaf;
    // optional: assemble pts from keyframes first
    const float4* arena;
    const KfDesc* descs;
    int inline1;               // 1 + index of this problem's first descriptor in PipeArgs::inline_descs (0: read `descs`)
    int K;
    float4* raw;
    // scratch
    unsigned* keys[2];
    unsigned* vals[2];
    int* hist;                 // kHistInts: 3 rotating block-histogram buffers + 256 digit totals
    unsigned long long* tile_state;   // ceil(n / kRunTile) + 1 look-back words (epoch << 32 | run heads of the tile)
    unsigned* enc;             // 6 ordered-uint min / max; re-armed by the kernel for the next launch
    VgParams* vp;              // out (diagnostics; the bounding box)
    // output
    float4* out;
    int* count;
    // optional: search grid over `out`
    int build_grid;
    int max_cells;
    GridParams* gp;
    int* cell_counts;          // all zero between launches: counted up in R2, counted down to zero again in G3
    int* cell_start;
    unsigned long long* grid_state;   // look-back words of the cell-count scan
    float4* cpts;
    int sort_only;             // stop after the sort (NDT target build): sorted pairs are in keys[1] / vals[1]
    int grid_only;             // no down-sampling: the search grid is built over the (assembled) cloud itself; out == pts
};

constexpr int kPipeMarks = 16;
constexpr int kInlineDescs = 40;      // keyframe descriptors that travel with the launch (40 x 64 B of kernel parameters)
struct PipeArgs {
    VgProb prob[kMaxProbs];
    int n_prob;
    // Small local maps (the usual case: 15-30 keyframes) carry their descriptors as kernel parameters: a host-to-device
    // copy in front of the launch would put a copy-engine hop of several microseconds at the start of every frame.
    int n_inline;
    unsigned epoch;              // launch counter: tags the look-back words, so they never need clearing
    KfDesc inline_descs[kInlineDescs];
    unsigned long long* marks;   // optional: %globaltimer (ns) at every phase boundary, written by block 0 (latency breakdown)
    int* sm_scratch;             // kSmScratch zeroed ints per launch in flight: per-SM arrival counters, the sort-block counter, the abort word
    int* error_flag;             // mapped host word: 1 when a look-back wait gave up, 2 when an input point was not finite
};
constexpr int kSmAbort = 513;         // index of the abort word in sm_scratch
constexpr int kSmScratch = 514;

LILOC_DEV int prob_n(const VgProb& q) { return q.n_dev ? min(__ldg(q.n_dev), q.n) : q.n; }

LILOC_DEV unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// VoxelGrid set-up from the ordered-uint bounding box.
LILOC_DEV VgParams vg_params_compute(const unsigned* __restrict__ enc, float leaf, int n) {
    VgParams p;
    p.inv_leaf = 1.0f / leaf;
    for (int a = 0; a < 3; ++a) { p.mn[a] = ord2f(enc[a]); p.mx[a] = ord2f(enc[3 + a]); }
    if (n <= 0) { for (int a = 0; a < 3; ++a) { p.mn[a] = 0.f; p.mx[a] = 0.f; } }
    const long long dx = (long long)((p.mx[0] - p.mn[0]) * p.inv_leaf) + 1;
    const long long dy = (long long)((p.mx[1] - p.mn[1]) * p.inv_leaf) + 1;
    const long long dz = (long long)((p.mx[2] - p.mn[2]) * p.inv_leaf) + 1;
    p.overflow = (dx * dy * dz > 2147483647LL) ? 1 : 0;
    int divb[3];
    for (int a = 0; a < 3; ++a) {
        p.minb[a] = (int)floorf(p.mn[a] * p.inv_leaf);
        const int maxb = (int)floorf(p.mx[a] * p.inv_leaf);
        divb[a] = maxb - p.minb[a] + 1;
    }
    p.mul[0] = 1; p.mul[1] = divb[0]; p.mul[2] = divb[0] * divb[1];
    const long long max_key = p.overflow ? (long long)(n > 0 ? n - 1 : 0) : (long long)divb[0] * divb[1] * divb[2] - 1;
    int bits = 1;
    while (bits < 32 && (max_key >> bits) != 0) ++bits;
    p.key_bits = bits;
    return p;
}

// Search-grid geometry from a bounding box: cell edge 2^k >= 1 m, one cell of padding, at most max_cells cells.
LILOC_DEV GridParams grid_params_compute(const float* mn, const float* mx, bool empty, int max_cells) {
    float lo[3], hi[3];
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
    if (g.ncells < 0) { g.ncells = 1; g.dim[0] = g.dim[1] = g.dim[2] = 1; g.inv_cell = 0.f; g.g0[0] = g.g0[1] = g.g0[2] = 0; }
    return g;
}

LILOC_DEV int voxel_key(const float4 p, const VgParams& vp, int i) {
    if (vp.overflow) return i;
    const int i0 = (int)(floorf(p.x * vp.inv_leaf) - (float)vp.minb[0]);
    const int i1 = (int)(floorf(p.y * vp.inv_leaf) - (float)vp.minb[1]);
    const int i2 = (int)(floorf(p.z * vp.inv_leaf) - (float)vp.minb[2]);
    return i0 + i1 * vp.mul[1] + i2 * vp.mul[2];
}

LILOC_DEV int block_sum_256(int v) {                        // all 256 threads; returns the block total to every thread
    __shared__ int s_w[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_w[w];
    return t;
}

// Decoupled look-back over per-tile totals (phases R2 and G2).  A tile publishes (epoch, total) as one 64-bit word, then
// sums the totals of every earlier tile, spinning on the words that are not there yet.  Tiles are handed out round-robin
// in ascending order and a block publishes before it waits, so with all blocks resident (cooperative launch) every
// wait ends: the earliest unpublished tile always belongs to a block that is not waiting on anything later.  This
// replaces a separate counting phase and its grid barrier (~4.5 us per launch at 220 k points).  The spin is bounded:
// a bug shows up as a wrong result in the parity tests, not as a hung GPU.
LILOC_DEV void pipe_publish(unsigned long long* state, int tile, int total, unsigned epoch) {
    if (threadIdx.x == 0) atomicExch(&state[tile], ((unsigned long long)epoch << 32) | (unsigned)total);
}
// sum of the published totals of tiles [lo, tile)
LILOC_DEV int pipe_lookback(unsigned long long* state, int tile, int total, unsigned epoch, int* error_flag, bool publish = true, int lo = 0) {
    if (publish) pipe_publish(state, tile, total, epoch);
    int part = 0;
    for (int u = lo + threadIdx.x; u < tile; u += kPipeThreads) {
        unsigned long long w;
        int spins = 0;
        do { w = *((volatile unsigned long long*)&state[u]); } while ((unsigned)(w >> 32) != epoch && ++spins < (1 << 22));
        if ((unsigned)(w >> 32) != epoch && error_flag) *(volatile int*)error_flag = 1;      // gave up: the result is not valid
        part += (int)(unsigned)(w & 0xffffffffull);
    }
    return block_sum_256(part);
}

// Two ways to get a block's digit offsets, chosen per launch from the problem sizes (uniform over the grid):
//  * small sorts (< kSortTwoLevelMin items): at most one sorting block per SM, and every sorting block reads every busy
//    block's histogram row (O(nb^2) L2 reads: 3.5 us per pass at 107 blocks) -- no extra phase, no extra barrier.  With two
//    blocks per SM resident, WHICH blocks sort is decided at run time: the first block to arrive on an SM takes a sort index.
//  * large sorts: every block sorts (half the tiles per block and pass); the row reads would double with the blocks, so the
//    offsets are computed ONCE per pass by a phase of their own (pipe_sort_offsets: one warp per digit scans that digit's
//    column over the blocks, in place) for one more grid barrier per pass.  Measured (profiles/frame_latency_r02.md): the
//    2.3 M-point Mid-360 map sorts in 148 us instead of 182 us; at 219 k / 836 k points the extra barriers cost more than
//    the shorter scatter saves (49 vs 39 us, 78 vs 74 us), hence the threshold.
struct SortGeom { int C, nb; };                            // items per sorting block (whole tiles), busy sort blocks
LILOC_DEV SortGeom sort_geom(int n, int nsort) {
    SortGeom g;
    const int gs = max(1, min(nsort, kSortBlocksMax));
    // whole 1024-item tiles per block: finer chunks put more blocks to work but every block reads every busy
    // block's histogram row (O(nb^2) L2 traffic), which measured slower (profiles/frame_latency_r01.md)
    g.C = (((n + gs - 1) / gs + kSortTile - 1) / kSortTile) * kSortTile;
    if (g.C < kSortTile) g.C = kSortTile;
    g.nb = (n + g.C - 1) / g.C;
    return g;
}

constexpr int kSortUnits = kSortRounds * (kPipeThreads / 32);      // (round, warp) pairs of a tile, in item order
struct PipeSmemSort {
    int offs[256];                                 // next free slot of every digit for this block
    unsigned short cnt[kSortUnits][256];           // items of digit d in unit u (zero between tiles)
    unsigned short pre[kSortUnits][256];           // exclusive prefix of cnt over the units
};
// a tile of sorted pairs plus a halo of the following kRunHalo pairs: a voxel run that starts in the tile continues there
constexpr int kRunHalo = 512;
struct PipeSmemRun { unsigned key[kRunTile + kRunHalo]; float4 pt[kRunTile + kRunHalo]; unsigned prev; };
union PipeSmem { PipeSmemSort sort; PipeSmemRun run; int hist[256]; };

// ---- P0 -------------------------------------------------------------------------------------------------------
LILOC_DEV void pipe_bbox(const PipeArgs& a, const VgProb& q) {
    const KfDesc* __restrict__ descs = q.inline1 ? &a.inline_descs[q.inline1 - 1] : q.descs;
    float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    // Input coordinates must be finite (include/liloc_gpu.h; the reference shuts down on a non-dense cloud,
    // src/imageProjection.cpp:395-399).  A NaN / Inf would leave the bounding box yet still get a voxel key and a search
    // cell -- out-of-range indices further down.  It is detected here, where every point passes once, and aborts the launch.
    bool bad = false;
    if (q.K > 0) {
        __shared__ int s_k0;
        for (int tile = blockIdx.x; tile * kAsmTile < prob_n(q); tile += gridDim.x) {
            const int base = tile * kAsmTile;
            __syncthreads();
            if (threadIdx.x == 0) {               // last descriptor with dst_off <= base
                int lo = 0, hi = q.K - 1;
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (descs[mid].dst_off <= base) lo = mid; else hi = mid - 1;
                }
                s_k0 = lo;
            }
            __syncthreads();
            // addresses first, then all loads of the thread in flight, then the transforms
            int k = s_k0;
            const KfDesc* d[kAsmPerThread];
            float4 p[kAsmPerThread];
#pragma unroll
            for (int j = 0; j < kAsmPerThread; ++j) {
                const int i = base + j * kPipeThreads + threadIdx.x;
                d[j] = nullptr;
                if (i < prob_n(q)) {
                    while (k + 1 < q.K && descs[k + 1].dst_off <= i) ++k;
                    d[j] = &descs[k];
                }
            }
#pragma unroll
            for (int j = 0; j < kAsmPerThread; ++j) {
                const int i = base + j * kPipeThreads + threadIdx.x;
                if (d[j]) p[j] = __ldg(&q.arena[d[j]->src_off + (i - d[j]->dst_off)]);
            }
#pragma unroll
            for (int j = 0; j < kAsmPerThread; ++j) {
                const int i = base + j * kPipeThreads + threadIdx.x;
                if (d[j]) {
                    const float4 o = apply_T(d[j]->T, p[j]);
                    q.raw[i] = o;
                    bad |= !(isfinite(o.x) && isfinite(o.y) && isfinite(o.z));
                    mn[0] = fminf(mn[0], o.x); mx[0] = fmaxf(mx[0], o.x);
                    mn[1] = fminf(mn[1], o.y); mx[1] = fmaxf(mx[1], o.y);
                    mn[2] = fminf(mn[2], o.z); mx[2] = fmaxf(mx[2], o.z);
                }
            }
        }
    } else {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < prob_n(q); i += gridDim.x * blockDim.x) {
            const float4 p = __ldg(&q.pts[i]);
            bad |= !(isfinite(p.x) && isfinite(p.y) && isfinite(p.z));
            mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
            mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
            mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
        }
    }
    if (bad) atomicOr(&a.sm_scratch[kSmAbort], 1);
    block_minmax_commit(mn, mx, q.enc);
    // histogram buffers 1 and 2 must be zero when the first scatter pass accumulates into them
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * kHistStride; i += gridDim.x * blockDim.x) q.hist[kHistStride + i] = 0;
}

// ---- P1: keys + histogram of the first pass ----------------------------------------------------------------------
LILOC_DEV void pipe_keys(const VgProb& q, const VgParams& vp, int npass, PipeSmem& sm, const int sidx, const int nsort) {
    const SortGeom g = sort_geom(prob_n(q), nsort);
    if (sidx < 0 || sidx >= g.nb) return;
    unsigned* keys = q.keys[(npass + 1) & 1];                  // so that the last pass lands in buffer 1
    sm.hist[threadIdx.x] = 0;
    __syncthreads();
    const int lo = sidx * g.C, hi = min(prob_n(q), lo + g.C);
#pragma unroll 1
    for (int i0 = lo + threadIdx.x; i0 < hi; i0 += 4 * kPipeThreads) {      // four point loads in flight per thread
        float4 p[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int i = i0 + u * kPipeThreads; if (i < hi) p[u] = q.pts[i]; }     // plain load: a map cloud was written by phase P0 of this launch
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * kPipeThreads;
            if (i < hi) {
                const unsigned key = (unsigned)voxel_key(p[u], vp, i);
                keys[i] = key;
                atomicAdd(&sm.hist[key & 255u], 1);
            }
        }
    }
    __syncthreads();
    q.hist[sidx * 256 + threadIdx.x] = sm.hist[threadIdx.x];
    __syncthreads();
}

// ---- digit offsets of a pass: per digit, the exclusive prefix of its counts over the blocks (in place) and its total --------
// One warp per digit; all loads of a lane are issued before the scan (a dependent loop would pay an L2 round trip per 32 blocks).
LILOC_DEV void pipe_sort_offsets(const VgProb& q, int pass, const int nsort) {
    const SortGeom g = sort_geom(prob_n(q), nsort);
    int* __restrict__ hcur = q.hist + (pass % 3) * kHistStride;
    int* __restrict__ tot = q.hist + 3 * kHistStride;
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * (kPipeThreads / 32) + (threadIdx.x >> 5), nwarps = gridDim.x * (kPipeThreads / 32);
    constexpr int kTrips = (kSortBlocksMax + 31) / 32;
    for (int d = gwarp; d < 256; d += nwarps) {
        int v[kTrips];
#pragma unroll
        for (int u = 0; u < kTrips; ++u) { const int b = u * 32 + lane; v[u] = b < g.nb ? __ldcg(&hcur[b * 256 + d]) : 0; }
        int run = 0;
#pragma unroll
        for (int u = 0; u < kTrips; ++u) {
            int inc = v[u];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            const int b = u * 32 + lane;
            if (b < g.nb) hcur[b * 256 + d] = run + inc - v[u];
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) tot[d] = run;
    }
}

// ---- one radix pass: stable in-tile ranking, scatter, next histogram ------------------------------------------------
LILOC_DEV void pipe_sort_pass(const VgProb& q, int npass, int pass, PipeSmem& sm, const int sidx, const int nsort, const bool two_level) {
    const SortGeom g = sort_geom(prob_n(q), nsort);
    if (sidx < 0 || sidx >= g.nb) return;
    const int src = (npass + 1 + pass) & 1, dst = src ^ 1;
    const unsigned* __restrict__ skeys = q.keys[src];
    const unsigned* __restrict__ svals = (pass == 0) ? nullptr : q.vals[src];      // pass 0: value = source index
    unsigned* __restrict__ dkeys = q.keys[dst];
    unsigned* __restrict__ dvals = q.vals[dst];
    const int* __restrict__ hcur = q.hist + (pass % 3) * kHistStride;
    int* __restrict__ hnext = (pass + 1 < npass) ? q.hist + ((pass + 1) % 3) * kHistStride : nullptr;
    int* __restrict__ hzero = q.hist + ((pass + 2) % 3) * kHistStride;
    const int shift = pass * 8;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    // digit offsets of this block: all lower digits of every block + this digit of the blocks before
    if (two_level) {        // scan of the digit totals + this block's row, turned into exclusive prefixes by pipe_sort_offsets
        const int tot = __ldcg(&q.hist[3 * kHistStride + t]);
        const int pre = __ldcg(&hcur[sidx * 256 + t]);
        int total;
        const int excl = block_exclusive_scan_256(tot, &total);
        sm.sort.offs[t] = excl + pre;
#pragma unroll
        for (int u = 0; u < kSortUnits; ++u) sm.sort.cnt[u][t] = 0;
    } else {
        // Warp w reads rows w, w + 8, ... whole (a lane: eight digits = two 16-byte loads, kRowsInFlight rows in flight), the eight
        // warps' partial sums meet in shared memory: 4x fewer load instructions and half the dependent round trips of the version
        // where thread t walked column t of every row.
        constexpr int kRowsInFlight = 4;
        int tot8[8], pre8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { tot8[j] = 0; pre8[j] = 0; }
#pragma unroll 1
        for (int b = warp; b < g.nb; b += 8 * kRowsInFlight) {
            int4 a[kRowsInFlight][2];
#pragma unroll
            for (int u = 0; u < kRowsInFlight; ++u) {
                const int bb = b + 8 * u;
                a[u][0] = make_int4(0, 0, 0, 0); a[u][1] = make_int4(0, 0, 0, 0);
                if (bb < g.nb) {
                    a[u][0] = __ldcg(reinterpret_cast<const int4*>(hcur + bb * 256 + 8 * lane));
                    a[u][1] = __ldcg(reinterpret_cast<const int4*>(hcur + bb * 256 + 8 * lane) + 1);
                }
            }
#pragma unroll
            for (int u = 0; u < kRowsInFlight; ++u) {
                const int v[8] = {a[u][0].x, a[u][0].y, a[u][0].z, a[u][0].w, a[u][1].x, a[u][1].y, a[u][1].z, a[u][1].w};
                const bool before = b + 8 * u < sidx;
#pragma unroll
                for (int j = 0; j < 8; ++j) { tot8[j] += v[j]; pre8[j] += before ? v[j] : 0; }
            }
        }
        int* stage = reinterpret_cast<int*>(&sm.sort.pre[0][0]);        // 2 x 8 x 256 ints: free until the first tile's prefix step
        static_assert(sizeof(sm.sort.pre) >= 2 * 8 * 256 * sizeof(int), "staging of the warps' partial digit sums");
#pragma unroll
        for (int j = 0; j < 8; ++j) { stage[warp * 256 + 8 * lane + j] = tot8[j]; stage[2048 + warp * 256 + 8 * lane + j] = pre8[j]; }
        __syncthreads();
        int pre = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { tot += stage[w * 256 + t]; pre += stage[2048 + w * 256 + t]; }
        int total;
        const int excl = block_exclusive_scan_256(tot, &total);
        sm.sort.offs[t] = excl + pre;
#pragma unroll
        for (int u = 0; u < kSortUnits; ++u) sm.sort.cnt[u][t] = 0;
    }
    __syncthreads();
    if (pass >= 1) hzero[sidx * 256 + t] = 0;                  // free since the previous barrier; needed two passes on
    const int lo = sidx * g.C, hi = min(prob_n(q), lo + g.C);
    int carry = 0;                                              // items of digit t in the tiles already scattered
    for (int base = lo; base < hi; base += kSortTile) {
        const int rounds = min(kSortRounds, (hi - base + kPipeThreads - 1) / kPipeThreads);
        unsigned k[kSortRounds], v[kSortRounds];
        int rank[kSortRounds];
#pragma unroll
        for (int r = 0; r < kSortRounds; ++r) {
            const int i = base + r * kPipeThreads + t;
            k[r] = 0u; v[r] = 0u;
            if (i < hi) { k[r] = skeys[i]; v[r] = svals ? svals[i] : (unsigned)i; }
        }
        // 1. stable rank inside every (round, warp) unit by warp match; the unit leader publishes the unit's count
#pragma unroll
        for (int r = 0; r < kSortRounds; ++r) {
            const int i = base + r * kPipeThreads + t;
            const bool valid = i < hi;
            const unsigned d = valid ? ((k[r] >> shift) & 255u) : 256u;
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            rank[r] = __popc(peers & ((1u << lane) - 1u));
            if (valid && rank[r] == 0) sm.sort.cnt[r * (kPipeThreads / 32) + warp][d] = (unsigned short)__popc(peers);
        }
        __syncthreads();
        // 2. per digit: exclusive prefix over the units (item order), counters cleared for the next tile
        {
            sm.sort.offs[t] += carry;
            int run = 0;
            const int units = rounds * (kPipeThreads / 32);
            for (int u = 0; u < units; ++u) {
                const int c = sm.sort.cnt[u][t];
                sm.sort.pre[u][t] = (unsigned short)run;
                sm.sort.cnt[u][t] = 0;
                run += c;
            }
            carry = run;
        }
        __syncthreads();
        // 3. scatter; the histogram of the next pass is accumulated at the destination block's row
#pragma unroll
        for (int r = 0; r < kSortRounds; ++r) {
            const int i = base + r * kPipeThreads + t;
            if (i < hi) {
                const unsigned d = (k[r] >> shift) & 255u;
                const int pos = sm.sort.offs[d] + sm.sort.pre[r * (kPipeThreads / 32) + warp][d] + rank[r];
                dkeys[pos] = k[r];
                dvals[pos] = v[r];
                if (hnext) atomicAdd(&hnext[(pos / g.C) * 256 + ((k[r] >> (shift + 8)) & 255u)], 1);
            }
        }
        __syncthreads();
    }
}

// ---- R2: centroids (+ search-cell count) ----------------------------------------------------------------------------
LILOC_DEV void pipe_centroids(const VgProb& q, const GridParams& gp, PipeSmem& sm, const unsigned epoch, int* error_flag) {
    const unsigned* __restrict__ skeys = q.keys[1];
    const unsigned* __restrict__ svals = q.vals[1];
    const int n = prob_n(q);
    const int n_tiles = (n + kRunTile - 1) / kRunTile;
    if (n_tiles == 0 && blockIdx.x == 0 && threadIdx.x == 0) *q.count = 0;
    // A block with several tiles publishes the head counts of ALL of them before it starts the first centroid walk:
    // otherwise its later tiles would be published only after the earlier walks, and the blocks would advance in
    // lock-step rounds paced by the slowest walk of every round (measured on the 836 k-point VLP-16 map: 59 -> 88 us).
    const bool ahead = n_tiles > (int)gridDim.x;
    if (ahead) {
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int base = tile * kRunTile;
            int c = 0;
#pragma unroll
            for (int j = 0; j < kRunPerThread; ++j) {
                const int i = base + j * kPipeThreads + threadIdx.x;
                if (i < n) c += (i == 0 || skeys[i] != skeys[i - 1]) ? 1 : 0;
            }
            pipe_publish(q.tile_state, tile, block_sum_256(c), epoch);
        }
    }
    // A block's later tiles do not sum the head counts of ALL earlier tiles again (tile 800 of a 836 k-point map: four
    // dependent L2 round trips per tile, O(tiles^2) loads in total): the offset of the block's previous tile is carried
    // and only the tiles in between are added -- at most one word per thread, one round trip.
    int prev_tile = -1, prev_end = 0;                           // prev_end = run heads in tiles [0, prev_tile]
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int base = tile * kRunTile;
        const int cnt = min(kRunTile, n - base);
        // The tile AND the kRunHalo pairs behind it are gathered into shared memory by all threads.  A run that straddles the
        // end of the tile used to be continued by its head thread alone through dependent global loads (key -> index ->
        // point: three L2 round trips per element, ~1.5 us each) -- every tile has such a run, and on maps where a voxel
        // collects points of many keyframes that tail was the longest piece of the phase.
        const int ext = min(kRunTile + kRunHalo, n - base);     // pairs available in shared memory
#pragma unroll
        for (int j = 0; j < (kRunTile + kRunHalo) / kPipeThreads; ++j) {
            const int l = j * kPipeThreads + threadIdx.x;
            if (l < ext) {
                sm.run.key[l] = skeys[base + l];
                sm.run.pt[l] = q.pts[svals[base + l]];
            }
        }
        if (threadIdx.x == 0) sm.run.prev = (base > 0) ? skeys[base - 1] : 0u;
        __syncthreads();
        const int l0 = threadIdx.x * kRunPerThread;
        unsigned char head[kRunPerThread];
        int c = 0;
#pragma unroll
        for (int j = 0; j < kRunPerThread; ++j) {
            const int l = l0 + j;
            bool h = false;
            if (l < cnt) h = (l == 0) ? (base == 0 || sm.run.key[0] != sm.run.prev) : (sm.run.key[l] != sm.run.key[l - 1]);
            head[j] = h ? 1 : 0;
            c += head[j];
        }
        int total;
        const int local_rank = block_exclusive_scan_256(c, &total);
        const int tile_offset = prev_end + pipe_lookback(q.tile_state, tile, total, epoch, error_flag, !ahead, prev_tile + 1);      // run heads in all earlier tiles
        prev_tile = tile; prev_end = tile_offset + total;
        int rank = tile_offset + local_rank;
        if (tile == n_tiles - 1 && threadIdx.x == 0) *q.count = tile_offset + total;
#pragma unroll
        for (int j = 0; j < kRunPerThread; ++j) {
            if (!head[j]) continue;
            const int l = l0 + j;
            const unsigned key = sm.run.key[l];
            float sx = 0.f, sy = 0.f, sz = 0.f, sw = 0.f;
            int e = l;
            // four elements per trip: the keys and points of a trip are loaded together, so a long run (a voxel seen by
            // many keyframes) costs the four dependent float adds per element, not a shared-memory latency per element
            bool more = true;
            while (more) {
                float4 pp[4]; bool same[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int ee = min(e + u, ext - 1);
                    same[u] = (e + u < ext) && sm.run.key[ee] == key;
                    pp[u] = sm.run.pt[ee];
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (more && same[u]) { sx += pp[u].x; sy += pp[u].y; sz += pp[u].z; sw += pp[u].w; ++e; }
                    else more = false;
                }
            }
            int len = e - l;
            if (e == ext) {                              // longer than the halo: the rest from global memory, four pairs per trip
                int gi = base + ext;
                bool go_on = gi < n;
                while (go_on) {
                    unsigned kk[4]; float4 pg[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) kk[u] = (gi + u < n) ? skeys[gi + u] : ~key;
#pragma unroll
                    for (int u = 0; u < 4; ++u) pg[u] = (gi + u < n && kk[u] == key) ? q.pts[svals[gi + u]] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (go_on && gi + u < n && kk[u] == key) { sx += pg[u].x; sy += pg[u].y; sz += pg[u].z; sw += pg[u].w; ++len; }
                        else go_on = false;
                    }
                    gi += 4;
                }
            }
            const float fc = (float)len;
            const float4 o = make_float4(sx / fc, sy / fc, sz / fc, sw / fc);
            q.out[rank] = o;
            if (q.build_grid) {
                const int cx = cell_coord(o.x, gp.inv_cell, gp.g0[0]);
                const int cy = cell_coord(o.y, gp.inv_cell, gp.g0[1]);
                const int cz = cell_coord(o.z, gp.inv_cell, gp.g0[2]);
                atomicAdd(&q.cell_counts[(cz * gp.dim[1] + cy) * gp.dim[0] + cx], 1);
            }
            ++rank;
        }
        __syncthreads();
    }
}

// grid_only problems: the cloud itself is the search set -- count its points per cell, publish the count
LILOC_DEV void pipe_cells_count_direct(const VgProb& q, const GridParams& gp) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *q.count = prob_n(q);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < prob_n(q); i += gridDim.x * blockDim.x) {
        const float4 p = q.out[i];
        const int cx = cell_coord(p.x, gp.inv_cell, gp.g0[0]);
        const int cy = cell_coord(p.y, gp.inv_cell, gp.g0[1]);
        const int cz = cell_coord(p.z, gp.inv_cell, gp.g0[2]);
        atomicAdd(&q.cell_counts[(cz * gp.dim[1] + cy) * gp.dim[0] + cx], 1);
    }
}

// ---- G2, G3: cell counts -> cell_start (one pass, look-back across tiles), points sorted by cell -----------------------------------------------------------
LILOC_DEV void pipe_cells_scan(const VgProb& q, const GridParams& gp, const unsigned epoch, int* error_flag) {
    const int n = gp.ncells;
    const int n_tiles = (n + kScanTile - 1) / kScanTile;
    int prev_tile = -1, prev_end = 0;                           // carried like in pipe_centroids
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int base = tile * kScanTile + threadIdx.x * kScanPerThread;
        const bool whole = base + kScanPerThread <= n;          // cell_counts / cell_start are 16-byte aligned and base is a multiple of 16
        int v[kScanPerThread];
        if (whole) {
#pragma unroll
            for (int u = 0; u < kScanPerThread / 4; ++u) {
                const int4 w = __ldcg(reinterpret_cast<const int4*>(q.cell_counts + base) + u);
                v[4 * u] = w.x; v[4 * u + 1] = w.y; v[4 * u + 2] = w.z; v[4 * u + 3] = w.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < kScanPerThread; ++j) v[j] = (base + j < n) ? __ldcg(&q.cell_counts[base + j]) : 0;
        }
        int c = 0;
#pragma unroll
        for (int j = 0; j < kScanPerThread; ++j) c += v[j];
        int total;
        const int local = block_exclusive_scan_256(c, &total);
        const int tile_offset = prev_end + pipe_lookback(q.grid_state, tile, total, epoch, error_flag, true, prev_tile + 1);
        prev_tile = tile; prev_end = tile_offset + total;
        int run = tile_offset + local;
        if (whole) {
#pragma unroll
            for (int u = 0; u < kScanPerThread / 4; ++u) {
                int4 w;
                w.x = run; run += v[4 * u]; w.y = run; run += v[4 * u + 1]; w.z = run; run += v[4 * u + 2]; w.w = run; run += v[4 * u + 3];
                reinterpret_cast<int4*>(q.cell_start + base)[u] = w;
            }
            if (base + kScanPerThread == n) q.cell_start[n] = run;
        } else {
#pragma unroll
            for (int j = 0; j < kScanPerThread; ++j) {
                if (base + j < n) q.cell_start[base + j] = run;
                run += v[j];
                if (base + j == n - 1) q.cell_start[n] = run;
            }
        }
    }
}

// The counters are counted DOWN while the points take their places: the order inside a cell is arbitrary either way (the 5-NN
// is a total order over (d2, index)), and every counter is back at zero when the phase ends -- nothing to clear for the next
// build (the previous version cleared ncells counters in phase P0 and wrote a cursor array in G2: 8 MB per frame at 1 M cells).
LILOC_DEV void pipe_cells_scatter(const VgProb& q, const GridParams& gp) {
    const int M = *q.count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
        const float4 p = q.out[i];
        const int cx = cell_coord(p.x, gp.inv_cell, gp.g0[0]);
        const int cy = cell_coord(p.y, gp.inv_cell, gp.g0[1]);
        const int cz = cell_coord(p.z, gp.inv_cell, gp.g0[2]);
        const int cell = (cz * gp.dim[1] + cy) * gp.dim[0] + cx;
        const int pos = __ldcg(&q.cell_start[cell]) + atomicSub(&q.cell_counts[cell], 1) - 1;
        q.cpts[pos] = make_float4(p.x, p.y, p.z, __int_as_float(i));
    }
}

// -----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPipeThreads, 3) k_voxel_pipeline(const __grid_constant__ PipeArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ __align__(16) PipeSmem sm;
    __shared__ VgParams s_vp[kMaxProbs];
    __shared__ GridParams s_gp[kMaxProbs];
    __shared__ int s_npass[kMaxProbs];
    const int np = a.n_prob;
    __shared__ int s_sort_idx, s_nsort;
    // sort mode of the launch (uniform over the grid: problem sizes and the grid size only)
    bool two_level = false;
    for (int p = 0; p < np; ++p) two_level |= a.prob[p].n >= kSortTwoLevelMin && !a.prob[p].grid_only;
    if (threadIdx.x == 0) {
        int idx = -1;
        if (two_level) {
            idx = (int)blockIdx.x < kSortBlocksMax ? (int)blockIdx.x : -1;
        } else {                                 // the first block to arrive on an SM becomes that SM's sorting block
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            const int slot = atomicAdd(&a.sm_scratch[smid & 511u], 1);
            if (slot == 0) { idx = atomicAdd(&a.sm_scratch[512], 1); if (idx >= kSortBlocksSmall) idx = -1; }
        }
        s_sort_idx = idx;
    }
    int mark = 0;
#define PIPE_MARK() do { if (a.marks && blockIdx.x == 0 && threadIdx.x == 0 && mark < kPipeMarks) a.marks[mark] = global_ns(); ++mark; } while (0)
    PIPE_MARK();                                                                          // 0 start
    for (int p = 0; p < np; ++p) pipe_bbox(a, a.prob[p]);                                   // P0
    grid.sync();
    PIPE_MARK();                                                                          // 1 after assemble + bbox
    if (__ldcg(&a.sm_scratch[kSmAbort]) != 0) {                                             // a non-finite input point: nothing below may run
        grid.sync();                                                                       // every block has read the word
        if (blockIdx.x == 0) {
            for (int i = threadIdx.x; i < kSmScratch; i += kPipeThreads) a.sm_scratch[i] = 0;
            if (threadIdx.x < 6 * np) { const int p = threadIdx.x / 6, w = threadIdx.x % 6; a.prob[p].enc[w] = w < 3 ? 0xFFFFFFFFu : 0u; }
            if (threadIdx.x == 0 && a.error_flag) *(volatile int*)a.error_flag = 2;
        }
        return;
    }

    if (threadIdx.x < np) {                                                               // P1
        const VgProb& q = a.prob[threadIdx.x];
        const VgParams vp = vg_params_compute(q.enc, q.leaf, prob_n(q));
        s_vp[threadIdx.x] = vp;
        s_npass[threadIdx.x] = (prob_n(q) > 0 && !q.grid_only) ? (vp.key_bits + 7) / 8 : 0;
        GridParams g{};
        if (q.build_grid) g = grid_params_compute(vp.mn, vp.mx, prob_n(q) <= 0, q.max_cells);
        s_gp[threadIdx.x] = g;
        if (blockIdx.x == 0) { *q.vp = vp; if (q.build_grid) *q.gp = g; }
    }
    __syncthreads();
    int max_pass = 0;
    for (int p = 0; p < np; ++p) max_pass = max(max_pass, s_npass[p]);
    if (threadIdx.x == 0) s_nsort = two_level ? min((int)gridDim.x, kSortBlocksMax) : min(__ldcg(&a.sm_scratch[512]), kSortBlocksSmall);
    __syncthreads();
    const int sidx = s_sort_idx, nsort = s_nsort;
    for (int p = 0; p < np; ++p) if (prob_n(a.prob[p]) > 0 && !a.prob[p].grid_only) pipe_keys(a.prob[p], s_vp[p], s_npass[p], sm, sidx, nsort);
    grid.sync();
    if (blockIdx.x == 0) for (int i = threadIdx.x; i < kSmAbort; i += kPipeThreads) a.sm_scratch[i] = 0;   // re-armed for the next launch
    PIPE_MARK();                                                                          // 2 after keys

    for (int pass = 0; pass < max_pass; ++pass) {                                         // radix passes: (digit offsets,) rank + scatter
        if (two_level) {
            for (int p = 0; p < np; ++p) if (pass < s_npass[p]) pipe_sort_offsets(a.prob[p], pass, nsort);
            grid.sync();
        }
        for (int p = 0; p < np; ++p) if (pass < s_npass[p]) pipe_sort_pass(a.prob[p], s_npass[p], pass, sm, sidx, nsort, two_level);
        grid.sync();
    }
    if (blockIdx.x == 0 && threadIdx.x < 6 * np) {                                        // re-arm the bounding boxes
        const int p = threadIdx.x / 6, w = threadIdx.x % 6;
        a.prob[p].enc[w] = w < 3 ? 0xFFFFFFFFu : 0u;
    }
    PIPE_MARK();                                                                          // 3 after the sort
    bool any_out = false, any_grid = false;
    for (int p = 0; p < np; ++p) { any_out |= !a.prob[p].sort_only; any_grid |= a.prob[p].build_grid != 0; }
    if (!any_out) return;

    PIPE_MARK();                                                                          // 4 (kept for the layout of the marks)
    for (int p = 0; p < np; ++p) {                                                        // R2 (R1, a counting phase, is folded into it)
        if (a.prob[p].grid_only) pipe_cells_count_direct(a.prob[p], s_gp[p]);
        else if (!a.prob[p].sort_only) pipe_centroids(a.prob[p], s_gp[p], sm, a.epoch, a.error_flag);
    }
    if (!any_grid) { PIPE_MARK(); return; }                                               // 5 (block 0's own end)
    grid.sync();
    PIPE_MARK();                                                                          // 5 after the centroids
    for (int p = 0; p < np; ++p) if (a.prob[p].build_grid) pipe_cells_scan(a.prob[p], s_gp[p], a.epoch, a.error_flag);   // G2
    grid.sync();
    PIPE_MARK();                                                                          // 6 after the cell scan
    for (int p = 0; p < np; ++p) if (a.prob[p].build_grid) pipe_cells_scatter(a.prob[p], s_gp[p]);   // G3
    PIPE_MARK();                                                                          // 7 block 0's own end
#undef PIPE_MARK
}

}  // namespace liloc
