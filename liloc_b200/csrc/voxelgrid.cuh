// liloc_b200/csrc/voxelgrid.cuh -- pcl::VoxelGrid<PointXYZI> semantics on the device, and the fused
// keyframe transform + concatenation that feeds it.
//
// Replaces (reference): downSizeFilterLocalMapSurf.filter  src/mapOptmization.cpp:954-955,
//                       downSizeFilterSurf.filter          src/mapOptmization.cpp:972-973,
//                       transformPointCloud + `+=` concat  include/utility.h:388-407, :944-950.
// Semantics restated in SURVEY.md A.1: float min/max -> int32 voxel index -> sort by index (stable
// in source order) -> one centroid per voxel, float sums accumulated sequentially in source order,
// output in ascending voxel index.  All kernels are HBM-streaming: 16-byte float4 accesses,
// coalesced, grid sized in multiples of the SM count.
#pragma once
#include "common.cuh"

namespace liloc {

struct KfDesc {            // one selected keyframe of the local map
    long long src_off;     // first point in the keyframe arena
    int n;                 // points
    int dst_off;           // first point in the concatenated cloud
    float T[12];           // pose as a 3x4 transform
};

struct VgParams {
    float inv_leaf;
    int minb[3];
    int mul[3];
    int overflow;          // dx*dy*dz > INT32_MAX: output = input unchanged (PCL warning path)
    int key_bits;          // bits needed by the largest voxel index (or by n-1 when overflow)
    float mn[3], mx[3];    // bounding box (also sizes the search grid)
};

constexpr int kMinMaxThreads = 256;

LILOC_DEV void block_minmax_commit(float mn[3], float mx[3], unsigned* __restrict__ enc) {
    __shared__ float s_mn[3][kMinMaxThreads / 32], s_mx[3][kMinMaxThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
        if (lane == 0) { s_mn[a][warp] = mn[a]; s_mx[a][warp] = mx[a]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int a = threadIdx.x;
        float lo = s_mn[a][0], hi = s_mx[a][0];
#pragma unroll
        for (int w = 1; w < kMinMaxThreads / 32; ++w) { lo = fminf(lo, s_mn[a][w]); hi = fmaxf(hi, s_mx[a][w]); }
        if (lo <= hi) {                       // at least one point seen by this block
            atomicMin(&enc[a], f2ord(lo));
            atomicMax(&enc[3 + a], f2ord(hi));
        }
    }
}

// enc[0..2] = ordered-min (init 0xFFFFFFFF), enc[3..5] = ordered-max (init 0)
__global__ void __launch_bounds__(kMinMaxThreads) k_minmax(const float4* __restrict__ pts, int n, unsigned* __restrict__ enc) {
    float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = __ldg(&pts[i]);
        mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
        mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
        mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
    }
    block_minmax_commit(mn, mx, enc);
}

// Keyframe transform + concatenation fused with the bounding-box pass of the VoxelGrid.
// Each block owns kAsmTile consecutive output points; a keyframe switch inside a tile is handled by
// advancing the descriptor index (tiles are much smaller than keyframes).
constexpr int kAsmPerThread = 4;
constexpr int kAsmTile = kMinMaxThreads * kAsmPerThread;

__global__ void __launch_bounds__(kMinMaxThreads)
k_assemble(const float4* __restrict__ arena, const KfDesc* __restrict__ descs, int K, int n_raw,
           float4* __restrict__ raw, unsigned* __restrict__ enc) {
    float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    __shared__ int s_k0;
    for (int tile = blockIdx.x; tile * kAsmTile < n_raw; tile += gridDim.x) {
        const int base = tile * kAsmTile;
        __syncthreads();
        if (threadIdx.x == 0) {               // last descriptor with dst_off <= base
            int lo = 0, hi = K - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (descs[mid].dst_off <= base) lo = mid; else hi = mid - 1;
            }
            s_k0 = lo;
        }
        __syncthreads();
        int k = s_k0;
#pragma unroll
        for (int j = 0; j < kAsmPerThread; ++j) {
            const int i = base + j * kMinMaxThreads + threadIdx.x;
            if (i < n_raw) {
                while (k + 1 < K && descs[k + 1].dst_off <= i) ++k;
                const KfDesc* d = &descs[k];
                const float4 p = __ldg(&arena[d->src_off + (i - d->dst_off)]);
                const float4 o = apply_T(d->T, p);
                raw[i] = o;
                mn[0] = fminf(mn[0], o.x); mx[0] = fmaxf(mx[0], o.x);
                mn[1] = fminf(mn[1], o.y); mx[1] = fmaxf(mx[1], o.y);
                mn[2] = fminf(mn[2], o.z); mx[2] = fmaxf(mx[2], o.z);
            }
        }
    }
    block_minmax_commit(mn, mx, enc);
}

// VoxelGrid set-up, one thread.
__global__ void k_vg_params(const unsigned* __restrict__ enc, float leaf, int n, VgParams* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    VgParams p;
    p.inv_leaf = 1.0f / leaf;
    for (int a = 0; a < 3; ++a) { p.mn[a] = ord2f(enc[a]); p.mx[a] = ord2f(enc[3 + a]); }
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
    long long max_key = p.overflow ? (long long)(n > 0 ? n - 1 : 0)
                                   : (long long)divb[0] * divb[1] * divb[2] - 1;
    int bits = 1;
    while (bits < 32 && (max_key >> bits) != 0) ++bits;
    p.key_bits = bits;
    *out = p;
}

// int32 voxel index per point; values = source index.
__global__ void __launch_bounds__(256)
k_vg_keys(const float4* __restrict__ pts, int n, const VgParams* __restrict__ vp, unsigned* __restrict__ keys, unsigned* __restrict__ vals) {
    const float inv = vp->inv_leaf;
    const int m0 = vp->minb[0], m1 = vp->minb[1], m2 = vp->minb[2];
    const int u1 = vp->mul[1], u2 = vp->mul[2];
    const int overflow = vp->overflow;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = __ldg(&pts[i]);
        int key;
        if (overflow) key = i;
        else {
            const int i0 = (int)(floorf(p.x * inv) - (float)m0);
            const int i1 = (int)(floorf(p.y * inv) - (float)m1);
            const int i2 = (int)(floorf(p.z * inv) - (float)m2);
            key = i0 + i1 * u1 + i2 * u2;
        }
        keys[i] = (unsigned)key;
        vals[i] = (unsigned)i;
    }
}

// ---- run heads -> ranks -> centroids ------------------------------------------------------------
constexpr int kRunThreads = 256;
constexpr int kRunPerThread = 4;
constexpr int kRunTile = kRunThreads * kRunPerThread;

LILOC_DEV int block_exclusive_scan_256(int v, int* total) {   // 256 threads
    __shared__ int s_w[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    int wbase = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { const int s = s_w[w]; if (w < warp) wbase += s; tot += s; }
    __syncthreads();
    *total = tot;
    return wbase + inc - v;
}

// number of run heads (sorted_keys[i] != sorted_keys[i-1]) per tile
__global__ void __launch_bounds__(kRunThreads)
k_run_count(const unsigned* __restrict__ skeys, int n, int* __restrict__ tile_counts) {
    const int base = blockIdx.x * kRunTile;
    int c = 0;
#pragma unroll
    for (int j = 0; j < kRunPerThread; ++j) {
        const int i = base + j * kRunThreads + threadIdx.x;
        if (i < n) c += (i == 0 || __ldg(&skeys[i]) != __ldg(&skeys[i - 1])) ? 1 : 0;
    }
    int total;
    block_exclusive_scan_256(c, &total);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}

// exclusive scan of the tile counts by one block; total -> *total_out
LILOC_DEV void scan_tiles_body(int* __restrict__ tile_counts, int n_tiles, int* __restrict__ total_out) {
    __shared__ int s_w[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n_tiles; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = (i < n_tiles) ? tile_counts[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        int wbase = 0, tot = 0;
        for (int w = 0; w < 32; ++w) { const int s = s_w[w]; if (w < warp) wbase += s; tot += s; }
        const int carry = s_carry;
        if (i < n_tiles) tile_counts[i] = carry + wbase + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = s_carry;
}
__global__ void __launch_bounds__(1024)
k_scan_tiles(int* __restrict__ tile_counts, int n_tiles, int* __restrict__ total_out) {
    scan_tiles_body(tile_counts, n_tiles, total_out);
}

// One centroid per run, written at the run's rank (ascending voxel index).  The tile's points are
// gathered into shared memory by all threads in parallel (coalesced key/index loads, 16-byte point
// gathers); the thread that owns a run head then walks the run sequentially out of shared memory:
// float sums in ascending source index (the canonical order).  Runs that cross the tile end continue
// from global memory.
__global__ void __launch_bounds__(kRunThreads)
k_run_centroids(const unsigned* __restrict__ skeys, const unsigned* __restrict__ svals, int n,
                const int* __restrict__ tile_offsets, const float4* __restrict__ pts, float4* __restrict__ out) {
    __shared__ unsigned s_key[kRunTile];
    __shared__ float4 s_pt[kRunTile];
    __shared__ unsigned s_prev;
    const int base = blockIdx.x * kRunTile;
    const int cnt = min(kRunTile, n - base);
#pragma unroll
    for (int j = 0; j < kRunPerThread; ++j) {
        const int l = j * kRunThreads + threadIdx.x;
        if (l < cnt) {
            s_key[l] = __ldg(&skeys[base + l]);
            s_pt[l] = __ldg(&pts[__ldg(&svals[base + l])]);
        }
    }
    if (threadIdx.x == 0) s_prev = (base > 0) ? __ldg(&skeys[base - 1]) : 0u;
    __syncthreads();
    const int l0 = threadIdx.x * kRunPerThread;
    unsigned char head[kRunPerThread];
    int c = 0;
#pragma unroll
    for (int j = 0; j < kRunPerThread; ++j) {
        const int l = l0 + j;
        bool h = false;
        if (l < cnt) h = (l == 0) ? (base == 0 || s_key[0] != s_prev) : (s_key[l] != s_key[l - 1]);
        head[j] = h ? 1 : 0;
        c += head[j];
    }
    int total;
    int rank = tile_offsets[blockIdx.x] + block_exclusive_scan_256(c, &total);
#pragma unroll
    for (int j = 0; j < kRunPerThread; ++j) {
        if (!head[j]) continue;
        const int l = l0 + j;
        const unsigned key = s_key[l];
        float sx = 0.f, sy = 0.f, sz = 0.f, sw = 0.f;
        int e = l;
        do {
            const float4 p = s_pt[e];
            sx += p.x; sy += p.y; sz += p.z; sw += p.w;
            ++e;
        } while (e < cnt && s_key[e] == key);
        int len = e - l;
        if (e == cnt) {                              // run may continue in the next tile(s)
            int g = base + cnt;
            while (g < n && __ldg(&skeys[g]) == key) {
                const float4 p = __ldg(&pts[__ldg(&svals[g])]);
                sx += p.x; sy += p.y; sz += p.z; sw += p.w;
                ++g; ++len;
            }
        }
        const float fc = (float)len;
        out[rank] = make_float4(sx / fc, sy / fc, sz / fc, sw / fc);
        ++rank;
    }
}

}  // namespace liloc
