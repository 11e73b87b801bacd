// liloc_b200/csrc/voxelgrid.cuh -- types and block-level helpers of the pcl::VoxelGrid<PointXYZI> restatement and of
// the keyframe transform + concatenation that feeds it.  The phases themselves live in pipeline.cuh (one cooperative
// kernel); this header holds what they share with ndt.cuh and liloc_gpu.cu.
//
// Replaces (reference): downSizeFilterLocalMapSurf.filter  src/mapOptmization.cpp:954-955,
//                       downSizeFilterSurf.filter          src/mapOptmization.cpp:972-973,
//                       transformPointCloud + `+=` concat  include/utility.h:388-407, :944-950.
// Semantics restated in SURVEY.md A.1: float min/max -> int32 voxel index -> sort by index (stable
// in source order) -> one centroid per voxel, float sums accumulated sequentially in source order,
// output in ascending voxel index.
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

// Block reduction of a bounding box, committed with ordered-uint atomics: enc[0..2] = min (armed with 0xFFFFFFFF),
// enc[3..5] = max (armed with 0).  All kMinMaxThreads threads call it.
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
    __syncthreads();                          // the staging arrays are reused by the next call (next problem of the phase)
}

constexpr int kAsmPerThread = 4;
constexpr int kAsmTile = kMinMaxThreads * kAsmPerThread;     // output points per block and step of the assemble phase

// tile of sorted (key, index) pairs handled by one block step of the run-head and centroid phases
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

}  // namespace liloc
