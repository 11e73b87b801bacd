// liloc_b200/csrc/common.cuh -- shared device helpers for the sm_100a scan-to-map path.
//
// Arithmetic contract (must stay in lock-step with the parity oracle, see DESIGN.md "Arithmetic"):
// this translation unit is compiled with -fmad=false, IEEE division and square root (nvcc
// defaults), so every float expression below rounds exactly like the reference's x86-64 build.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define LILOC_DEV __device__ __forceinline__

namespace liloc {

// ---- order-preserving float <-> uint encoding, for atomicMin/atomicMax on floats -----------------
LILOC_DEV unsigned f2ord(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float ord2f(unsigned u) {
    const unsigned v = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(v);
#else
    float f; memcpy(&f, &v, 4); return f;
#endif
}

// ---- pcl::getTransformation (src/mapOptmization.cpp:404-406) -------------------------
// sin/cos evaluated in double and rounded to float (canonical choice shared with the oracle).
__host__ __device__ inline void pose_to_T(const float* pose6, float* T) {
    const float roll = pose6[0], pitch = pose6[1], yaw = pose6[2];
    const float A = (float)cos((double)yaw),   B = (float)sin((double)yaw);
    const float C = (float)cos((double)pitch), D = (float)sin((double)pitch);
    const float E = (float)cos((double)roll),  F = (float)sin((double)roll);
    const float DE = D * E, DF = D * F;
    T[0] = A * C;  T[1] = A * DF - B * E;  T[2]  = B * F + A * DE;  T[3]  = pose6[3];
    T[4] = B * C;  T[5] = A * E + B * DF;  T[6]  = B * DE - A * F;  T[7]  = pose6[4];
    T[8] = -D;     T[9] = C * F;           T[10] = C * E;           T[11] = pose6[5];
}

// pointAssociateToMap (src/mapOptmization.cpp:388-393) / transformPointCloud (include/utility.h:400-405)
LILOC_DEV float4 apply_T(const float* __restrict__ T, const float4 p) {
    float4 o;
    o.x = T[0] * p.x + T[1] * p.y + T[2]  * p.z + T[3];
    o.y = T[4] * p.x + T[5] * p.y + T[6]  * p.z + T[7];
    o.z = T[8] * p.x + T[9] * p.y + T[10] * p.z + T[11];
    o.w = p.w;
    return o;
}

LILOC_DEV float dist2(const float4 q, const float4 p) {
    const float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
    return ((dx * dx) + (dy * dy)) + (dz * dz);
}

// (d2, index) total order: the canonical kNN tie-break
LILOC_DEV bool nn_less(float da, int ia, float db, int ib) { return da < db || (da == db && ia < ib); }

// ---- warp helpers ---------------------------------------------------------------------------------
LILOC_DEV double shfl_xor_d(double v, int m) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, m);
    hi = __shfl_xor_sync(0xffffffffu, hi, m);
    return __hiloint2double(hi, lo);
}
LILOC_DEV double shfl_down_d(double v, int d) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_down_sync(0xffffffffu, lo, d);
    hi = __shfl_down_sync(0xffffffffu, hi, d);
    return __hiloint2double(hi, lo);
}

}  // namespace liloc
