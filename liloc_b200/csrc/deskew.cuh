// liloc_b200/csrc/deskew.cuh -- imageProjection's per-point work on the device.
//
// Replaces (reference): projectPointCloud (src/imageProjection.cpp:645-673) with deskewPoint (:615-643),
// findRotation (:573-597) and findPosition (:599-613, always zero).  The raw sweep (x, y, z, intensity, ring, time)
// is uploaded once; the range / ring / stride filters, the stable compaction (push_back order of fullCloud) and the
// IMU-interpolated rotation of every surviving point run here, and the de-skewed cloud stays on the device as the input
// of downsampleCurrentScan (the first phase of the voxel pipeline reads its size from device memory).
//
// One cooperative launch, one grid barrier:
//   A  per 256-point tile: keep flags -> tile count
//   -- grid.sync --
//   B  every block: offset of its first tile (sum of the earlier tile counts), the sweep's first surviving point
//      (it fixes transStartInverse, :628-631), then per tile: stable rank of the survivors, transform, store.
// Arithmetic (bit-exact against oracle/deskew_oracle.cpp): double interpolation of the integrated gyro angles narrowed
// to float, pcl::getTransformation with sin/cos in double rounded to float, Eigen's Affine3f inverse / product
// evaluation order (3-term sums as a0 + (a1 + a2)), -fmad=false.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"

namespace liloc {

namespace cg = cooperative_groups;

constexpr int kDeskewThreads = 256;
constexpr int kImuMax = 2000;                 // queueLength (src/imageProjection.cpp:61)

struct DeskewArgs {
    const float4* pts;           // x, y, z, intensity
    const int2* rt;              // ring, time (float bits)
    int n;
    const double* imu;           // 4 x imu_stride: time | rotX | rotY | rotZ
    int imu_stride;
    int imu_last;                // imuPointerCur
    double t_cur;                // timeScanCur
    int deskew;                  // deskewFlag != -1 && imuAvailable
    float rmin, rmax;
    int n_scan, ds_rate, pf_num;
    int* tile_count;             // ceil(n / 256)
    float4* out;
    int* n_out;
    int* n_out_host;             // the same count into mapped host memory (no copy behind the launch)
    unsigned long long* marks;   // optional: %globaltimer (ns) of block 0 at start / after the barrier / at its end
};

struct Aff3 { float l[3][3]; float t[3]; };

LILOC_DEV unsigned long long deskew_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

LILOC_DEV float sum3(float a0, float a1, float a2) { return a0 + (a1 + a2); }

LILOC_DEV Aff3 aff_from_rpy(float roll, float pitch, float yaw) {      // pcl::getTransformation(0, 0, 0, roll, pitch, yaw)
    const float pose6[6] = {roll, pitch, yaw, 0.f, 0.f, 0.f};
    float T[12];
    pose_to_T(pose6, T);
    Aff3 a;
#pragma unroll
    for (int i = 0; i < 3; ++i) { a.l[i][0] = T[4 * i]; a.l[i][1] = T[4 * i + 1]; a.l[i][2] = T[4 * i + 2]; a.t[i] = T[4 * i + 3]; }
    return a;
}

#define LILOC_COF(m, i, j) ((m)[((i) + 1) % 3][((j) + 1) % 3] * (m)[((i) + 2) % 3][((j) + 2) % 3] - (m)[((i) + 1) % 3][((j) + 2) % 3] * (m)[((i) + 2) % 3][((j) + 1) % 3])

LILOC_DEV Aff3 aff_inverse(const Aff3& a) {                              // Eigen::Affine3f::inverse()
    const float c00 = LILOC_COF(a.l, 0, 0), c10 = LILOC_COF(a.l, 1, 0), c20 = LILOC_COF(a.l, 2, 0);
    const float det = sum3(c00 * a.l[0][0], c10 * a.l[1][0], c20 * a.l[2][0]);
    const float inv = 1.0f / det;
    Aff3 r;
    r.l[0][0] = c00 * inv;                  r.l[0][1] = c10 * inv;                  r.l[0][2] = c20 * inv;
    r.l[1][0] = LILOC_COF(a.l, 0, 1) * inv; r.l[1][1] = LILOC_COF(a.l, 1, 1) * inv; r.l[1][2] = LILOC_COF(a.l, 2, 1) * inv;
    r.l[2][0] = LILOC_COF(a.l, 0, 2) * inv; r.l[2][1] = LILOC_COF(a.l, 1, 2) * inv; r.l[2][2] = LILOC_COF(a.l, 2, 2) * inv;
#pragma unroll
    for (int i = 0; i < 3; ++i) r.t[i] = sum3((-r.l[i][0]) * a.t[0], (-r.l[i][1]) * a.t[1], (-r.l[i][2]) * a.t[2]);
    return r;
}
#undef LILOC_COF

LILOC_DEV Aff3 aff_mul(const Aff3& a, const Aff3& b) {                   // Eigen::Affine3f * Eigen::Affine3f
    Aff3 r;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) r.l[i][j] = sum3(a.l[i][0] * b.l[0][j], a.l[i][1] * b.l[1][j], a.l[i][2] * b.l[2][j]);
        r.t[i] = sum3(a.l[i][0] * b.t[0], a.l[i][1] * b.t[1], a.l[i][2] * b.t[2]) + a.t[i];
    }
    return r;
}

// findRotation (:573-597).  s_time: imuTime staged in shared memory; the three angle arrays are read from global memory.
// The reference walks the samples from the front until the first one later than the point; when the stamps are
// non-decreasing (`sorted`: the normal case) that index is found by bisection instead, with the same result.
LILOC_DEV void find_rotation(const double* s_time, const double* __restrict__ imu, int stride, int last, bool sorted, double pointTime,
                             float* rx, float* ry, float* rz) {
    int front = 0;
    if (sorted) {
        int lo = 0, hi = last;                      // first index in [0, last) with pointTime < time, else last
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (pointTime < s_time[mid]) hi = mid; else lo = mid + 1;
        }
        front = lo;
    } else {
        while (front < last) {
            if (pointTime < s_time[front]) break;
            ++front;
        }
    }
    const double* RX = imu + stride; const double* RY = imu + 2 * stride; const double* RZ = imu + 3 * stride;
    if (pointTime > s_time[front] || front == 0) {
        *rx = (float)__ldg(RX + front); *ry = (float)__ldg(RY + front); *rz = (float)__ldg(RZ + front);
    } else {
        const int back = front - 1;
        const double span = s_time[front] - s_time[back];
        const double rf = (pointTime - s_time[back]) / span;
        const double rb = (s_time[front] - pointTime) / span;
        *rx = (float)(__ldg(RX + front) * rf + __ldg(RX + back) * rb);
        *ry = (float)(__ldg(RY + front) * rf + __ldg(RY + back) * rb);
        *rz = (float)(__ldg(RZ + front) * rf + __ldg(RZ + back) * rb);
    }
}

// filters of projectPointCloud (:654-668)
LILOC_DEV bool deskew_keep(const DeskewArgs& a, int i, const float4 p, int ring) {
    const float range = sqrtf(p.x * p.x + p.y * p.y + p.z * p.z);
    // written so that a non-finite point fails (NaN compares false): the reference refuses non-dense clouds altogether
    // (src/imageProjection.cpp:395-399); for finite input this is the same test as `range < min || range > max` (:655)
    if (!(range >= a.rmin && range <= a.rmax)) return false;
    if (ring < 0 || ring >= a.n_scan) return false;
    if (ring % a.ds_rate != 0) return false;
    if (i % a.pf_num != 0) return false;
    return true;
}

__global__ void __launch_bounds__(kDeskewThreads, 2) k_deskew(const __grid_constant__ DeskewArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double s_time[kImuMax];
    __shared__ int s_w[8];
    __shared__ int s_red[2][8];
    __shared__ Aff3 s_start_inv;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_tiles = (a.n + kDeskewThreads - 1) / kDeskewThreads;
    const int tpb = (n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;         // contiguous tiles per block
    const int t0 = min(n_tiles, (int)blockIdx.x * tpb), t1 = min(n_tiles, t0 + tpb);

    if (a.marks && blockIdx.x == 0 && threadIdx.x == 0) a.marks[0] = deskew_ns();
    bool sorted = true;
    if (a.deskew) {
        for (int k = threadIdx.x; k <= a.imu_last; k += kDeskewThreads) {
            const double tk = __ldg(a.imu + k);
            s_time[k] = tk;
            if (k < a.imu_last && !(tk <= __ldg(a.imu + k + 1))) sorted = false;
        }
    }

    // ---- A: tile counts ----
    for (int t = t0; t < t1; ++t) {
        const int i = t * kDeskewThreads + threadIdx.x;
        bool keep = false;
        if (i < a.n) keep = deskew_keep(a, i, __ldg(a.pts + i), __ldg(a.rt + i).x);
        const int c = __syncthreads_count(keep);
        if (threadIdx.x == 0) a.tile_count[t] = c;
    }
    sorted = __syncthreads_and(sorted) != 0;
    grid.sync();
    if (a.marks && blockIdx.x == 0 && threadIdx.x == 0) a.marks[1] = deskew_ns();

    // ---- B: offsets, the sweep's first surviving point, transform + store ----
    int before = 0, total = 0, first_tile = INT_MAX;
    for (int u = threadIdx.x; u < n_tiles; u += kDeskewThreads) {
        const int c = a.tile_count[u];
        total += c;
        if (u < t0) before += c;
        if (c > 0) first_tile = min(first_tile, u);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        before += __shfl_xor_sync(0xffffffffu, before, o);
        total += __shfl_xor_sync(0xffffffffu, total, o);
        first_tile = min(first_tile, __shfl_xor_sync(0xffffffffu, first_tile, o));
    }
    if (lane == 0) { s_red[0][warp] = before; s_red[1][warp] = total; s_w[warp] = first_tile; }
    __syncthreads();
    before = 0; total = 0; first_tile = INT_MAX;
#pragma unroll
    for (int w = 0; w < 8; ++w) { before += s_red[0][w]; total += s_red[1][w]; first_tile = min(first_tile, s_w[w]); }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) { *a.n_out = total; *a.n_out_host = total; }
    if (t0 >= t1 || total == 0) {
        if (a.marks && blockIdx.x == 0 && threadIdx.x == 0) a.marks[2] = deskew_ns();
        return;
    }

    if (a.deskew) {
        // first survivor of the first non-empty tile: warp 0 scans that tile with ballots
        if (warp == 0) {
            int first = -1;
            for (int k = 0; k < kDeskewThreads && first < 0; k += 32) {
                const int i = first_tile * kDeskewThreads + k + lane;
                bool keep = false;
                if (i < a.n) keep = deskew_keep(a, i, __ldg(a.pts + i), __ldg(a.rt + i).x);
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                if (m) first = first_tile * kDeskewThreads + k + __ffs(m) - 1;
            }
            if (lane == 0) {
                float rx, ry, rz;
                find_rotation(s_time, a.imu, a.imu_stride, a.imu_last, sorted, a.t_cur + (double)__int_as_float(__ldg(a.rt + first).y), &rx, &ry, &rz);
                s_start_inv = aff_inverse(aff_from_rpy(rx, ry, rz));
            }
        }
        __syncthreads();
    }

    int base = before;
    for (int t = t0; t < t1; ++t) {
        const int i = t * kDeskewThreads + threadIdx.x;
        bool keep = false;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        int2 rt = make_int2(0, 0);
        if (i < a.n) { p = __ldg(a.pts + i); rt = __ldg(a.rt + i); keep = deskew_keep(a, i, p, rt.x); }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_w[warp] = __popc(m);
        __syncthreads();
        int wbase = 0, tile_total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { const int c = s_w[w]; if (w < warp) wbase += c; tile_total += c; }
        if (keep) {
            float4 o = p;
            if (a.deskew) {
                float rx, ry, rz;
                find_rotation(s_time, a.imu, a.imu_stride, a.imu_last, sorted, a.t_cur + (double)__int_as_float(rt.y), &rx, &ry, &rz);
                const Aff3 bt = aff_mul(s_start_inv, aff_from_rpy(rx, ry, rz));
                o.x = bt.l[0][0] * p.x + bt.l[0][1] * p.y + bt.l[0][2] * p.z + bt.t[0];
                o.y = bt.l[1][0] * p.x + bt.l[1][1] * p.y + bt.l[1][2] * p.z + bt.t[1];
                o.z = bt.l[2][0] * p.x + bt.l[2][1] * p.y + bt.l[2][2] * p.z + bt.t[2];
            }
            a.out[base + wbase + __popc(m & ((1u << lane) - 1u))] = o;
        }
        base += tile_total;
        __syncthreads();
    }
    if (a.marks && blockIdx.x == 0 && threadIdx.x == 0) a.marks[2] = deskew_ns();
}

}  // namespace liloc
