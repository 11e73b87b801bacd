// liloc_b200/csrc/liloc_gpu.cu -- context, memory and the C ABI declared in include/liloc_gpu.h.
//
// One context = one GPU, one stream, one caller thread (mirrors `mtx`, src/mapOptmization.cpp:255).
// Everything the path touches lives in HBM for the lifetime of the context: the keyframe arena
// (body-frame float4 clouds), the concatenated local map, its down-sampled form, the search grid,
// the down-sampled scan and the 6x6 state.  Per frame the host sends the raw scan and K keyframe
// descriptors (80 bytes each) and reads back one result struct.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/liloc_gpu.h"
#include "common.cuh"
#include "voxelgrid.cuh"
#include "knn_grid.cuh"
#include "s2m.cuh"
#include "ndt.cuh"

using namespace liloc;

namespace {

thread_local std::string g_create_error;

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;       // elements
};

struct VgInst {            // scratch of one VoxelGrid instance (the reference has three filters, :78-80)
    DevBuf<unsigned> keys_a, keys_b, vals_a, vals_b;
    DevBuf<unsigned char> cub_tmp;
    DevBuf<int> tile_counts;
    unsigned* d_enc = nullptr;     // 6 ordered-uint min/max
    VgParams* d_vp = nullptr;
    int* d_count = nullptr;
    int bits_hint = 32;            // radix-sort key bits expected for the next frame (verified after the fact)
    int bits_used = 32;
};

struct HostFrameInfo {     // pinned, written by async D2H
    VgParams map_vp;
    VgParams scan_vp;
    int M;
    int Q_all;
    S2mResult res;
};

}  // namespace

struct liloc_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    long long launches = 0;

    // keyframe arena
    DevBuf<float4> arena;
    size_t arena_used = 0;
    struct KfEntry { size_t off; int n; };
    std::unordered_map<int, KfEntry> kf;

    // local map
    DevBuf<float4> raw;            // concatenated, map frame
    DevBuf<float4> map;            // down-sampled (voxel order)
    int n_raw = 0;
    int M = -1;                    // host copy, -1 unknown
    bool have_map = false;
    bool have_grid = false;
    DevBuf<KfDesc> descs;
    KfDesc* h_descs = nullptr;     // pinned
    size_t h_descs_cap = 0;

    // scan
    DevBuf<float4> scan_raw;
    DevBuf<float4> scan;           // down-sampled
    int Q_all = -1;
    int n_scan_last = 0;
    const float4* scan_src_last = nullptr;
    float scan_leaf_last = 0.f, map_leaf_last = 0.f;
    long long sort_redos = 0;
    bool have_scan = false;

    // VoxelGrid instances: local map (main stream) and current scan (second stream, overlaps the map branch)
    VgInst vg_map, vg_scan;
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int* d_M = nullptr;            // = vg_map.d_count
    int* d_Q = nullptr;            // = vg_scan.d_count
    DevBuf<int> tile_counts;       // search-grid scan scratch

    // search grid
    GridParams gp{};
    DevBuf<int> cell_counts, cell_start, cell_cursor;
    DevBuf<float4> cpts;
    int max_cells = 1 << 24;

    // scan2map
    DevBuf<double> partial;
    float* d_pose = nullptr;       // 6
    S2mState* d_state = nullptr;
    S2mResult* d_result = nullptr;
    int s2m_max_blocks = 0;
    int s2m_nq_last = 0;

    // NDT: cached targets (prior submaps), sources (scans), job state
    struct NdtTargetHost { DevBuf<int> table; DevBuf<NdtLeaf> leaves; NdtTargetView view; int n_leaves; float resolution; int slot; };
    struct NdtSourceHost { DevBuf<float4> own; NdtSourceView view; int slot; };
    std::unordered_map<int, NdtTargetHost> ndt_targets;
    std::unordered_map<int, NdtSourceHost> ndt_sources;
    DevBuf<NdtTargetView> d_tviews;
    DevBuf<NdtSourceView> d_sviews;
    DevBuf<NdtJob> d_jobs;
    DevBuf<double> ndt_partial;
    int* d_ndt_active = nullptr;
    bool ndt_views_dirty = true;

    // kernel-level test scratch
    DevBuf<float4> tmp_pts;
    DevBuf<int> tmp_i;
    DevBuf<float> tmp_f;
    DevBuf<unsigned char> tmp_b;

    HostFrameInfo* h_info = nullptr;   // pinned

    // timing
    bool timing = false;
    cudaEvent_t ev[8]{};
    liloc_timing last_timing{};
    liloc_stats stats{};
    long long launches_at_reset = 0;
};

namespace {

int fail(liloc_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? LILOC_ERR_OOM : LILOC_ERR_CUDA,     \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define KCHECK()                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = cudaGetLastError();                                                       \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, LILOC_ERR_CUDA, "kernel launch failed: %s (%s:%d)",                   \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                               \
        ++ctx->launches;                                                                           \
    } while (0)

template <typename T>
int ensure(liloc_ctx* ctx, DevBuf<T>& b, size_t n, bool keep = false) {
    if (n <= b.cap) return 0;
    size_t cap = b.cap ? b.cap : 1024;
    while (cap < n) cap *= 2;
    T* np = nullptr;
    CU(cudaMalloc(&np, cap * sizeof(T)));
    if (b.p) {
        if (ctx->stream2) CU(cudaStreamSynchronize(ctx->stream2));
        if (keep) CU(cudaMemcpyAsync(np, b.p, b.cap * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaFree(b.p));
    }
    b.p = np; b.cap = cap;
    return 0;
}

inline int blocks_for(const liloc_ctx* ctx, long long n, int per_block, int waves = 8) {
    long long b = (n + per_block - 1) / per_block;
    const long long cap = (long long)ctx->num_sms * waves;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

__global__ void k_init_enc(unsigned* enc) {
    if (threadIdx.x < 3) enc[threadIdx.x] = 0xFFFFFFFFu;
    else if (threadIdx.x < 6) enc[threadIdx.x] = 0u;
}
__global__ void k_set_int(int* p, int v) { *p = v; }
__global__ void k_pose_to_T(const float* pose6, float* T) { if (threadIdx.x == 0) pose_to_T(pose6, T); }

// device-side evaluation of the small dense algebra, one system per thread (kernel-level parity tests)
__global__ void k_debug_lm6(const float* A, const float* b, int n, float* x, float* w, float* P, int* deg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float V[36];
    solve6_qr(A + 36 * i, b + 6 * i, x + 6 * i);
    eigen6_jacobi(A + 36 * i, w + 6 * i, V);
    deg[i] = degeneracy_projector(A + 36 * i, P + 36 * i);
}
__global__ void k_debug_plane(const float* pts, int n, float* x) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float P[5][3], X[3];
    for (int r = 0; r < 5; ++r) for (int c = 0; c < 3; ++c) P[r][c] = pts[15 * i + 3 * r + c];
    plane_lsq_5x3(P, X);
    x[3 * i] = X[0]; x[3 * i + 1] = X[1]; x[3 * i + 2] = X[2];
}

// VoxelGrid of a device cloud on stream `st`.  The bounding box must already be in vi.d_enc when have_minmax.
// `bits` = radix-sort key bits (32 = always safe; fewer = one 8-bit pass less, verified by the caller against
// VgParams.key_bits after the fact).
int vg_run(liloc_ctx* ctx, VgInst& vi, cudaStream_t st, const float4* d_pts, int n, float leaf, bool have_minmax,
           DevBuf<float4>& out, int bits, bool sort_only = false) {
    if (n <= 0) {
        k_set_int<<<1, 1, 0, st>>>(vi.d_count, 0); KCHECK();
        return 0;
    }
    int rc;
    if (!sort_only && (rc = ensure(ctx, out, (size_t)n))) return rc;
    if ((rc = ensure(ctx, vi.keys_a, (size_t)n))) return rc;
    if ((rc = ensure(ctx, vi.keys_b, (size_t)n))) return rc;
    if ((rc = ensure(ctx, vi.vals_a, (size_t)n))) return rc;
    if ((rc = ensure(ctx, vi.vals_b, (size_t)n))) return rc;
    const int n_tiles = (n + kRunTile - 1) / kRunTile;
    if ((rc = ensure(ctx, vi.tile_counts, (size_t)n_tiles + 1))) return rc;
    if (bits < 1) bits = 1;
    if (bits > 32) bits = 32;
    vi.bits_used = bits;
    size_t tmp_bytes = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, vi.keys_a.p, vi.keys_b.p, vi.vals_a.p, vi.vals_b.p, n, 0, 32, st));
    if ((rc = ensure(ctx, vi.cub_tmp, tmp_bytes))) return rc;

    if (!have_minmax) {
        k_init_enc<<<1, 32, 0, st>>>(vi.d_enc); KCHECK();
        k_minmax<<<blocks_for(ctx, n, kMinMaxThreads * 4), kMinMaxThreads, 0, st>>>(d_pts, n, vi.d_enc); KCHECK();
    }
    k_vg_params<<<1, 1, 0, st>>>(vi.d_enc, leaf, n, vi.d_vp); KCHECK();
    k_vg_keys<<<blocks_for(ctx, n, 256 * 4), 256, 0, st>>>(d_pts, n, vi.d_vp, vi.keys_a.p, vi.vals_a.p); KCHECK();
    size_t tb = vi.cub_tmp.cap;
    CU(cub::DeviceRadixSort::SortPairs(vi.cub_tmp.p, tb, vi.keys_a.p, vi.keys_b.p, vi.vals_a.p, vi.vals_b.p, n, 0, bits, st));
    if (sort_only) return 0;
    k_run_count<<<n_tiles, kRunThreads, 0, st>>>(vi.keys_b.p, n, vi.tile_counts.p); KCHECK();
    k_scan_tiles<<<1, 1024, 0, st>>>(vi.tile_counts.p, n_tiles, vi.d_count); KCHECK();
    k_run_centroids<<<n_tiles, kRunThreads, 0, st>>>(vi.keys_b.p, vi.vals_b.p, n, vi.tile_counts.p, d_pts, out.p); KCHECK();
    return 0;
}

inline int next_bits_hint(int key_bits) {       // one bit of head-room, rounded up to whole 8-bit radix passes
    int b = ((key_bits + 1 + 7) / 8) * 8;
    return b > 32 ? 32 : b;
}

// Search-grid geometry from the local map's bounding box: cell edge 2^k >= 1 m, one cell of padding.
GridParams grid_geometry(const float mn[3], const float mx[3], int max_cells) {
    GridParams g{};
    for (int k = 0; k < 24; ++k) {
        const float inv = 1.0f / (float)(1 << k);
        long long cells = 1;
        for (int a = 0; a < 3; ++a) {
            g.g0[a] = (int)floorf(mn[a] * inv) - 1;
            const int g1 = (int)floorf(mx[a] * inv) + 1;
            g.dim[a] = g1 - g.g0[a] + 1;
            cells *= g.dim[a];
        }
        g.inv_cell = inv;
        if (cells <= max_cells) { g.ncells = (int)cells; return g; }
    }
    g.ncells = -1;
    return g;
}

int grid_build(liloc_ctx* ctx, int M, const float mn[3], const float mx[3]) {
    int rc;
    ctx->have_grid = false;
    float lo[3] = {mn[0], mn[1], mn[2]}, hi[3] = {mx[0], mx[1], mx[2]};
    if (M <= 0) { for (int a = 0; a < 3; ++a) { lo[a] = 0.f; hi[a] = 0.f; } }
    ctx->gp = grid_geometry(lo, hi, ctx->max_cells);
    if (ctx->gp.ncells < 0) return fail(ctx, LILOC_ERR_CAPACITY, "local map extent does not fit the search grid");
    const int nc = ctx->gp.ncells;
    if ((rc = ensure(ctx, ctx->cell_counts, (size_t)nc))) return rc;
    if ((rc = ensure(ctx, ctx->cell_start, (size_t)nc + 1))) return rc;
    if ((rc = ensure(ctx, ctx->cell_cursor, (size_t)nc))) return rc;
    if ((rc = ensure(ctx, ctx->cpts, (size_t)(M > 0 ? M : 1)))) return rc;
    const int n_tiles = (nc + kScanTile - 1) / kScanTile;
    if ((rc = ensure(ctx, ctx->tile_counts, (size_t)n_tiles + 1))) return rc;
    k_zero_int<<<blocks_for(ctx, nc, 256 * 4), 256, 0, ctx->stream>>>(ctx->cell_counts.p, nc); KCHECK();
    if (M > 0) { k_grid_count<<<blocks_for(ctx, M, 256), 256, 0, ctx->stream>>>(ctx->map.p, M, ctx->gp, ctx->cell_counts.p); KCHECK(); }
    k_scan_reduce<<<n_tiles, kScanThreads, 0, ctx->stream>>>(ctx->cell_counts.p, nc, ctx->tile_counts.p); KCHECK();
    k_scan_tiles<<<1, 1024, 0, ctx->stream>>>(ctx->tile_counts.p, n_tiles, ctx->tile_counts.p + n_tiles); KCHECK();
    k_scan_down<<<n_tiles, kScanThreads, 0, ctx->stream>>>(ctx->cell_counts.p, nc, ctx->tile_counts.p, ctx->cell_start.p, ctx->cell_cursor.p); KCHECK();
    if (M > 0) { k_grid_scatter<<<blocks_for(ctx, M, 256), 256, 0, ctx->stream>>>(ctx->map.p, M, ctx->gp, ctx->cell_cursor.p, ctx->cpts.p); KCHECK(); }
    ctx->have_grid = true;
    return 0;
}

GridView grid_view(const liloc_ctx* ctx) {
    GridView g{ctx->cell_start.p, ctx->cpts.p, ctx->gp};
    return g;
}

void tick(liloc_ctx* ctx, int i) { if (ctx->timing) cudaEventRecord(ctx->ev[i], ctx->stream); }

// enqueue: keyframe descriptors -> transform + concat + bbox -> VoxelGrid; async D2H of {params, M}
int localmap_enqueue(liloc_ctx* ctx, const int* kf_ids, const float* poses6, int K, float leaf, int bits = 32) {
    if (K < 0 || (K > 0 && (!kf_ids || !poses6)) || !(leaf > 0.f)) return fail(ctx, LILOC_ERR_ARG, "localmap: bad arguments");
    int rc;
    ctx->have_map = false; ctx->have_grid = false; ctx->M = -1;
    if ((size_t)K > ctx->h_descs_cap) {
        if (ctx->h_descs) { CU(cudaStreamSynchronize(ctx->stream)); CU(cudaFreeHost(ctx->h_descs)); ctx->h_descs = nullptr; }
        size_t cap = 256; while (cap < (size_t)K) cap *= 2;
        CU(cudaMallocHost(&ctx->h_descs, cap * sizeof(KfDesc)));
        ctx->h_descs_cap = cap;
    } else if (ctx->h_descs) {
        CU(cudaStreamSynchronize(ctx->stream));      // previous async copy out of the pinned staging must be done
    }
    long long total = 0;
    for (int k = 0; k < K; ++k) {
        auto it = ctx->kf.find(kf_ids[k]);
        if (it == ctx->kf.end()) return fail(ctx, LILOC_ERR_ARG, "localmap: unknown keyframe id %d", kf_ids[k]);
        KfDesc& d = ctx->h_descs[k];
        d.src_off = (long long)it->second.off; d.n = it->second.n; d.dst_off = (int)total;
        pose_to_T(poses6 + 6 * (size_t)k, d.T);
        total += it->second.n;
        if (total > 0x7fffffffLL) return fail(ctx, LILOC_ERR_CAPACITY, "localmap: more than 2^31 points");
    }
    ctx->n_raw = (int)total;
    tick(ctx, 0);
    if (total > 0) {
        if ((rc = ensure(ctx, ctx->descs, (size_t)K))) return rc;
        if ((rc = ensure(ctx, ctx->raw, (size_t)total))) return rc;
        CU(cudaMemcpyAsync(ctx->descs.p, ctx->h_descs, (size_t)K * sizeof(KfDesc), cudaMemcpyHostToDevice, ctx->stream));
        ctx->stats.h2d_bytes += (long long)K * (long long)sizeof(KfDesc);
        k_init_enc<<<1, 32, 0, ctx->stream>>>(ctx->vg_map.d_enc); KCHECK();
        k_assemble<<<blocks_for(ctx, total, kAsmTile), kMinMaxThreads, 0, ctx->stream>>>(ctx->arena.p, ctx->descs.p, K, (int)total, ctx->raw.p, ctx->vg_map.d_enc); KCHECK();
    }
    tick(ctx, 1);
    ctx->map_leaf_last = leaf;
    if ((rc = vg_run(ctx, ctx->vg_map, ctx->stream, ctx->raw.p, (int)total, leaf, true, ctx->map, bits))) return rc;
    tick(ctx, 2);
    if (total > 0) CU(cudaMemcpyAsync(&ctx->h_info->map_vp, ctx->vg_map.d_vp, sizeof(VgParams), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(&ctx->h_info->M, ctx->d_M, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->stats.d2h_bytes += (long long)sizeof(VgParams) + 4;
    ctx->have_map = true;
    return 0;
}

// after a stream sync: pick up M and the bounding box, build the search grid
int localmap_finish(liloc_ctx* ctx) {
    ctx->M = ctx->h_info->M;
    return grid_build(ctx, ctx->M, ctx->h_info->map_vp.mn, ctx->h_info->map_vp.mx);
}

int scan_enqueue(liloc_ctx* ctx, const liloc_point* pts, int n, bool on_device, float leaf, cudaStream_t st = nullptr, int bits = 32) {
    if (!st) st = ctx->stream;
    if (n < 0 || (n > 0 && !pts) || !(leaf > 0.f)) return fail(ctx, LILOC_ERR_ARG, "scan_put: bad arguments");
    int rc;
    ctx->have_scan = false; ctx->Q_all = -1;
    const float4* src = reinterpret_cast<const float4*>(pts);
    if (!on_device && n > 0) {
        if ((rc = ensure(ctx, ctx->scan_raw, (size_t)n))) return rc;
        CU(cudaMemcpyAsync(ctx->scan_raw.p, pts, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, st));
        ctx->stats.h2d_bytes += (long long)n * (long long)sizeof(float4);
        src = ctx->scan_raw.p;
    }
    ctx->scan_src_last = src; ctx->scan_leaf_last = leaf;
    if ((rc = vg_run(ctx, ctx->vg_scan, st, src, n, leaf, false, ctx->scan, bits))) return rc;
    if (n > 0) CU(cudaMemcpyAsync(&ctx->h_info->scan_vp, ctx->vg_scan.d_vp, sizeof(VgParams), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&ctx->h_info->Q_all, ctx->d_Q, sizeof(int), cudaMemcpyDeviceToHost, st));
    ctx->stats.d2h_bytes += 4 + (long long)sizeof(VgParams);
    ctx->n_scan_last = n;
    ctx->have_scan = true;
    return 0;
}

int s2m_enqueue(liloc_ctx* ctx, const float* pose6, const liloc_s2m_opts* o) {
    if (!ctx->have_map || !ctx->have_grid) return fail(ctx, LILOC_ERR_STATE, "scan2map: no local map");
    if (!ctx->have_scan || ctx->Q_all < 0) return fail(ctx, LILOC_ERR_STATE, "scan2map: no scan");
    if (o->stride < 1 || o->max_iters < 1) return fail(ctx, LILOC_ERR_ARG, "scan2map: bad options");
    int rc;
    const int nq = (ctx->Q_all + o->stride - 1) / o->stride;
    int blocks = (nq + kQPB - 1) / kQPB;
    if (blocks > ctx->s2m_max_blocks) blocks = ctx->s2m_max_blocks;
    if (blocks < 1) blocks = 1;
    if ((rc = ensure(ctx, ctx->partial, (size_t)2 * blocks * kNSums))) return rc;
    CU(cudaMemcpyAsync(ctx->d_pose, pose6, 6 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += 24;
    ctx->s2m_nq_last = nq;
    S2mArgs a{};
    a.scan = ctx->scan.p; a.q_all = ctx->d_Q; a.map = ctx->map.p; a.map_count = ctx->d_M;
    a.grid = grid_view(ctx);
    a.stride = o->stride; a.max_iters = o->max_iters; a.min_sel = o->min_sel; a.min_points = o->min_points;
    a.partial = ctx->partial.p; a.pose_io = ctx->d_pose; a.state = ctx->d_state; a.result = ctx->d_result;
    void* args[] = {&a};
    CU(cudaLaunchCooperativeKernel((void*)k_scan2map, dim3(blocks), dim3(kS2mThreads), args, 0, ctx->stream));
    ++ctx->launches;
    CU(cudaMemcpyAsync(&ctx->h_info->res, ctx->d_result, sizeof(S2mResult), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->stats.d2h_bytes += (long long)sizeof(S2mResult);
    ++ctx->stats.s2m_launches;
    return 0;
}

int s2m_collect(liloc_ctx* ctx, float* pose6, liloc_s2m_result* res) {
    const S2mResult& r = ctx->h_info->res;
    static_assert(sizeof(liloc_s2m_result) == 8 * sizeof(int) + 72 * sizeof(float), "result layout");
    if (res) std::memcpy(res, &r, sizeof(liloc_s2m_result));
    if (!r.skipped) std::memcpy(pose6, r.pose, 6 * sizeof(float));
    ctx->stats.s2m_iters += r.iters;
    ctx->stats.alg_bytes_s2m += (double)r.iters * (96.0 * ctx->s2m_nq_last + 216.0);
    return r.skipped ? LILOC_SKIPPED : LILOC_OK;
}

const liloc_s2m_opts kDefaultOpts = {20, 2, 50, 30, {0, 0, 0, 0}};

}  // namespace

extern "C" {

const char* liloc_version(void) { return "liloc_b200 0.1 (sm_100a)"; }

const char* liloc_last_error(const liloc_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int liloc_create(const liloc_config* cfg, liloc_ctx** out) {
    liloc_ctx* ctx = nullptr;       // CU() reports into g_create_error while ctx == nullptr
    if (!cfg || !out) return fail(nullptr, LILOC_ERR_ARG, "liloc_create: null argument");
    *out = nullptr;
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, LILOC_ERR_CUDA, "liloc_create: no CUDA device %d (found %d)", cfg->device, ndev);
    CU(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major < 10) return fail(nullptr, LILOC_ERR_CUDA, "liloc_create: device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major, prop.minor);
    liloc_ctx* c = new liloc_ctx();
    c->device = cfg->device;
    c->num_sms = prop.multiProcessorCount;
    ctx = c;
    auto bail = [&](int rc) { g_create_error = c->err; liloc_destroy(c); return rc; };
#define CUB_(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fail(c, LILOC_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); return bail(LILOC_ERR_CUDA); } } while (0)
    CUB_(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUB_(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    CUB_(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CUB_(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    for (VgInst* vi : {&c->vg_map, &c->vg_scan}) {
        CUB_(cudaMalloc(&vi->d_enc, 6 * sizeof(unsigned)));
        CUB_(cudaMalloc(&vi->d_vp, sizeof(VgParams)));
        CUB_(cudaMalloc(&vi->d_count, sizeof(int)));
    }
    c->d_M = c->vg_map.d_count;
    c->d_Q = c->vg_scan.d_count;
    CUB_(cudaMalloc(&c->d_pose, 6 * sizeof(float)));
    CUB_(cudaMalloc(&c->d_state, sizeof(S2mState)));
    CUB_(cudaMalloc(&c->d_result, sizeof(S2mResult)));
    CUB_(cudaMemset(c->d_state, 0, sizeof(S2mState)));
    CUB_(cudaMemset(c->d_result, 0, sizeof(S2mResult)));
    CUB_(cudaMemset(c->d_M, 0, sizeof(int)));
    CUB_(cudaMemset(c->d_Q, 0, sizeof(int)));
    CUB_(cudaMalloc(&c->d_ndt_active, sizeof(int)));
    CUB_(cudaMallocHost(&c->h_info, sizeof(HostFrameInfo)));
    std::memset(c->h_info, 0, sizeof(HostFrameInfo));
    for (auto& e : c->ev) CUB_(cudaEventCreate(&e));
    int per_sm = 0;
    CUB_(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_scan2map, kS2mThreads, 0));
    if (per_sm < 1) { fail(c, LILOC_ERR_CUDA, "k_scan2map cannot be resident"); return bail(LILOC_ERR_CUDA); }
    c->s2m_max_blocks = per_sm * c->num_sms;
#undef CUB_
    const int scan_cap = cfg->max_scan_points > 0 ? cfg->max_scan_points : 16 * 1800;
    const int raw_cap = cfg->max_raw_points > 0 ? cfg->max_raw_points : (1 << 20);
    int rc;
    if ((rc = ensure(c, c->scan_raw, (size_t)scan_cap)) || (rc = ensure(c, c->scan, (size_t)scan_cap)) ||
        (rc = ensure(c, c->raw, (size_t)raw_cap)) || (rc = ensure(c, c->map, (size_t)raw_cap)) ||
        (rc = ensure(c, c->vg_map.keys_a, (size_t)raw_cap)) || (rc = ensure(c, c->vg_map.keys_b, (size_t)raw_cap)) ||
        (rc = ensure(c, c->vg_map.vals_a, (size_t)raw_cap)) || (rc = ensure(c, c->vg_map.vals_b, (size_t)raw_cap)) ||
        (rc = ensure(c, c->vg_scan.keys_a, (size_t)scan_cap)) || (rc = ensure(c, c->vg_scan.keys_b, (size_t)scan_cap)) ||
        (rc = ensure(c, c->vg_scan.vals_a, (size_t)scan_cap)) || (rc = ensure(c, c->vg_scan.vals_b, (size_t)scan_cap)) ||
        (rc = ensure(c, c->arena, (size_t)raw_cap)) || (rc = ensure(c, c->cpts, (size_t)raw_cap)) ||
        (rc = ensure(c, c->partial, (size_t)2 * c->s2m_max_blocks * kNSums)))
        return bail(rc);
    *out = c;
    return LILOC_OK;
}

void liloc_destroy(liloc_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    auto fr = [](void* p) { if (p) cudaFree(p); };
    fr(c->arena.p); fr(c->raw.p); fr(c->map.p); fr(c->descs.p); fr(c->scan_raw.p); fr(c->scan.p);
    for (VgInst* vi : {&c->vg_map, &c->vg_scan}) {
        fr(vi->keys_a.p); fr(vi->keys_b.p); fr(vi->vals_a.p); fr(vi->vals_b.p); fr(vi->cub_tmp.p); fr(vi->tile_counts.p);
        fr(vi->d_enc); fr(vi->d_vp); fr(vi->d_count);
    }
    fr(c->tile_counts.p);
    if (c->stream2) { cudaStreamSynchronize(c->stream2); cudaStreamDestroy(c->stream2); }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    fr(c->cell_counts.p); fr(c->cell_start.p); fr(c->cell_cursor.p); fr(c->cpts.p);
    fr(c->partial.p); fr(c->d_pose); fr(c->d_state); fr(c->d_result);
    fr(c->tmp_pts.p); fr(c->tmp_i.p); fr(c->tmp_f.p); fr(c->tmp_b.p);
    for (auto& kv : c->ndt_targets) { fr(kv.second.table.p); fr(kv.second.leaves.p); }
    for (auto& kv : c->ndt_sources) fr(kv.second.own.p);
    fr(c->d_tviews.p); fr(c->d_sviews.p); fr(c->d_jobs.p); fr(c->ndt_partial.p); fr(c->d_ndt_active);
    if (c->h_descs) cudaFreeHost(c->h_descs);
    if (c->h_info) cudaFreeHost(c->h_info);
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

#define ENTER(ctx)                                                                     \
    if (!(ctx)) return LILOC_ERR_ARG;                                                  \
    { cudaError_t e_ = cudaSetDevice((ctx)->device);                                   \
      if (e_ != cudaSuccess) return fail((ctx), LILOC_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e_)); }

int liloc_keyframe_put(liloc_ctx* ctx, int kf_id, const liloc_point* pts, int n) {
    ENTER(ctx);
    if (n < 0 || (n > 0 && !pts)) return fail(ctx, LILOC_ERR_ARG, "keyframe_put: bad arguments");
    int rc;
    if ((rc = ensure(ctx, ctx->arena, ctx->arena_used + (size_t)n, true))) return rc;
    if (n > 0) CU(cudaMemcpyAsync(ctx->arena.p + ctx->arena_used, pts, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    ctx->kf[kf_id] = liloc_ctx::KfEntry{ctx->arena_used, n};
    ctx->arena_used += (size_t)n;
    return LILOC_OK;
}

int liloc_keyframe_put_from_scan(liloc_ctx* ctx, int kf_id) {
    ENTER(ctx);
    if (!ctx->have_scan) return fail(ctx, LILOC_ERR_STATE, "keyframe_put_from_scan: no scan");
    if (ctx->Q_all < 0) { CU(cudaStreamSynchronize(ctx->stream)); ctx->Q_all = ctx->h_info->Q_all; }
    const int n = ctx->Q_all;
    int rc;
    if ((rc = ensure(ctx, ctx->arena, ctx->arena_used + (size_t)n, true))) return rc;
    if (n > 0) CU(cudaMemcpyAsync(ctx->arena.p + ctx->arena_used, ctx->scan.p, (size_t)n * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->kf[kf_id] = liloc_ctx::KfEntry{ctx->arena_used, n};
    ctx->arena_used += (size_t)n;
    return LILOC_OK;
}

int liloc_keyframe_size(liloc_ctx* ctx, int kf_id) {
    if (!ctx) return LILOC_ERR_ARG;
    auto it = ctx->kf.find(kf_id);
    return it == ctx->kf.end() ? LILOC_ERR_ARG : it->second.n;
}

int liloc_keyframe_clear(liloc_ctx* ctx) {
    ENTER(ctx);
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->kf.clear(); ctx->arena_used = 0;
    return LILOC_OK;
}

int liloc_localmap_build(liloc_ctx* ctx, const int* kf_ids, const float* poses6, int K, float leaf, int* M_out, int* n_raw_out) {
    ENTER(ctx);
    int rc;
    if ((rc = localmap_enqueue(ctx, kf_ids, poses6, K, leaf))) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    if ((rc = localmap_finish(ctx))) return rc;
    if (M_out) *M_out = ctx->M;
    if (n_raw_out) *n_raw_out = ctx->n_raw;
    return LILOC_OK;
}

int liloc_localmap_set(liloc_ctx* ctx, const liloc_point* pts, int n, float leaf, int* M_out) {
    ENTER(ctx);
    if (n < 0 || (n > 0 && !pts) || !(leaf > 0.f)) return fail(ctx, LILOC_ERR_ARG, "localmap_set: bad arguments");
    int rc;
    ctx->have_map = false; ctx->have_grid = false; ctx->M = -1;
    if ((rc = ensure(ctx, ctx->raw, (size_t)(n > 0 ? n : 1)))) return rc;
    if (n > 0) CU(cudaMemcpyAsync(ctx->raw.p, pts, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    ctx->n_raw = n;
    if ((rc = vg_run(ctx, ctx->vg_map, ctx->stream, ctx->raw.p, n, leaf, false, ctx->map, 32))) return rc;
    if (n > 0) CU(cudaMemcpyAsync(&ctx->h_info->map_vp, ctx->vg_map.d_vp, sizeof(VgParams), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(&ctx->h_info->M, ctx->d_M, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->have_map = true;
    if ((rc = localmap_finish(ctx))) return rc;
    if (M_out) *M_out = ctx->M;
    return LILOC_OK;
}

int liloc_localmap_get(liloc_ctx* ctx, liloc_point* out, int cap, int* M_out) {
    ENTER(ctx);
    if (!ctx->have_map || ctx->M < 0) return fail(ctx, LILOC_ERR_STATE, "localmap_get: no local map");
    if (M_out) *M_out = ctx->M;
    if (out) {
        if (cap < ctx->M) return fail(ctx, LILOC_ERR_CAPACITY, "localmap_get: cap %d < M %d", cap, ctx->M);
        if (ctx->M > 0) CU(cudaMemcpyAsync(out, ctx->map.p, (size_t)ctx->M * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return LILOC_OK;
}

int liloc_scan_put(liloc_ctx* ctx, const liloc_point* pts, int n, float leaf, int* Q_all_out) {
    ENTER(ctx);
    int rc;
    if ((rc = scan_enqueue(ctx, pts, n, false, leaf))) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->Q_all = ctx->h_info->Q_all;
    if (Q_all_out) *Q_all_out = ctx->Q_all;
    return LILOC_OK;
}

int liloc_scan_put_dev(liloc_ctx* ctx, const liloc_point* pts_dev, int n, float leaf, int* Q_all_out) {
    ENTER(ctx);
    int rc;
    if ((rc = scan_enqueue(ctx, pts_dev, n, true, leaf))) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->Q_all = ctx->h_info->Q_all;
    if (Q_all_out) *Q_all_out = ctx->Q_all;
    return LILOC_OK;
}

int liloc_scan_ds_get(liloc_ctx* ctx, liloc_point* out, int cap, int* Q_all_out) {
    ENTER(ctx);
    if (!ctx->have_scan || ctx->Q_all < 0) return fail(ctx, LILOC_ERR_STATE, "scan_ds_get: no scan");
    if (Q_all_out) *Q_all_out = ctx->Q_all;
    if (out) {
        if (cap < ctx->Q_all) return fail(ctx, LILOC_ERR_CAPACITY, "scan_ds_get: cap %d < Q_all %d", cap, ctx->Q_all);
        if (ctx->Q_all > 0) CU(cudaMemcpyAsync(out, ctx->scan.p, (size_t)ctx->Q_all * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return LILOC_OK;
}

int liloc_scan2map(liloc_ctx* ctx, float* pose6, const liloc_s2m_opts* opts, liloc_s2m_result* res) {
    ENTER(ctx);
    if (!pose6) return fail(ctx, LILOC_ERR_ARG, "scan2map: null pose");
    const liloc_s2m_opts* o = opts ? opts : &kDefaultOpts;
    int rc;
    tick(ctx, 4);
    if ((rc = s2m_enqueue(ctx, pose6, o))) return rc;
    tick(ctx, 5);
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->timing) { float ms = 0.f; cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]); ctx->last_timing.s2m_ms = ms; }
    return s2m_collect(ctx, pose6, res);
}

int liloc_frame(liloc_ctx* ctx, const int* kf_ids, const float* kf_poses6, int K, float map_leaf,
                const liloc_point* scan, int n_scan, int scan_on_device, float scan_leaf,
                float* pose6, const liloc_s2m_opts* opts, liloc_s2m_result* res) {
    ENTER(ctx);
    if (!pose6) return fail(ctx, LILOC_ERR_ARG, "frame: null pose");
    const liloc_s2m_opts* o = opts ? opts : &kDefaultOpts;
    int rc;
    // fork: the scan branch (H2D + VoxelGrid) runs on stream2 while the map branch runs on the main stream
    CU(cudaEventRecord(ctx->ev_fork, ctx->stream));
    CU(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
    if (ctx->timing) cudaEventRecord(ctx->ev[6], ctx->stream2);
    if ((rc = scan_enqueue(ctx, scan, n_scan, scan_on_device != 0, scan_leaf, ctx->stream2, ctx->vg_scan.bits_hint))) return rc;
    if (ctx->timing) cudaEventRecord(ctx->ev[7], ctx->stream2);
    CU(cudaEventRecord(ctx->ev_join, ctx->stream2));
    if ((rc = localmap_enqueue(ctx, kf_ids, kf_poses6, K, map_leaf, ctx->vg_map.bits_hint))) return rc;    // ev 0,1,2
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));                                              // join
    tick(ctx, 3);
    CU(cudaStreamSynchronize(ctx->stream));          // sync 1: M, bounding box, Q_all
    // The sorts used the previous frame's key width; if this frame needed more bits, redo with the full 32.
    if (ctx->n_raw > 0 && ctx->h_info->map_vp.key_bits > ctx->vg_map.bits_used) {
        ++ctx->sort_redos;
        if ((rc = vg_run(ctx, ctx->vg_map, ctx->stream, ctx->raw.p, ctx->n_raw, map_leaf, true, ctx->map, 32))) return rc;
        CU(cudaMemcpyAsync(&ctx->h_info->M, ctx->d_M, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    if (n_scan > 0 && ctx->h_info->scan_vp.key_bits > ctx->vg_scan.bits_used) {
        ++ctx->sort_redos;
        if ((rc = vg_run(ctx, ctx->vg_scan, ctx->stream, ctx->scan_src_last, n_scan, scan_leaf, true, ctx->scan, 32))) return rc;
        CU(cudaMemcpyAsync(&ctx->h_info->Q_all, ctx->d_Q, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    if (ctx->n_raw > 0) ctx->vg_map.bits_hint = next_bits_hint(ctx->h_info->map_vp.key_bits);
    if (n_scan > 0) ctx->vg_scan.bits_hint = next_bits_hint(ctx->h_info->scan_vp.key_bits);
    ctx->Q_all = ctx->h_info->Q_all;
    if ((rc = localmap_finish(ctx))) return rc;
    tick(ctx, 4);
    if ((rc = s2m_enqueue(ctx, pose6, o))) return rc;
    tick(ctx, 5);
    CU(cudaStreamSynchronize(ctx->stream));          // sync 2: result
    if (ctx->timing) {
        liloc_timing& t = ctx->last_timing;
        cudaEventElapsedTime(&t.assemble_ms, ctx->ev[0], ctx->ev[1]);
        cudaEventElapsedTime(&t.map_voxel_ms, ctx->ev[1], ctx->ev[2]);
        cudaEventElapsedTime(&t.scan_ms, ctx->ev[6], ctx->ev[7]);
        cudaEventElapsedTime(&t.grid_ms, ctx->ev[3], ctx->ev[4]);
        cudaEventElapsedTime(&t.s2m_ms, ctx->ev[4], ctx->ev[5]);
        cudaEventElapsedTime(&t.total_ms, ctx->ev[0], ctx->ev[5]);
        liloc_stats& st = ctx->stats;
        st.assemble_ms += t.assemble_ms; st.map_voxel_ms += t.map_voxel_ms; st.scan_ms += t.scan_ms;
        st.grid_ms += t.grid_ms; st.s2m_ms += t.s2m_ms; st.total_ms += t.total_ms;
    }
    ++ctx->stats.frames;
    ctx->stats.alg_bytes_map += 16.0 * ctx->n_raw + 16.0 * (ctx->M > 0 ? ctx->M : 0);
    ctx->stats.alg_bytes_scan += 16.0 * ctx->n_scan_last + 16.0 * (ctx->Q_all > 0 ? ctx->Q_all : 0);
    return s2m_collect(ctx, pose6, res);
}

int liloc_voxelgrid(liloc_ctx* ctx, const liloc_point* in, int n, float leaf, liloc_point* out, int cap, int* m_out) {
    ENTER(ctx);
    if (n < 0 || (n > 0 && (!in || !out)) || !(leaf > 0.f) || !m_out) return fail(ctx, LILOC_ERR_ARG, "voxelgrid: bad arguments");
    int rc;
    if ((rc = ensure(ctx, ctx->tmp_pts, (size_t)(n > 0 ? 2 * (size_t)n : 2)))) return rc;
    DevBuf<float4> outb; outb.p = ctx->tmp_pts.p + n; outb.cap = (size_t)(n > 0 ? n : 1);
    if (n > 0) CU(cudaMemcpyAsync(ctx->tmp_pts.p, in, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = ensure(ctx, ctx->tmp_i, 1))) return rc;
    {   // run on the scan instance's scratch but keep its result count (d_Q) untouched
        int* keep = ctx->vg_scan.d_count;
        ctx->vg_scan.d_count = ctx->tmp_i.p;
        rc = vg_run(ctx, ctx->vg_scan, ctx->stream, ctx->tmp_pts.p, n, leaf, false, outb, 32);
        ctx->vg_scan.d_count = keep;
        if (rc) return rc;
    }
    int m = 0;
    CU(cudaMemcpyAsync(&m, ctx->tmp_i.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *m_out = m;
    if (m > cap) return fail(ctx, LILOC_ERR_CAPACITY, "voxelgrid: cap %d < m %d", cap, m);
    if (m > 0) CU(cudaMemcpyAsync(out, outb.p, (size_t)m * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return LILOC_OK;
}

int liloc_knn5(liloc_ctx* ctx, const liloc_point* queries, int nq, int* idx5, float* d2_5, int* found) {
    ENTER(ctx);
    if (nq < 0 || (nq > 0 && (!queries || !idx5 || !d2_5 || !found))) return fail(ctx, LILOC_ERR_ARG, "knn5: bad arguments");
    if (!ctx->have_grid) return fail(ctx, LILOC_ERR_STATE, "knn5: no local map");
    if (nq == 0) return LILOC_OK;
    int rc;
    if ((rc = ensure(ctx, ctx->tmp_pts, (size_t)nq))) return rc;
    if ((rc = ensure(ctx, ctx->tmp_i, (size_t)nq * 6))) return rc;
    if ((rc = ensure(ctx, ctx->tmp_f, (size_t)nq * 5))) return rc;
    CU(cudaMemcpyAsync(ctx->tmp_pts.p, queries, (size_t)nq * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    k_knn5<<<blocks_for(ctx, (long long)nq * kKnnLanes, 256), 256, 0, ctx->stream>>>(ctx->tmp_pts.p, nq, grid_view(ctx), ctx->tmp_i.p, ctx->tmp_f.p, ctx->tmp_i.p + (size_t)nq * 5); KCHECK();
    CU(cudaMemcpyAsync(idx5, ctx->tmp_i.p, (size_t)nq * 5 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(found, ctx->tmp_i.p + (size_t)nq * 5, (size_t)nq * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(d2_5, ctx->tmp_f.p, (size_t)nq * 5 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return LILOC_OK;
}

int liloc_surf_pass(liloc_ctx* ctx, const float* pose6, int stride, liloc_point* coeff, unsigned char* flag) {
    ENTER(ctx);
    if (!pose6 || stride < 1 || !coeff || !flag) return fail(ctx, LILOC_ERR_ARG, "surf_pass: bad arguments");
    if (!ctx->have_grid) return fail(ctx, LILOC_ERR_STATE, "surf_pass: no local map");
    if (!ctx->have_scan || ctx->Q_all < 0) return fail(ctx, LILOC_ERR_STATE, "surf_pass: no scan");
    const int Q = ctx->Q_all;
    if (Q == 0) return LILOC_OK;
    int rc;
    if ((rc = ensure(ctx, ctx->tmp_pts, (size_t)Q))) return rc;
    if ((rc = ensure(ctx, ctx->tmp_b, (size_t)Q))) return rc;
    if ((rc = ensure(ctx, ctx->tmp_f, 6))) return rc;
    CU(cudaMemsetAsync(ctx->tmp_pts.p, 0, (size_t)Q * sizeof(float4), ctx->stream));
    CU(cudaMemsetAsync(ctx->tmp_b.p, 0, (size_t)Q, ctx->stream));
    CU(cudaMemcpyAsync(ctx->tmp_f.p, pose6, 6 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    S2mArgs a{};
    a.scan = ctx->scan.p; a.q_all = ctx->d_Q; a.map = ctx->map.p; a.map_count = ctx->d_M;
    a.grid = grid_view(ctx); a.stride = stride;
    const int nq = (Q + stride - 1) / stride;
    k_surf_pass<<<blocks_for(ctx, nq, kQPB), kS2mThreads, 0, ctx->stream>>>(a, ctx->tmp_f.p, ctx->tmp_pts.p, ctx->tmp_b.p); KCHECK();
    CU(cudaMemcpyAsync(coeff, ctx->tmp_pts.p, (size_t)Q * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(flag, ctx->tmp_b.p, (size_t)Q, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return LILOC_OK;
}

int liloc_pose_to_T(liloc_ctx* ctx, const float* pose6, float* T12) {
    ENTER(ctx);
    if (!pose6 || !T12) return fail(ctx, LILOC_ERR_ARG, "pose_to_T: null argument");
    int rc;
    if ((rc = ensure(ctx, ctx->tmp_f, 32))) return rc;
    CU(cudaMemcpyAsync(ctx->tmp_f.p, pose6, 6 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    k_pose_to_T<<<1, 32, 0, ctx->stream>>>(ctx->tmp_f.p, ctx->tmp_f.p + 8); KCHECK();
    CU(cudaMemcpyAsync(T12, ctx->tmp_f.p + 8, 12 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return LILOC_OK;
}

int liloc_debug_lm6(liloc_ctx* ctx, const float* A36, const float* b6, int n, float* x6, float* w6, float* P36, int* deg) {
    ENTER(ctx);
    if (n <= 0 || !A36 || !b6 || !x6 || !w6 || !P36 || !deg) return fail(ctx, LILOC_ERR_ARG, "debug_lm6: bad arguments");
    int rc;
    if ((rc = ensure(ctx, ctx->tmp_f, (size_t)n * (36 + 6 + 6 + 6 + 36)))) return rc;
    if ((rc = ensure(ctx, ctx->tmp_i, (size_t)n))) return rc;
    float* dA = ctx->tmp_f.p; float* db = dA + 36 * (size_t)n; float* dx = db + 6 * (size_t)n; float* dw = dx + 6 * (size_t)n; float* dP = dw + 6 * (size_t)n;
    CU(cudaMemcpyAsync(dA, A36, (size_t)n * 36 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(db, b6, (size_t)n * 6 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    k_debug_lm6<<<(n + 31) / 32, 32, 0, ctx->stream>>>(dA, db, n, dx, dw, dP, ctx->tmp_i.p); KCHECK();
    CU(cudaMemcpyAsync(x6, dx, (size_t)n * 6 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(w6, dw, (size_t)n * 6 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(P36, dP, (size_t)n * 36 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(deg, ctx->tmp_i.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return LILOC_OK;
}

int liloc_debug_plane(liloc_ctx* ctx, const float* pts15, int n, float* x3) {
    ENTER(ctx);
    if (n <= 0 || !pts15 || !x3) return fail(ctx, LILOC_ERR_ARG, "debug_plane: bad arguments");
    int rc;
    if ((rc = ensure(ctx, ctx->tmp_f, (size_t)n * 18))) return rc;
    CU(cudaMemcpyAsync(ctx->tmp_f.p, pts15, (size_t)n * 15 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    k_debug_plane<<<(n + 63) / 64, 64, 0, ctx->stream>>>(ctx->tmp_f.p, n, ctx->tmp_f.p + 15 * (size_t)n); KCHECK();
    CU(cudaMemcpyAsync(x3, ctx->tmp_f.p + 15 * (size_t)n, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return LILOC_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// NDT
// ---------------------------------------------------------------------------------------------------------
namespace {

NdtOptsDev ndt_opts_dev(const liloc_ndt_opts* o, float resolution) {
    NdtOptsDev d{};
    const double c1 = 10.0 * (1.0 - o->outlier_ratio), c2 = o->outlier_ratio / pow((double)resolution, 3);
    const double d3 = -log(c2);
    d.d1 = -log(c1 + c2) - d3;
    d.d2 = -2.0 * log((-log(c1 * exp(-0.5) + c2) - d3) / d.d1);
    d.step_max = o->step_size; d.step_min = o->trans_eps / 2.0; d.eps = o->trans_eps;
    d.max_iterations = o->max_iterations; d.search7 = o->search == LILOC_NDT_DIRECT7 ? 1 : 0;
    d.n_align = o->n_align > 0 ? o->n_align : 1;
    return d;
}

const liloc_ndt_opts kDefaultNdtOpts = {0.1, 0.55, 0.01, 35, LILOC_NDT_DIRECT1, 1, 0};

int ndt_sync_views(liloc_ctx* ctx) {
    if (!ctx->ndt_views_dirty) return 0;
    int rc;
    std::vector<NdtTargetView> tv(ctx->ndt_targets.size());
    std::vector<NdtSourceView> sv(ctx->ndt_sources.size());
    int i = 0;
    for (auto& kv : ctx->ndt_targets) { kv.second.slot = i; tv[i++] = kv.second.view; }
    i = 0;
    for (auto& kv : ctx->ndt_sources) { kv.second.slot = i; sv[i++] = kv.second.view; }
    if ((rc = ensure(ctx, ctx->d_tviews, tv.size() ? tv.size() : 1))) return rc;
    if ((rc = ensure(ctx, ctx->d_sviews, sv.size() ? sv.size() : 1))) return rc;
    if (!tv.empty()) CU(cudaMemcpyAsync(ctx->d_tviews.p, tv.data(), tv.size() * sizeof(NdtTargetView), cudaMemcpyHostToDevice, ctx->stream));
    if (!sv.empty()) CU(cudaMemcpyAsync(ctx->d_sviews.p, sv.data(), sv.size() * sizeof(NdtSourceView), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));      // the staging vectors die here
    ctx->ndt_views_dirty = false;
    return 0;
}

}  // namespace

extern "C" {

int liloc_ndt_target_put(liloc_ctx* ctx, int target_id, const liloc_point* pts, int n, float resolution, int* n_leaves_out) {
    ENTER(ctx);
    if (n <= 0 || !pts || !(resolution > 0.f)) return fail(ctx, LILOC_ERR_ARG, "ndt_target_put: bad arguments");
    int rc;
    if ((rc = ensure(ctx, ctx->tmp_pts, (size_t)n))) return rc;
    CU(cudaMemcpyAsync(ctx->tmp_pts.p, pts, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    DevBuf<float4> none;
    if ((rc = vg_run(ctx, ctx->vg_map, ctx->stream, ctx->tmp_pts.p, n, resolution, false, none, 32, true))) return rc;
    VgParams vp;
    CU(cudaMemcpyAsync(&vp, ctx->vg_map.d_vp, sizeof(VgParams), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (vp.overflow) return fail(ctx, LILOC_ERR_CAPACITY, "ndt_target_put: leaf size too small for the target extent (PCL overflow guard)");
    liloc_ctx::NdtTargetHost& T = ctx->ndt_targets[target_id];
    int divb[3];
    divb[0] = vp.mul[1]; divb[1] = vp.mul[1] ? vp.mul[2] / vp.mul[1] : 1;
    divb[2] = (int)floorf(vp.mx[2] * vp.inv_leaf) - vp.minb[2] + 1;
    const long long ncell = (long long)divb[0] * divb[1] * divb[2];
    if (ncell > (1LL << 28)) return fail(ctx, LILOC_ERR_CAPACITY, "ndt_target_put: %lld voxels exceed the dense table limit", ncell);
    if ((rc = ensure(ctx, T.table, (size_t)ncell))) return rc;
    if ((rc = ensure(ctx, T.leaves, (size_t)(n / 6 + 1)))) return rc;
    k_fill_int<<<blocks_for(ctx, ncell, 256 * 4), 256, 0, ctx->stream>>>(T.table.p, ncell, -1); KCHECK();
    CU(cudaMemsetAsync(ctx->d_ndt_active, 0, sizeof(int), ctx->stream));
    k_ndt_leaves<<<blocks_for(ctx, n, 256), 256, 0, ctx->stream>>>(ctx->vg_map.keys_b.p, ctx->vg_map.vals_b.p, n, ctx->tmp_pts.p, T.leaves.p, T.table.p, ctx->d_ndt_active); KCHECK();
    int nl = 0;
    CU(cudaMemcpyAsync(&nl, ctx->d_ndt_active, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    T.n_leaves = nl; T.resolution = resolution;
    T.view.table = T.table.p; T.view.leaves = T.leaves.p; T.view.inv_leaf = vp.inv_leaf;
    for (int a = 0; a < 3; ++a) { T.view.minb[a] = vp.minb[a]; T.view.divb[a] = divb[a]; T.view.maxb[a] = vp.minb[a] + divb[a] - 1; }
    ctx->ndt_views_dirty = true;
    if (n_leaves_out) *n_leaves_out = nl;
    return LILOC_OK;
}

int liloc_ndt_source_put(liloc_ctx* ctx, int source_id, const liloc_point* pts, int n, int on_device) {
    ENTER(ctx);
    if (n <= 0 || !pts) return fail(ctx, LILOC_ERR_ARG, "ndt_source_put: bad arguments");
    liloc_ctx::NdtSourceHost& S = ctx->ndt_sources[source_id];
    if (on_device) {
        S.view.pts = reinterpret_cast<const float4*>(pts);
    } else {
        int rc;
        if ((rc = ensure(ctx, S.own, (size_t)n))) return rc;
        CU(cudaMemcpyAsync(S.own.p, pts, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
        ctx->stats.h2d_bytes += (long long)n * 16;
        S.view.pts = S.own.p;
    }
    S.view.n = n;
    ctx->ndt_views_dirty = true;
    return LILOC_OK;
}

int liloc_ndt_clear(liloc_ctx* ctx) {
    ENTER(ctx);
    CU(cudaStreamSynchronize(ctx->stream));
    for (auto& kv : ctx->ndt_targets) { if (kv.second.table.p) cudaFree(kv.second.table.p); if (kv.second.leaves.p) cudaFree(kv.second.leaves.p); }
    for (auto& kv : ctx->ndt_sources) if (kv.second.own.p) cudaFree(kv.second.own.p);
    ctx->ndt_targets.clear(); ctx->ndt_sources.clear(); ctx->ndt_views_dirty = true;
    return LILOC_OK;
}

int liloc_ndt_align_batch(liloc_ctx* ctx, const liloc_ndt_job* jobs, int n_jobs, const liloc_ndt_opts* opts, liloc_ndt_out* outs) {
    ENTER(ctx);
    if (n_jobs <= 0 || !jobs || !outs) return fail(ctx, LILOC_ERR_ARG, "ndt_align_batch: bad arguments");
    const liloc_ndt_opts* o = opts ? opts : &kDefaultNdtOpts;
    int rc;
    if ((rc = ndt_sync_views(ctx))) return rc;
    std::vector<NdtJob> hj((size_t)n_jobs);
    int max_n = 0;
    float resolution = -1.f;
    for (int j = 0; j < n_jobs; ++j) {
        auto ti = ctx->ndt_targets.find(jobs[j].target_id);
        auto si = ctx->ndt_sources.find(jobs[j].source_id);
        if (ti == ctx->ndt_targets.end()) return fail(ctx, LILOC_ERR_ARG, "ndt_align_batch: unknown target id %d", jobs[j].target_id);
        if (si == ctx->ndt_sources.end()) return fail(ctx, LILOC_ERR_ARG, "ndt_align_batch: unknown source id %d", jobs[j].source_id);
        if (resolution < 0.f) resolution = ti->second.resolution;
        else if (resolution != ti->second.resolution) return fail(ctx, LILOC_ERR_ARG, "ndt_align_batch: all targets of one batch must share the resolution");
        std::memset(&hj[j], 0, sizeof(NdtJob));
        hj[j].target = ti->second.slot; hj[j].source = si->second.slot;
        std::memcpy(hj[j].guess, jobs[j].guess, 12 * sizeof(float));     // top 3 rows of the row-major 4x4
        if (si->second.view.n > max_n) max_n = si->second.view.n;
    }
    const NdtOptsDev od = ndt_opts_dev(o, resolution);
    const int tiles = (max_n + kNdtThreads - 1) / kNdtThreads;
    if ((rc = ensure(ctx, ctx->d_jobs, (size_t)n_jobs))) return rc;
    if ((rc = ensure(ctx, ctx->ndt_partial, (size_t)n_jobs * tiles * kNdtSums))) return rc;
    CU(cudaMemcpyAsync(ctx->d_jobs.p, hj.data(), (size_t)n_jobs * sizeof(NdtJob), cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += (long long)n_jobs * (long long)sizeof(NdtJob);
    k_ndt_begin<<<(n_jobs + 63) / 64, 64, 0, ctx->stream>>>(ctx->d_jobs.p, n_jobs); KCHECK();
    // worst case per align pass: first sweep + (max_iterations + 2) x (1 + 10) line-search sweeps
    const int max_sweeps = od.n_align * (1 + (od.max_iterations + 2) * 11);
    const int check_every = 4;
    int* h_active = &ctx->h_info->Q_all;                   // pinned scratch int
    int done_sweeps = 0;
    while (done_sweeps < max_sweeps) {
        for (int k = 0; k < check_every; ++k) {
            k_ndt_sweep<<<dim3(tiles, n_jobs), kNdtThreads, 0, ctx->stream>>>(ctx->d_jobs.p, ctx->d_tviews.p, ctx->d_sviews.p, od, tiles, ctx->ndt_partial.p); KCHECK();
            CU(cudaMemsetAsync(ctx->d_ndt_active, 0, sizeof(int), ctx->stream));
            k_ndt_update<<<n_jobs, 32, 0, ctx->stream>>>(ctx->d_jobs.p, ctx->d_sviews.p, od, tiles, ctx->ndt_partial.p, ctx->d_ndt_active); KCHECK();
            ++done_sweeps;
        }
        CU(cudaMemcpyAsync(h_active, ctx->d_ndt_active, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (*h_active == 0) break;
    }
    CU(cudaMemcpyAsync(hj.data(), ctx->d_jobs.p, (size_t)n_jobs * sizeof(NdtJob), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->stats.d2h_bytes += (long long)n_jobs * (long long)sizeof(NdtJob);
    for (int j = 0; j < n_jobs; ++j) {
        liloc_ndt_out& r = outs[j];
        std::memcpy(r.T, hj[j].Tfinal, 12 * sizeof(float));
        r.T[12] = 0.f; r.T[13] = 0.f; r.T[14] = 0.f; r.T[15] = 1.f;
        r.score = hj[j].score; r.iterations = hj[j].total_iterations; r.converged = hj[j].converged; r.sweeps = hj[j].total_sweeps;
        r.n_source = ctx->ndt_sources[jobs[j].source_id].view.n;
    }
    return LILOC_OK;
}

int liloc_ndt_target_get(liloc_ctx* ctx, int target_id, double* means, float* icovs, int* counts, int* keys, int cap) {
    ENTER(ctx);
    auto ti = ctx->ndt_targets.find(target_id);
    if (ti == ctx->ndt_targets.end()) return fail(ctx, LILOC_ERR_ARG, "ndt_target_get: unknown target id %d", target_id);
    const int n = ti->second.n_leaves;
    if (!means) return n;
    std::vector<NdtLeaf> h((size_t)(n > 0 ? n : 1));
    if (n > 0) CU(cudaMemcpyAsync(h.data(), ti->second.leaves.p, (size_t)n * sizeof(NdtLeaf), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n && i < cap; ++i) {
        for (int a = 0; a < 3; ++a) means[3 * i + a] = h[i].mean[a];
        for (int q = 0; q < 9; ++q) icovs[9 * i + q] = h[i].icov[q];
        counts[i] = h[i].n; keys[i] = h[i].pad;
    }
    return n;
}

int liloc_ndt_derivatives(liloc_ctx* ctx, int target_id, int source_id, const double* p6, const liloc_ndt_opts* opts, double* out28) {
    ENTER(ctx);
    if (!p6 || !out28) return fail(ctx, LILOC_ERR_ARG, "ndt_derivatives: null argument");
    const liloc_ndt_opts* o = opts ? opts : &kDefaultNdtOpts;
    int rc;
    if ((rc = ndt_sync_views(ctx))) return rc;
    auto ti = ctx->ndt_targets.find(target_id);
    auto si = ctx->ndt_sources.find(source_id);
    if (ti == ctx->ndt_targets.end() || si == ctx->ndt_sources.end()) return fail(ctx, LILOC_ERR_ARG, "ndt_derivatives: unknown target or source id");
    NdtJob J;
    std::memset(&J, 0, sizeof(J));
    J.target = ti->second.slot; J.source = si->second.slot; J.phase = NDT_AWAIT_FIRST;
    for (int i = 0; i < 6; ++i) { J.p[i] = p6[i]; J.x_t[i] = p6[i]; }
    const NdtOptsDev od = ndt_opts_dev(o, ti->second.resolution);
    const int n = si->second.view.n, tiles = (n + kNdtThreads - 1) / kNdtThreads;
    if ((rc = ensure(ctx, ctx->d_jobs, 1))) return rc;
    if ((rc = ensure(ctx, ctx->ndt_partial, (size_t)tiles * kNdtSums))) return rc;
    CU(cudaMemcpyAsync(ctx->d_jobs.p, &J, sizeof(NdtJob), cudaMemcpyHostToDevice, ctx->stream));
    k_ndt_sweep<<<dim3(tiles, 1), kNdtThreads, 0, ctx->stream>>>(ctx->d_jobs.p, ctx->d_tviews.p, ctx->d_sviews.p, od, tiles, ctx->ndt_partial.p); KCHECK();
    std::vector<double> part((size_t)tiles * kNdtSums);
    CU(cudaMemcpyAsync(part.data(), ctx->ndt_partial.p, part.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < kNdtSums; ++k) { double v = 0.0; for (int t = 0; t < tiles; ++t) v += part[(size_t)t * kNdtSums + k]; out28[k] = v; }
    return LILOC_OK;
}

}  // extern "C"

extern "C" {

int liloc_last_timing(liloc_ctx* ctx, liloc_timing* t) {
    if (!ctx || !t) return LILOC_ERR_ARG;
    *t = ctx->last_timing;
    return LILOC_OK;
}

int liloc_set_timing(liloc_ctx* ctx, int enabled) {
    if (!ctx) return LILOC_ERR_ARG;
    ctx->timing = enabled != 0;
    return LILOC_OK;
}

long long liloc_kernel_launches(const liloc_ctx* ctx) { return ctx ? ctx->launches : 0; }

int liloc_last_clocks(liloc_ctx* ctx, int* out, int max_iters) {
    if (!ctx || !out || max_iters < 0) return LILOC_ERR_ARG;
    const int n = max_iters < 32 ? max_iters : 32;
    std::memcpy(out, ctx->h_info->res.clk, (size_t)n * 8 * sizeof(int));
    return n;
}

int liloc_stats_get(const liloc_ctx* ctx, liloc_stats* out) {
    if (!ctx || !out) return LILOC_ERR_ARG;
    *out = ctx->stats;
    out->launches = ctx->launches - ctx->launches_at_reset;
    return LILOC_OK;
}

int liloc_stats_reset(liloc_ctx* ctx) {
    if (!ctx) return LILOC_ERR_ARG;
    ctx->stats = liloc_stats{};
    ctx->launches_at_reset = ctx->launches;
    return LILOC_OK;
}

void* liloc_stream(liloc_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int liloc_last_trace(liloc_ctx* ctx, float* out, int max_iters) {
    if (!ctx || !out || max_iters < 0) return LILOC_ERR_ARG;
    const int n = max_iters < 32 ? max_iters : 32;
    std::memcpy(out, ctx->h_info->res.trace, (size_t)n * 16 * sizeof(float));
    return n;
}

}  // extern "C"
