// liloc_b200/csrc/liloc_gpu.cu -- context, memory and the C ABI declared in include/liloc_gpu.h.
//
// One context = one GPU, one caller thread (mirrors `mtx`, src/mapOptmization.cpp:255).
// Everything the path touches lives in HBM for the lifetime of the context: the keyframe arena
// (body-frame float4 clouds), the concatenated local maps, their down-sampled form, the search grids,
// the down-sampled scans and the 6x6 state.  Per frame the host sends the raw scan(s) and K keyframe
// descriptors (64 bytes each) and reads back one result struct.
//
// A frame (liloc_frame / liloc_frame_features) is THREE kernel launches and one host synchronisation: the
// cooperative point-cloud pipeline (pipeline.cuh) once for the local maps (main stream) and once for the scans
// (second stream, so the scan's H2D copy overlaps the map work), then the cooperative registration kernel
// (s2m.cuh).  Every size that is only known on the device (M, Q_all, the search-grid geometry) is consumed on
// the device.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <climits>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/liloc_gpu.h"
#include "common.cuh"
#include "voxelgrid.cuh"
#include "knn_grid.cuh"
#include "pipeline.cuh"
#include "s2m.cuh"
#include "ndt.cuh"
#include "deskew.cuh"
#include "nccl_shim.h"

using namespace liloc;

namespace {

thread_local std::string g_create_error;

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;       // elements
};

struct VgInst {            // scratch of one VoxelGrid instance (the reference has three filters, :78-80)
    DevBuf<unsigned> keys_a, keys_b, vals_a, vals_b;      // after the sort: keys_b / vals_b hold the sorted pairs
    DevBuf<int> hist;              // 3 rotating block-histogram buffers of the radix sort
    DevBuf<unsigned long long> tile_state;   // look-back words of the centroid phase (pipeline.cuh)
    unsigned* d_enc = nullptr;     // 6 ordered-uint min/max (armed at creation, re-armed by the kernel)
    VgParams* d_vp = nullptr;
    int* d_count = nullptr;
};

constexpr int kClasses = 2;        // LILOC_SURF, LILOC_CORNER

// One feature class: its local map, search grid and current scan.  Class 0 (surf) is the reference's only
// class; class 1 (corner) is the upstream LIO-SAM superset the north_star names.
struct FeatClass {
    // local map
    DevBuf<float4> raw;            // concatenated, map frame
    DevBuf<float4> map;            // down-sampled (voxel order)
    DevBuf<float4> cpts;           // the same points sorted by search cell
    DevBuf<KfDesc> descs;
    KfDesc* h_descs = nullptr;     // pinned
    size_t h_descs_cap = 0;
    int n_raw = 0;
    int M = -1;                    // host copy, -1 unknown
    bool have_map = false;
    float map_leaf_last = 0.f;
    // search grid (geometry on the device)
    GridParams* d_gp = nullptr;
    DevBuf<int> cell_counts, cell_start;
    DevBuf<unsigned long long> grid_state;   // look-back words of the cell-count scan
    // scan
    DevBuf<float4> scan_raw;
    DevBuf<float4> scan;           // down-sampled
    int Q_all = -1;
    int q_prev = -1;               // Q_all of the previous frame (sizes the registration launch; never affects results)
    int n_scan_last = 0;
    const float4* scan_src_last = nullptr;
    float scan_leaf_last = 0.f;
    bool have_scan = false;
    VgInst vg_map, vg_scan;
    int* d_M = nullptr;            // = vg_map.d_count
    int* d_Q = nullptr;            // = vg_scan.d_count
};

struct HostClassInfo { VgParams map_vp, scan_vp; GridParams gp; int M, Q_all; };
struct HostFrameInfo {     // pinned, written by async D2H
    HostClassInfo c[kClasses];
    S2mResult res;
    unsigned long long marks[2 * kPipeMarks + kS2mMarks];
    int scratch;
    int dk_n[2];                         // de-skewed points of the two sweep slots
    unsigned long long dk_marks[3];      // de-skew launch: block 0 at start / after the barrier / end
    unsigned done_seq;                   // written last by the registration kernel: sequence number of the finished frame
    int pipe_error;                      // set by the point-cloud pipeline: 1 a look-back wait gave up, 2 a non-finite input point
};

}  // namespace

struct liloc_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;       // surf map branch, registration, everything step-wise
    cudaStream_t stream2 = nullptr;      // surf scan branch of a fused frame
    cudaStream_t stream3 = nullptr;      // corner branch of a fused frame
    cudaEvent_t ev_fork = nullptr, ev_join2 = nullptr, ev_join3 = nullptr;
    std::string err;
    long long launches = 0;
    int pipe_failed = 0;                 // error code of a pipeline launch that failed since the last check (0: none)
    unsigned pipe_epoch = 0;             // launch counter of the point-cloud pipeline (tags its look-back words)

    // keyframe arena
    DevBuf<float4> arena;
    size_t arena_used = 0;
    struct KfEntry {
        size_t off[kClasses]; int n[kClasses];
        float pose6[6]; float T[12]; bool have_T;     // the last pose this keyframe was placed with and its 3x4 transform (six double sin / cos per pose)
    };
    std::unordered_map<int, KfEntry> kf;

    FeatClass fc[kClasses];
    FeatClass aux;                       // scratch class of liloc_icp_fitness (its own cloud + search grid)
    DevBuf<float> aux_d2;
    int max_cells = 1 << 24;
    long long sort_redos = 0;
    bool wide_maps = true;               // local-map launch with two blocks per SM (LILOC_PIPE_NARROW=1 in the environment: one)
    long long wide_min = 100000;         // ... for local maps above this many raw points (LILOC_PIPE_WIDE_MIN in the environment)
    bool map_cache = false;              // liloc_config.map_cache
    long long map_cache_hits = 0;
    struct MapKey { std::vector<int> ids; std::vector<float> poses; float leaf = 0.f; bool valid = false; } map_key[kClasses];
    bool corner_scan_current = false;    // the last scan upload carried corner features too

    // scan2map
    DevBuf<double> partial;
    float* d_pose = nullptr;       // 6
    S2mState* d_state[2] = {nullptr, nullptr};   // double-buffered: a frame that has to be redone must not see its own update
    int state_cur = 0;
    S2mResult* d_result_host = nullptr;          // device alias of h_info->res (mapped pinned memory)
    unsigned long long* d_marks_host = nullptr;  // device alias of h_info->marks
    unsigned* d_done_host = nullptr;             // device alias of h_info->done_seq
    int* d_pipe_error_host = nullptr;            // device alias of h_info->pipe_error
    unsigned frame_seq = 0;                      // sequence number of the last enqueued registration
    bool poll_result = true;                     // wait for a frame by polling done_seq (LILOC_NO_POLL=1: cudaStreamSynchronize)
    int s2m_max_blocks = 0;
    int s2m_blocks_full = 0, pipe_blocks_full = 0, pipe_blocks_wide_full = 0;   // launch sizes of a context that has the GPU to itself
    int share = 1;                       // liloc_set_share: this context is one of `share` driving the GPU concurrently
    int pipe_blocks = 0;                 // blocks of a narrow point-cloud pipeline launch (one per SM)
    int pipe_blocks_wide = 0;            // blocks of a wide one (two per SM when three fit, so wide + narrow co-reside)
    int* d_sm_scratch = nullptr;             // 2 x kSmScratch ints, zero between launches (pipeline.cuh)
    unsigned long long* d_marks = nullptr;   // 2 x kPipeMarks phase timestamps (maps launch, scans launch) + kS2mMarks
    int s2m_nq_last = 0;
    int s2m_nq_corner_last = 0;

    // NDT: cached targets (prior submaps), sources (scans), job state
    struct NdtTargetHost { DevBuf<int> table; DevBuf<NdtLeaf> leaves; NdtTargetView view; int n_leaves; float resolution; int slot; int n_points; };
    DevBuf<KfDesc> ndt_descs;
    struct NdtSourceHost { DevBuf<float4> own; NdtSourceView view; int slot; };
    std::unordered_map<int, NdtTargetHost> ndt_targets;
    std::unordered_map<int, NdtSourceHost> ndt_sources;
    DevBuf<NdtTargetView> d_tviews;
    DevBuf<NdtSourceView> d_sviews;
    DevBuf<NdtJob> d_jobs;
    DevBuf<double> ndt_partial;
    int* d_ndt_active = nullptr;
    bool ndt_views_dirty = true;
    liloc_ndt_stats ndt_stats{};
    DevBuf<NdtHot> ndt_hot;
    DevBuf<NdtSync> ndt_sync;
    DevBuf<unsigned long long> ndt_q;    // [0..3] head / tail / finished / error words, then the ticket ring
    DevBuf<float> ndt_rows;              // packed results {pose6, score / n, iterations} of the last batch, 8 floats per job
    unsigned long long* h_ndt_hdr = nullptr;    // pinned copy of the queue header (counters + profile words) of the last batch
    NdtJob* h_ndt_jobs = nullptr;               // pinned staging of the job records (up and, when asked for, back)
    size_t h_ndt_jobs_cap = 0;
    int ndt_pending_jobs = 0, ndt_pending_blocks = 0, ndt_pending_cells = 1;
    int ndt_max_blocks = 0;
    int ndt_grain_div = 2;               // tickets aimed at per block (LILOC_NDT_GRAIN_DIV in the environment: diagnostics)

    // multi-GPU: one communicator per context (one context per GPU / process); the only collective is the all-gather of
    // packed result rows
    NcclShim::comm_t comm = nullptr;
    int comm_rank = 0, comm_world = 1;
    DevBuf<float> gather_send, gather_recv;
    float* h_gather = nullptr;           // pinned staging of the gathered rows
    size_t h_gather_cap = 0;             // floats

    // de-skew (imageProjection): raw sweep in, de-skewed cloud out.  Two output slots: the current sweep (what
    // LILOC_SCAN_DESKEWED refers to) and one staged ahead of it (LILOC_DESKEW_AHEAD), so that the next sweep can be uploaded
    // and de-skewed on stream3 while the current frame runs.
    struct DkSweep {
        DevBuf<float4> out;
        int* d_n = nullptr;              // device count
        int* d_n_host = nullptr;         // device alias of h_info->dk_n[slot]
        int n_upper = -1;                // raw points of the sweep (-1: empty slot)
        int n = -1;                      // host copy of the count, -1 while unknown
        cudaEvent_t ev_done = nullptr;   // recorded behind the de-skew launch
    } dk[2];
    int dk_cur = -1, dk_staged = -1;
    // liloc_scan_prefetch: the NEXT frame's surf scan (host memory) goes up on stream3 behind the current frame's launches,
    // into one of two buffers (the other may be the current frame's scan)
    struct ScanPrefetch {
        DevBuf<float4> buf[2];
        const liloc_point* hint = nullptr;   // hinted host cloud whose copy is not enqueued yet (nullptr: none)
        int hint_n = 0;
        const liloc_point* host = nullptr;   // host cloud whose copy is enqueued and not consumed yet (ev says when it is done)
        int n = 0;
        bool issued = false;
        int slot = 0;                        // ... into buf[slot]
        int in_use = -1;                     // buffer the frame in flight reads (-1: none)
        cudaEvent_t ev = nullptr;
    } pf;
    DevBuf<float4> dk_pts;
    DevBuf<int2> dk_rt;
    DevBuf<int> dk_tiles;
    double* d_dk_imu = nullptr;          // 4 x kImuMax
    double* h_dk_imu = nullptr;          // pinned staging of the same, two slots
    int dk_slot = 0;
    cudaEvent_t ev_dk_slot[2]{};
    cudaEvent_t ev_dk_fork2 = nullptr;
    unsigned long long* d_dk_marks = nullptr;   // device alias of h_info->dk_marks (timing)
    int dk_blocks_per_sm = 1;
    float dk_ms = 0.f;

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
        if (ctx->stream3) CU(cudaStreamSynchronize(ctx->stream3));
        if (keep) CU(cudaMemcpyAsync(np, b.p, b.cap * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaFree(b.p));
    }
    b.p = np; b.cap = cap;
    return 0;
}

// look-back words (pipeline.cuh) are tagged with the launch counter, which starts at 1: a fresh buffer must not hold a
// stale tag by accident, so it is cleared once when it is (re)allocated
template <typename T>
int ensure_zeroed(liloc_ctx* ctx, DevBuf<T>& b, size_t n) {
    const size_t before = b.cap;
    int rc = ensure(ctx, b, n);
    if (rc) return rc;
    if (b.cap != before) {
        CU(cudaMemsetAsync(b.p, 0, b.cap * sizeof(T), ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

inline int blocks_for(const liloc_ctx* ctx, long long n, int per_block, int waves = 8) {
    long long b = (n + per_block - 1) / per_block;
    const long long cap = (long long)ctx->num_sms * waves;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

inline bool bad_class(int cls) { return cls < 0 || cls >= kClasses; }

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
__global__ void k_debug_eigen3(const float* A, int n, float* w, float* V) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    eigen3_jacobi(A + 9 * i, w + 3 * i, V + 9 * i);
}
// cornerOptimization's loop body on explicit neighbours: pts15 = 5 x 3 neighbours, sel = pointSel
__global__ void k_debug_line(const float* pts, const float* sel3, int n, float4* coeff, int* keep) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float P[5][3];
    for (int r = 0; r < 5; ++r) for (int c = 0; c < 3; ++c) P[r][c] = pts[15 * i + 3 * r + c];
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 sel = make_float4(sel3[3 * i], sel3[3 * i + 1], sel3[3 * i + 2], 0.f);
    keep[i] = corner_residual(P, 0.f, sel, &c) ? 1 : 0;
    coeff[i] = c;
}

// ---- problems of the cooperative point-cloud pipeline (pipeline.cuh) ------------------------------------------
// Scratch and output of one VoxelGrid over n points.
int prob_voxelgrid(liloc_ctx* ctx, VgInst& vi, const float4* d_pts, int n, float leaf, DevBuf<float4>* out, VgProb* q) {
    int rc;
    const size_t cap = (size_t)(n > 0 ? n : 1);
    if (out && (rc = ensure(ctx, *out, cap))) return rc;
    if ((rc = ensure(ctx, vi.keys_a, cap))) return rc;
    if ((rc = ensure(ctx, vi.keys_b, cap))) return rc;
    if ((rc = ensure(ctx, vi.vals_a, cap))) return rc;
    if ((rc = ensure(ctx, vi.vals_b, cap))) return rc;
    if ((rc = ensure(ctx, vi.hist, (size_t)kHistInts))) return rc;
    if ((rc = ensure_zeroed(ctx, vi.tile_state, (size_t)(n / kRunTile) + 2))) return rc;
    *q = VgProb{};
    q->pts = d_pts; q->n = n; q->leaf = leaf;
    q->keys[0] = vi.keys_a.p; q->keys[1] = vi.keys_b.p; q->vals[0] = vi.vals_a.p; q->vals[1] = vi.vals_b.p;
    q->hist = vi.hist.p; q->tile_state = vi.tile_state.p; q->enc = vi.d_enc; q->vp = vi.d_vp;
    q->out = out ? out->p : nullptr; q->count = vi.d_count;
    q->sort_only = out ? 0 : 1;
    return 0;
}

// Search structure over the class's down-sampled map (replaces kdtreeSurfFromMap->setInputCloud, :1188).
int prob_add_grid(liloc_ctx* ctx, FeatClass& fc, int m_upper, VgProb* q) {
    int rc;
    const int nc_max = ctx->max_cells;
    if (!fc.cell_counts.p) {                          // the kernel keeps the counters zero outside the last grid
        if ((rc = ensure(ctx, fc.cell_counts, (size_t)nc_max))) return rc;
        CU(cudaMemsetAsync(fc.cell_counts.p, 0, fc.cell_counts.cap * sizeof(int), ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    if ((rc = ensure(ctx, fc.cell_start, (size_t)nc_max + 1))) return rc;
    if ((rc = ensure_zeroed(ctx, fc.grid_state, (size_t)(nc_max / kScanTile) + 2))) return rc;
    if ((rc = ensure(ctx, fc.cpts, (size_t)(m_upper > 0 ? m_upper : 1)))) return rc;
    q->build_grid = 1; q->max_cells = nc_max;
    q->gp = fc.d_gp; q->cell_counts = fc.cell_counts.p; q->cell_start = fc.cell_start.p; q->grid_state = fc.grid_state.p; q->cpts = fc.cpts.p;
    return 0;
}

// `wide`: two blocks per SM (local maps above ~1.5 M raw points, whose streaming phases want the extra loads in flight;
// the concurrent scan launch then waits for a free slot); otherwise one block per SM, so that the map and the scan
// launch are co-resident (2 blocks of 96 registers x 256 threads per SM).
int pipe_launch(liloc_ctx* ctx, cudaStream_t st, PipeArgs& a, unsigned long long* marks, bool wide = false) {
    if (a.n_prob <= 0) return 0;
    a.marks = marks;
    a.epoch = ++ctx->pipe_epoch;
    a.error_flag = ctx->d_pipe_error_host;
    a.sm_scratch = ctx->d_sm_scratch + (st == ctx->stream2 ? kSmScratch : 0);     // the two launches of a frame overlap
    void* args[] = {&a};
    const int blocks = wide ? ctx->pipe_blocks_wide : ctx->pipe_blocks;
    CU(cudaLaunchCooperativeKernel((void*)k_voxel_pipeline, dim3((unsigned)blocks), dim3(kPipeThreads), args, 0, st));
    ++ctx->launches;
    return 0;
}

GridView grid_view(const FeatClass& fc) {
    GridView g{fc.cell_start.p, fc.cpts.p, fc.d_gp};
    return g;
}

void tick(liloc_ctx* ctx, int i, cudaStream_t st) { if (ctx->timing) cudaEventRecord(ctx->ev[i], st); }

// Local-map problem of class `cls`: keyframe descriptors (H2D on `st`) -> transform + concat + bbox -> VoxelGrid ->
// search grid.  Only fills *q; the caller launches the pipeline on the same stream.
int localmap_problem(liloc_ctx* ctx, int cls, cudaStream_t st, const int* kf_ids, const float* poses6, int K, float leaf, VgProb* q,
                     PipeArgs* pa = nullptr) {
    FeatClass& fc = ctx->fc[cls];
    if (K < 0 || (K > 0 && (!kf_ids || !poses6)) || !(leaf > 0.f)) return fail(ctx, LILOC_ERR_ARG, "localmap: bad arguments");
    int rc;
    fc.have_map = false; fc.M = -1;
    ctx->map_key[cls].valid = false;
    if ((size_t)K > fc.h_descs_cap) {
        if (fc.h_descs) { CU(cudaFreeHost(fc.h_descs)); fc.h_descs = nullptr; }
        size_t cap = 256; while (cap < (size_t)K) cap *= 2;
        CU(cudaMallocHost(&fc.h_descs, cap * sizeof(KfDesc)));
        fc.h_descs_cap = cap;
    }
    long long total = 0;
    for (int k = 0; k < K; ++k) {
        auto it = ctx->kf.find(kf_ids[k]);
        if (it == ctx->kf.end()) return fail(ctx, LILOC_ERR_ARG, "localmap: unknown keyframe id %d", kf_ids[k]);
        KfDesc& d = fc.h_descs[k];
        d.src_off = (long long)it->second.off[cls]; d.n = it->second.n[cls]; d.dst_off = (int)total;
        liloc_ctx::KfEntry& e = it->second;
        if (!e.have_T || std::memcmp(e.pose6, poses6 + 6 * (size_t)k, sizeof(e.pose6)) != 0) {      // key poses rarely move
            std::memcpy(e.pose6, poses6 + 6 * (size_t)k, sizeof(e.pose6));
            pose_to_T(e.pose6, e.T);
            e.have_T = true;
        }
        std::memcpy(d.T, e.T, sizeof(d.T));
        total += it->second.n[cls];
        if (total > 0x7fffffffLL) return fail(ctx, LILOC_ERR_CAPACITY, "localmap: more than 2^31 points");
    }
    fc.n_raw = (int)total;
    if ((rc = ensure(ctx, fc.descs, (size_t)(K > 0 ? K : 1)))) return rc;
    if ((rc = ensure(ctx, fc.raw, (size_t)(total > 0 ? total : 1)))) return rc;
    int inline1 = 0;
    if (K > 0) {
        if (pa && pa->n_inline + K <= kInlineDescs) {          // descriptors as kernel parameters (pipeline.cuh)
            std::memcpy(pa->inline_descs + pa->n_inline, fc.h_descs, (size_t)K * sizeof(KfDesc));
            inline1 = pa->n_inline + 1;
            pa->n_inline += K;
        } else {
            CU(cudaMemcpyAsync(fc.descs.p, fc.h_descs, (size_t)K * sizeof(KfDesc), cudaMemcpyHostToDevice, st));
        }
        ctx->stats.h2d_bytes += (long long)K * (long long)sizeof(KfDesc);
    }
    fc.map_leaf_last = leaf;
    if ((rc = prob_voxelgrid(ctx, fc.vg_map, fc.raw.p, (int)total, leaf, &fc.map, q))) return rc;
    q->arena = ctx->arena.p; q->descs = fc.descs.p; q->inline1 = inline1; q->K = total > 0 ? K : 0; q->raw = fc.raw.p;
    if ((rc = prob_add_grid(ctx, fc, (int)total, q))) return rc;
    fc.have_map = true;
    return 0;
}

// Scan problem of class `cls`: H2D on `st` (LILOC_SCAN_HOST), a caller-owned device cloud (LILOC_SCAN_DEVICE) or the
// de-skewed sweep this context holds (LILOC_SCAN_DESKEWED: its size is read from device memory) -> VoxelGrid.
int scan_problem(liloc_ctx* ctx, int cls, cudaStream_t st, const liloc_point* pts, int n, int mode, float leaf, VgProb* q) {
    FeatClass& fc = ctx->fc[cls];
    const int* n_dev = nullptr;
    if (mode == LILOC_SCAN_DESKEWED) {
        if (cls != LILOC_SURF) return fail(ctx, LILOC_ERR_ARG, "scan_put: the de-skewed sweep is the surf-class scan");
        if (ctx->dk_cur < 0) return fail(ctx, LILOC_ERR_STATE, "scan_put: no de-skewed sweep (call liloc_deskew first)");
        liloc_ctx::DkSweep& S = ctx->dk[ctx->dk_cur];
        pts = reinterpret_cast<const liloc_point*>(S.out.p);
        n = S.n >= 0 ? S.n : S.n_upper;
        if (S.n < 0) n_dev = S.d_n;
        CU(cudaStreamWaitEvent(st, S.ev_done, 0));
    } else if (mode != LILOC_SCAN_HOST && mode != LILOC_SCAN_DEVICE) {
        return fail(ctx, LILOC_ERR_ARG, "scan_put: bad scan_on_device %d", mode);
    }
    if (n < 0 || (n > 0 && !pts) || !(leaf > 0.f)) return fail(ctx, LILOC_ERR_ARG, "scan_put: bad arguments");
    int rc;
    fc.have_scan = false; fc.Q_all = -1;
    const float4* src = reinterpret_cast<const float4*>(pts);
    if (mode == LILOC_SCAN_HOST && n > 0) {
        liloc_ctx::ScanPrefetch& pf = ctx->pf;
        if (cls == LILOC_SURF && pf.issued && pf.host == pts && pf.n == n) {
            // uploaded ahead by liloc_scan_prefetch (bytes counted there): wait for that copy instead of making one
            CU(cudaStreamWaitEvent(st, pf.ev, 0));
            src = pf.buf[pf.slot].p;
            pf.in_use = pf.slot; pf.issued = false;
        } else {
            if (cls == LILOC_SURF) pf.in_use = -1;
            if ((rc = ensure(ctx, fc.scan_raw, (size_t)n))) return rc;
            CU(cudaMemcpyAsync(fc.scan_raw.p, pts, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, st));
            ctx->stats.h2d_bytes += (long long)n * (long long)sizeof(float4);
            src = fc.scan_raw.p;
        }
    }
    fc.scan_src_last = src; fc.scan_leaf_last = leaf;
    if ((rc = prob_voxelgrid(ctx, fc.vg_scan, src, n, leaf, &fc.scan, q))) return rc;
    q->n_dev = n_dev;
    fc.n_scan_last = n;
    fc.have_scan = true;
    return 0;
}

// async D2H of the device-side counts of a class (picked up by class_collect after a synchronisation)
int class_fetch(liloc_ctx* ctx, int cls, cudaStream_t st, bool map, bool scan) {
    FeatClass& fc = ctx->fc[cls];
    HostClassInfo& hi = ctx->h_info->c[cls];
    if (map) { CU(cudaMemcpyAsync(&hi.M, fc.d_M, sizeof(int), cudaMemcpyDeviceToHost, st)); ctx->stats.d2h_bytes += 4; }
    if (scan) { CU(cudaMemcpyAsync(&hi.Q_all, fc.d_Q, sizeof(int), cudaMemcpyDeviceToHost, st)); ctx->stats.d2h_bytes += 4; }
    return 0;
}

const liloc_s2m_opts kDefaultOpts = {20, 2, 50, 30, 0, 10, 1, {0}};

FeatView feat_view(const FeatClass& fc, int stride) {
    FeatView v{};
    v.scan = fc.scan.p; v.q_all = fc.d_Q; v.map = fc.map.p; v.map_count = fc.d_M;
    v.grid = grid_view(fc); v.stride = stride < 1 ? 1 : stride;
    return v;
}

int s2m_args(liloc_ctx* ctx, const liloc_s2m_opts* o, S2mArgs* a) {
    const bool corner = o->enable_corner != 0;
    for (int cls = 0; cls < (corner ? 2 : 1); ++cls) {
        if (!ctx->fc[cls].have_map) return fail(ctx, LILOC_ERR_STATE, "scan2map: no %s local map", cls ? "corner" : "surf");
        if (!ctx->fc[cls].have_scan) return fail(ctx, LILOC_ERR_STATE, "scan2map: no %s scan", cls ? "corner" : "surf");
    }
    if (o->stride < 1 || o->max_iters < 1) return fail(ctx, LILOC_ERR_ARG, "scan2map: bad options");
    *a = S2mArgs{};
    a->f[0] = feat_view(ctx->fc[0], o->stride);
    a->f[1] = feat_view(ctx->fc[1], o->stride_corner);
    a->enable_corner = corner ? 1 : 0;
    a->max_iters = o->max_iters; a->min_sel = o->min_sel; a->min_points = o->min_points; a->min_corner = o->min_points_corner;
    return 0;
}

// Enqueue the registration kernel on the main stream.  Q_all is only known on the device here, so the launch is
// sized from host-side upper bounds (the raw scan sizes); idle blocks only take part in the grid barrier.
int s2m_enqueue(liloc_ctx* ctx, const float* pose6, const liloc_s2m_opts* o) {
    int rc;
    S2mArgs a;
    if ((rc = s2m_args(ctx, o, &a))) return rc;
    auto upper = [](const FeatClass& fc) {
        if (fc.Q_all >= 0) return fc.Q_all;
        const int guess = fc.q_prev > 0 ? fc.q_prev + fc.q_prev / 4 + kQPB : fc.n_scan_last;
        return guess < fc.n_scan_last ? guess : fc.n_scan_last;
    };
    long long nq = (upper(ctx->fc[0]) + a.f[0].stride - 1) / a.f[0].stride;
    if (a.enable_corner) nq += (upper(ctx->fc[1]) + a.f[1].stride - 1) / a.f[1].stride + kQPB;
    long long blocks = (nq + kQPB - 1) / kQPB;
    if (blocks > ctx->s2m_max_blocks) blocks = ctx->s2m_max_blocks;
    if (blocks < 1) blocks = 1;
    if ((rc = ensure(ctx, ctx->partial, (size_t)2 * blocks * kNSums))) return rc;
    // the initial guess is a kernel parameter: a 24-byte H2D copy between the map launch and this one would put a
    // copy-engine hop (several microseconds) on the frame's critical path
    for (int i = 0; i < 6; ++i) a.pose_in[i] = pose6[i];
    ctx->stats.h2d_bytes += 24;
    a.partial = ctx->partial.p; a.pose_io = ctx->d_pose;
    a.state = ctx->d_state[ctx->state_cur]; a.state_out = ctx->d_state[ctx->state_cur ^ 1]; a.result = ctx->d_result_host;
    a.marks = ctx->timing ? ctx->d_marks_host + 2 * kPipeMarks : nullptr;
    a.done_flag = ctx->d_done_host; a.done_value = ++ctx->frame_seq;
    void* args[] = {&a};
    CU(cudaLaunchCooperativeKernel((void*)k_scan2map, dim3((unsigned)blocks), dim3(kS2mThreads), args, 0, ctx->stream));
    ++ctx->launches;
    ctx->stats.d2h_bytes += (long long)sizeof(S2mResult);        // written by the kernel into mapped host memory
    ++ctx->stats.s2m_launches;
    return 0;
}

// after the stream sync: commit the 6x6 state, hand out the result
int s2m_collect(liloc_ctx* ctx, float* pose6, liloc_s2m_result* res, const liloc_s2m_opts* o) {
    const S2mResult& r = ctx->h_info->res;
    static_assert(sizeof(liloc_s2m_result) == 8 * sizeof(int) + 72 * sizeof(float), "result layout");
    ctx->state_cur ^= 1;
    if (res) std::memcpy(res, &r, sizeof(liloc_s2m_result));
    if (!r.skipped) std::memcpy(pose6, r.pose, 6 * sizeof(float));
    const int st0 = o->stride < 1 ? 1 : o->stride, st1 = o->stride_corner < 1 ? 1 : o->stride_corner;
    ctx->s2m_nq_last = (r.q_all + st0 - 1) / st0;
    ctx->s2m_nq_corner_last = o->enable_corner ? (r.extra[0] + st1 - 1) / st1 : 0;
    ctx->stats.s2m_iters += r.iters;
    ctx->stats.alg_bytes_s2m += (double)r.iters * (96.0 * (ctx->s2m_nq_last + ctx->s2m_nq_corner_last) + 216.0);
    return r.skipped ? LILOC_SKIPPED : LILOC_OK;
}

// After a synchronisation behind a pipeline launch: did the launch report a failure (mapped host word)?  A failed launch
// leaves no usable map / scan / memoised key behind: a later liloc_scan2map must not run on half-made state.
int pipe_check(liloc_ctx* ctx) {
    const int e = ctx->h_info->pipe_error;
    if (!e) return 0;
    ctx->h_info->pipe_error = 0;
    if (e == 2) { ctx->err = "a point with a non-finite coordinate in an input cloud (all input coordinates must be finite)"; ctx->pipe_failed = LILOC_ERR_ARG; }
    else { ctx->err = "point-cloud pipeline: a look-back wait gave up (results of this call are not valid)"; ctx->pipe_failed = LILOC_ERR_CUDA; }
    for (FeatClass& fc : ctx->fc) { fc.have_map = false; fc.have_scan = false; fc.M = -1; fc.Q_all = -1; }
    for (auto& k : ctx->map_key) k.valid = false;
    if (e != 2) {      // a launch that gave up half-way may have left search-cell counters behind (the kernel counts them back to zero itself)
        for (FeatClass& fc : ctx->fc) if (fc.cell_counts.p) cudaMemsetAsync(fc.cell_counts.p, 0, fc.cell_counts.cap * sizeof(int), ctx->stream);
        if (ctx->aux.cell_counts.p) cudaMemsetAsync(ctx->aux.cell_counts.p, 0, ctx->aux.cell_counts.cap * sizeof(int), ctx->stream);
    }
    return ctx->pipe_failed;
}

// pick up the device-side counts after a synchronisation
void class_collect(liloc_ctx* ctx, int cls, bool map, bool scan) {
    FeatClass& fc = ctx->fc[cls];
    pipe_check(ctx);
    if (map) fc.M = ctx->h_info->c[cls].M;
    if (scan) { fc.Q_all = ctx->h_info->c[cls].Q_all; fc.q_prev = fc.Q_all; }
}

}  // namespace

extern "C" {

const char* liloc_version(void) { return "liloc_b200 0.2 (sm_100a)"; }

const char* liloc_last_error(const liloc_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

// S contexts sharing one GPU (concurrent sequences, one host thread each): the cooperative launches of one frame fill the
// co-residency capacity of the device on their own (2 + 1 pipeline blocks per SM, 2 registration blocks per SM), so launches
// of the other contexts queue behind them; with 1 / S of the blocks each, the launches of S frames are resident together.
// Results do not depend on the launch size (every sum has a fixed order).
static void apply_share(liloc_ctx* c, int S) {
    if (S < 1) S = 1;
    c->share = S;
    c->s2m_max_blocks = std::max(1, c->s2m_blocks_full / S);
    c->pipe_blocks = std::max(1, c->pipe_blocks_full / S);
    c->pipe_blocks_wide = std::max(1, c->pipe_blocks_wide_full / S);
}

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
    c->map_cache = cfg->map_cache != 0;
    { const char* e = getenv("LILOC_PIPE_NARROW"); c->wide_maps = !(e && e[0] == '1'); }
    { const char* e = getenv("LILOC_NO_POLL"); c->poll_result = !(e && e[0] == '1'); }
    { const char* e = getenv("LILOC_PIPE_WIDE_MIN"); if (e && *e) c->wide_min = atoll(e); }
    { const char* e = getenv("LILOC_NDT_GRAIN_DIV"); if (e && *e && atoi(e) > 0) c->ndt_grain_div = atoi(e); }
    ctx = c;
    auto bail = [&](int rc) { g_create_error = c->err; liloc_destroy(c); return rc; };
#define CUB_(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fail(c, LILOC_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); return bail(LILOC_ERR_CUDA); } } while (0)
    CUB_(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUB_(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    CUB_(cudaStreamCreateWithFlags(&c->stream3, cudaStreamNonBlocking));
    CUB_(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CUB_(cudaEventCreateWithFlags(&c->ev_join2, cudaEventDisableTiming));
    CUB_(cudaEventCreateWithFlags(&c->ev_join3, cudaEventDisableTiming));
    CUB_(cudaEventCreateWithFlags(&c->pf.ev, cudaEventDisableTiming));
    for (FeatClass* fcp : {&c->fc[0], &c->fc[1], &c->aux}) {
        FeatClass& fc = *fcp;
        for (VgInst* vi : {&fc.vg_map, &fc.vg_scan}) {
            CUB_(cudaMalloc(&vi->d_enc, 6 * sizeof(unsigned)));
            CUB_(cudaMalloc(&vi->d_vp, sizeof(VgParams)));
            CUB_(cudaMalloc(&vi->d_count, sizeof(int)));
            CUB_(cudaMemset(vi->d_vp, 0, sizeof(VgParams)));
            { const unsigned arm[6] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u};
              CUB_(cudaMemcpy(vi->d_enc, arm, sizeof(arm), cudaMemcpyHostToDevice)); }
            CUB_(cudaMemset(vi->d_count, 0, sizeof(int)));
        }
        fc.d_M = fc.vg_map.d_count;
        fc.d_Q = fc.vg_scan.d_count;
        CUB_(cudaMalloc(&fc.d_gp, sizeof(GridParams)));
        CUB_(cudaMemset(fc.d_gp, 0, sizeof(GridParams)));
    }
    CUB_(cudaMalloc(&c->d_pose, 6 * sizeof(float)));
    for (auto& s : c->d_state) { CUB_(cudaMalloc(&s, sizeof(S2mState))); CUB_(cudaMemset(s, 0, sizeof(S2mState))); }
    CUB_(cudaMalloc(&c->d_ndt_active, sizeof(int)));
    // pinned AND mapped: the registration kernel and the phase marks write their results straight into host memory
    // (posted PCIe writes), so no device-to-host copy sits between the last kernel of a frame and the host's wake-up
    CUB_(cudaHostAlloc(&c->h_info, sizeof(HostFrameInfo), cudaHostAllocMapped));
    std::memset(c->h_info, 0, sizeof(HostFrameInfo));
    {
        HostFrameInfo* dev_alias = nullptr;
        CUB_(cudaHostGetDevicePointer(&dev_alias, c->h_info, 0));
        c->d_result_host = &dev_alias->res;
        c->d_marks_host = dev_alias->marks;
        for (int k = 0; k < 2; ++k) c->dk[k].d_n_host = &dev_alias->dk_n[k];
        c->d_done_host = &dev_alias->done_seq;
        c->d_pipe_error_host = &dev_alias->pipe_error;
        c->d_dk_marks = dev_alias->dk_marks;
    }
    for (auto& e : c->ev) CUB_(cudaEventCreate(&e));
    int per_sm = 0;
    CUB_(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_scan2map, kS2mThreads, 0));
    if (per_sm < 1) { fail(c, LILOC_ERR_CUDA, "k_scan2map cannot be resident"); return bail(LILOC_ERR_CUDA); }
    c->s2m_max_blocks = per_sm * c->num_sms;
    CUB_(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_voxel_pipeline, kPipeThreads, 0));
    if (per_sm < 1) { fail(c, LILOC_ERR_CUDA, "k_voxel_pipeline cannot be resident"); return bail(LILOC_ERR_CUDA); }
    c->pipe_blocks = c->num_sms;
    c->pipe_blocks_wide = per_sm >= 3 ? 2 * c->num_sms : c->num_sms;      // wide (2 per SM) + narrow (1 per SM) must co-reside
    c->s2m_blocks_full = c->s2m_max_blocks; c->pipe_blocks_full = c->pipe_blocks; c->pipe_blocks_wide_full = c->pipe_blocks_wide;
    {
        int share = cfg->share;
        if (share <= 0) { const char* e = std::getenv("LILOC_SHARE"); share = e ? std::atoi(e) : 1; }
        apply_share(c, share);
    }
    CUB_(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ndt_align, kNdtThreads, 0));
    if (per_sm < 1) { fail(c, LILOC_ERR_CUDA, "k_ndt_align cannot be resident"); return bail(LILOC_ERR_CUDA); }
    c->ndt_max_blocks = per_sm * c->num_sms;
    CUB_(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_deskew, kDeskewThreads, 0));
    if (per_sm < 1) { fail(c, LILOC_ERR_CUDA, "k_deskew cannot be resident"); return bail(LILOC_ERR_CUDA); }
    c->dk_blocks_per_sm = per_sm >= 2 ? 2 : 1;
    CUB_(cudaMallocHost(&c->h_ndt_hdr, 32 * sizeof(unsigned long long)));
    CUB_(cudaMalloc(&c->d_sm_scratch, 2 * kSmScratch * sizeof(int)));
    CUB_(cudaMemset(c->d_sm_scratch, 0, 2 * kSmScratch * sizeof(int)));
    CUB_(cudaMalloc(&c->d_dk_imu, 4 * kImuMax * sizeof(double)));
    CUB_(cudaMallocHost(&c->h_dk_imu, 2 * 4 * kImuMax * sizeof(double)));
    for (auto& e : c->ev_dk_slot) CUB_(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& S : c->dk) {
        CUB_(cudaMalloc(&S.d_n, sizeof(int)));
        CUB_(cudaMemset(S.d_n, 0, sizeof(int)));
        CUB_(cudaEventCreateWithFlags(&S.ev_done, cudaEventDisableTiming));
    }
    CUB_(cudaEventCreateWithFlags(&c->ev_dk_fork2, cudaEventDisableTiming));
    CUB_(cudaMalloc(&c->d_marks, (2 * kPipeMarks + kS2mMarks) * sizeof(unsigned long long)));
    CUB_(cudaMemset(c->d_marks, 0, (2 * kPipeMarks + kS2mMarks) * sizeof(unsigned long long)));
#undef CUB_
    const int scan_cap = cfg->max_scan_points > 0 ? cfg->max_scan_points : 16 * 1800;
    const int raw_cap = cfg->max_raw_points > 0 ? cfg->max_raw_points : (1 << 20);
    FeatClass& s = c->fc[0];
    int rc;
    if ((rc = ensure(c, s.scan_raw, (size_t)scan_cap)) || (rc = ensure(c, s.scan, (size_t)scan_cap)) ||
        (rc = ensure(c, s.raw, (size_t)raw_cap)) || (rc = ensure(c, s.map, (size_t)raw_cap)) ||
        (rc = ensure(c, s.vg_map.keys_a, (size_t)raw_cap)) || (rc = ensure(c, s.vg_map.keys_b, (size_t)raw_cap)) ||
        (rc = ensure(c, s.vg_map.vals_a, (size_t)raw_cap)) || (rc = ensure(c, s.vg_map.vals_b, (size_t)raw_cap)) ||
        (rc = ensure(c, s.vg_scan.keys_a, (size_t)scan_cap)) || (rc = ensure(c, s.vg_scan.keys_b, (size_t)scan_cap)) ||
        (rc = ensure(c, s.vg_scan.vals_a, (size_t)scan_cap)) || (rc = ensure(c, s.vg_scan.vals_b, (size_t)scan_cap)) ||
        (rc = ensure(c, c->arena, (size_t)raw_cap)) || (rc = ensure(c, s.cpts, (size_t)raw_cap)) ||
        (rc = ensure(c, c->partial, (size_t)2 * c->s2m_max_blocks * kNSums)))
        return bail(rc);
    *out = c;
    return LILOC_OK;
}

void liloc_destroy(liloc_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (cudaStream_t st : {c->stream, c->stream2, c->stream3}) if (st) cudaStreamSynchronize(st);
    auto fr = [](void* p) { if (p) cudaFree(p); };
    fr(c->arena.p);
    fr(c->aux_d2.p);
    for (FeatClass* fcp : {&c->fc[0], &c->fc[1], &c->aux}) {
        FeatClass& fc = *fcp;
        fr(fc.raw.p); fr(fc.map.p); fr(fc.cpts.p); fr(fc.descs.p); fr(fc.scan_raw.p); fr(fc.scan.p);
        fr(fc.cell_counts.p); fr(fc.cell_start.p); fr(fc.grid_state.p); fr(fc.d_gp);
        if (fc.h_descs) cudaFreeHost(fc.h_descs);
        for (VgInst* vi : {&fc.vg_map, &fc.vg_scan}) {
            fr(vi->keys_a.p); fr(vi->keys_b.p); fr(vi->vals_a.p); fr(vi->vals_b.p); fr(vi->hist.p); fr(vi->tile_state.p);
            fr(vi->d_enc); fr(vi->d_vp); fr(vi->d_count);
        }
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join2) cudaEventDestroy(c->ev_join2);
    if (c->ev_join3) cudaEventDestroy(c->ev_join3);
    if (c->pf.ev) cudaEventDestroy(c->pf.ev);
    fr(c->pf.buf[0].p); fr(c->pf.buf[1].p);
    fr(c->d_marks); fr(c->d_sm_scratch);
    fr(c->dk_pts.p); fr(c->dk_rt.p); fr(c->dk_tiles.p); fr(c->d_dk_imu);
    for (auto& S : c->dk) { fr(S.out.p); fr(S.d_n); if (S.ev_done) cudaEventDestroy(S.ev_done); }
    if (c->ev_dk_fork2) cudaEventDestroy(c->ev_dk_fork2);
    if (c->h_dk_imu) cudaFreeHost(c->h_dk_imu);
    for (auto& e : c->ev_dk_slot) if (e) cudaEventDestroy(e);
    fr(c->partial.p); fr(c->d_pose); fr(c->d_state[0]); fr(c->d_state[1]);
    fr(c->tmp_pts.p); fr(c->tmp_i.p); fr(c->tmp_f.p); fr(c->tmp_b.p);
    for (auto& kv : c->ndt_targets) { fr(kv.second.table.p); fr(kv.second.leaves.p); }
    for (auto& kv : c->ndt_sources) fr(kv.second.own.p);
    fr(c->d_tviews.p); fr(c->d_sviews.p); fr(c->d_jobs.p); fr(c->ndt_partial.p); fr(c->d_ndt_active);
    fr(c->ndt_descs.p);
    fr(c->ndt_hot.p); fr(c->ndt_sync.p); fr(c->ndt_q.p); fr(c->ndt_rows.p);
    if (c->h_ndt_hdr) cudaFreeHost(c->h_ndt_hdr);
    if (c->h_ndt_jobs) cudaFreeHost(c->h_ndt_jobs);
    fr(c->gather_send.p); fr(c->gather_recv.p);
    if (c->h_gather) cudaFreeHost(c->h_gather);
    if (c->comm) { nccl().CommDestroy(c->comm); c->comm = nullptr; }
    if (c->h_info) cudaFreeHost(c->h_info);
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    for (cudaStream_t st : {c->stream2, c->stream3, c->stream}) if (st) cudaStreamDestroy(st);
    delete c;
}

#define ENTER(ctx)                                                                     \
    if (!(ctx)) return LILOC_ERR_ARG;                                                  \
    { cudaError_t e_ = cudaSetDevice((ctx)->device);                                   \
      if (e_ != cudaSuccess) return fail((ctx), LILOC_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e_)); }

// ---- keyframe store -----------------------------------------------------------------------------------
int liloc_keyframe_put_features(liloc_ctx* ctx, int kf_id, const liloc_point* surf, int n_surf, const liloc_point* corner, int n_corner) {
    ENTER(ctx);
    if (n_surf < 0 || n_corner < 0 || (n_surf > 0 && !surf) || (n_corner > 0 && !corner)) return fail(ctx, LILOC_ERR_ARG, "keyframe_put: bad arguments");
    int rc;
    if ((rc = ensure(ctx, ctx->arena, ctx->arena_used + (size_t)n_surf + (size_t)n_corner, true))) return rc;
    liloc_ctx::KfEntry e{};
    e.off[0] = ctx->arena_used; e.n[0] = n_surf;
    e.off[1] = ctx->arena_used + (size_t)n_surf; e.n[1] = n_corner;
    if (n_surf > 0) CU(cudaMemcpyAsync(ctx->arena.p + e.off[0], surf, (size_t)n_surf * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    if (n_corner > 0) CU(cudaMemcpyAsync(ctx->arena.p + e.off[1], corner, (size_t)n_corner * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));          // the caller may free its buffers on return
    if (ctx->kf.count(kf_id)) for (auto& k : ctx->map_key) k.valid = false;      // an id re-used for new data
    ctx->kf[kf_id] = e;
    ctx->arena_used += (size_t)n_surf + (size_t)n_corner;
    return LILOC_OK;
}

int liloc_keyframe_put(liloc_ctx* ctx, int kf_id, const liloc_point* pts, int n) {
    return liloc_keyframe_put_features(ctx, kf_id, pts, n, nullptr, 0);
}

int liloc_keyframe_put_from_scan(liloc_ctx* ctx, int kf_id) {
    ENTER(ctx);
    FeatClass& s = ctx->fc[0];
    FeatClass& cr = ctx->fc[1];
    if (!s.have_scan || s.Q_all < 0) return fail(ctx, LILOC_ERR_STATE, "keyframe_put_from_scan: no scan");
    const bool corner = ctx->corner_scan_current && cr.have_scan && cr.Q_all >= 0;
    const int n0 = s.Q_all, n1 = corner ? cr.Q_all : 0;
    int rc;
    if ((rc = ensure(ctx, ctx->arena, ctx->arena_used + (size_t)n0 + (size_t)n1, true))) return rc;
    liloc_ctx::KfEntry e{};
    e.off[0] = ctx->arena_used; e.n[0] = n0; e.off[1] = ctx->arena_used + (size_t)n0; e.n[1] = n1;
    if (n0 > 0) CU(cudaMemcpyAsync(ctx->arena.p + e.off[0], s.scan.p, (size_t)n0 * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
    if (n1 > 0) CU(cudaMemcpyAsync(ctx->arena.p + e.off[1], cr.scan.p, (size_t)n1 * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
    if (ctx->kf.count(kf_id)) for (auto& k : ctx->map_key) k.valid = false;
    ctx->kf[kf_id] = e;
    ctx->arena_used += (size_t)n0 + (size_t)n1;
    return LILOC_OK;
}

// priorSession_->keyCloudVec_.push_back(copy of curCloud_) (anchorOptimize.hpp:279-282, :318-321): curCloud_ is
// laserCloudSurfLast, the FULL de-skewed scan (src/mapOptmization.cpp:1342) -- device-to-device from wherever the current
// frame's raw scan lives (the upload buffer, the caller's device cloud or the de-skew slot).
int liloc_keyframe_put_from_raw_scan(liloc_ctx* ctx, int kf_id) {
    ENTER(ctx);
    FeatClass& s = ctx->fc[0];
    if (!s.scan_src_last || s.n_scan_last <= 0) return fail(ctx, LILOC_ERR_STATE, "keyframe_put_from_raw_scan: no scan");
    const int n = s.n_scan_last;
    int rc;
    if ((rc = ensure(ctx, ctx->arena, ctx->arena_used + (size_t)n, true))) return rc;
    liloc_ctx::KfEntry e{};
    e.off[0] = ctx->arena_used; e.n[0] = n; e.off[1] = ctx->arena_used + (size_t)n; e.n[1] = 0;
    CU(cudaStreamSynchronize(ctx->stream2));         // the scan branch of the frame that brought the scan
    CU(cudaMemcpyAsync(ctx->arena.p + e.off[0], s.scan_src_last, (size_t)n * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
    if (ctx->kf.count(kf_id)) for (auto& k : ctx->map_key) k.valid = false;
    ctx->kf[kf_id] = e;
    ctx->arena_used += (size_t)n;
    return LILOC_OK;
}

int liloc_keyframe_size(liloc_ctx* ctx, int kf_id) {
    if (!ctx) return LILOC_ERR_ARG;
    auto it = ctx->kf.find(kf_id);
    return it == ctx->kf.end() ? LILOC_ERR_ARG : it->second.n[0];
}

int liloc_keyframe_size_class(liloc_ctx* ctx, int cls, int kf_id) {
    if (!ctx || bad_class(cls)) return LILOC_ERR_ARG;
    auto it = ctx->kf.find(kf_id);
    return it == ctx->kf.end() ? LILOC_ERR_ARG : it->second.n[cls];
}

int liloc_keyframe_get(liloc_ctx* ctx, int cls, int kf_id, liloc_point* out, int cap, int* n_out) {
    ENTER(ctx);
    if (bad_class(cls)) return fail(ctx, LILOC_ERR_ARG, "keyframe_get: bad class %d", cls);
    auto it = ctx->kf.find(kf_id);
    if (it == ctx->kf.end()) return fail(ctx, LILOC_ERR_ARG, "keyframe_get: unknown keyframe id %d", kf_id);
    const int n = it->second.n[cls];
    if (n_out) *n_out = n;
    if (out) {
        if (cap < n) return fail(ctx, LILOC_ERR_CAPACITY, "keyframe_get: cap %d < n %d", cap, n);
        if (n > 0) CU(cudaMemcpyAsync(out, ctx->arena.p + it->second.off[cls], (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return LILOC_OK;
}

int liloc_keyframe_clear(liloc_ctx* ctx) {
    ENTER(ctx);
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->kf.clear(); ctx->arena_used = 0;
    for (auto& k : ctx->map_key) k.valid = false;
    return LILOC_OK;
}

// ---- local map ----------------------------------------------------------------------------------------
int liloc_localmap_build_class(liloc_ctx* ctx, int cls, const int* kf_ids, const float* poses6, int K, float leaf, int* M_out, int* n_raw_out) {
    ENTER(ctx);
    if (bad_class(cls)) return fail(ctx, LILOC_ERR_ARG, "localmap_build: bad class %d", cls);
    int rc;
    PipeArgs pa{};
    if ((rc = localmap_problem(ctx, cls, ctx->stream, kf_ids, poses6, K, leaf, &pa.prob[0], &pa))) return rc;
    pa.n_prob = 1;
    if ((rc = pipe_launch(ctx, ctx->stream, pa, nullptr))) return rc;
    if ((rc = class_fetch(ctx, cls, ctx->stream, true, false))) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    class_collect(ctx, cls, true, false);
    if (ctx->pipe_failed) { const int e = ctx->pipe_failed; ctx->pipe_failed = 0; return e; }
    if (M_out) *M_out = ctx->fc[cls].M;
    if (n_raw_out) *n_raw_out = ctx->fc[cls].n_raw;
    return LILOC_OK;
}

int liloc_localmap_build(liloc_ctx* ctx, const int* kf_ids, const float* poses6, int K, float leaf, int* M_out, int* n_raw_out) {
    return liloc_localmap_build_class(ctx, LILOC_SURF, kf_ids, poses6, K, leaf, M_out, n_raw_out);
}

int liloc_localmap_set_class(liloc_ctx* ctx, int cls, const liloc_point* pts, int n, float leaf, int* M_out) {
    ENTER(ctx);
    if (bad_class(cls) || n < 0 || (n > 0 && !pts) || !(leaf > 0.f)) return fail(ctx, LILOC_ERR_ARG, "localmap_set: bad arguments");
    FeatClass& fc = ctx->fc[cls];
    int rc;
    fc.have_map = false; fc.M = -1;
    ctx->map_key[cls].valid = false;
    if ((rc = ensure(ctx, fc.raw, (size_t)(n > 0 ? n : 1)))) return rc;
    if (n > 0) CU(cudaMemcpyAsync(fc.raw.p, pts, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    fc.n_raw = n;
    PipeArgs pa{};
    if ((rc = prob_voxelgrid(ctx, fc.vg_map, fc.raw.p, n, leaf, &fc.map, &pa.prob[0]))) return rc;
    if ((rc = prob_add_grid(ctx, fc, n, &pa.prob[0]))) return rc;
    pa.n_prob = 1;
    if ((rc = pipe_launch(ctx, ctx->stream, pa, nullptr))) return rc;
    if ((rc = class_fetch(ctx, cls, ctx->stream, true, false))) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    fc.have_map = true;
    class_collect(ctx, cls, true, false);
    if (ctx->pipe_failed) { const int e = ctx->pipe_failed; ctx->pipe_failed = 0; return e; }
    if (M_out) *M_out = fc.M;
    return LILOC_OK;
}

int liloc_localmap_set(liloc_ctx* ctx, const liloc_point* pts, int n, float leaf, int* M_out) {
    return liloc_localmap_set_class(ctx, LILOC_SURF, pts, n, leaf, M_out);
}

int liloc_localmap_get_class(liloc_ctx* ctx, int cls, liloc_point* out, int cap, int* M_out) {
    ENTER(ctx);
    if (bad_class(cls)) return fail(ctx, LILOC_ERR_ARG, "localmap_get: bad class %d", cls);
    FeatClass& fc = ctx->fc[cls];
    if (!fc.have_map || fc.M < 0) return fail(ctx, LILOC_ERR_STATE, "localmap_get: no local map");
    if (M_out) *M_out = fc.M;
    if (out) {
        if (cap < fc.M) return fail(ctx, LILOC_ERR_CAPACITY, "localmap_get: cap %d < M %d", cap, fc.M);
        if (fc.M > 0) CU(cudaMemcpyAsync(out, fc.map.p, (size_t)fc.M * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return LILOC_OK;
}

int liloc_localmap_get(liloc_ctx* ctx, liloc_point* out, int cap, int* M_out) { return liloc_localmap_get_class(ctx, LILOC_SURF, out, cap, M_out); }

// ---- current scan -------------------------------------------------------------------------------------
static int scan_put_impl(liloc_ctx* ctx, int cls, const liloc_point* pts, int n, int on_device, float leaf, int* Q_all_out) {
    ENTER(ctx);
    if (bad_class(cls)) return fail(ctx, LILOC_ERR_ARG, "scan_put: bad class %d", cls);
    int rc;
    PipeArgs pa{};
    if ((rc = scan_problem(ctx, cls, ctx->stream, pts, n, on_device, leaf, &pa.prob[0]))) return rc;
    pa.n_prob = 1;
    if ((rc = pipe_launch(ctx, ctx->stream, pa, nullptr))) return rc;
    if ((rc = class_fetch(ctx, cls, ctx->stream, false, true))) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    class_collect(ctx, cls, false, true);
    if (ctx->pipe_failed) { const int e = ctx->pipe_failed; ctx->pipe_failed = 0; return e; }
    if (on_device == LILOC_SCAN_DESKEWED) { liloc_ctx::DkSweep& S = ctx->dk[ctx->dk_cur]; S.n = ctx->h_info->dk_n[ctx->dk_cur]; ctx->fc[cls].n_scan_last = S.n; }
    if (cls == LILOC_SURF) ctx->corner_scan_current = false;      // a new frame starts with its surf scan
    else ctx->corner_scan_current = true;
    if (Q_all_out) *Q_all_out = ctx->fc[cls].Q_all;
    return LILOC_OK;
}

int liloc_scan_put(liloc_ctx* ctx, const liloc_point* pts, int n, float leaf, int* Q_all_out) { return scan_put_impl(ctx, LILOC_SURF, pts, n, LILOC_SCAN_HOST, leaf, Q_all_out); }
int liloc_scan_put_dev(liloc_ctx* ctx, const liloc_point* pts_dev, int n, float leaf, int* Q_all_out) { return scan_put_impl(ctx, LILOC_SURF, pts_dev, n, LILOC_SCAN_DEVICE, leaf, Q_all_out); }
int liloc_scan_put_class(liloc_ctx* ctx, int cls, const liloc_point* pts, int n, int on_device, float leaf, int* Q_all_out) {
    return scan_put_impl(ctx, cls, pts, n, on_device, leaf, Q_all_out);
}

int liloc_scan_ds_get_class(liloc_ctx* ctx, int cls, liloc_point* out, int cap, int* Q_all_out) {
    ENTER(ctx);
    if (bad_class(cls)) return fail(ctx, LILOC_ERR_ARG, "scan_ds_get: bad class %d", cls);
    FeatClass& fc = ctx->fc[cls];
    if (!fc.have_scan || fc.Q_all < 0) return fail(ctx, LILOC_ERR_STATE, "scan_ds_get: no scan");
    if (Q_all_out) *Q_all_out = fc.Q_all;
    if (out) {
        if (cap < fc.Q_all) return fail(ctx, LILOC_ERR_CAPACITY, "scan_ds_get: cap %d < Q_all %d", cap, fc.Q_all);
        if (fc.Q_all > 0) CU(cudaMemcpyAsync(out, fc.scan.p, (size_t)fc.Q_all * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return LILOC_OK;
}

int liloc_scan_ds_get(liloc_ctx* ctx, liloc_point* out, int cap, int* Q_all_out) { return liloc_scan_ds_get_class(ctx, LILOC_SURF, out, cap, Q_all_out); }

// ---- de-skew (imageProjection::projectPointCloud, src/imageProjection.cpp:645-673) ---------------------------------
// Runs on stream3, a stream of its own: H2D of the raw sweep and of the IMU arrays, one cooperative launch, the count
// into mapped host memory.  Nothing is synchronised unless the caller asks for the count.  The result goes into the sweep
// slot that is not the current one and becomes the current sweep at once -- or, with LILOC_DESKEW_AHEAD, stays staged
// until liloc_deskew_advance: the next sweep is then uploaded and de-skewed while the current frame runs.
int liloc_deskew(liloc_ctx* ctx, const liloc_point* pts, const liloc_ring_time* ring_time, int n, const liloc_deskew_params* p, int* n_out) {
    ENTER(ctx);
    if (n < 0 || (n > 0 && (!pts || !ring_time)) || !p) return fail(ctx, LILOC_ERR_ARG, "deskew: bad arguments");
    if (p->downsample_rate < 1 || p->point_filter_num < 1) return fail(ctx, LILOC_ERR_ARG, "deskew: downsample_rate and point_filter_num must be >= 1");
    const bool deskew = p->deskew_flag != -1 && p->imu_available != 0;
    if (deskew && (p->imu_pointer_cur < 0 || p->imu_pointer_cur >= kImuMax || !p->imu_time || !p->imu_rot_x || !p->imu_rot_y || !p->imu_rot_z))
        return fail(ctx, LILOC_ERR_ARG, "deskew: bad IMU arrays (imu_pointer_cur %d, queue length %d)", p->imu_pointer_cur, kImuMax);
    const bool ahead = p->reserved[0] == LILOC_DESKEW_AHEAD;
    int rc;
    cudaStream_t st = ctx->stream3;
    const int slot = ctx->dk_cur < 0 ? 0 : (ctx->dk_cur ^ 1);             // never the current sweep's slot
    liloc_ctx::DkSweep& S = ctx->dk[slot];
    const size_t cap = (size_t)(n > 0 ? n : 1);
    if ((rc = ensure(ctx, ctx->dk_pts, cap)) || (rc = ensure(ctx, ctx->dk_rt, cap)) || (rc = ensure(ctx, S.out, cap)) ||
        (rc = ensure(ctx, ctx->dk_tiles, cap / kDeskewThreads + 2))) return rc;
    // whatever was enqueued so far on the frame's streams may still read this slot (a sweep two de-skews back)
    CU(cudaEventRecord(ctx->ev_fork, ctx->stream));
    CU(cudaStreamWaitEvent(st, ctx->ev_fork, 0));
    CU(cudaEventRecord(ctx->ev_dk_fork2, ctx->stream2));
    CU(cudaStreamWaitEvent(st, ctx->ev_dk_fork2, 0));
    if (n > 0) {
        CU(cudaMemcpyAsync(ctx->dk_pts.p, pts, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ctx->dk_rt.p, ring_time, (size_t)n * sizeof(int2), cudaMemcpyHostToDevice, st));
        ctx->stats.h2d_bytes += (long long)n * 24;
    }
    if (deskew) {
        const int m = p->imu_pointer_cur + 1;
        const int stg = ctx->dk_slot; ctx->dk_slot ^= 1;           // two pinned staging slots, each guarded by an event
        CU(cudaEventSynchronize(ctx->ev_dk_slot[stg]));
        double* stage = ctx->h_dk_imu + (size_t)stg * 4 * kImuMax;
        const double* src[4] = {p->imu_time, p->imu_rot_x, p->imu_rot_y, p->imu_rot_z};
        for (int k = 0; k < 4; ++k) std::memcpy(stage + (size_t)k * m, src[k], (size_t)m * sizeof(double));
        CU(cudaMemcpyAsync(ctx->d_dk_imu, stage, (size_t)4 * m * sizeof(double), cudaMemcpyHostToDevice, st));      // one copy: 4 arrays of m
        CU(cudaEventRecord(ctx->ev_dk_slot[stg], st));
        ctx->stats.h2d_bytes += (long long)m * 32;
    }
    DeskewArgs a{};
    a.pts = ctx->dk_pts.p; a.rt = ctx->dk_rt.p; a.n = n;
    a.imu = ctx->d_dk_imu; a.imu_stride = p->imu_pointer_cur + 1; a.imu_last = p->imu_pointer_cur; a.t_cur = p->time_scan_cur; a.deskew = deskew ? 1 : 0;
    a.rmin = p->lidar_min_range; a.rmax = p->lidar_max_range;
    a.n_scan = p->n_scan; a.ds_rate = p->downsample_rate; a.pf_num = p->point_filter_num;
    a.tile_count = ctx->dk_tiles.p; a.out = S.out.p; a.n_out = S.d_n; a.n_out_host = S.d_n_host;
    const int n_tiles = (n + kDeskewThreads - 1) / kDeskewThreads;
    // a sweep staged ahead shares the GPU with the current frame's launches: one block per SM leaves them room
    const int max_blocks = ctx->num_sms * (ahead ? 1 : ctx->dk_blocks_per_sm);
    const int blocks = n_tiles < 1 ? 1 : (n_tiles < max_blocks ? n_tiles : max_blocks);
    a.marks = ctx->timing ? ctx->d_dk_marks : nullptr;
    void* args[] = {&a};
    CU(cudaLaunchCooperativeKernel((void*)k_deskew, dim3((unsigned)blocks), dim3(kDeskewThreads), args, 0, st));
    ++ctx->launches;
    ctx->stats.d2h_bytes += 4;                                   // the count, written by the kernel into mapped host memory
    CU(cudaEventRecord(S.ev_done, st));
    S.n_upper = n; S.n = -1;
    if (ahead) {
        ctx->dk_staged = slot;
    } else {
        ctx->dk_cur = slot; ctx->dk_staged = -1;
        ctx->fc[LILOC_SURF].have_scan = false; ctx->fc[LILOC_SURF].Q_all = -1;
    }
    if (n_out) {
        CU(cudaStreamSynchronize(st));
        S.n = ctx->h_info->dk_n[slot];
        *n_out = S.n;
    }
    return LILOC_OK;
}

// The sweep staged by a LILOC_DESKEW_AHEAD call becomes the current one.
int liloc_deskew_advance(liloc_ctx* ctx) {
    ENTER(ctx);
    if (ctx->dk_staged < 0) return fail(ctx, LILOC_ERR_STATE, "deskew_advance: no sweep staged ahead");
    ctx->dk_cur = ctx->dk_staged; ctx->dk_staged = -1;
    ctx->fc[LILOC_SURF].have_scan = false; ctx->fc[LILOC_SURF].Q_all = -1;
    return LILOC_OK;
}

static int deskew_count(liloc_ctx* ctx) {          // host copy of the current sweep's size (synchronises when still unknown)
    if (ctx->dk_cur < 0) return fail(ctx, LILOC_ERR_STATE, "no de-skewed sweep (call liloc_deskew first)");
    liloc_ctx::DkSweep& S = ctx->dk[ctx->dk_cur];
    if (S.n < 0) {
        CU(cudaEventSynchronize(S.ev_done));
        S.n = ctx->h_info->dk_n[ctx->dk_cur];
    }
    return 0;
}

int liloc_deskewed_get(liloc_ctx* ctx, liloc_point* out, int cap, int* n_out) {
    ENTER(ctx);
    int rc;
    if ((rc = deskew_count(ctx))) return rc;
    const liloc_ctx::DkSweep& S = ctx->dk[ctx->dk_cur];
    if (n_out) *n_out = S.n;
    if (out) {
        if (cap < S.n) return fail(ctx, LILOC_ERR_CAPACITY, "deskewed_get: cap %d < n %d", cap, S.n);
        CU(cudaEventSynchronize(S.ev_done));
        if (S.n > 0) CU(cudaMemcpy(out, S.out.p, (size_t)S.n * sizeof(float4), cudaMemcpyDeviceToHost));
        ctx->stats.d2h_bytes += (long long)S.n * 16;
    }
    return LILOC_OK;
}

// phase timestamps (ns, %globaltimer) of block 0 in the last de-skew launch: start, after the barrier, end
int liloc_last_deskew_marks(liloc_ctx* ctx, unsigned long long* out3) {
    ENTER(ctx);
    if (!out3) return fail(ctx, LILOC_ERR_ARG, "last_deskew_marks: null");
    CU(cudaStreamSynchronize(ctx->stream3));
    std::memcpy(out3, ctx->h_info->dk_marks, 3 * sizeof(unsigned long long));
    return LILOC_OK;
}

int liloc_last_deskew_ms(liloc_ctx* ctx, float* ms) {
    ENTER(ctx);
    if (!ms) return fail(ctx, LILOC_ERR_ARG, "last_deskew_ms: null");
    if (ctx->timing && (ctx->dk_cur >= 0 || ctx->dk_staged >= 0)) {      // in-kernel duration (block 0, %globaltimer)
        CU(cudaStreamSynchronize(ctx->stream3));
        const unsigned long long* mk = ctx->h_info->dk_marks;
        ctx->dk_ms = mk[2] > mk[0] ? (float)((double)(mk[2] - mk[0]) * 1e-6) : 0.f;
    }
    *ms = ctx->dk_ms;
    return LILOC_OK;
}

// ---- registration -------------------------------------------------------------------------------------
int liloc_scan2map(liloc_ctx* ctx, float* pose6, const liloc_s2m_opts* opts, liloc_s2m_result* res) {
    ENTER(ctx);
    if (!pose6) return fail(ctx, LILOC_ERR_ARG, "scan2map: null pose");
    const liloc_s2m_opts* o = opts ? opts : &kDefaultOpts;
    int rc;
    tick(ctx, 4, ctx->stream);
    if ((rc = s2m_enqueue(ctx, pose6, o))) return rc;
    tick(ctx, 5, ctx->stream);
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->timing) { float ms = 0.f; cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]); ctx->last_timing.s2m_ms = ms; }
    return s2m_collect(ctx, pose6, res, o);
}

// Wait for the registration kernel enqueued last.  The kernel writes its result into mapped host memory and then the
// frame's sequence number; polling that word returns a few microseconds before cudaStreamSynchronize would (which waits
// for the kernel to retire).  Everything enqueued afterwards is stream-ordered behind the kernel anyway.  If the word
// does not arrive (a faulted kernel never writes it) the stream is synchronised and the error surfaces there.
static int frame_wait(liloc_ctx* ctx) {
    if (ctx->poll_result) {
        const volatile unsigned* flag = &ctx->h_info->done_seq;
        const unsigned want = ctx->frame_seq;
        const auto t0 = std::chrono::steady_clock::now();
        int spins = 0;
        while (*flag != want) {
            if ((++spins & 1023) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(20)) break;
        }
        if (*flag == want) return 0;
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// liloc_scan_prefetch's copy: on stream3 (nothing of a host-scan frame runs there), into the buffer the frame in flight
// does not read.  The buffer it overwrites was last read two frames ago, and every frame ends with a host wait.
static int prefetch_issue(liloc_ctx* ctx) {
    liloc_ctx::ScanPrefetch& pf = ctx->pf;
    if (!pf.hint) return 0;
    const int slot = pf.in_use == 0 ? 1 : 0;         // an upload nobody consumed may be overwritten: it is replaced here
    int rc;
    if ((rc = ensure(ctx, pf.buf[slot], (size_t)pf.hint_n))) return rc;
    CU(cudaMemcpyAsync(pf.buf[slot].p, pf.hint, (size_t)pf.hint_n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream3));
    CU(cudaEventRecord(pf.ev, ctx->stream3));
    ctx->stats.h2d_bytes += (long long)pf.hint_n * (long long)sizeof(float4);
    pf.host = pf.hint; pf.n = pf.hint_n; pf.slot = slot; pf.issued = true;
    pf.hint = nullptr;
    return 0;
}

int liloc_scan_prefetch(liloc_ctx* ctx, const liloc_point* pts, int n) {
    ENTER(ctx);
    if (n < 0 || (n > 0 && !pts)) return fail(ctx, LILOC_ERR_ARG, "scan_prefetch: bad arguments");
    liloc_ctx::ScanPrefetch& pf = ctx->pf;
    if (pf.issued && pf.host == pts && pf.n == n) { pf.hint = nullptr; return LILOC_OK; }      // already on its way
    pf.hint = n > 0 ? pts : nullptr; pf.hint_n = n;
    return LILOC_OK;
}

int liloc_frame_features(liloc_ctx* ctx, const liloc_frame_in* in, float* pose6, const liloc_s2m_opts* opts, liloc_s2m_result* res) {
    ENTER(ctx);
    if (!pose6 || !in) return fail(ctx, LILOC_ERR_ARG, "frame: null argument");
    const liloc_s2m_opts* o = opts ? opts : &kDefaultOpts;
    const bool corner = o->enable_corner != 0;
    const int dev = in->scan_on_device;
    int rc;
    const int ncls = corner ? 2 : 1;
    // scans: H2D + VoxelGrid on stream2, so the copy overlaps the map pipeline on the main stream
    CU(cudaEventRecord(ctx->ev_fork, ctx->stream));
    CU(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
    PipeArgs ps{}, pm{};
    for (int cls = 0; cls < ncls; ++cls)
        if ((rc = scan_problem(ctx, cls, ctx->stream2, in->scan[cls], in->n_scan[cls],
                               (dev == LILOC_SCAN_DESKEWED && cls != LILOC_SURF) ? LILOC_SCAN_HOST : dev, in->scan_leaf[cls], &ps.prob[cls]))) return rc;
    ps.n_prob = ncls;
    // No CUDA events inside a frame, not even with timing on: an event record between two launches costs ~2.5 us of
    // device time each (measured: 151 -> 145 us per frame without the two around the join).  The phase durations come
    // from the %globaltimer marks block 0 of every launch writes into mapped host memory.
    // local maps: descriptors -> assemble -> VoxelGrid -> search grid, one launch for all classes.  With map_cache a
    // class whose (ids, poses, leaf) are bit-identical to the map it already holds is left as it is.
    for (int cls = 0; cls < ncls; ++cls) {
        liloc_ctx::MapKey& key = ctx->map_key[cls];
        const size_t K = (size_t)(in->K > 0 ? in->K : 0);
        const bool hit = ctx->map_cache && key.valid && ctx->fc[cls].have_map && key.leaf == in->map_leaf[cls] && key.ids.size() == K &&
                         (K == 0 || (std::memcmp(key.ids.data(), in->kf_ids, K * sizeof(int)) == 0 &&
                                     std::memcmp(key.poses.data(), in->kf_poses6, K * 6 * sizeof(float)) == 0));
        if (hit) { ++ctx->map_cache_hits; continue; }
        if ((rc = localmap_problem(ctx, cls, ctx->stream, in->kf_ids, in->kf_poses6, in->K, in->map_leaf[cls], &pm.prob[pm.n_prob], &pm))) {
            // the scans were staged (and their upload enqueued) but no pipeline launch will follow: they are not usable
            for (int c2 = 0; c2 < ncls; ++c2) { ctx->fc[c2].have_scan = false; ctx->fc[c2].Q_all = -1; }
            return rc;
        }
        ++pm.n_prob;
        if (ctx->map_cache) {
            key.ids.assign(in->kf_ids, in->kf_ids + K); key.poses.assign(in->kf_poses6, in->kf_poses6 + 6 * K);
            key.leaf = in->map_leaf[cls]; key.valid = true;
        }
    }
    // measured (profiles/frame_latency_r01.md): two blocks per SM pay off for the streaming phases above ~1.5 M raw
    // points (Mid-360 map: 394 -> 332 us); below that one block per SM is as fast or faster (centroid gathers)
    const bool wide = ctx->wide_maps && (long long)ctx->fc[0].n_raw + (corner ? ctx->fc[1].n_raw : 0) > ctx->wide_min;
    const bool map_launched = pm.n_prob > 0;
    if ((rc = pipe_launch(ctx, ctx->stream, pm, ctx->timing ? ctx->d_marks_host : nullptr, wide))) return rc;
    // The scan launch is issued AFTER the map launch: its upload was enqueued first (above) and takes longer than the host
    // needs to get here, while the map launch -- the longer branch when the scan is already on the device -- would
    // otherwise start a launch-call later.
    if ((rc = pipe_launch(ctx, ctx->stream2, ps, ctx->timing ? ctx->d_marks_host + kPipeMarks : nullptr))) return rc;
    CU(cudaEventRecord(ctx->ev_join2, ctx->stream2));
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_join2, 0));                                         // join
    if ((rc = s2m_enqueue(ctx, pose6, o))) return rc;
    if ((rc = prefetch_issue(ctx))) return rc;       // the hinted next scan goes up while this frame computes
    if ((rc = frame_wait(ctx))) return rc;           // the frame's only host synchronisation
    {   // counts come back inside the result struct
        const S2mResult& r = ctx->h_info->res;
        ctx->h_info->c[0].M = r.map_points; ctx->h_info->c[0].Q_all = r.q_all;
        ctx->h_info->c[1].M = r.extra[1];   ctx->h_info->c[1].Q_all = r.extra[0];
    }
    for (int cls = 0; cls < (corner ? 2 : 1); ++cls) class_collect(ctx, cls, true, true);
    if (dev == LILOC_SCAN_DESKEWED) { liloc_ctx::DkSweep& S = ctx->dk[ctx->dk_cur]; S.n = ctx->h_info->dk_n[ctx->dk_cur]; ctx->fc[LILOC_SURF].n_scan_last = S.n; }
    ctx->corner_scan_current = corner;
    if (ctx->timing) {
        liloc_timing& t = ctx->last_timing;
        const unsigned long long* mk = ctx->h_info->marks;       // phase boundaries of the maps launch (ns)
        auto span = [&](int i, int j) { return (map_launched && mk[j] > mk[i]) ? (float)((double)(mk[j] - mk[i]) * 1e-6) : 0.f; };
        t.assemble_ms = span(0, 1);                               // transform + concat + bounding box
        t.map_voxel_ms = span(1, 5);                              // keys, radix sort, run count, centroids
        t.grid_ms = span(5, 7);                                   // search-grid scan + scatter
        const unsigned long long* ms_ = mk + kPipeMarks;         // scans launch: [0] start .. [5] block 0's end (no search grid)
        const unsigned long long* mr = mk + 2 * kPipeMarks;      // registration: [0] start .. [kS2mMarks - 1] end
        auto dur = [](unsigned long long a0, unsigned long long a1) { return a1 > a0 ? (float)((double)(a1 - a0) * 1e-6) : 0.f; };
        t.scan_ms = dur(ms_[0], ms_[5]);
        t.s2m_ms = dur(mr[0], mr[kS2mMarks - 1]);
        const unsigned long long first = map_launched ? (mk[0] < ms_[0] ? mk[0] : ms_[0]) : ms_[0];
        t.total_ms = dur(first, mr[kS2mMarks - 1]);              // first kernel's first instruction .. last kernel's last
        liloc_stats& st = ctx->stats;
        st.assemble_ms += t.assemble_ms; st.map_voxel_ms += t.map_voxel_ms; st.scan_ms += t.scan_ms;
        st.grid_ms += t.grid_ms; st.s2m_ms += t.s2m_ms; st.total_ms += t.total_ms;
    }
    ++ctx->stats.frames;
    for (int cls = 0; cls < (corner ? 2 : 1); ++cls) {
        const FeatClass& fc = ctx->fc[cls];
        if (map_launched) ctx->stats.alg_bytes_map += 16.0 * fc.n_raw + 16.0 * (fc.M > 0 ? fc.M : 0);
        ctx->stats.alg_bytes_scan += 16.0 * fc.n_scan_last + 16.0 * (fc.Q_all > 0 ? fc.Q_all : 0);
    }
    if (ctx->pipe_failed) { const int e = ctx->pipe_failed; ctx->pipe_failed = 0; return e; }
    return s2m_collect(ctx, pose6, res, o);
}

int liloc_frame(liloc_ctx* ctx, const int* kf_ids, const float* kf_poses6, int K, float map_leaf,
                const liloc_point* scan, int n_scan, int scan_on_device, float scan_leaf,
                float* pose6, const liloc_s2m_opts* opts, liloc_s2m_result* res) {
    liloc_frame_in in{};
    in.kf_ids = kf_ids; in.kf_poses6 = kf_poses6; in.K = K;
    in.map_leaf[0] = map_leaf; in.scan[0] = scan; in.n_scan[0] = n_scan; in.scan_leaf[0] = scan_leaf;
    in.scan_on_device = scan_on_device;
    liloc_s2m_opts o = opts ? *opts : kDefaultOpts;
    o.enable_corner = 0;                             // this entry point is the reference's surf-only frame
    return liloc_frame_features(ctx, &in, pose6, &o, res);
}

int liloc_last_s2m_extra(liloc_ctx* ctx, liloc_s2m_extra* out) {
    if (!ctx || !out) return LILOC_ERR_ARG;
    std::memset(out, 0, sizeof(*out));
    const S2mResult& r = ctx->h_info->res;
    out->q_corner = r.extra[0]; out->map_points_corner = r.extra[1]; out->tiles_corner = r.extra[2]; out->tiles_surf = r.extra[3];
    out->sort_redos = (int)ctx->sort_redos;
    return LILOC_OK;
}

// ---- kernel-level entry points ------------------------------------------------------------------------
int liloc_voxelgrid(liloc_ctx* ctx, const liloc_point* in, int n, float leaf, liloc_point* out, int cap, int* m_out) {
    ENTER(ctx);
    if (n < 0 || (n > 0 && (!in || !out)) || !(leaf > 0.f) || !m_out) return fail(ctx, LILOC_ERR_ARG, "voxelgrid: bad arguments");
    int rc;
    if ((rc = ensure(ctx, ctx->tmp_pts, (size_t)(n > 0 ? 2 * (size_t)n : 2)))) return rc;
    DevBuf<float4> outb; outb.p = ctx->tmp_pts.p + n; outb.cap = (size_t)(n > 0 ? n : 1);
    if (n > 0) CU(cudaMemcpyAsync(ctx->tmp_pts.p, in, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = ensure(ctx, ctx->tmp_i, 1))) return rc;
    {   // run on the scan instance's scratch but keep its result count (d_Q) untouched
        VgInst& vi = ctx->fc[0].vg_scan;
        PipeArgs pa{};
        if ((rc = prob_voxelgrid(ctx, vi, ctx->tmp_pts.p, n, leaf, &outb, &pa.prob[0]))) return rc;
        pa.prob[0].count = ctx->tmp_i.p;
        pa.n_prob = 1;
        if ((rc = pipe_launch(ctx, ctx->stream, pa, nullptr))) return rc;
    }
    int m = 0;
    CU(cudaMemcpyAsync(&m, ctx->tmp_i.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (pipe_check(ctx)) { const int e = ctx->pipe_failed; ctx->pipe_failed = 0; return e; }
    *m_out = m;
    if (m > cap) return fail(ctx, LILOC_ERR_CAPACITY, "voxelgrid: cap %d < m %d", cap, m);
    if (m > 0) CU(cudaMemcpyAsync(out, outb.p, (size_t)m * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return LILOC_OK;
}

int liloc_knn5_class(liloc_ctx* ctx, int cls, const liloc_point* queries, int nq, int* idx5, float* d2_5, int* found) {
    ENTER(ctx);
    if (bad_class(cls) || nq < 0 || (nq > 0 && (!queries || !idx5 || !d2_5 || !found))) return fail(ctx, LILOC_ERR_ARG, "knn5: bad arguments");
    FeatClass& fc = ctx->fc[cls];
    if (!fc.have_map) return fail(ctx, LILOC_ERR_STATE, "knn5: no local map");
    if (nq == 0) return LILOC_OK;
    int rc;
    if ((rc = ensure(ctx, ctx->tmp_pts, (size_t)nq))) return rc;
    if ((rc = ensure(ctx, ctx->tmp_i, (size_t)nq * 6))) return rc;
    if ((rc = ensure(ctx, ctx->tmp_f, (size_t)nq * 5))) return rc;
    CU(cudaMemcpyAsync(ctx->tmp_pts.p, queries, (size_t)nq * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    k_knn5<<<blocks_for(ctx, (long long)nq * kKnnLanes, 256), 256, 0, ctx->stream>>>(ctx->tmp_pts.p, nq, grid_view(fc), ctx->tmp_i.p, ctx->tmp_f.p, ctx->tmp_i.p + (size_t)nq * 5); KCHECK();
    CU(cudaMemcpyAsync(idx5, ctx->tmp_i.p, (size_t)nq * 5 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(found, ctx->tmp_i.p + (size_t)nq * 5, (size_t)nq * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(d2_5, ctx->tmp_f.p, (size_t)nq * 5 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return LILOC_OK;
}

int liloc_knn5(liloc_ctx* ctx, const liloc_point* queries, int nq, int* idx5, float* d2_5, int* found) {
    return liloc_knn5_class(ctx, LILOC_SURF, queries, nq, idx5, d2_5, found);
}

// liloc_knn5_class with the caller's upper bound of every query's fifth squared distance (the pruning the registration kernel
// applies from its second iteration on); test surface.
int liloc_debug_knn5_bounded(liloc_ctx* ctx, int cls, const liloc_point* queries, int nq, const float* bounds, int* idx5, float* d2_5, int* found) {
    ENTER(ctx);
    if (bad_class(cls) || nq < 0 || (nq > 0 && (!queries || !bounds || !idx5 || !d2_5 || !found))) return fail(ctx, LILOC_ERR_ARG, "knn5_bounded: bad arguments");
    FeatClass& fc = ctx->fc[cls];
    if (!fc.have_map) return fail(ctx, LILOC_ERR_STATE, "knn5_bounded: no local map");
    if (nq == 0) return LILOC_OK;
    int rc;
    if ((rc = ensure(ctx, ctx->tmp_pts, (size_t)nq))) return rc;
    if ((rc = ensure(ctx, ctx->tmp_i, (size_t)nq * 6))) return rc;
    if ((rc = ensure(ctx, ctx->tmp_f, (size_t)nq * 6))) return rc;
    CU(cudaMemcpyAsync(ctx->tmp_pts.p, queries, (size_t)nq * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->tmp_f.p + (size_t)nq * 5, bounds, (size_t)nq * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    k_knn5<<<blocks_for(ctx, (long long)nq * kKnnLanes, 256), 256, 0, ctx->stream>>>(ctx->tmp_pts.p, nq, grid_view(fc), ctx->tmp_i.p, ctx->tmp_f.p, ctx->tmp_i.p + (size_t)nq * 5, ctx->tmp_f.p + (size_t)nq * 5); KCHECK();
    CU(cudaMemcpyAsync(idx5, ctx->tmp_i.p, (size_t)nq * 5 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(found, ctx->tmp_i.p + (size_t)nq * 5, (size_t)nq * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(d2_5, ctx->tmp_f.p, (size_t)nq * 5 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return LILOC_OK;
}

int liloc_feature_pass(liloc_ctx* ctx, int cls, const float* pose6, int stride, liloc_point* coeff, unsigned char* flag) {
    ENTER(ctx);
    if (bad_class(cls) || !pose6 || stride < 1 || !coeff || !flag) return fail(ctx, LILOC_ERR_ARG, "feature_pass: bad arguments");
    FeatClass& fc = ctx->fc[cls];
    if (!fc.have_map) return fail(ctx, LILOC_ERR_STATE, "feature_pass: no local map");
    if (!fc.have_scan || fc.Q_all < 0) return fail(ctx, LILOC_ERR_STATE, "feature_pass: no scan");
    const int Q = fc.Q_all;
    if (Q == 0) return LILOC_OK;
    int rc;
    if ((rc = ensure(ctx, ctx->tmp_pts, (size_t)Q))) return rc;
    if ((rc = ensure(ctx, ctx->tmp_b, (size_t)Q))) return rc;
    if ((rc = ensure(ctx, ctx->tmp_f, 6))) return rc;
    CU(cudaMemsetAsync(ctx->tmp_pts.p, 0, (size_t)Q * sizeof(float4), ctx->stream));
    CU(cudaMemsetAsync(ctx->tmp_b.p, 0, (size_t)Q, ctx->stream));
    CU(cudaMemcpyAsync(ctx->tmp_f.p, pose6, 6 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    S2mArgs a{};
    a.f[cls] = feat_view(fc, stride);
    a.enable_corner = cls;
    const int nq = (Q + stride - 1) / stride;
    k_feature_pass<<<blocks_for(ctx, nq, kQPB), kS2mThreads, 0, ctx->stream>>>(a, cls, ctx->tmp_f.p, ctx->tmp_pts.p, ctx->tmp_b.p); KCHECK();
    CU(cudaMemcpyAsync(coeff, ctx->tmp_pts.p, (size_t)Q * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(flag, ctx->tmp_b.p, (size_t)Q, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return LILOC_OK;
}

int liloc_surf_pass(liloc_ctx* ctx, const float* pose6, int stride, liloc_point* coeff, unsigned char* flag) {
    return liloc_feature_pass(ctx, LILOC_SURF, pose6, stride, coeff, flag);
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

int liloc_debug_eigen3(liloc_ctx* ctx, const float* A9, int n, float* w3, float* V9) {
    ENTER(ctx);
    if (n <= 0 || !A9 || !w3 || !V9) return fail(ctx, LILOC_ERR_ARG, "debug_eigen3: bad arguments");
    int rc;
    if ((rc = ensure(ctx, ctx->tmp_f, (size_t)n * 21))) return rc;
    float* dA = ctx->tmp_f.p; float* dw = dA + 9 * (size_t)n; float* dV = dw + 3 * (size_t)n;
    CU(cudaMemcpyAsync(dA, A9, (size_t)n * 9 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    k_debug_eigen3<<<(n + 63) / 64, 64, 0, ctx->stream>>>(dA, n, dw, dV); KCHECK();
    CU(cudaMemcpyAsync(w3, dw, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(V9, dV, (size_t)n * 9 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return LILOC_OK;
}

int liloc_debug_line(liloc_ctx* ctx, const float* pts15, const float* sel3, int n, liloc_point* coeff, int* keep) {
    ENTER(ctx);
    if (n <= 0 || !pts15 || !sel3 || !coeff || !keep) return fail(ctx, LILOC_ERR_ARG, "debug_line: bad arguments");
    int rc;
    if ((rc = ensure(ctx, ctx->tmp_f, (size_t)n * 18))) return rc;
    if ((rc = ensure(ctx, ctx->tmp_pts, (size_t)n))) return rc;
    if ((rc = ensure(ctx, ctx->tmp_i, (size_t)n))) return rc;
    CU(cudaMemcpyAsync(ctx->tmp_f.p, pts15, (size_t)n * 15 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->tmp_f.p + 15 * (size_t)n, sel3, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    k_debug_line<<<(n + 63) / 64, 64, 0, ctx->stream>>>(ctx->tmp_f.p, ctx->tmp_f.p + 15 * (size_t)n, n, ctx->tmp_pts.p, ctx->tmp_i.p); KCHECK();
    CU(cudaMemcpyAsync(coeff, ctx->tmp_pts.p, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(keep, ctx->tmp_i.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return LILOC_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// Global-map down-sample of the visualisation thread
// ---------------------------------------------------------------------------------------------------------
extern "C" {

int liloc_globalmap_build(liloc_ctx* ctx, const int* kf_ids, const float* poses6, int K, float leaf, liloc_point* out, int cap, int* M_out) {
    ENTER(ctx);
    if (K < 0 || (K > 0 && (!kf_ids || !poses6)) || !(leaf > 0.f) || !M_out) return fail(ctx, LILOC_ERR_ARG, "globalmap_build: bad arguments");
    FeatClass& ax = ctx->aux;                        // its own buffers: the frame's local map stays untouched
    int rc;
    std::vector<KfDesc> h((size_t)(K > 0 ? K : 1));
    long long total = 0;
    for (int k = 0; k < K; ++k) {
        auto it = ctx->kf.find(kf_ids[k]);
        if (it == ctx->kf.end()) return fail(ctx, LILOC_ERR_ARG, "globalmap_build: unknown keyframe id %d", kf_ids[k]);
        h[k].src_off = (long long)it->second.off[0]; h[k].n = it->second.n[0]; h[k].dst_off = (int)total;
        pose_to_T(poses6 + 6 * (size_t)k, h[k].T);
        total += it->second.n[0];
        if (total > 0x7fffffffLL) return fail(ctx, LILOC_ERR_CAPACITY, "globalmap_build: more than 2^31 points");
    }
    if ((rc = ensure(ctx, ax.descs, (size_t)(K > 0 ? K : 1)))) return rc;
    if ((rc = ensure(ctx, ax.raw, (size_t)(total > 0 ? total : 1)))) return rc;
    if (K > 0) CU(cudaMemcpyAsync(ax.descs.p, h.data(), (size_t)K * sizeof(KfDesc), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));          // the staging vector dies at return
    PipeArgs pa{};
    VgProb& q = pa.prob[0];
    if ((rc = prob_voxelgrid(ctx, ax.vg_map, ax.raw.p, (int)total, leaf, &ax.map, &q))) return rc;
    q.arena = ctx->arena.p; q.descs = ax.descs.p; q.K = total > 0 ? K : 0; q.raw = ax.raw.p;
    pa.n_prob = 1;
    if ((rc = pipe_launch(ctx, ctx->stream, pa, nullptr, ctx->wide_maps && total > ctx->wide_min))) return rc;
    int m = 0;
    CU(cudaMemcpyAsync(&m, ax.vg_map.d_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (pipe_check(ctx)) { const int e = ctx->pipe_failed; ctx->pipe_failed = 0; return e; }
    *M_out = m;
    if (out) {
        if (m > cap) return fail(ctx, LILOC_ERR_CAPACITY, "globalmap_build: cap %d < M %d", cap, m);
        if (m > 0) CU(cudaMemcpyAsync(out, ax.map.p, (size_t)m * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return LILOC_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// ICP fitness score
// ---------------------------------------------------------------------------------------------------------
namespace {

// cloud 1 (n1 points, transform T1) against cloud 2 (n2 points, transform T2); both already on the device
int icp_fitness_dev(liloc_ctx* ctx, const float4* d1, int n1, const float* pose6_1, const float4* d2, int n2, const float* pose6_2, float* score_out) {
    if (n1 <= 0 || n2 <= 0) return fail(ctx, LILOC_ERR_ARG, "icp_fitness: empty cloud");
    static const float kIdentity6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int rc;
    FeatClass& ax = ctx->aux;
    KfDesc desc{};
    desc.src_off = 0; desc.n = n2; desc.dst_off = 0;
    pose_to_T(pose6_2 ? pose6_2 : kIdentity6, desc.T);
    float T1[12];
    pose_to_T(pose6_1 ? pose6_1 : kIdentity6, T1);
    if ((rc = ensure(ctx, ax.descs, 1))) return rc;
    if ((rc = ensure(ctx, ax.raw, (size_t)n2))) return rc;
    if ((rc = ensure(ctx, ctx->aux_d2, (size_t)n1 + 16))) return rc;
    if ((rc = ensure(ctx, ctx->tmp_f, 32))) return rc;
    CU(cudaMemcpyAsync(ax.descs.p, &desc, sizeof(KfDesc), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->tmp_f.p, T1, sizeof(T1), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));          // desc / T1 are stack variables
    PipeArgs pa{};
    VgProb& q = pa.prob[0];
    if ((rc = prob_voxelgrid(ctx, ax.vg_map, ax.raw.p, n2, 1.0f, &ax.raw, &q))) return rc;       // leaf unused: grid_only
    q.arena = d2; q.descs = ax.descs.p; q.K = 1; q.raw = ax.raw.p; q.grid_only = 1; q.sort_only = 0;
    if ((rc = prob_add_grid(ctx, ax, n2, &q))) return rc;
    pa.n_prob = 1;
    if ((rc = pipe_launch(ctx, ctx->stream, pa, nullptr))) return rc;
    k_nn1_dist2<<<blocks_for(ctx, (long long)n1 * kKnnLanes, 256), 256, 0, ctx->stream>>>(d1, n1, ctx->tmp_f.p, grid_view(ax), ctx->aux_d2.p); KCHECK();
    k_seq_mean<<<1, 32, 0, ctx->stream>>>(ctx->aux_d2.p, n1, ctx->aux_d2.p + n1); KCHECK();
    CU(cudaMemcpyAsync(score_out, ctx->aux_d2.p + n1, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (pipe_check(ctx)) { const int e = ctx->pipe_failed; ctx->pipe_failed = 0; return e; }
    return LILOC_OK;
}

}  // namespace

extern "C" {

int liloc_icp_fitness(liloc_ctx* ctx, const liloc_point* cloud1, int n1, const float* pose6_1,
                      const liloc_point* cloud2, int n2, const float* pose6_2, float* score_out) {
    ENTER(ctx);
    if (n1 <= 0 || n2 <= 0 || !cloud1 || !cloud2 || !score_out) return fail(ctx, LILOC_ERR_ARG, "icp_fitness: bad arguments");
    int rc;
    if ((rc = ensure(ctx, ctx->tmp_pts, (size_t)n1 + (size_t)n2))) return rc;
    CU(cudaMemcpyAsync(ctx->tmp_pts.p, cloud1, (size_t)n1 * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->tmp_pts.p + n1, cloud2, (size_t)n2 * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    return icp_fitness_dev(ctx, ctx->tmp_pts.p, n1, pose6_1, ctx->tmp_pts.p + n1, n2, pose6_2, score_out);
}

int liloc_icp_fitness_resident(liloc_ctx* ctx, int kf_id, const float* pose6_kf, int ndt_source_id, const float* pose6_src, float* score_out) {
    ENTER(ctx);
    if (!score_out) return fail(ctx, LILOC_ERR_ARG, "icp_fitness: null argument");
    auto ki = ctx->kf.find(kf_id);
    auto si = ctx->ndt_sources.find(ndt_source_id);
    if (ki == ctx->kf.end()) return fail(ctx, LILOC_ERR_ARG, "icp_fitness: unknown keyframe id %d", kf_id);
    if (si == ctx->ndt_sources.end()) return fail(ctx, LILOC_ERR_ARG, "icp_fitness: unknown source id %d", ndt_source_id);
    return icp_fitness_dev(ctx, ctx->arena.p + ki->second.off[0], ki->second.n[0], pose6_kf, si->second.view.pts, si->second.view.n, pose6_src, score_out);
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
    d.max_iterations = o->max_iterations;
    d.search = o->search == LILOC_NDT_DIRECT7 ? 1 : (o->search == LILOC_NDT_KDTREE ? 2 : 0);
    d.radius2 = resolution * resolution;
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

// Voxel-Gaussian grid of the n points at ctx->tmp_pts (already on the device, or assembled there by the pipeline when
// kf != nullptr): voxel keys + stable sort (k_voxel_pipeline, sort_only), one leaf per voxel run, dense cell table.
static int ndt_target_build_into(liloc_ctx* ctx, liloc_ctx::NdtTargetHost& T, int n, float resolution, const VgProb* assemble, int* n_leaves_out) {
    int rc;
    {
        PipeArgs pa{};
        if ((rc = prob_voxelgrid(ctx, ctx->fc[0].vg_map, ctx->tmp_pts.p, n, resolution, nullptr, &pa.prob[0]))) return rc;
        if (assemble) { pa.prob[0].arena = assemble->arena; pa.prob[0].descs = assemble->descs; pa.prob[0].K = assemble->K; pa.prob[0].raw = ctx->tmp_pts.p; }
        pa.n_prob = 1;
        if ((rc = pipe_launch(ctx, ctx->stream, pa, nullptr))) return rc;
    }
    VgParams vp;
    CU(cudaMemcpyAsync(&vp, ctx->fc[0].vg_map.d_vp, sizeof(VgParams), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (pipe_check(ctx)) { const int e = ctx->pipe_failed; ctx->pipe_failed = 0; return e; }
    if (vp.overflow) return fail(ctx, LILOC_ERR_CAPACITY, "ndt_target_put: leaf size too small for the target extent (PCL overflow guard)");
    int divb[3];
    divb[0] = vp.mul[1]; divb[1] = vp.mul[1] ? vp.mul[2] / vp.mul[1] : 1;
    divb[2] = (int)floorf(vp.mx[2] * vp.inv_leaf) - vp.minb[2] + 1;
    const long long ncell = (long long)divb[0] * divb[1] * divb[2];
    if (ncell > (1LL << 28)) return fail(ctx, LILOC_ERR_CAPACITY, "ndt_target_put: %lld voxels exceed the dense table limit", ncell);
    if ((rc = ensure(ctx, T.table, (size_t)ncell))) return rc;
    if ((rc = ensure(ctx, T.leaves, (size_t)(n / 6 + 1)))) return rc;
    k_fill_int<<<blocks_for(ctx, ncell, 256 * 4), 256, 0, ctx->stream>>>(T.table.p, ncell, -1); KCHECK();
    CU(cudaMemsetAsync(ctx->d_ndt_active, 0, sizeof(int), ctx->stream));
    k_ndt_leaves<<<blocks_for(ctx, n, 256), 256, 0, ctx->stream>>>(ctx->fc[0].vg_map.keys_b.p, ctx->fc[0].vg_map.vals_b.p, n, ctx->tmp_pts.p, T.leaves.p, T.table.p, ctx->d_ndt_active); KCHECK();
    int nl = 0;
    CU(cudaMemcpyAsync(&nl, ctx->d_ndt_active, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    T.n_leaves = nl; T.resolution = resolution; T.n_points = n;
    T.view.table = T.table.p; T.view.leaves = T.leaves.p; T.view.inv_leaf = vp.inv_leaf;
    for (int a = 0; a < 3; ++a) { T.view.minb[a] = vp.minb[a]; T.view.divb[a] = divb[a]; T.view.maxb[a] = vp.minb[a] + divb[a] - 1; }
    if (n_leaves_out) *n_leaves_out = nl;
    return LILOC_OK;
}

// A failed build must not leave a half-made target behind (its view would be dereferenced on the device): the entry of
// `target_id` is replaced on success and removed on failure.
static int ndt_target_build(liloc_ctx* ctx, int target_id, int n, float resolution, const VgProb* assemble, int* n_leaves_out) {
    liloc_ctx::NdtTargetHost& T = ctx->ndt_targets[target_id];
    T.view = NdtTargetView{};
    ctx->ndt_views_dirty = true;
    const int rc = ndt_target_build_into(ctx, T, n, resolution, assemble, n_leaves_out);
    if (rc != LILOC_OK) {
        cudaStreamSynchronize(ctx->stream);
        if (T.table.p) cudaFree(T.table.p);
        if (T.leaves.p) cudaFree(T.leaves.p);
        ctx->ndt_targets.erase(target_id);
    }
    return rc;
}

int liloc_ndt_target_put(liloc_ctx* ctx, int target_id, const liloc_point* pts, int n, float resolution, int* n_leaves_out) {
    ENTER(ctx);
    if (n <= 0 || !pts || !(resolution > 0.f)) return fail(ctx, LILOC_ERR_ARG, "ndt_target_put: bad arguments");
    int rc;
    if ((rc = ensure(ctx, ctx->tmp_pts, (size_t)n))) return rc;
    CU(cudaMemcpyAsync(ctx->tmp_pts.p, pts, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    return ndt_target_build(ctx, target_id, n, resolution, nullptr, n_leaves_out);
}

// generateSubMaps (include/dataManager/dataLoader.hpp:306-348) for one submap, on the device: the K prior keyframes
// (already in the arena, liloc_keyframe_put) are transformed by their prior poses and concatenated -- the submap the
// reference keeps un-down-sampled (:332-336) -- and its voxel-Gaussian grid is built.  Nothing crosses PCIe.
int liloc_ndt_target_put_keyframes(liloc_ctx* ctx, int target_id, const int* kf_ids, const float* poses6, int K, float resolution,
                                   int* n_leaves_out, int* n_points_out) {
    ENTER(ctx);
    if (K <= 0 || !kf_ids || !poses6 || !(resolution > 0.f)) return fail(ctx, LILOC_ERR_ARG, "ndt_target_put_keyframes: bad arguments");
    int rc;
    std::vector<KfDesc> h((size_t)K);
    long long total = 0;
    for (int k = 0; k < K; ++k) {
        auto it = ctx->kf.find(kf_ids[k]);
        if (it == ctx->kf.end()) return fail(ctx, LILOC_ERR_ARG, "ndt_target_put_keyframes: unknown keyframe id %d", kf_ids[k]);
        h[k].src_off = (long long)it->second.off[0]; h[k].n = it->second.n[0]; h[k].dst_off = (int)total;
        pose_to_T(poses6 + 6 * (size_t)k, h[k].T);
        total += it->second.n[0];
        if (total > 0x7fffffffLL) return fail(ctx, LILOC_ERR_CAPACITY, "ndt_target_put_keyframes: more than 2^31 points");
    }
    if (total <= 0) return fail(ctx, LILOC_ERR_ARG, "ndt_target_put_keyframes: empty submap");
    if ((rc = ensure(ctx, ctx->ndt_descs, (size_t)K))) return rc;
    if ((rc = ensure(ctx, ctx->tmp_pts, (size_t)total))) return rc;
    CU(cudaMemcpyAsync(ctx->ndt_descs.p, h.data(), (size_t)K * sizeof(KfDesc), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));          // the staging vector dies at return
    VgProb asmb{};
    asmb.arena = ctx->arena.p; asmb.descs = ctx->ndt_descs.p; asmb.K = K;
    if (n_points_out) *n_points_out = (int)total;
    return ndt_target_build(ctx, target_id, (int)total, resolution, &asmb, n_leaves_out);
}

// The submap addNewPrior cuts after every submap_size added keyframes (anchorOptimize.hpp:341-377): transform + concatenate
// like generateSubMaps, but this one IS voxel-filtered (downSizeFilterSurf at 0.2 writes into `submap` itself, :363-367)
// before it becomes an NDT target.  Two pipeline launches on the device: assemble + VoxelGrid(filter_leaf) in the
// auxiliary buffers, then keys + sort of the filtered cloud at `resolution`.
int liloc_ndt_target_put_keyframes_filtered(liloc_ctx* ctx, int target_id, const int* kf_ids, const float* poses6, int K, float filter_leaf,
                                            float resolution, int* n_leaves_out, int* n_points_out) {
    ENTER(ctx);
    if (K <= 0 || !kf_ids || !poses6 || !(resolution > 0.f) || !(filter_leaf > 0.f)) return fail(ctx, LILOC_ERR_ARG, "ndt_target_put_keyframes_filtered: bad arguments");
    int M = 0, rc;
    if ((rc = liloc_globalmap_build(ctx, kf_ids, poses6, K, filter_leaf, nullptr, 0, &M))) return rc;
    if (M <= 0) return fail(ctx, LILOC_ERR_ARG, "ndt_target_put_keyframes_filtered: empty submap");
    if ((rc = ensure(ctx, ctx->tmp_pts, (size_t)M))) return rc;
    CU(cudaMemcpyAsync(ctx->tmp_pts.p, ctx->aux.map.p, (size_t)M * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
    if (n_points_out) *n_points_out = M;
    return ndt_target_build(ctx, target_id, M, resolution, nullptr, n_leaves_out);
}

int liloc_ndt_target_size(liloc_ctx* ctx, int target_id) {
    if (!ctx) return LILOC_ERR_ARG;
    auto it = ctx->ndt_targets.find(target_id);
    return it == ctx->ndt_targets.end() ? LILOC_ERR_ARG : it->second.n_points;
}

int liloc_ndt_source_put(liloc_ctx* ctx, int source_id, const liloc_point* pts, int n, int on_device) {
    ENTER(ctx);
    if (on_device == LILOC_SCAN_DESKEWED) {          // the de-skewed sweep this context holds (setInputSource(laserCloudSurfLast), :274)
        int rc;
        if ((rc = deskew_count(ctx))) return rc;
        pts = reinterpret_cast<const liloc_point*>(ctx->dk[ctx->dk_cur].out.p);
        n = ctx->dk[ctx->dk_cur].n;
        CU(cudaStreamWaitEvent(ctx->stream, ctx->dk[ctx->dk_cur].ev_done, 0));
    }
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

// One batch of alignments = one launch of the queue-driven persistent kernel (ndt.cuh), in two halves so that a caller can
// put the gather between them and synchronise ONCE:
//   ndt_enqueue_batch  job records up (pinned staging), queue memory cleared, the kernel, the queue header (counters +
//                      profile words) and -- when `fetch_jobs` -- the job records back into pinned staging; nothing waits
//   ndt_finish_batch   after the caller synchronised the stream: error words, counters, job records
// The packed 8-float rows stay in ctx->ndt_rows (device) for the gather.
constexpr int kNdtHdrWords = 4 + 16;      // head / tail / finished / error, then the 16 profile words; the ticket ring follows

static int ndt_enqueue_batch(liloc_ctx* ctx, const liloc_ndt_job* jobs, int n_jobs, const liloc_ndt_opts* o, bool fetch_jobs) {
    int rc;
    if ((rc = ndt_sync_views(ctx))) return rc;
    if (n_jobs > kNdtMaxJobs) return fail(ctx, LILOC_ERR_CAPACITY, "ndt_align_batch: more than %d jobs in one batch", kNdtMaxJobs);
    if ((size_t)n_jobs > ctx->h_ndt_jobs_cap) {
        if (ctx->h_ndt_jobs) { CU(cudaFreeHost(ctx->h_ndt_jobs)); ctx->h_ndt_jobs = nullptr; ctx->h_ndt_jobs_cap = 0; }
        size_t cap = 256; while (cap < (size_t)n_jobs) cap *= 2;
        CU(cudaMallocHost(&ctx->h_ndt_jobs, cap * sizeof(NdtJob)));
        ctx->h_ndt_jobs_cap = cap;
    }
    NdtJob* hj = ctx->h_ndt_jobs;
    std::memset(hj, 0, (size_t)n_jobs * sizeof(NdtJob));
    int max_n = 0;
    long long total_tiles = 0;
    float resolution = -1.f;
    int last_t = INT_MIN, last_s = INT_MIN;
    const liloc_ctx::NdtTargetHost* tp = nullptr; const liloc_ctx::NdtSourceHost* sp = nullptr;
    for (int j = 0; j < n_jobs; ++j) {
        if (jobs[j].target_id != last_t || !tp) {           // neighbouring jobs mostly share their target / source: one hash look-up per change
            auto ti = ctx->ndt_targets.find(jobs[j].target_id);
            if (ti == ctx->ndt_targets.end()) return fail(ctx, LILOC_ERR_ARG, "ndt_align_batch: unknown target id %d", jobs[j].target_id);
            tp = &ti->second; last_t = jobs[j].target_id;
        }
        if (jobs[j].source_id != last_s || !sp) {
            auto si = ctx->ndt_sources.find(jobs[j].source_id);
            if (si == ctx->ndt_sources.end()) return fail(ctx, LILOC_ERR_ARG, "ndt_align_batch: unknown source id %d", jobs[j].source_id);
            sp = &si->second; last_s = jobs[j].source_id;
        }
        if (!tp->view.table) return fail(ctx, LILOC_ERR_STATE, "ndt_align_batch: target id %d was not built", jobs[j].target_id);
        if (resolution < 0.f) resolution = tp->resolution;
        else if (resolution != tp->resolution) return fail(ctx, LILOC_ERR_ARG, "ndt_align_batch: all targets of one batch must share the resolution");
        hj[j].target = tp->slot; hj[j].source = sp->slot;
        hj[j].n_tiles = (sp->view.n + kNdtTile - 1) / kNdtTile;
        if (hj[j].n_tiles > kNdtMaxTilesPerJob) return fail(ctx, LILOC_ERR_CAPACITY, "ndt_align_batch: source %d has more than %d points", jobs[j].source_id, kNdtMaxTilesPerJob * kNdtTile);
        std::memcpy(hj[j].guess, jobs[j].guess, 12 * sizeof(float));     // top 3 rows of the row-major 4x4
        if (sp->view.n > max_n) max_n = sp->view.n;
        total_tiles += hj[j].n_tiles;
    }
    const NdtOptsDev od = ndt_opts_dev(o, resolution);
    const int tiles = (max_n + kNdtTile - 1) / kNdtTile;
    long long blocks = total_tiles < ctx->ndt_max_blocks ? total_tiles : ctx->ndt_max_blocks;
    if (blocks < 1) blocks = 1;
    size_t cap = 1024;
    while (cap < 2 * ((size_t)total_tiles + (size_t)blocks)) cap *= 2;
    if ((rc = ensure(ctx, ctx->d_jobs, (size_t)n_jobs))) return rc;
    if ((rc = ensure(ctx, ctx->ndt_partial, (size_t)n_jobs * tiles * kNdtSums))) return rc;
    if ((rc = ensure(ctx, ctx->ndt_hot, (size_t)n_jobs))) return rc;
    if ((rc = ensure(ctx, ctx->ndt_sync, (size_t)n_jobs))) return rc;
    if ((rc = ensure(ctx, ctx->ndt_rows, (size_t)n_jobs * 8))) return rc;
    if ((rc = ensure(ctx, ctx->ndt_q, kNdtHdrWords + cap))) return rc;
    CU(cudaMemcpyAsync(ctx->d_jobs.p, hj, (size_t)n_jobs * sizeof(NdtJob), cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += (long long)n_jobs * (long long)sizeof(NdtJob);
    CU(cudaMemsetAsync(ctx->ndt_q.p, 0, (kNdtHdrWords + cap) * sizeof(unsigned long long), ctx->stream));     // counters, profile words, every ticket tag
    NdtAlignArgs aa{};
    aa.jobs = ctx->d_jobs.p; aa.n_jobs = n_jobs; aa.targets = ctx->d_tviews.p; aa.sources = ctx->d_sviews.p; aa.o = od;
    aa.tiles_max = tiles; aa.partial = ctx->ndt_partial.p; aa.hot = ctx->ndt_hot.p; aa.sync = ctx->ndt_sync.p;
    unsigned* qw = reinterpret_cast<unsigned*>(ctx->ndt_q.p);
    aa.q.head = qw; aa.q.tail = qw + 2; aa.q.finished = reinterpret_cast<int*>(qw + 4); aa.q.error = reinterpret_cast<int*>(qw + 6);
    aa.q.ring = ctx->ndt_q.p + kNdtHdrWords; aa.q.mask = (unsigned)(cap - 1);
    aa.rows = ctx->ndt_rows.p;
    aa.prof = ctx->timing ? ctx->ndt_q.p + 4 : nullptr;
    aa.grain_div = ctx->ndt_grain_div;
    k_ndt_align<<<(unsigned)blocks, kNdtThreads, 0, ctx->stream>>>(aa); KCHECK();
    CU(cudaMemcpyAsync(ctx->h_ndt_hdr, ctx->ndt_q.p, kNdtHdrWords * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    if (fetch_jobs) {
        CU(cudaMemcpyAsync(hj, ctx->d_jobs.p, (size_t)n_jobs * sizeof(NdtJob), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->stats.d2h_bytes += (long long)n_jobs * (long long)sizeof(NdtJob);
    }
    ctx->ndt_pending_jobs = n_jobs; ctx->ndt_pending_blocks = (int)blocks;
    ctx->ndt_pending_cells = od.search == 2 ? 27 : (od.search == 1 ? 7 : 1);
    return LILOC_OK;
}

// after the stream was synchronised
static int ndt_finish_batch(liloc_ctx* ctx) {
    const unsigned long long* h = ctx->h_ndt_hdr;
    const int n_jobs = ctx->ndt_pending_jobs;
    if ((int)(h[3] & 0xFFFFFFFFu) != 0) return fail(ctx, LILOC_ERR_CUDA, "ndt_align_batch: a block gave up waiting for a ticket (queue stalled)");
    if ((int)(h[2] & 0xFFFFFFFFu) != n_jobs) return fail(ctx, LILOC_ERR_CUDA, "ndt_align_batch: %d of %d jobs retired", (int)(h[2] & 0xFFFFFFFFu), n_jobs);
    liloc_ndt_stats& ns = ctx->ndt_stats;
    ns.batches += 1; ns.jobs += n_jobs; ns.launches += 1; ns.blocks += ctx->ndt_pending_blocks;
    if (ctx->timing) {
        const unsigned long long* pr = h + 4;
        const unsigned long long t0 = ~pr[0];
        ns.kernel_ms += pr[1] > t0 ? (double)(pr[1] - t0) * 1e-6 : 0.0;
        ns.block_sweep_ms += (double)pr[2] * 1e-6; ns.block_update_ms += (double)pr[3] * 1e-6; ns.block_wait_ms += (double)pr[4] * 1e-6;
        ns.tickets += (long long)pr[5]; ns.updates += (long long)pr[6];
        for (int k = 0; k < 4; ++k) ns.update_cycles[k] += (double)pr[8 + k];
        ns.job_sweeps += (long long)pr[12];
        if ((long long)pr[13] > ns.max_job_sweeps) ns.max_job_sweeps = (long long)pr[13];
        // per sweep and job, the scan once (16 P), a (mean, inverse covariance) gather per point and cell (36 P c)
        // and 28 doubles out; pr[14] = sum over jobs of sweeps x P
        ns.alg_bytes += (double)pr[14] * (16.0 + 36.0 * ctx->ndt_pending_cells) + (double)pr[12] * 224.0;
    }
    return LILOC_OK;
}

static void ndt_fill_outs(liloc_ctx* ctx, const liloc_ndt_job* jobs, int n_jobs, liloc_ndt_out* outs) {
    const NdtJob* hj = ctx->h_ndt_jobs;
    for (int j = 0; j < n_jobs; ++j) {
        liloc_ndt_out& r = outs[j];
        std::memcpy(r.T, hj[j].Tfinal, 12 * sizeof(float));
        r.T[12] = 0.f; r.T[13] = 0.f; r.T[14] = 0.f; r.T[15] = 1.f;
        r.score = hj[j].score; r.iterations = hj[j].total_iterations; r.converged = hj[j].converged; r.sweeps = hj[j].total_sweeps;
        r.n_source = ctx->ndt_sources[jobs[j].source_id].view.n;
    }
}

int liloc_ndt_align_batch(liloc_ctx* ctx, const liloc_ndt_job* jobs, int n_jobs, const liloc_ndt_opts* opts, liloc_ndt_out* outs) {
    ENTER(ctx);
    if (n_jobs <= 0 || !jobs || !outs) return fail(ctx, LILOC_ERR_ARG, "ndt_align_batch: bad arguments");
    const liloc_ndt_opts* o = opts ? opts : &kDefaultNdtOpts;
    int rc;
    if ((rc = ndt_enqueue_batch(ctx, jobs, n_jobs, o, true))) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    if ((rc = ndt_finish_batch(ctx))) return rc;
    ndt_fill_outs(ctx, jobs, n_jobs, outs);
    return LILOC_OK;
}

// ---- multi-GPU: sharded batches + the final gather (SURVEY.md 8b liloc_gather_results, 8e) ---------------------------------
int liloc_comm_unique_id(void* id128) {
    if (!id128) return LILOC_ERR_ARG;
    const std::string e = nccl().load();
    if (!e.empty()) { g_create_error = e; return LILOC_ERR_STATE; }
    NcclShim::unique_id id;
    const int r = nccl().GetUniqueId(&id);
    if (r != 0) { g_create_error = std::string("ncclGetUniqueId: ") + nccl().GetErrorString(r); return LILOC_ERR_CUDA; }
    std::memcpy(id128, &id, sizeof(id));
    return LILOC_OK;
}

int liloc_comm_init(liloc_ctx* ctx, int rank, int world, const void* id128) {
    ENTER(ctx);
    if (world < 1 || rank < 0 || rank >= world || (world > 1 && !id128)) return fail(ctx, LILOC_ERR_ARG, "comm_init: bad arguments");
    if (ctx->comm) { nccl().CommDestroy(ctx->comm); ctx->comm = nullptr; }
    ctx->comm_rank = rank; ctx->comm_world = world;
    if (world == 1) return LILOC_OK;
    const std::string e = nccl().load();
    if (!e.empty()) return fail(ctx, LILOC_ERR_STATE, "comm_init: %s", e.c_str());
    NcclShim::unique_id id;
    std::memcpy(&id, id128, sizeof(id));
    const int r = nccl().CommInitRank(&ctx->comm, world, id, rank);
    if (r != 0) { ctx->comm = nullptr; return fail(ctx, LILOC_ERR_CUDA, "ncclCommInitRank: %s", nccl().GetErrorString(r)); }
    return LILOC_OK;
}

int liloc_comm_destroy(liloc_ctx* ctx) {
    ENTER(ctx);
    if (ctx->comm) { CU(cudaStreamSynchronize(ctx->stream)); nccl().CommDestroy(ctx->comm); ctx->comm = nullptr; }
    ctx->comm_rank = 0; ctx->comm_world = 1;
    return LILOC_OK;
}

// the share of `rank`: item indices in the order their rows travel (no context, no device: pure index arithmetic)
int liloc_shard_indices(int n_items, int rank, int world, int policy, int* out, int cap) {
    if (n_items < 0 || world < 1 || rank < 0 || rank >= world || (policy != LILOC_SHARD_BLOCKS && policy != LILOC_SHARD_ROUND_ROBIN)) return LILOC_ERR_ARG;
    int n = 0;
    if (policy == LILOC_SHARD_ROUND_ROBIN) {
        for (int i = rank; i < n_items; i += world, ++n) if (out && n < cap) out[n] = i;
    } else {
        const int base = n_items / world, rem = n_items % world;
        const int lo = rank * base + (rank < rem ? rank : rem), hi = lo + base + (rank < rem ? 1 : 0);
        for (int i = lo; i < hi; ++i, ++n) if (out && n < cap) out[n] = i;
    }
    return n;
}

// gathered rows [world][per_rank][8] (rank-major, share order) -> rows_out[n_items][8] in item order
int liloc_unshard_rows(const float* rows_all, int n_items, int world, int per_rank, int policy, float* rows_out) {
    if (n_items < 0 || world < 1 || per_rank < 0 || (n_items > 0 && (!rows_all || !rows_out)) || (long long)per_rank * world < n_items) return LILOC_ERR_ARG;
    std::vector<int> idx((size_t)(per_rank > 0 ? per_rank : 1));
    for (int r = 0; r < world; ++r) {
        const int n = liloc_shard_indices(n_items, r, world, policy, idx.data(), per_rank);
        if (n < 0 || n > per_rank) return LILOC_ERR_ARG;
        for (int k = 0; k < n; ++k) std::memcpy(rows_out + (size_t)idx[k] * 8, rows_all + ((size_t)r * per_rank + k) * 8, 8 * sizeof(float));
    }
    return LILOC_OK;
}

int liloc_comm_rank(const liloc_ctx* ctx) { return ctx ? ctx->comm_rank : LILOC_ERR_ARG; }
int liloc_comm_world(const liloc_ctx* ctx) { return ctx ? ctx->comm_world : LILOC_ERR_ARG; }

// all-gather of `per_rank` rows of 8 floats from every rank's device buffer `d_send` -> host rows_all[world][per_rank][8]
static int gather_rows_dev(liloc_ctx* ctx, const float* d_send, int per_rank, float* rows_all) {
    const int W = ctx->comm_world;
    const size_t cnt = (size_t)per_rank * 8;
    int rc;
    if (cnt == 0) return LILOC_OK;
    if (cnt * W > ctx->h_gather_cap) {
        if (ctx->h_gather) { CU(cudaFreeHost(ctx->h_gather)); ctx->h_gather = nullptr; ctx->h_gather_cap = 0; }
        size_t cap = 4096; while (cap < cnt * W) cap *= 2;
        CU(cudaMallocHost(&ctx->h_gather, cap * sizeof(float)));
        ctx->h_gather_cap = cap;
    }
    const float* src = d_send;
    if (W > 1) {
        if (!ctx->comm) return fail(ctx, LILOC_ERR_STATE, "gather: liloc_comm_init was not called");
        if ((rc = ensure(ctx, ctx->gather_recv, cnt * W))) return rc;
        const int r = nccl().AllGather(d_send, ctx->gather_recv.p, cnt, NcclShim::kFloat, ctx->comm, ctx->stream);
        if (r != 0) return fail(ctx, LILOC_ERR_CUDA, "ncclAllGather: %s", nccl().GetErrorString(r));
        src = ctx->gather_recv.p;
    }
    CU(cudaMemcpyAsync(ctx->h_gather, src, cnt * W * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->stats.d2h_bytes += (long long)(cnt * W * sizeof(float));
    std::memcpy(rows_all, ctx->h_gather, cnt * W * sizeof(float));
    return LILOC_OK;
}

int liloc_gather_results(liloc_ctx* ctx, const float* rows_local, int n_local, int per_rank, float* rows_all) {
    ENTER(ctx);
    if (per_rank < 0 || n_local < 0 || n_local > per_rank || (n_local > 0 && !rows_local) || (per_rank > 0 && !rows_all)) return fail(ctx, LILOC_ERR_ARG, "gather_results: bad arguments");
    if (per_rank == 0) return LILOC_OK;
    int rc;
    if ((rc = ensure(ctx, ctx->gather_send, (size_t)per_rank * 8))) return rc;
    CU(cudaMemsetAsync(ctx->gather_send.p, 0, (size_t)per_rank * 8 * sizeof(float), ctx->stream));
    if (n_local > 0) CU(cudaMemcpyAsync(ctx->gather_send.p, rows_local, (size_t)n_local * 8 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += (long long)n_local * 32;
    return gather_rows_dev(ctx, ctx->gather_send.p, per_rank, rows_all);
}

int liloc_ndt_align_batch_sharded(liloc_ctx* ctx, const liloc_ndt_job* items, int n_items, const liloc_ndt_opts* opts, int policy,
                                  float* rows_out, liloc_ndt_out* local_outs) {
    ENTER(ctx);
    if (n_items <= 0 || !items || !rows_out || (policy != LILOC_SHARD_BLOCKS && policy != LILOC_SHARD_ROUND_ROBIN)) return fail(ctx, LILOC_ERR_ARG, "ndt_align_batch_sharded: bad arguments");
    const liloc_ndt_opts* o = opts ? opts : &kDefaultNdtOpts;
    const int W = ctx->comm_world, R = ctx->comm_rank;
    const int per = (n_items + W - 1) / W;
    // this rank's share, in the order its rows are sent
    std::vector<liloc_ndt_job> mine;
    auto share_of = [&](int r, std::vector<int>& idx) {
        idx.resize((size_t)per);
        idx.resize((size_t)liloc_shard_indices(n_items, r, W, policy, idx.data(), per));
    };
    std::vector<int> idx;
    share_of(R, idx);
    for (int i : idx) mine.push_back(items[i]);
    int rc;
    if ((rc = ensure(ctx, ctx->ndt_rows, (size_t)per * 8))) return rc;
    // enqueue everything -- the batch kernel, the padding of a short share, the all-gather, the copies back -- and wait ONCE
    if (!mine.empty() && (rc = ndt_enqueue_batch(ctx, mine.data(), (int)mine.size(), o, local_outs != nullptr))) return rc;
    if ((int)mine.size() < per)         // pad rows of a short share
        CU(cudaMemsetAsync(ctx->ndt_rows.p + mine.size() * 8, 0, ((size_t)per - mine.size()) * 8 * sizeof(float), ctx->stream));
    std::vector<float> all((size_t)W * per * 8);
    if ((rc = gather_rows_dev(ctx, ctx->ndt_rows.p, per, all.data()))) return rc;          // synchronises the stream
    if (!mine.empty()) {
        if ((rc = ndt_finish_batch(ctx))) return rc;
        if (local_outs) ndt_fill_outs(ctx, mine.data(), (int)mine.size(), local_outs);
    }
    if (liloc_unshard_rows(all.data(), n_items, W, per, policy, rows_out) != LILOC_OK) return fail(ctx, LILOC_ERR_ARG, "ndt_align_batch_sharded: unshard failed");
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
    const int n = si->second.view.n, tiles = (n + kNdtTile - 1) / kNdtTile;
    if ((rc = ensure(ctx, ctx->d_jobs, 1))) return rc;
    if ((rc = ensure(ctx, ctx->ndt_partial, (size_t)tiles * kNdtSums))) return rc;
    CU(cudaMemcpyAsync(ctx->d_jobs.p, &J, sizeof(NdtJob), cudaMemcpyHostToDevice, ctx->stream));
    k_ndt_sweep<<<tiles, kNdtThreads, 0, ctx->stream>>>(ctx->d_jobs.p, ctx->d_tviews.p, ctx->d_sviews.p, od, ctx->ndt_partial.p); KCHECK();
    std::vector<double> part((size_t)tiles * kNdtSums);
    CU(cudaMemcpyAsync(part.data(), ctx->ndt_partial.p, part.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < kNdtSums; ++k) { double v = 0.0; for (int t = 0; t < tiles; ++t) v += part[(size_t)t * kNdtSums + k]; out28[k] = v; }
    return LILOC_OK;
}

}  // extern "C"

extern "C" {

int liloc_ndt_stats_get(const liloc_ctx* ctx, liloc_ndt_stats* out) {
    if (!ctx || !out) return LILOC_ERR_ARG;
    *out = ctx->ndt_stats;
    return LILOC_OK;
}

int liloc_ndt_stats_reset(liloc_ctx* ctx) {
    if (!ctx) return LILOC_ERR_ARG;
    ctx->ndt_stats = liloc_ndt_stats{};
    return LILOC_OK;
}

int liloc_last_marks(liloc_ctx* ctx, unsigned long long* out, int cap) {
    if (!ctx || !out || cap < 0) return LILOC_ERR_ARG;
    const int n = cap < 2 * kPipeMarks + kS2mMarks ? cap : 2 * kPipeMarks + kS2mMarks;
    std::memcpy(out, ctx->h_info->marks, (size_t)n * sizeof(unsigned long long));
    return n;
}

int liloc_set_share(liloc_ctx* ctx, int share) {
    if (!ctx) return LILOC_ERR_ARG;
    const int prev = ctx->share;
    apply_share(ctx, share);
    return prev;
}

long long liloc_set_map_cache(liloc_ctx* ctx, int enabled) {
    if (!ctx) return LILOC_ERR_ARG;
    ctx->map_cache = enabled != 0;
    if (!ctx->map_cache) for (auto& k : ctx->map_key) k.valid = false;
    return ctx->map_cache_hits;
}

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
