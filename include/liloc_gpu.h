/* include/liloc_gpu.h -- C ABI of the B200-native scan-to-map registration path of LiLoc.
 *
 * The reference (Yixin-F/LiLoc, /root/reference) has no plugin or FFI layer: the path is a set of
 * member functions of `class mapOptimization` (src/mapOptmization.cpp:18) that call PCL, Eigen and
 * OpenCV in-process, plus the pcl::Registration virtual interface for NDT (:118).  This header is
 * the boundary a maintainer would bind instead: each entry point names the reference lines it
 * replaces.  Plain pointers and sizes only; no C++ types, no exceptions, no global state.
 *
 * Conventions
 *   point        {x, y, z, intensity} float32, 16 bytes (pcl::PointXYZI without its padding,
 *                include/utility.h:83).  All input coordinates must be finite.
 *   pose6        [roll, pitch, yaw, x, y, z] float32 = transformTobeMapped (src/mapOptmization.cpp:85),
 *                R = Rz(yaw) Ry(pitch) Rx(roll)  (pcl::getTransformation, :404-406).
 *   return code  0 success; > 0 benign status (LILOC_SKIPPED ...); < 0 error, text from
 *                liloc_last_error().  A context is NOT thread-safe (the reference serialises the
 *                path under `mtx`, :255); use one context per calling thread / per GPU.
 *   host/device  pointers are HOST pointers unless the name ends in _dev.
 *
 * There is no CPU fallback: every compute entry point fails with LILOC_ERR_CUDA when no sm_100
 * device is usable.
 */
#ifndef LILOC_GPU_H
#define LILOC_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LILOC_OK              0
#define LILOC_SKIPPED         1   /* scan2map gate Q_all <= min_points (src/mapOptmization.cpp:1187,1205) */
#define LILOC_ERR_ARG        -1
#define LILOC_ERR_CUDA       -2
#define LILOC_ERR_OOM        -3
#define LILOC_ERR_STATE      -4   /* e.g. scan2map before a map / scan exists */
#define LILOC_ERR_CAPACITY   -5

typedef struct { float x, y, z, intensity; } liloc_point;

typedef struct liloc_ctx liloc_ctx;

/* Feature classes.  LILOC_SURF is the reference's only class (the whole de-skewed scan, plane residuals,
 * surfOptimization :981-1048).  LILOC_CORNER is the upstream LIO-SAM cornerOptimization the north_star names
 * (line residuals); the reference dropped it (only remnants: laserCloudCornerTemp :947,950), so it is an
 * optional superset: off by default, and every entry point without a class argument is surf-only. */
#define LILOC_SURF   0
#define LILOC_CORNER 1

typedef struct {
    int device;             /* CUDA device ordinal */
    int max_scan_points;    /* N_SCAN * Horizon_SCAN, sizes scratch like allocateMemory() :228-230 */
    int max_raw_points;     /* initial capacity of the concatenated local map (grows on demand) */
    int map_cache;          /* [0] 1 = keep the local map + search grid of the previous frame when the selected keyframes,
                               their poses and the leaf size are bit-identical (the result is a pure function of those
                               inputs, so outputs do not change; the reference recomputes every frame).  See also
                               liloc_set_map_cache. */
    int share;              /* [0 -> 1] this context is one of `share` contexts (sequences) driving the same GPU at once, one
                               host thread each: every launch takes 1 / share of the blocks it would otherwise use, so that
                               the frames of the sequences are co-resident (throughput mode; results are unchanged).  0: the
                               LILOC_SHARE environment variable, else 1.  See also liloc_set_share. */
    int reserved[3];
} liloc_config;

/* scan2MapOptimization knobs.  Reference values in brackets. */
typedef struct {
    int max_iters;          /* [20]  loop bound, src/mapOptmization.cpp:1190 */
    int stride;             /* [2]   every 2nd down-sampled point, :985 */
    int min_sel;            /* [50]  LMOptimization bail-out, :1080 */
    int min_points;         /* [30]  gate is Q_all > min_points, :1187 (upstream surfFeatureMinValidNum 100) */
    int enable_corner;      /* [0]   1 = also use the LILOC_CORNER class (upstream LIO-SAM: 30 iterations, stride 1) */
    int min_points_corner;  /* [10]  upstream edgeFeatureMinValidNum: gate is Q_corner > min_points_corner */
    int stride_corner;      /* [1] */
    int reserved[1];
} liloc_s2m_opts;

typedef struct {
    int   iters;            /* loop iterations executed (1..max_iters) */
    int   converged;        /* LMOptimization returned true, :1177-1179 */
    int   n_sel;            /* correspondences in the last executed iteration (laserCloudSelNum, :1079) */
    int   is_degenerate;    /* member isDegenerate (:90), consumed by publishOdometry (:1573) */
    int   too_few;          /* last iteration had fewer than min_sel correspondences */
    int   skipped;          /* gate failed: pose untouched */
    int   q_all;            /* laserCloudSurfLastDSNum */
    int   map_points;       /* laserCloudSurfFromMapDSNum */
    float matP[36];         /* member matP (:91), row-major */
    float AtA0[36];         /* J^T J of iteration 0 (diagnostic; input of the degeneracy check) */
} liloc_s2m_result;

/* ---- lifetime: allocateMemory() :213-244, VoxelGrid set-up :158-160 ------------------------- */
int  liloc_create(const liloc_config* cfg, liloc_ctx** out);
int  liloc_set_share(liloc_ctx* ctx, int share);     /* returns the previous value; takes effect from the next launch */
void liloc_destroy(liloc_ctx* ctx);
const char* liloc_last_error(const liloc_ctx* ctx);   /* ctx may be NULL: last create() error */
const char* liloc_version(void);

/* ---- keyframe store: surfCloudKeyFrames.push_back :1385,1457 + laserCloudMapContainer :69 -------
 * The library keeps a device copy of every keyframe cloud (body frame); the caller keeps its own. */
int liloc_keyframe_put(liloc_ctx* ctx, int kf_id, const liloc_point* pts_body, int n);
/* thisSurfKeyFrame = copy of laserCloudSurfLastDS (:1382,1454): device-to-device, no host traffic */
int liloc_keyframe_put_from_scan(liloc_ctx* ctx, int kf_id);
/* priorSession_->keyCloudVec_.push_back(copy of curCloud_) in AnchorOptimization::addNewPrior (include/egoOptimization/
 * anchorOptimize.hpp:279-282, :318-321): the FULL de-skewed scan of the current frame (laserCloudSurfLast,
 * src/mapOptmization.cpp:1342) becomes keyframe kf_id, device-to-device */
int liloc_keyframe_put_from_raw_scan(liloc_ctx* ctx, int kf_id);
int liloc_keyframe_size(liloc_ctx* ctx, int kf_id);    /* points, or <0 if unknown id */
int liloc_keyframe_clear(liloc_ctx* ctx);              /* laserCloudMapContainer.clear() + new session */

/* ---- local map: extractCloud :936-961 + transformPointCloud include/utility.h:388-407 -----------
 * Transforms the K selected keyframes by their poses and concatenates them IN THE GIVEN ORDER
 * (duplicates allowed, :924-931), applies pcl::VoxelGrid semantics at `leaf`
 * (surroundingKeyframeMapLeafSize) and builds the search structure that replaces
 * kdtreeSurfFromMap->setInputCloud (:1188).  The map stays on the device.
 * M_out / n_raw_out may be NULL (then no host synchronisation happens here). */
int liloc_localmap_build(liloc_ctx* ctx, const int* kf_ids, const float* poses6, int K, float leaf,
                         int* M_out, int* n_raw_out);
/* Same, from an explicit map-frame cloud (tests; prior submaps): VoxelGrid + search structure. */
int liloc_localmap_set(liloc_ctx* ctx, const liloc_point* pts_map, int n, float leaf, int* M_out);
int liloc_localmap_get(liloc_ctx* ctx, liloc_point* out, int cap, int* M_out);   /* laserCloudSurfFromMapDS */

/* ---- current scan: downsampleCurrentScan :970-975 ---------------------------------------------- */
int liloc_scan_put(liloc_ctx* ctx, const liloc_point* pts_body, int n, float leaf, int* Q_all_out);
int liloc_scan_put_dev(liloc_ctx* ctx, const liloc_point* pts_body_dev, int n, float leaf, int* Q_all_out);
int liloc_scan_ds_get(liloc_ctx* ctx, liloc_point* out, int cap, int* Q_all_out);  /* laserCloudSurfLastDS */

/* ---- de-skew: imageProjection::projectPointCloud src/imageProjection.cpp:645-673 ------
 * The step before the path: the raw sweep is uploaded once; the range / ring / stride filters (:654-668), the
 * IMU-interpolated rotation of every surviving point (deskewPoint :615-643, findRotation :573-597) and the
 * compaction into fullCloud (push_back order) run on the device.  The de-skewed cloud (`cloud_deskewed`) stays there:
 * pass LILOC_SCAN_DESKEWED as scan_on_device / on_device to liloc_frame, liloc_frame_features, liloc_scan_put_class or
 * liloc_ndt_source_put to consume it without a host round trip (its size is read from device memory).
 * The IMU arrays are what imuDeskewInfo (:441-496) fills on the host: imu_pointer_cur + 1 valid entries (<= 2000). */
typedef struct { int ring; float time; } liloc_ring_time;      /* PointXYZIRT::ring / ::time (src/imageProjection.cpp:5-16) */
typedef struct {
    double time_scan_cur;                                      /* timeScanCur (:414) */
    const double* imu_time; const double* imu_rot_x; const double* imu_rot_y; const double* imu_rot_z;
    int imu_pointer_cur;                                       /* after the decrement of :490 */
    int deskew_flag;                                           /* -1: no time channel (:400-410) */
    int imu_available;                                         /* cloudInfo.imuAvailable (:495) */
    float lidar_min_range, lidar_max_range;                    /* [1.0, 1000.0] include/utility.h:257-258 */
    int n_scan, downsample_rate, point_filter_num;             /* N_SCAN, [1], [3] include/utility.h:254-255 */
    int reserved[2];                                           /* [0]: 0, or LILOC_DESKEW_AHEAD */
} liloc_deskew_params;
#define LILOC_SCAN_HOST 0
#define LILOC_SCAN_DEVICE 1
#define LILOC_SCAN_DESKEWED 2
/* params->reserved[0] = LILOC_DESKEW_AHEAD: the sweep is de-skewed into a second slot and stays STAGED while
 * LILOC_SCAN_DESKEWED keeps referring to the current sweep; liloc_deskew_advance() makes the staged sweep current.
 * Lets the next sweep upload and de-skew (on a stream of its own) while the current frame runs -- the reference's
 * imageProjection node is two sweeps ahead of mapOptimization too (cloudQueue, :271-276). */
#define LILOC_DESKEW_AHEAD 1
/* n_out may be NULL: then nothing is synchronised and the count stays on the device. */
int liloc_deskew(liloc_ctx* ctx, const liloc_point* pts, const liloc_ring_time* ring_time, int n,
                 const liloc_deskew_params* params, int* n_out);
int liloc_deskew_advance(liloc_ctx* ctx);
int liloc_deskewed_get(liloc_ctx* ctx, liloc_point* out, int cap, int* n_out);   /* fullCloud, for publishClouds (:675-679) */
int liloc_last_deskew_marks(liloc_ctx* ctx, unsigned long long* out3);            /* ns: block 0 at start / after the barrier / end */
int liloc_last_deskew_ms(liloc_ctx* ctx, float* ms);                             /* device time of the last de-skew launch (liloc_set_timing) */

/* ---- registration: scan2MapOptimization :1183-1201 (everything except transformUpdate :1202) -----
 * pose6: in = initial guess (updateInitialGuess :831-900), out = optimised transformTobeMapped.
 * All iterations run on the device without a host round trip.  isDegenerate / matP persist in the
 * context across calls exactly like the class members (:90-91). */
int liloc_scan2map(liloc_ctx* ctx, float* pose6, const liloc_s2m_opts* opts, liloc_s2m_result* res);

/* One frame in one call: extractCloud + downsampleCurrentScan + scan2MapOptimization with a single
 * host synchronisation at the end (laserCloudInfoHandler :311-315).  scan may be a host pointer
 * (scan_on_device = 0) or a device pointer (1). */
int liloc_frame(liloc_ctx* ctx, const int* kf_ids, const float* kf_poses6, int K, float map_leaf,
                const liloc_point* scan, int n_scan, int scan_on_device, float scan_leaf,
                float* pose6, const liloc_s2m_opts* opts, liloc_s2m_result* res);

/* The same with both feature classes (north_star superset; the reference frame is liloc_frame).  Arrays are
 * indexed by class: [LILOC_SURF], [LILOC_CORNER].  The corner entries are read only when opts->enable_corner. */
typedef struct {
    const int* kf_ids; const float* kf_poses6; int K;
    float map_leaf[2];                   /* surroundingKeyframeMapLeafSize | upstream mappingCornerLeafSize */
    const liloc_point* scan[2]; int n_scan[2];
    float scan_leaf[2];                  /* mappingSurfLeafSize | mappingCornerLeafSize */
    int scan_on_device;
} liloc_frame_in;
int liloc_frame_features(liloc_ctx* ctx, const liloc_frame_in* in, float* pose6, const liloc_s2m_opts* opts, liloc_s2m_result* res);
/* Replaying a recorded sequence (the reference plays a bag: the next cloud_info is already in the subscriber queue, :254's
 * queue of 5), the caller may name the NEXT frame's surf scan before making the current frame's call: its host->device copy
 * is then enqueued on a side stream behind the current frame's launches and overlaps that frame's kernels, and the next
 * liloc_frame / liloc_frame_features with this very (pointer, n) and scan_on_device = 0 finds it uploaded.  The host
 * memory must stay unchanged until that call returns.  A hint no frame consumes costs one wasted copy, never a wrong scan. */
int liloc_scan_prefetch(liloc_ctx* ctx, const liloc_point* next_scan, int n);
typedef struct { int q_corner, map_points_corner, tiles_corner, tiles_surf, sort_redos; int reserved[3]; } liloc_s2m_extra;
int liloc_last_s2m_extra(liloc_ctx* ctx, liloc_s2m_extra* out);

/* Class-indexed variants of the step-wise entry points above (cls = LILOC_SURF | LILOC_CORNER). */
int liloc_keyframe_put_features(liloc_ctx* ctx, int kf_id, const liloc_point* surf, int n_surf, const liloc_point* corner, int n_corner);
int liloc_keyframe_size_class(liloc_ctx* ctx, int cls, int kf_id);
/* host copy of a stored keyframe cloud (the save_map / save_session services write them as PCDs/<k>.pcd, :495-517, :602-611) */
int liloc_keyframe_get(liloc_ctx* ctx, int cls, int kf_id, liloc_point* out, int cap, int* n_out);
int liloc_localmap_build_class(liloc_ctx* ctx, int cls, const int* kf_ids, const float* poses6, int K, float leaf, int* M_out, int* n_raw_out);
int liloc_localmap_set_class(liloc_ctx* ctx, int cls, const liloc_point* pts_map, int n, float leaf, int* M_out);
int liloc_localmap_get_class(liloc_ctx* ctx, int cls, liloc_point* out, int cap, int* M_out);
int liloc_scan_put_class(liloc_ctx* ctx, int cls, const liloc_point* pts_body, int n, int on_device, float leaf, int* Q_all_out);
int liloc_scan_ds_get_class(liloc_ctx* ctx, int cls, liloc_point* out, int cap, int* Q_all_out);

/* ---- relo-mode NDT registration: pcl::Registration interface of `registration` (:118) ----------------
 * Replaces ndt_omp's pclomp::NormalDistributionsTransform as LiLoc configures it (:175-199): a target is a
 * prior-session submap (setInputTarget, :273, anchorOptimize.hpp:394), a source is the full de-skewed scan
 * (setInputSource, :274, anchorOptimize.hpp:395), a job is one align() from an initial guess (:284,
 * anchorOptimize.hpp:404).  Jobs are independent and run batched; n_align = 2 reproduces the two
 * back-to-back align calls of the relocalisation init (:282-290).  The voxel-covariance grid of a target is
 * built once per id and cached (the reference rebuilds it on every setInputTarget). */
#define LILOC_NDT_DIRECT1 0
#define LILOC_NDT_DIRECT7 1
#define LILOC_NDT_KDTREE 2     /* pclomp::KDTREE: every leaf whose centroid is closer than `resolution` to the transformed point */
typedef struct {
    double step_size;       /* [0.1]  pclomp default */
    double outlier_ratio;   /* [0.55] pclomp default */
    double trans_eps;       /* ndtEpsilon (:178) */
    int max_iterations;     /* [35]   pclomp default */
    int search;             /* LILOC_NDT_DIRECT1 | LILOC_NDT_DIRECT7 | LILOC_NDT_KDTREE  (regMethod, :181-199) */
    int n_align;            /* [1]; 2 for the relocalisation init */
    int reserved;
} liloc_ndt_opts;
typedef struct { int target_id, source_id; float guess[16]; } liloc_ndt_job;     /* guess: row-major 4x4 (init_guess, :276-280) */
typedef struct {
    float T[16];            /* getFinalTransformation(), row-major 4x4 */
    double score;           /* sum of per-point scores; trans_probability = score / n_source */
    int iterations;         /* nr_iterations_, summed over the n_align passes */
    int converged;
    int sweeps;             /* derivative sweeps executed (incl. line-search evaluations) */
    int n_source;
} liloc_ndt_out;
int liloc_ndt_target_put(liloc_ctx* ctx, int target_id, const liloc_point* pts_map, int n, float resolution, int* n_leaves_out);
/* generateSubMaps (include/dataManager/dataLoader.hpp:306-348) for one submap on the device: the K prior-session
 * keyframes (stored with liloc_keyframe_put) are transformed by their prior poses and concatenated -- the reference
 * keeps the submap un-down-sampled (:332-336) -- and become NDT target `target_id`.  Nothing crosses PCIe. */
int liloc_ndt_target_put_keyframes(liloc_ctx* ctx, int target_id, const int* kf_ids, const float* poses6, int K, float resolution,
                                   int* n_leaves_out, int* n_points_out);
/* The submap AnchorOptimization::addNewPrior cuts after every submap_size keyframes it added (anchorOptimize.hpp:341-377):
 * like generateSubMaps, but voxel-filtered at filter_leaf (0.2, :363-367) before it becomes the NDT target. */
int liloc_ndt_target_put_keyframes_filtered(liloc_ctx* ctx, int target_id, const int* kf_ids, const float* poses6, int K, float filter_leaf,
                                            float resolution, int* n_leaves_out, int* n_points_out);
int liloc_ndt_target_size(liloc_ctx* ctx, int target_id);      /* points of the target cloud, or <0 if unknown id */
int liloc_ndt_source_put(liloc_ctx* ctx, int source_id, const liloc_point* pts_body, int n, int on_device);
int liloc_ndt_align_batch(liloc_ctx* ctx, const liloc_ndt_job* jobs, int n_jobs, const liloc_ndt_opts* opts, liloc_ndt_out* outs);
int liloc_ndt_clear(liloc_ctx* ctx);

/* ---- multi-GPU: one process (or thread) and one context per GPU -------------------------------------------------------
 * What shards on this path is naturally independent work: relocalisation hypotheses
 * (src/mapOptmization.cpp:261-305 run for many initial poses), one scan against K prior submaps and offline batches of
 * (scan, submap, guess) triples (include/egoOptimization/anchorOptimize.hpp:382-421).  Every rank runs its share with NO
 * data-path collective; the only exchange is one ncclAllGather of 8 floats per item on the context's stream:
 *   row = {roll, pitch, yaw, x, y, z (RotMtoEuler of the final transform, include/math_tools.h:92-111),
 *          score / n_source (trans_probability), iterations}
 * packed on the device by the block that retires the job -- no host staging.  NCCL is bound at run time (dlopen of
 * libnccl.so.2); the launcher distributes the 128-byte unique id (MPI, torch.distributed, a file ...). */
#define LILOC_COMM_ID_BYTES 128
#define LILOC_SHARD_BLOCKS 0          /* contiguous blocks of the item list (lists sorted by submap: grids stay local) */
#define LILOC_SHARD_ROUND_ROBIN 1     /* items rank, rank + world, ... (lists whose cost correlates with position) */
int liloc_comm_unique_id(void* id128);                                        /* rank 0: ncclGetUniqueId */
int liloc_comm_init(liloc_ctx* ctx, int rank, int world, const void* id128);   /* ncclCommInitRank on the context's device; world 1: no NCCL */
int liloc_comm_destroy(liloc_ctx* ctx);
/* The share of `rank` under a policy: item indices in the order their rows travel (returns the count; out may be NULL).
 * Pure index arithmetic: needs neither a context nor a device. */
int liloc_shard_indices(int n_items, int rank, int world, int policy, int* out, int cap);
/* gathered rows [world][per_rank][8] (rank-major, each rank's share in liloc_shard_indices order) -> rows_out[n_items][8] */
int liloc_unshard_rows(const float* rows_all, int n_items, int world, int per_rank, int policy, float* rows_out);
int liloc_comm_rank(const liloc_ctx* ctx);
int liloc_comm_world(const liloc_ctx* ctx);
/* `items` is the FULL list, identical on every rank; targets / sources named by a rank's share must be resident in its
 * context.  rows_out[n_items][8] comes back complete on every rank.  local_outs (may be NULL): full results of this
 * rank's share, in share order.  Without liloc_comm_init (world 1) the whole list runs here. */
int liloc_ndt_align_batch_sharded(liloc_ctx* ctx, const liloc_ndt_job* items, int n_items, const liloc_ndt_opts* opts, int policy,
                                  float* rows_out, liloc_ndt_out* local_outs);
/* The gather alone: every rank contributes n_local <= per_rank rows of 8 floats (host), rows_all[world][per_rank][8] comes
 * back on every rank (short shares zero-padded).  Used for the per-frame results of lio replicas. */
int liloc_gather_results(liloc_ctx* ctx, const float* rows_local, int n_local, int per_rank, float* rows_all);
/* leaf records of a target (tests): means[3n] double, icovs[9n] float, counts[n], keys[n] (voxel index) */
int liloc_ndt_target_get(liloc_ctx* ctx, int target_id, double* means, float* icovs, int* counts, int* keys, int cap);
/* one derivative sweep at pose p6 = [tx,ty,tz,rx,ry,rz] (tests): out28 = score, gradient[6], Hessian upper[21] */
int liloc_ndt_derivatives(liloc_ctx* ctx, int target_id, int source_id, const double* p6, const liloc_ndt_opts* opts, double* out28);

/* ---- global map of the visualisation thread: publishGlobalMap (src/mapOptmization.cpp:698-753) ----------------------------
 * transformPointCloud of the K selected keyframes (:744-746) + VoxelGrid(globalMapVisualizationLeafSize) (:749-750) in
 * buffers of their own (the frame's local map is not touched).  out may be NULL (count only); cap in points. */
int liloc_globalmap_build(liloc_ctx* ctx, const int* kf_ids, const float* poses6, int K, float leaf, liloc_point* out, int cap, int* M_out);

/* ---- ICP fitness score: AnchorOptimization::getICPFitnessScore (include/egoOptimization/anchorOptimize.hpp:465-481) ----
 * Mean SQUARED distance from every point of cloud 1 to its nearest neighbour in cloud 2 (pcl::registration::
 * CorrespondenceEstimation::determineCorrespondences with its default, unbounded, maximum distance), summed sequentially
 * in source order in float like the reference.  Each cloud is first transformed by its pose6 (transformPointCloud,
 * :425, :439; NULL = identity).  The `_resident` form takes cloud 1 from the keyframe store and cloud 2 from an NDT
 * source -- what addScanMatchingFactor compares (:427-442): a prior vertex' key cloud against the current scan --
 * so nothing crosses PCIe. */
int liloc_icp_fitness(liloc_ctx* ctx, const liloc_point* cloud1, int n1, const float* pose6_1,
                      const liloc_point* cloud2, int n2, const float* pose6_2, float* score_out);
int liloc_icp_fitness_resident(liloc_ctx* ctx, int kf_id, const float* pose6_kf, int ndt_source_id, const float* pose6_src, float* score_out);

/* Cumulative NDT counters since liloc_create / liloc_ndt_stats_reset.  A batch is ONE launch of the persistent,
 * queue-driven kernel k_ndt_align (no per-sweep launches, no grid barriers): a ticket is a run of 1024-point tiles of one
 * derivative sweep of one job; the block that completes a sweep runs the job's Newton / More-Thuente step and publishes
 * the next sweep's tickets.  The *_ms fields are recorded only while timing is enabled (%globaltimer): kernel_ms = first
 * block's start .. last block's end, block_*_ms = summed over the `blocks` of the launches.  alg_bytes are the ALGORITHMIC
 * bytes of SURVEY.md 8(d): per derivative sweep and job 16 P + 36 P c + 224 (P source points, c cells per point). */
typedef struct {
    long long batches, jobs, launches, blocks, tickets, updates, job_sweeps, max_job_sweeps;
    double kernel_ms, alg_bytes;
    double block_sweep_ms, block_update_ms, block_wait_ms;
    double update_cycles[4];        /* SM cycles inside the updates: reduce + load, state machine, pose tables, store + publish */
} liloc_ndt_stats;
int liloc_ndt_stats_get(const liloc_ctx* ctx, liloc_ndt_stats* out);
int liloc_ndt_stats_reset(liloc_ctx* ctx);

/* ---- kernel-level entry points (used by the parity tests; no reference caller) ------------------ */
/* pcl::VoxelGrid<PointXYZI>::filter on a host cloud -> host cloud (cap >= n). */
int liloc_voxelgrid(liloc_ctx* ctx, const liloc_point* in, int n, float leaf, liloc_point* out, int cap, int* m_out);
/* Exact sorted 5-NN of nq map-frame queries against the current local map
 * (kdtreeSurfFromMap->nearestKSearch :992).  Ties broken by ascending index.  Only neighbours with
 * d2 < 1.0 are searched for (the reference discards everything else at :1002): for a query whose
 * 5th neighbour is not within 1 m, idx5[5*i+4] == -1 and found[i] < 5. */
int liloc_knn5(liloc_ctx* ctx, const liloc_point* queries_map, int nq, int* idx5, float* d2_5, int* found);
/* surfOptimization at a fixed pose (:981-1048): per scan point i (multiples of stride) the flag
 * laserCloudOriSurfFlag[i] and coefficient coeffSelSurfVec[i]; arrays have Q_all entries. */
int liloc_surf_pass(liloc_ctx* ctx, const float* pose6, int stride, liloc_point* coeff, unsigned char* flag);
int liloc_knn5_class(liloc_ctx* ctx, int cls, const liloc_point* queries_map, int nq, int* idx5, float* d2_5, int* found);
/* Test surface: liloc_knn5_class with bounds[i] = an upper bound of query i's fifth squared distance (what the registration
 * kernel derives from the previous iteration's neighbours to skip cells; knn_grid.cuh).  With a valid bound the result is
 * the unbounded one. */
int liloc_debug_knn5_bounded(liloc_ctx* ctx, int cls, const liloc_point* queries_map, int nq, const float* bounds, int* idx5, float* d2_5, int* found);
/* surfOptimization (cls 0) / upstream cornerOptimization (cls 1) at a fixed pose, like liloc_surf_pass. */
int liloc_feature_pass(liloc_ctx* ctx, int cls, const float* pose6, int stride, liloc_point* coeff, unsigned char* flag);
/* pcl::getTransformation as the device evaluates it: top 3x4 block, row-major. */
int liloc_pose_to_T(liloc_ctx* ctx, const float* pose6, float* T12);

/* The small dense algebra as the device evaluates it, n independent systems (one thread each):
 * cv::solve(DECOMP_QR) :1130, cv::eigen :1138 (w descending) and the degeneracy projector :1141-1153. */
int liloc_debug_lm6(liloc_ctx* ctx, const float* A36, const float* b6, int n, float* x6, float* w6, float* P36, int* deg);
/* Eigen colPivHouseholderQr().solve of n 5x3 systems A x = -1 (:1009); pts15 = n x 5 x 3 row-major. */
int liloc_debug_plane(liloc_ctx* ctx, const float* pts15, int n, float* x3);

/* cv::eigen of n symmetric 3x3 matrices as cornerOptimization uses it: w descending, eigenvectors as rows of V. */
int liloc_debug_eigen3(liloc_ctx* ctx, const float* A9, int n, float* w3, float* V9);
/* cornerOptimization's loop body after the kNN on explicit neighbours: pts15 = n x 5 x 3, sel3 = n x pointSel. */
int liloc_debug_line(liloc_ctx* ctx, const float* pts15, const float* sel3, int n, liloc_point* coeff, int* keep);

/* ---- timing: milliseconds of the phases of the last liloc_frame (from the in-kernel %globaltimer marks below;
 *      the step-wise liloc_scan2map uses CUDA events) ------------------------------------------------------ */
typedef struct { float assemble_ms, map_voxel_ms, grid_ms, scan_ms, s2m_ms, total_ms; } liloc_timing;
int liloc_last_timing(liloc_ctx* ctx, liloc_timing* t);
/* Local-map memoisation (liloc_config.map_cache) on / off at run time; returns the number of frames that reused the map. */
long long liloc_set_map_cache(liloc_ctx* ctx, int enabled);
int liloc_set_timing(liloc_ctx* ctx, int enabled);
/* Device-side latency breakdown of the last liloc_frame / liloc_frame_features made while timing is enabled:
 * %globaltimer nanoseconds written by block 0 at the phase boundaries of the three launches --
 * [0..15] local-map pipeline (start, assemble+bbox, keys, sort, (unused), centroids, cell scan, end),
 * [16..31] scan pipeline (same phases), [32] registration start, then per iteration i: [33+3i] tiles done,
 * [34+3i] past the grid barrier, [35+3i] pose updated, [129] end of the registration kernel.  The kernels write them
 * straight into mapped host memory; liloc_last_timing is derived from them (no CUDA events inside a frame).
 * Returns the number of entries copied. */
int liloc_last_marks(liloc_ctx* ctx, unsigned long long* out, int cap);
/* Cumulative counters since liloc_create / liloc_stats_reset.  Phase times are the liloc_timing milliseconds
 * summed over liloc_frame calls made while timing is enabled; alg_bytes_* are the ALGORITHMIC bytes of
 * assemble+map VoxelGrid 16 N_raw + 16 M, scan VoxelGrid 16 N_scan + 16 Q_all,
 * scan2map I (96 Q + 216) with Q = ceil(Q_all / stride) and I the iterations actually executed. */
typedef struct {
    long long frames, launches, s2m_launches, s2m_iters;
    long long h2d_bytes, d2h_bytes;
    double assemble_ms, map_voxel_ms, grid_ms, scan_ms, s2m_ms, total_ms;
    double alg_bytes_map, alg_bytes_scan, alg_bytes_s2m;
} liloc_stats;
int liloc_stats_get(const liloc_ctx* ctx, liloc_stats* out);
int liloc_stats_reset(liloc_ctx* ctx);
void* liloc_stream(liloc_ctx* ctx);                       /* the context's cudaStream_t */
long long liloc_kernel_launches(const liloc_ctx* ctx);   /* kernels launched by this context so far */
/* Per-iteration trace of the last scan2map: rows of 16 floats {n_sel, X[6], pose_after[6], AtB[0..2]}
 * for the first min(max_iters, 32) iterations; returns the number of rows copied. */
int liloc_last_trace(liloc_ctx* ctx, float* out, int max_iters);
/* SM-clock deltas of the phases of each iteration of the last scan2map (rows of 8 ints: tiles, grid barrier,
 * reduction, solve, transform refresh); all zero unless the library was built with -DLILOC_S2M_PROFILE. */
int liloc_last_clocks(liloc_ctx* ctx, int* out, int max_iters);

#ifdef __cplusplus
}
#endif
#endif /* LILOC_GPU_H */
