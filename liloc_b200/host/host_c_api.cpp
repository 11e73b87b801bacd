// liloc_b200/host/host_c_api.cpp -- C wrapper around the host class so that tests and bench.py can drive
// the same per-frame callback the reference's ROS spinner drives (src/mapOptmization.cpp:246, :1628).
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "image_projection.hpp"
#include "map_optimization.hpp"
#include "session_io.hpp"

using liloc_host::mapOptimization;

extern "C" {

struct liloc_host_cloud_info {           // mirror of liloc_host::CloudInfo (plain C layout)
    double stamp;
    int imuAvailable, odomAvailable;
    float imuRollInit, imuPitchInit, imuYawInit;
    float initialGuessX, initialGuessY, initialGuessZ, initialGuessRoll, initialGuessPitch, initialGuessYaw;
    const liloc_point* cloud_deskewed;
    int cloud_size;
    int cloud_on_device;
};

struct liloc_host_report {               // mirror of liloc_host::FrameReport
    int processed, status;
    float pose[6];
    float guess[6];
    int is_degenerate, iters, converged, n_sel, q_all, map_points, n_raw, n_selected_keyframes, keyframe_id;
    double reg_ms;
    int relo_initialized, relo_submap, relo_iterations, relo_converged;
    float relo_pose[6];
    int relo_n_vertices, relo_vertex[3];
    float relo_fitness[3];
    int prior_added, prior_new_submap;
    int relo_hypothesis; float relo_score;
};

static thread_local std::string g_err;

void* liloc_host_create(const char* yaml_path, const char* mode_override, int device) {
    auto* h = new (std::nothrow) mapOptimization(yaml_path ? yaml_path : "", mode_override, device);
    if (!h) { g_err = "out of memory"; return nullptr; }
    if (!h->ok()) { g_err = h->error(); delete h; return nullptr; }
    return h;
}
void liloc_host_destroy(void* h) { delete static_cast<mapOptimization*>(h); }
const char* liloc_host_error(void* h) { return h ? static_cast<mapOptimization*>(h)->error().c_str() : g_err.c_str(); }
void* liloc_host_ctx(void* h) { return static_cast<mapOptimization*>(h)->context(); }

static void to_report(const liloc_host::FrameReport& r, liloc_host_report* o) {
    o->processed = r.processed; o->status = r.status; std::memcpy(o->pose, r.pose, sizeof(o->pose)); std::memcpy(o->guess, r.guess, sizeof(o->guess));
    o->is_degenerate = r.is_degenerate; o->iters = r.iters; o->converged = r.converged; o->n_sel = r.n_sel;
    o->q_all = r.q_all; o->map_points = r.map_points; o->n_raw = r.n_raw; o->n_selected_keyframes = r.n_selected_keyframes;
    o->keyframe_id = r.keyframe_id; o->reg_ms = r.reg_ms;
    o->relo_initialized = r.relo_initialized; o->relo_submap = r.relo_submap; o->relo_iterations = r.relo_iterations; o->relo_converged = r.relo_converged;
    std::memcpy(o->relo_pose, r.relo_pose, sizeof(o->relo_pose));
    o->relo_n_vertices = r.relo_n_vertices;
    std::memcpy(o->relo_vertex, r.relo_vertex, sizeof(o->relo_vertex)); std::memcpy(o->relo_fitness, r.relo_fitness, sizeof(o->relo_fitness));
    o->prior_added = r.prior_added; o->prior_new_submap = r.prior_new_submap;
    o->relo_hypothesis = r.relo_hypothesis; o->relo_score = r.relo_score;
}

static liloc_host::CloudInfo to_info(const liloc_host_cloud_info& m) {
    liloc_host::CloudInfo c;
    c.stamp = m.stamp; c.imuAvailable = m.imuAvailable; c.odomAvailable = m.odomAvailable;
    c.imuRollInit = m.imuRollInit; c.imuPitchInit = m.imuPitchInit; c.imuYawInit = m.imuYawInit;
    c.initialGuessX = m.initialGuessX; c.initialGuessY = m.initialGuessY; c.initialGuessZ = m.initialGuessZ;
    c.initialGuessRoll = m.initialGuessRoll; c.initialGuessPitch = m.initialGuessPitch; c.initialGuessYaw = m.initialGuessYaw;
    c.cloud_deskewed = m.cloud_deskewed; c.cloud_size = m.cloud_size; c.cloud_on_device = m.cloud_on_device;
    return c;
}

// one callback (laserCloudInfoHandler).  Returns 0, or <0 when the registration reported an error.
int liloc_host_handle(void* h, const liloc_host_cloud_info* msg, liloc_host_report* out) {
    auto* m = static_cast<mapOptimization*>(h);
    const liloc_host::FrameReport r = m->laserCloudInfoHandler(to_info(*msg));
    if (out) to_report(r, out);
    return r.status < 0 ? r.status : 0;
}

// F callbacks back to back (what ros::spin() does with a queue of messages)
int liloc_host_run(void* h, const liloc_host_cloud_info* msgs, int F, liloc_host_report* outs) {
    auto* m = static_cast<mapOptimization*>(h);
    for (int i = 0; i < F; ++i) {
        // the next message is already in the queue: its scan goes up while this frame's kernels run (liloc_scan_prefetch)
        if (i + 1 < F && !msgs[i + 1].cloud_on_device && msgs[i + 1].cloud_size > 0)
            liloc_scan_prefetch(m->context(), msgs[i + 1].cloud_deskewed, msgs[i + 1].cloud_size);
        const int rc = liloc_host_handle(h, msgs + i, outs ? outs + i : nullptr);
        if (rc < 0) return rc;
    }
    return 0;
}

// relo mode: prior session = n key poses (rows [roll, pitch, yaw, x, y, z, time]) + their key clouds (body frame)
int liloc_host_load_prior(void* h, const double* keyposes7, int n, const liloc_point* clouds, const long long* offsets) {
    auto* m = static_cast<mapOptimization*>(h);
    std::vector<liloc_host::PointTypePose> kp((size_t)(n > 0 ? n : 0));
    for (int i = 0; i < n; ++i) {
        const double* r = keyposes7 + 7 * i;
        kp[i] = liloc_host::PointTypePose{(float)r[3], (float)r[4], (float)r[5], (float)i, (float)r[0], (float)r[1], (float)r[2], r[6]};
    }
    return m->loadPriorSession(kp.data(), n, clouds, offsets) ? (int)m->prior().SubMapCenteriod.size() : -1;
}
int liloc_host_load_prior_dir(void* h, const char* dir) {
    auto* m = static_cast<mapOptimization*>(h);
    return m->loadPriorSessionDir(dir ? dir : "") ? (int)m->prior().SubMapCenteriod.size() : -1;
}
int liloc_host_save_session(void* h, const char* dir) { return static_cast<mapOptimization*>(h)->saveSession(dir ? dir : "") ? 0 : -1; }
// ... at req.resolution (save_mapRequest / save_sessionRequest)
int liloc_host_save_session_res(void* h, const char* dir, float resolution) { return static_cast<mapOptimization*>(h)->saveSession(dir ? dir : "", resolution) ? 0 : -1; }
int liloc_host_save_poses_pcd(const char* path, const double* keyposes7, int n) {
    std::vector<liloc_host::PointTypePose> kp((size_t)(n > 0 ? n : 0));
    for (int i = 0; i < n; ++i) { const double* r = keyposes7 + 7 * i; kp[i] = liloc_host::PointTypePose{(float)r[3], (float)r[4], (float)r[5], (float)i, (float)r[0], (float)r[1], (float)r[2], r[6]}; }
    std::string e;
    return liloc_host::savePosesPCDBinary(path, kp, e) ? 0 : -1;
}

// session file format, host only (no device context needed): parse a PCD / a g2o; used by the CPU tests
int liloc_host_load_pcd(const char* path, liloc_point* out, int cap, char* err, int err_cap) {
    std::vector<liloc_point> pts; std::string e;
    if (!liloc_host::loadPCD(path, pts, e)) { if (err && err_cap > 0) { std::strncpy(err, e.c_str(), err_cap - 1); err[err_cap - 1] = 0; } return -1; }
    if (out && (int)pts.size() <= cap && !pts.empty()) std::memcpy(out, pts.data(), pts.size() * sizeof(liloc_point));
    return (int)pts.size();
}
int liloc_host_save_pcd(const char* path, const liloc_point* pts, int n) { std::string e; return liloc_host::savePCDBinary(path, pts, n, e) ? 0 : -1; }
int liloc_host_load_g2o(const char* path, double* keyposes7, int cap) {
    std::vector<liloc_host::PointTypePose> kp; std::string e;
    if (!liloc_host::loadSessionGraph(path, kp, e)) return -1;
    for (int i = 0; i < (int)kp.size() && i < cap; ++i) {
        double* o = keyposes7 + 7 * i;
        o[0] = kp[i].roll; o[1] = kp[i].pitch; o[2] = kp[i].yaw; o[3] = kp[i].x; o[4] = kp[i].y; o[5] = kp[i].z; o[6] = kp[i].intensity;
    }
    return (int)kp.size();
}
int liloc_host_save_g2o(const char* path, const double* keyposes7, int n) {
    std::vector<liloc_host::PointTypePose> kp((size_t)(n > 0 ? n : 0));
    for (int i = 0; i < n; ++i) { const double* r = keyposes7 + 7 * i; kp[i] = liloc_host::PointTypePose{(float)r[3], (float)r[4], (float)r[5], (float)i, (float)r[0], (float)r[1], (float)r[2], r[6]}; }
    std::string e;
    return liloc_host::saveSessionGraph(path, kp, e) ? 0 : -1;
}

// throughput mode: this object is one of `share` objects (sequences) processed concurrently on its GPU (liloc_set_share)
int liloc_host_set_share(void* h, int share) { return liloc_set_share(static_cast<mapOptimization*>(h)->context(), share); }
void liloc_host_set_relo_hypotheses(void* h, int n_yaw, int n_xy, float radius) { static_cast<mapOptimization*>(h)->setReloHypotheses(n_yaw, n_xy, radius); }
int liloc_host_comm_init(void* h, int rank, int world, const void* id128) { return static_cast<mapOptimization*>(h)->commInit(rank, world, id128); }
void liloc_host_hypothesis_pose(const float* center6, int n_yaw, int n_xy, float radius, int hyp, float* out6) {
    mapOptimization::hypothesisPose(center6, n_yaw, n_xy, radius, hyp, out6);
}
void liloc_host_initialpose(void* h, float x, float y, float yaw) { static_cast<mapOptimization*>(h)->initialposeHandler(x, y, yaw); }
int liloc_host_submap_info(void* h, int s, float* centroid3, int* first, int* count, int* points) {
    const auto& p = static_cast<mapOptimization*>(h)->prior();
    if (s < 0 || s >= (int)p.SubMapCenteriod.size()) return -1;
    centroid3[0] = p.SubMapCenteriod[s].x; centroid3[1] = p.SubMapCenteriod[s].y; centroid3[2] = p.SubMapCenteriod[s].z;
    *first = p.submap_first[s]; *count = p.submap_count[s]; *points = p.submap_points[s];
    return 0;
}

// publishGlobalMap (:698-753): returns the number of points (copied into out when cap suffices), or <0
int liloc_host_global_map(void* h, liloc_point* out, int cap) {
    std::vector<liloc_point> g;
    if (!static_cast<mapOptimization*>(h)->publishGlobalMap(g)) return -1;
    if (out && (int)g.size() <= cap && !g.empty()) std::memcpy(out, g.data(), g.size() * sizeof(liloc_point));
    return (int)g.size();
}

void liloc_host_reset(void* h) { static_cast<mapOptimization*>(h)->reset(); }

void liloc_host_set_s2m_opts(void* h, int max_iters, int stride, int min_sel, int min_points) {
    auto* m = static_cast<mapOptimization*>(h);
    m->s2m_opts.max_iters = max_iters; m->s2m_opts.stride = stride; m->s2m_opts.min_sel = min_sel; m->s2m_opts.min_points = min_points;
}

int liloc_host_num_keyframes(void* h) { return (int)static_cast<mapOptimization*>(h)->cloudKeyPoses6D.size(); }

// key poses as rows [roll, pitch, yaw, x, y, z, time]
int liloc_host_keyposes(void* h, double* out7, int cap) {
    auto* m = static_cast<mapOptimization*>(h);
    const int n = (int)m->cloudKeyPoses6D.size();
    for (int i = 0; i < n && i < cap; ++i) {
        const auto& p = m->cloudKeyPoses6D[i];
        double* o = out7 + 7 * i;
        o[0] = p.roll; o[1] = p.pitch; o[2] = p.yaw; o[3] = p.x; o[4] = p.y; o[5] = p.z; o[6] = p.time;
    }
    return n;
}

int liloc_host_last_selection(void* h, int* ids, int cap) {
    const auto& s = static_cast<mapOptimization*>(h)->lastSelection();
    for (int i = 0; i < (int)s.size() && i < cap; ++i) ids[i] = s[i];
    return (int)s.size();
}

// parsed parameters, for the config tests: returns 0 and writes *out, or -1 for an unknown key
static int param_of(const liloc_host::ParamServer* m, const char* key, double* out) {
    const std::string k(key);
#define P(name) if (k == #name) { *out = (double)m->name; return 0; }
    P(numberOfCores) P(N_SCAN) P(Horizon_SCAN) P(downsampleRate) P(point_filter_num) P(lidarMinRange) P(lidarMaxRange)
    P(imuType) P(imuRate) P(imuAccNoise) P(imuGyrNoise) P(imuAccBiasN) P(imuGyrBiasN) P(imuGravity) P(imuRPYWeight)
    P(z_tollerance) P(rotation_tollerance) P(ndtResolution) P(ndtEpsilon)
    P(timeInterval) P(mappingProcessInterval) P(mappingSurfLeafSize) P(surroundingKeyframeMapLeafSize)
    P(surroundingkeyframeAddingDistThreshold) P(surroundingkeyframeAddingAngleThreshold) P(surroundingKeyframeDensity)
    P(surroundingKeyframeSearchRadius) P(globalMapVisualizationSearchRadius) P(globalMapVisualizationPoseDensity)
    P(globalMapVisualizationLeafSize) P(have_ring_time_channel) P(correct)
#undef P
    if (k == "mode") { *out = m->mode == liloc_host::ModeType::LIO ? 0 : 1; return 0; }
    if (k == "sensor") { *out = (double)(int)m->sensor; return 0; }
    if (k == "regMethod.DIRECT1") { *out = m->regMethod == "DIRECT1"; return 0; }
    if (k.rfind("extrinsicRot.", 0) == 0) { const size_t i = std::stoul(k.substr(13)); if (i < m->extRotV.size()) { *out = m->extRotV[i]; return 0; } return -1; }
    if (k.rfind("extrinsicRPY.", 0) == 0) { const size_t i = std::stoul(k.substr(13)); if (i < m->extRPYV.size()) { *out = m->extRPYV[i]; return 0; } return -1; }
    if (k.rfind("extrinsicTrans.", 0) == 0) { const size_t i = std::stoul(k.substr(15)); if (i < m->extTransV.size()) { *out = m->extTransV[i]; return 0; } return -1; }
    return -1;
}
int liloc_host_param(void* h, const char* key, double* out) { return param_of(static_cast<mapOptimization*>(h), key, out); }

// ParamServer alone (no GPU needed): parse `yaml_path` and read one parameter. -2: yaml unusable.
int liloc_host_yaml_param(const char* yaml_path, const char* key, double* out) {
    liloc_host::ParamServer ps;
    if (!ps.load(yaml_path)) { g_err = ps.param_error; return -2; }
    return param_of(&ps, key, out);
}

// ---- host-logic hooks for the CPU tests (no device work) ----
void liloc_host_debug_add_keypose(void* h, const float* pose6, double time) {
    auto* m = static_cast<mapOptimization*>(h);
    const float id = (float)m->cloudKeyPoses3D.size();
    m->cloudKeyPoses3D.push_back(liloc_host::PointType{pose6[3], pose6[4], pose6[5], id});
    m->cloudKeyPoses6D.push_back(liloc_host::PointTypePose{pose6[3], pose6[4], pose6[5], id, pose6[0], pose6[1], pose6[2], time});
}
int liloc_host_debug_extract_nearby(void* h, double stamp, int* ids, int cap) {
    auto* m = static_cast<mapOptimization*>(h);
    m->timeLaserInfoCur = stamp;
    m->extractSurroundingKeyFrames();
    return liloc_host_last_selection(h, ids, cap);
}
void liloc_host_debug_nearby_plain(void* h, int on) { static_cast<mapOptimization*>(h)->nearby_plain_only = on != 0; }
void liloc_host_debug_set_pose(void* h, const float* pose6) { std::memcpy(static_cast<mapOptimization*>(h)->transformTobeMapped, pose6, 24); }
void liloc_host_debug_get_pose(void* h, float* pose6) { std::memcpy(pose6, static_cast<mapOptimization*>(h)->transformTobeMapped, 24); }
void liloc_host_debug_update_initial_guess(void* h, const liloc_host_cloud_info* msg) {
    auto* m = static_cast<mapOptimization*>(h);
    m->cloudInfo = to_info(*msg); m->timeLaserInfoCur = msg->stamp;
    m->updateInitialGuess();
}
// init_guess exactly as the class builds it for registration->align (eulerToRotation, src/mapOptmization.cpp:276-280): tests feed the
// oracle the same 16 floats
void liloc_host_debug_euler_to_matrix(const float* pose6, float* T16) {
    liloc_host::eulerToMatrix4(pose6[0], pose6[1], pose6[2], pose6[3], pose6[4], pose6[5], T16);
}
int liloc_host_debug_save_frame(void* h) { return static_cast<mapOptimization*>(h)->saveFrame() ? 1 : 0; }
void liloc_host_debug_transform_update(void* h, const liloc_host_cloud_info* msg) {
    auto* m = static_cast<mapOptimization*>(h);
    m->cloudInfo = to_info(*msg);
    m->transformUpdate();
}
int liloc_host_debug_voxelgrid_small(const float* in4, int n, float leaf, float* out4) {
    std::vector<liloc_host::PointType> v(n);
    std::memcpy(v.data(), in4, sizeof(float) * 4 * (size_t)n);
    const auto o = liloc_host::voxelGridSmall(v, leaf);
    std::memcpy(out4, o.data(), sizeof(float) * 4 * o.size());
    return (int)o.size();
}

// ---- ImageProjection (src/imageProjection.cpp): the node in front of mapOptimization, sharing its device context ----
void* liloc_host_ip_create(const char* yaml_path, void* map_optimization) {
    auto* m = static_cast<mapOptimization*>(map_optimization);
    auto* ip = new (std::nothrow) liloc_host::ImageProjection(yaml_path ? yaml_path : "", m ? m->context() : nullptr);
    if (!ip) { g_err = "out of memory"; return nullptr; }
    if (!ip->ok()) { g_err = ip->error(); delete ip; return nullptr; }
    return ip;
}
void liloc_host_ip_destroy(void* ip) { delete static_cast<liloc_host::ImageProjection*>(ip); }
const char* liloc_host_ip_error(void* ip) { return ip ? static_cast<liloc_host::ImageProjection*>(ip)->error().c_str() : g_err.c_str(); }
// msg11 = stamp, angular_velocity xyz, linear_acceleration xyz, orientation xyzw
void liloc_host_ip_imu(void* ip, const double* msg11) {
    liloc_host::ImuMsg m{};
    m.stamp = msg11[0];
    for (int i = 0; i < 3; ++i) { m.angular_velocity[i] = msg11[1 + i]; m.linear_acceleration[i] = msg11[4 + i]; }
    for (int i = 0; i < 4; ++i) m.orientation[i] = msg11[7 + i];
    static_cast<liloc_host::ImageProjection*>(ip)->imuHandler(m);
}
// msg9 = stamp, position xyz, orientation xyzw, pose.covariance[0]
void liloc_host_ip_odom(void* ip, const double* msg9) {
    liloc_host::OdomMsg m{};
    m.stamp = msg9[0];
    for (int i = 0; i < 3; ++i) m.position[i] = msg9[1 + i];
    for (int i = 0; i < 4; ++i) m.orientation[i] = msg9[4 + i];
    m.covariance0 = msg9[8];
    static_cast<liloc_host::ImageProjection*>(ip)->odometryHandler(m);
}
static void from_info(const liloc_host::CloudInfo& c, liloc_host_cloud_info* m) {
    m->stamp = c.stamp; m->imuAvailable = c.imuAvailable; m->odomAvailable = c.odomAvailable;
    m->imuRollInit = c.imuRollInit; m->imuPitchInit = c.imuPitchInit; m->imuYawInit = c.imuYawInit;
    m->initialGuessX = c.initialGuessX; m->initialGuessY = c.initialGuessY; m->initialGuessZ = c.initialGuessZ;
    m->initialGuessRoll = c.initialGuessRoll; m->initialGuessPitch = c.initialGuessPitch; m->initialGuessYaw = c.initialGuessYaw;
    m->cloud_deskewed = c.cloud_deskewed; m->cloud_size = c.cloud_size; m->cloud_on_device = c.cloud_on_device;
}
// cloudHandler: 1 = a cloud_info was produced (the reference publishes it), 0 = nothing yet (queue filling / waiting for IMU), <0 error
int liloc_host_ip_cloud(void* ip, double stamp, const liloc_point* pts, const liloc_ring_time* ring_time, int n, int has_time, liloc_host_cloud_info* out) {
    auto* p = static_cast<liloc_host::ImageProjection*>(ip);
    liloc_host::CloudInfo ci;
    const bool produced = p->cloudHandler(stamp, pts, ring_time, n, has_time != 0, &ci);
    if (!p->ok()) return LILOC_ERR_STATE;
    if (produced && out) from_info(ci, out);
    return produced ? 1 : 0;
}
struct liloc_host_livox_point { float x, y, z; unsigned char reflectivity, tag, line; unsigned offset_time; };
int liloc_host_ip_cloud_livox(void* ip, double stamp, const liloc_host_livox_point* pts, int n, liloc_host_cloud_info* out) {
    static_assert(sizeof(liloc_host_livox_point) == sizeof(liloc_host::LivoxPoint), "layout");
    auto* p = static_cast<liloc_host::ImageProjection*>(ip);
    liloc_host::CloudInfo ci;
    const bool produced = p->cloudHandlerLivox(stamp, reinterpret_cast<const liloc_host::LivoxPoint*>(pts), n, &ci);
    if (!p->ok()) return LILOC_ERR_STATE;
    if (produced && out) from_info(ci, out);
    return produced ? 1 : 0;
}
// the inputs of the last published sweep's de-skew call: imu4 = time | rotX | rotY | rotZ (each imuPointerCur + 1 long, cap
// entries of room per array), the sweep's points and ring / time as handed to liloc_deskew.  Returns the number of points.
int liloc_host_ip_last_sweep(void* ip, double* imu4, int imu_cap, int* imuPointerCur, liloc_point* pts, liloc_ring_time* rt, int cap) {
    auto* p = static_cast<liloc_host::ImageProjection*>(ip);
    const int last = p->lastImuPointerCur(), m = last >= 0 ? last + 1 : 0;
    if (imuPointerCur) *imuPointerCur = last;
    if (imu4 && m <= imu_cap) for (int a = 0; a < 4; ++a) for (int k = 0; k < m; ++k) imu4[(size_t)a * imu_cap + k] = p->lastImu()[(size_t)a * m + k];
    const int n = (int)p->lastSweepPoints().size();
    if (pts && rt && n <= cap) { std::memcpy(pts, p->lastSweepPoints().data(), (size_t)n * sizeof(liloc_point)); std::memcpy(rt, p->lastSweepRingTime().data(), (size_t)n * sizeof(liloc_ring_time)); }
    return n;
}

void liloc_host_ip_times(void* ip, double* timeScanCur, double* timeScanEnd, int* deskewFlag) {
    auto* p = static_cast<liloc_host::ImageProjection*>(ip);
    if (timeScanCur) *timeScanCur = p->timeScanCur;
    if (timeScanEnd) *timeScanEnd = p->timeScanEnd;
    if (deskewFlag) *deskewFlag = p->deskewFlag;
}

}  // extern "C"
