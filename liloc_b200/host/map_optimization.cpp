// liloc_b200/host/map_optimization.cpp -- see map_optimization.hpp.
#include "map_optimization.hpp"
#include "session_io.hpp"

#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstring>
#include <filesystem>

namespace liloc_host {

namespace {
inline float pointDistance(const PointType& a, const PointType& b) {       // include/utility.h pointDistance(p1, p2)
    return std::sqrt((a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y) + (a.z - b.z) * (a.z - b.z));
}
inline float sqDist(const PointType& a, const PointType& b) {
    const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return ((dx * dx) + (dy * dy)) + (dz * dz);
}
double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}
}  // namespace

std::vector<PointType> voxelGridSmall(const std::vector<PointType>& in, float leaf) {
    std::vector<PointType> out;
    if (in.empty()) return out;
    const float inv = 1.0f / leaf;
    float mn[3] = {in[0].x, in[0].y, in[0].z}, mx[3] = {in[0].x, in[0].y, in[0].z};
    for (const auto& p : in) {
        mn[0] = std::min(mn[0], p.x); mx[0] = std::max(mx[0], p.x);
        mn[1] = std::min(mn[1], p.y); mx[1] = std::max(mx[1], p.y);
        mn[2] = std::min(mn[2], p.z); mx[2] = std::max(mx[2], p.z);
    }
    const long long dx = (long long)((mx[0] - mn[0]) * inv) + 1, dy = (long long)((mx[1] - mn[1]) * inv) + 1, dz = (long long)((mx[2] - mn[2]) * inv) + 1;
    if (dx * dy * dz > (long long)INT_MAX) return in;
    int minb[3], divb[3];
    for (int a = 0; a < 3; ++a) { minb[a] = (int)std::floor(mn[a] * inv); divb[a] = (int)std::floor(mx[a] * inv) - minb[a] + 1; }
    std::vector<std::pair<int, int>> keys(in.size());
    for (size_t i = 0; i < in.size(); ++i) {
        const int i0 = (int)(std::floor(in[i].x * inv) - (float)minb[0]);
        const int i1 = (int)(std::floor(in[i].y * inv) - (float)minb[1]);
        const int i2 = (int)(std::floor(in[i].z * inv) - (float)minb[2]);
        keys[i] = {i0 + i1 * divb[0] + i2 * divb[0] * divb[1], (int)i};
    }
    std::sort(keys.begin(), keys.end());
    for (size_t i = 0; i < keys.size();) {
        size_t j = i; float sx = 0, sy = 0, sz = 0, sw = 0;
        for (; j < keys.size() && keys[j].first == keys[i].first; ++j) { const auto& p = in[keys[j].second]; sx += p.x; sy += p.y; sz += p.z; sw += p.intensity; }
        const float c = (float)(j - i);
        out.push_back(PointType{sx / c, sy / c, sz / c, sw / c});
        i = j;
    }
    return out;
}

mapOptimization::mapOptimization(const std::string& yaml_path, const char* mode_override, int device) {
    if (!load(yaml_path)) { error_ = param_error; return; }
    if (mode_override && *mode_override) {
        const std::string m(mode_override);
        if (m == "lio") mode = ModeType::LIO; else if (m == "relo") mode = ModeType::RELO;
        else { error_ = "Invalid Mode Type (must be either 'lio' or 'relo'): " + m; return; }
    }
    if (device < 0) return;                                  // host logic only (CPU tests): no device context
    liloc_config cfg{};
    cfg.device = device;
    cfg.max_scan_points = N_SCAN * Horizon_SCAN;          // allocateMemory() :228-230
    cfg.max_raw_points = 1 << 21;
    if (liloc_create(&cfg, &ctx_) != LILOC_OK) { error_ = liloc_last_error(nullptr); ctx_ = nullptr; }
}

void mapOptimization::reset() {
    cloudKeyPoses3D.clear(); cloudKeyPoses6D.clear(); laserCloudMapContainer.clear(); reg_time.clear();
    for (float& v : transformTobeMapped) v = 0.f;
    for (float& v : transformTobeMappedInit) v = 0.f;
    isDegenerate = false; systemInitialized = false; poseInitialized = false;
    timeLastProcessing_ = -1; lastImuPreTransAvailable_ = false;
    lastImuTransformation_ = Affine3f::Identity(); lastImuPreTransformation_ = Affine3f::Identity();
    sel_ids_.clear(); sel_poses_.clear(); have_selection_ = false;
    nearby_ds_.clear(); nearby_n_ = 0; by_x_.clear(); by_x_n_ = 0;
    prior_ = PriorSession{};
    if (ctx_) { liloc_keyframe_clear(ctx_); liloc_ndt_clear(ctx_); }
}

mapOptimization::~mapOptimization() { if (ctx_) liloc_destroy(ctx_); }

FrameReport mapOptimization::laserCloudInfoHandler(const CloudInfo& msgIn) {
    rep_ = FrameReport{};
    timeLaserInfoCur = msgIn.stamp;
    cloudInfo = msgIn;                                                         // :252-253
    if (timeLaserInfoCur - timeLastProcessing_ < timeInterval) return rep_;    // :257
    if (mode == ModeType::RELO && !poseInitialized) {                          // :261-305
        if (!systemInitialized) return rep_;                                   // "Wait for Initialized Pose ..."
        if (!relocalizeInit()) { rep_.status = LILOC_ERR_STATE; return rep_; }
    }
    const double t0 = now_ms();
    have_selection_ = false;
    updateInitialGuess();                  // :309
    std::memcpy(rep_.guess, transformTobeMapped, sizeof(rep_.guess));
    extractSurroundingKeyFrames();         // :311
    downsampleCurrentScan();               // :313
    scan2MapOptimization();                // :315
    rep_.reg_ms = now_ms() - t0;           // :317-318
    reg_time.push_back(rep_.reg_ms);
    if (mode == ModeType::LIO) saveLIOKeyFramesAndFactor();                    // :327-330
    else saveRELOKeyFramesAndFactor();                                         // :331-334
    rep_.processed = 1;
    std::memcpy(rep_.pose, transformTobeMapped, sizeof(rep_.pose));
    rep_.is_degenerate = isDegenerate ? 1 : 0;                                 // publishOdometry :1573
    timeLastProcessing_ = timeLaserInfoCur;                                    // :352
    return rep_;
}

void mapOptimization::updateInitialGuess() {
    incrementalOdometryAffineFront = trans2Affine3f(transformTobeMapped);      // :832
    if (cloudKeyPoses3D.empty()) {                                             // :835
        if (mode == ModeType::LIO) {
            transformTobeMapped[0] = cloudInfo.imuRollInit;
            transformTobeMapped[1] = cloudInfo.imuPitchInit;
            transformTobeMapped[2] = cloudInfo.imuYawInit;
            lastImuTransformation_ = getTransformation(0, 0, 0, cloudInfo.imuRollInit, cloudInfo.imuPitchInit, cloudInfo.imuYawInit);
        } else {
            const Affine3f transImuInit = getTransformation(0, 0, 0, cloudInfo.initialGuessRoll, cloudInfo.initialGuessPitch, cloudInfo.initialGuessYaw);
            const Affine3f transCbInit = getTransformation(transformTobeMappedInit[3], transformTobeMappedInit[4], transformTobeMappedInit[5],
                                                           transformTobeMappedInit[0], transformTobeMappedInit[1], transformTobeMappedInit[2]);
            const Affine3f transInit = transCbInit * transImuInit;
            getTranslationAndEulerAngles(transInit, transformTobeMapped[3], transformTobeMapped[4], transformTobeMapped[5],
                                         transformTobeMapped[0], transformTobeMapped[1], transformTobeMapped[2]);
            lastImuTransformation_ = getTransformation(0, 0, 0, transformTobeMapped[0], transformTobeMapped[1], transformTobeMapped[2]);
        }
        return;
    }
    if (cloudInfo.odomAvailable) {                                             // :865-884
        const Affine3f transBack = getTransformation(cloudInfo.initialGuessX, cloudInfo.initialGuessY, cloudInfo.initialGuessZ,
                                                     cloudInfo.initialGuessRoll, cloudInfo.initialGuessPitch, cloudInfo.initialGuessYaw);
        if (!lastImuPreTransAvailable_) {
            lastImuPreTransformation_ = transBack;
            lastImuPreTransAvailable_ = true;
        } else {
            const Affine3f transIncre = lastImuPreTransformation_.inverse() * transBack;
            const Affine3f transFinal = trans2Affine3f(transformTobeMapped) * transIncre;
            getTranslationAndEulerAngles(transFinal, transformTobeMapped[3], transformTobeMapped[4], transformTobeMapped[5],
                                         transformTobeMapped[0], transformTobeMapped[1], transformTobeMapped[2]);
            lastImuPreTransformation_ = transBack;
            lastImuTransformation_ = getTransformation(0, 0, 0, cloudInfo.imuRollInit, cloudInfo.imuPitchInit, cloudInfo.imuYawInit);
            return;
        }
    }
    if (cloudInfo.imuAvailable && imuType) {                                   // :887-899
        const Affine3f transBack = getTransformation(0, 0, 0, cloudInfo.imuRollInit, cloudInfo.imuPitchInit, cloudInfo.imuYawInit);
        const Affine3f transIncre = lastImuTransformation_.inverse() * transBack;
        const Affine3f transFinal = trans2Affine3f(transformTobeMapped) * transIncre;
        getTranslationAndEulerAngles(transFinal, transformTobeMapped[3], transformTobeMapped[4], transformTobeMapped[5],
                                     transformTobeMapped[0], transformTobeMapped[1], transformTobeMapped[2]);
        lastImuTransformation_ = transBack;
    }
}

void mapOptimization::extractSurroundingKeyFrames() {
    if (cloudKeyPoses3D.empty()) return;       // :964-965
    extractNearby();
}

// extractNearby's pose-only part, the way the reference does it step by step: radiusSearch (ascending distance) ->
// VoxelGrid -> nearestKSearch(1) per centroid.  Used when the fast version below declines.
void mapOptimization::extractNearbyPlain(const PointType& last) {
    const float r2 = surroundingKeyframeSearchRadius * surroundingKeyframeSearchRadius;
    std::vector<std::pair<float, int>> hits;                                           // :908-914
    hits.reserve(cloudKeyPoses3D.size());
    for (int i = 0; i < (int)cloudKeyPoses3D.size(); ++i) {
        const float d = sqDist(last, cloudKeyPoses3D[i]);
        if (d < r2) hits.push_back({d, i});
    }
    std::sort(hits.begin(), hits.end());
    std::vector<PointType> surroundingKeyPoses;
    surroundingKeyPoses.reserve(hits.size());
    for (const auto& h : hits) surroundingKeyPoses.push_back(cloudKeyPoses3D[h.second]);
    nearby_ds_ = voxelGridSmall(surroundingKeyPoses, surroundingKeyframeDensity);      // :916-917
    for (auto& pt : nearby_ds_) {                                                      // :918-922
        int best = 0; float bd = sqDist(pt, cloudKeyPoses3D[0]);
        for (int i = 1; i < (int)cloudKeyPoses3D.size(); ++i) { const float d = sqDist(pt, cloudKeyPoses3D[i]); if (d < bd) { bd = d; best = i; } }
        pt.intensity = cloudKeyPoses3D[best].intensity;
    }
}

// The same result with one sort and windowed 1-NN searches (the plain version is ~30 us at a few hundred key poses, on
// every frame that follows a new keyframe).  (1) The VoxelGrid sums a voxel's poses in the order of its input, which is
// ascending (distance, index): sorting the hits once by (voxel, distance, index) gives the same groups in the same
// order.  (2) The nearest key pose of a centroid is searched in a list of the key poses sorted by x, outwards from the
// centroid's x until the x offset alone exceeds the best distance; ties go to the lowest index like a front-to-back scan.
bool mapOptimization::extractNearbyFast(const PointType& last) {
    const int N = (int)cloudKeyPoses3D.size();
    // x-sorted index, extended by the poses added since the last call (rebuilt when the list shrank)
    if (by_x_n_ > (size_t)N) { by_x_.clear(); by_x_n_ = 0; }
    for (int i = (int)by_x_n_; i < N; ++i) {
        const std::pair<float, int> e{cloudKeyPoses3D[i].x, i};
        by_x_.insert(std::upper_bound(by_x_.begin(), by_x_.end(), e), e);
    }
    by_x_n_ = (size_t)N;
    const float r2 = surroundingKeyframeSearchRadius * surroundingKeyframeSearchRadius;
    using Hit = HitScratch;
    std::vector<Hit>& hits = hit_scratch_;
    hits.clear();
    float mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
    for (int i = 0; i < N; ++i) {
        const PointType& p = cloudKeyPoses3D[i];
        const float d = sqDist(last, p);
        if (!(d < r2)) continue;
        if (hits.empty()) { mn[0] = mx[0] = p.x; mn[1] = mx[1] = p.y; mn[2] = mx[2] = p.z; }
        else {
            mn[0] = std::min(mn[0], p.x); mx[0] = std::max(mx[0], p.x);
            mn[1] = std::min(mn[1], p.y); mx[1] = std::max(mx[1], p.y);
            mn[2] = std::min(mn[2], p.z); mx[2] = std::max(mx[2], p.z);
        }
        hits.push_back(Hit{0, d, i});
    }
    nearby_ds_.clear();
    if (hits.empty()) return true;
    // pcl::VoxelGrid parameters, as in voxelGridSmall
    const float inv = 1.0f / surroundingKeyframeDensity;
    const long long dx = (long long)((mx[0] - mn[0]) * inv) + 1, dy = (long long)((mx[1] - mn[1]) * inv) + 1, dz = (long long)((mx[2] - mn[2]) * inv) + 1;
    if (dx * dy * dz > (long long)INT_MAX) return false;                   // PCL's "leaf size too small" path: plain version
    int minb[3], divb[3];
    for (int a = 0; a < 3; ++a) { minb[a] = (int)std::floor(mn[a] * inv); divb[a] = (int)std::floor(mx[a] * inv) - minb[a] + 1; }
    for (Hit& h : hits) {
        const PointType& p = cloudKeyPoses3D[h.i];
        const int i0 = (int)(std::floor(p.x * inv) - (float)minb[0]);
        const int i1 = (int)(std::floor(p.y * inv) - (float)minb[1]);
        const int i2 = (int)(std::floor(p.z * inv) - (float)minb[2]);
        h.key = i0 + i1 * divb[0] + i2 * divb[0] * divb[1];
    }
    std::sort(hits.begin(), hits.end(), [](const Hit& a, const Hit& b) {
        if (a.key != b.key) return a.key < b.key;
        if (a.d != b.d) return a.d < b.d;
        return a.i < b.i;
    });
    for (size_t g0 = 0; g0 < hits.size();) {
        size_t g1 = g0; float sx = 0, sy = 0, sz = 0, sw = 0;
        for (; g1 < hits.size() && hits[g1].key == hits[g0].key; ++g1) { const PointType& p = cloudKeyPoses3D[hits[g1].i]; sx += p.x; sy += p.y; sz += p.z; sw += p.intensity; }
        const float c = (float)(g1 - g0);
        PointType pt{sx / c, sy / c, sz / c, sw / c};
        // nearest key pose: outwards from the centroid's x in the x-sorted list
        int best = -1; float bd = 3.4e38f;
        auto consider = [&](int i) { const float d = sqDist(pt, cloudKeyPoses3D[i]); if (d < bd || (d == bd && i < best)) { bd = d; best = i; } };
        const size_t pos = (size_t)(std::lower_bound(by_x_.begin(), by_x_.end(), std::pair<float, int>{pt.x, -1}) - by_x_.begin());
        size_t r = pos; long long l = (long long)pos - 1;
        while (r < by_x_.size() || l >= 0) {
            bool moved = false;
            if (r < by_x_.size()) { const float ex = by_x_[r].first - pt.x; if (ex * ex <= bd) { consider(by_x_[r].second); ++r; moved = true; } else r = by_x_.size(); }
            if (l >= 0) { const float ex = pt.x - by_x_[(size_t)l].first; if (ex * ex <= bd) { consider(by_x_[(size_t)l].second); --l; moved = true; } else l = -1; }
            if (!moved) break;
        }
        pt.intensity = cloudKeyPoses3D[best].intensity;
        nearby_ds_.push_back(pt);
        g0 = g1;
    }
    return true;
}

void mapOptimization::extractNearby() {
    const PointType& last = cloudKeyPoses3D.back();
    // The radius search, its VoxelGrid thinning and the 1-NN index recovery (:908-922) depend only on the key poses, which
    // change when a keyframe is added (~40 % of the frames): their result is kept until then.  Only the "last 5 s" tail
    // (:924-931) depends on the frame's stamp.
    const PointType& first = cloudKeyPoses3D.front();
    const bool cached = nearby_n_ == cloudKeyPoses3D.size() && nearby_last_.x == last.x && nearby_last_.y == last.y && nearby_last_.z == last.z &&
                        nearby_first_.x == first.x && nearby_first_.y == first.y && nearby_first_.z == first.z;
    if (!cached || nearby_plain_only) {
        if (nearby_plain_only || !extractNearbyFast(last)) extractNearbyPlain(last);
        nearby_n_ = cloudKeyPoses3D.size(); nearby_last_ = last; nearby_first_ = first;
    }
    std::vector<PointType> surroundingKeyPosesDS = nearby_ds_;
    for (int i = (int)cloudKeyPoses3D.size() - 1; i >= 0; --i) {               // :924-931 (duplicates possible)
        if (timeLaserInfoCur - cloudKeyPoses6D[i].time < 5.0) surroundingKeyPosesDS.push_back(cloudKeyPoses3D[i]);
        else break;
    }
    extractCloud(surroundingKeyPosesDS);
}

void mapOptimization::extractCloud(const std::vector<PointType>& cloudToExtract) {
    sel_ids_.clear(); sel_poses_.clear();
    for (const auto& p : cloudToExtract) {
        if (pointDistance(p, cloudKeyPoses3D.back()) > surroundingKeyframeSearchRadius) continue;      // :939
        const int thisKeyInd = (int)p.intensity;
        auto it = laserCloudMapContainer.find(thisKeyInd);
        if (it == laserCloudMapContainer.end())                                                       // :946-951 cache miss
            it = laserCloudMapContainer.emplace(thisKeyInd, cloudKeyPoses6D[thisKeyInd]).first;
        const PointTypePose& kp = it->second;      // the pose the cached transformed cloud was made with
        sel_ids_.push_back(thisKeyInd);
        const float pose6[6] = {kp.roll, kp.pitch, kp.yaw, kp.x, kp.y, kp.z};
        sel_poses_.insert(sel_poses_.end(), pose6, pose6 + 6);
    }
    have_selection_ = true;                        // transform + concat + VoxelGrid (:944-956) run in scan2MapOptimization's device call
    if (laserCloudMapContainer.size() > 1000) laserCloudMapContainer.clear();                          // :959-960
}

void mapOptimization::downsampleCurrentScan() {
    if (!ctx_) { error_ = "no device context"; rep_.status = LILOC_ERR_STATE; return; }
    // VoxelGrid(mappingSurfLeafSize) of cloud_deskewed (:970-975).  With keyframes present it is fused into the
    // device call issued by scan2MapOptimization(); without, it runs here because the first keyframe needs it.
    if (!have_selection_) {
        int q = 0;
        // cloud_on_device: LILOC_SCAN_HOST, LILOC_SCAN_DEVICE or LILOC_SCAN_DESKEWED (the sweep ImageProjection left on the device)
        const int rc = liloc_scan_put_class(ctx_, LILOC_SURF, cloudInfo.cloud_deskewed, cloudInfo.cloud_size, cloudInfo.cloud_on_device, mappingSurfLeafSize, &q);
        if (rc < 0) { error_ = liloc_last_error(ctx_); rep_.status = rc; return; }
        laserCloudSurfLastDSNum = q;
        rep_.q_all = q;
    }
}

void mapOptimization::scan2MapOptimization() {
    if (cloudKeyPoses3D.empty()) return;           // :1184
    if (!ctx_) { error_ = "no device context"; rep_.status = LILOC_ERR_STATE; return; }
    liloc_s2m_result res{};
    const int K = (int)sel_ids_.size();
    const int rc = liloc_frame(ctx_, sel_ids_.data(), sel_poses_.data(), K, surroundingKeyframeMapLeafSize,
                               cloudInfo.cloud_deskewed, cloudInfo.cloud_size, cloudInfo.cloud_on_device, mappingSurfLeafSize,
                               transformTobeMapped, &s2m_opts, &res);
    rep_.status = rc;
    if (rc < 0) { error_ = liloc_last_error(ctx_); return; }
    laserCloudSurfLastDSNum = res.q_all;
    laserCloudSurfFromMapDSNum = res.map_points;
    rep_.q_all = res.q_all; rep_.map_points = res.map_points; rep_.n_selected_keyframes = K;
    rep_.iters = res.iters; rep_.converged = res.converged; rep_.n_sel = res.n_sel;
    if (rc == LILOC_SKIPPED) return;               // "Not enough features" (:1205): no transformUpdate
    isDegenerate = res.is_degenerate != 0;
    std::memcpy(matP, res.matP, sizeof(matP));
    transformUpdate();                             // :1202
}

void mapOptimization::transformUpdate() {
    if (cloudInfo.imuAvailable && imuType) {       // :1210-1227
        if (std::abs(cloudInfo.imuPitchInit) < 1.4f) {
            const double imuWeight = imuRPYWeight;
            double rollMid, pitchMid, yawMid;
            Quat tq = quatFromRPY(transformTobeMapped[0], 0, 0), iq = quatFromRPY(cloudInfo.imuRollInit, 0, 0);
            quatToRPY(slerp(tq, iq, imuWeight), rollMid, pitchMid, yawMid);
            transformTobeMapped[0] = (float)rollMid;
            tq = quatFromRPY(0, transformTobeMapped[1], 0); iq = quatFromRPY(0, cloudInfo.imuPitchInit, 0);
            quatToRPY(slerp(tq, iq, imuWeight), rollMid, pitchMid, yawMid);
            transformTobeMapped[1] = (float)pitchMid;
        }
    }
    transformTobeMapped[0] = constraintTransformation(transformTobeMapped[0], rotation_tollerance);    // :1229-1231
    transformTobeMapped[1] = constraintTransformation(transformTobeMapped[1], rotation_tollerance);
    transformTobeMapped[5] = constraintTransformation(transformTobeMapped[5], z_tollerance);
    incrementalOdometryAffineBack = trans2Affine3f(transformTobeMapped);                               // :1233
}

float mapOptimization::constraintTransformation(float value, float limit) {
    if (value < -limit) value = -limit;
    if (value > limit) value = limit;
    return value;
}

bool mapOptimization::saveFrame() {
    if (cloudKeyPoses3D.empty()) return true;
    const PointTypePose& b = cloudKeyPoses6D.back();
    const Affine3f transStart = getTransformation(b.x, b.y, b.z, b.roll, b.pitch, b.yaw);
    const Affine3f transFinal = trans2Affine3f(transformTobeMapped);
    const Affine3f transBetween = transStart.inverse() * transFinal;
    float x, y, z, roll, pitch, yaw;
    getTranslationAndEulerAngles(transBetween, x, y, z, roll, pitch, yaw);
    if (std::abs(roll) < surroundingkeyframeAddingAngleThreshold && std::abs(pitch) < surroundingkeyframeAddingAngleThreshold &&
        std::abs(yaw) < surroundingkeyframeAddingAngleThreshold && std::sqrt(x * x + y * y + z * z) < surroundingkeyframeAddingDistThreshold)
        return false;
    return true;
}

void mapOptimization::saveLIOKeyFramesAndFactor() {
    if (!saveFrame()) return;                      // :1392
    if (rep_.status < 0) return;
    const int id = (int)cloudKeyPoses3D.size();
    PointType p3{transformTobeMapped[3], transformTobeMapped[4], transformTobeMapped[5], (float)id};   // :1423-1427
    PointTypePose p6{p3.x, p3.y, p3.z, p3.intensity, transformTobeMapped[0], transformTobeMapped[1], transformTobeMapped[2], timeLaserInfoCur};
    if (liloc_keyframe_put_from_scan(ctx_, id) < 0) { error_ = liloc_last_error(ctx_); return; }       // :1453-1457
    cloudKeyPoses3D.push_back(p3);
    cloudKeyPoses6D.push_back(p6);
    rep_.keyframe_id = id;
}

bool mapOptimization::publishGlobalMap(std::vector<liloc_point>& globalMapKeyFramesDS) {
    globalMapKeyFramesDS.clear();
    if (cloudKeyPoses3D.empty()) return true;                                   // :714-715
    if (!ctx_) { error_ = "no device context"; return false; }
    const PointType& last = cloudKeyPoses3D.back();
    const float r2 = globalMapVisualizationSearchRadius * globalMapVisualizationSearchRadius;
    std::vector<std::pair<float, int>> hits;                                    // radiusSearch (:729), ascending distance
    for (int i = 0; i < (int)cloudKeyPoses3D.size(); ++i) { const float d = sqDist(last, cloudKeyPoses3D[i]); if (d < r2) hits.push_back({d, i}); }
    std::sort(hits.begin(), hits.end());
    std::vector<PointType> globalMapKeyPoses;
    for (const auto& h : hits) globalMapKeyPoses.push_back(cloudKeyPoses3D[h.second]);
    std::vector<PointType> ds = voxelGridSmall(globalMapKeyPoses, globalMapVisualizationPoseDensity);   // :735-738
    std::vector<int> ids; std::vector<float> poses;
    for (auto& pt : ds) {                                                        // nearestKSearch(pt, 1) (:739-742)
        int best = 0; float bd = sqDist(pt, cloudKeyPoses3D[0]);
        for (int i = 1; i < (int)cloudKeyPoses3D.size(); ++i) { const float d = sqDist(pt, cloudKeyPoses3D[i]); if (d < bd) { bd = d; best = i; } }
        pt.intensity = cloudKeyPoses3D[best].intensity;
        if (pointDistance(pt, last) > globalMapVisualizationSearchRadius) continue;               // :745-746
        const int k = (int)pt.intensity;
        const PointTypePose& p = cloudKeyPoses6D[k];
        ids.push_back(k);
        const float pose6[6] = {p.roll, p.pitch, p.yaw, p.x, p.y, p.z};
        poses.insert(poses.end(), pose6, pose6 + 6);
    }
    int M = 0;
    if (liloc_globalmap_build(ctx_, ids.data(), poses.data(), (int)ids.size(), globalMapVisualizationLeafSize, nullptr, 0, &M) < 0) { error_ = liloc_last_error(ctx_); return false; }
    globalMapKeyFramesDS.resize((size_t)M);
    if (M > 0 && liloc_globalmap_build(ctx_, ids.data(), poses.data(), (int)ids.size(), globalMapVisualizationLeafSize, globalMapKeyFramesDS.data(), M, &M) < 0) {
        error_ = liloc_last_error(ctx_); return false;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------------
// relo mode
// ---------------------------------------------------------------------------------------------------------
bool mapOptimization::loadPriorSession(const PointTypePose* keyPoses, int n, const liloc_point* clouds, const long long* offsets) {
    if (!ctx_) { error_ = "no device context"; return false; }
    if (n <= 0 || !keyPoses || !clouds || !offsets) { error_ = "loadPriorSession: bad arguments"; return false; }
    prior_ = PriorSession{};
    prior_.KeyPoses6D.assign(keyPoses, keyPoses + n);
    prior_.prior_size = n; prior_.end = n;
    for (int i = 0; i < n; ++i) {                                               // loadKeyCloud (dataLoader.hpp:283-304)
        const long long cnt = offsets[i + 1] - offsets[i];
        if (liloc_keyframe_put(ctx_, kPriorKeyframeBase + i, clouds + offsets[i], (int)cnt) < 0) { error_ = liloc_last_error(ctx_); return false; }
    }
    // generateSubMaps (dataLoader.hpp:306-348): every submap_size key clouds, transformed by their key poses and
    // concatenated (NOT down-sampled: the filtered copy is overwritten, :332-336); centroid = mean key position
    float x = 0.f, y = 0.f, z = 0.f;
    int count = 0, first = 0;
    std::vector<int> ids; std::vector<float> poses;
    for (int i = 0; i < n; ++i) {
        ++count;
        const PointTypePose& p = prior_.KeyPoses6D[i];
        ids.push_back(kPriorKeyframeBase + i);
        const float pose6[6] = {p.roll, p.pitch, p.yaw, p.x, p.y, p.z};
        poses.insert(poses.end(), pose6, pose6 + 6);
        x += p.x; y += p.y; z += p.z;
        if (count % prior_.submap_size == 0 || i == n - 1) {
            const int sid = (int)prior_.SubMapCenteriod.size();
            int leaves = 0, pts = 0;
            if (liloc_ndt_target_put_keyframes(ctx_, sid, ids.data(), poses.data(), count, ndtResolution, &leaves, &pts) < 0) {
                error_ = liloc_last_error(ctx_); return false;
            }
            prior_.SubMapCenteriod.push_back(PointType{x / (float)count, y / (float)count, z / (float)count, (float)sid});
            prior_.submap_first.push_back(first); prior_.submap_count.push_back(count); prior_.submap_points.push_back(pts);
            prior_.submap_vertexes.emplace_back();      // the shared vertex cloud
            first = i + 1; count = 0; x = y = z = 0.f; ids.clear(); poses.clear();
        }
    }
    return true;
}

bool mapOptimization::loadPriorSessionDir(const std::string& dir) {
    std::vector<PointTypePose> kp;
    std::vector<std::vector<liloc_point>> clouds;
    if (!loadSessionGraph(dir + "/singlesession_posegraph.g2o", kp, error_)) return false;
    if (!loadKeyClouds(dir + "/PCDs", clouds, error_)) return false;
    if (clouds.size() != kp.size()) { error_ = "session " + dir + ": " + std::to_string(kp.size()) + " vertices but " + std::to_string(clouds.size()) + " key clouds"; return false; }
    std::vector<liloc_point> flat;
    std::vector<long long> offs(1, 0);
    for (const auto& c : clouds) { flat.insert(flat.end(), c.begin(), c.end()); offs.push_back((long long)flat.size()); }
    return loadPriorSession(kp.data(), (int)kp.size(), flat.data(), offs.data());
}

// save_map (lio, :419-541) / save_session (relo, :543-635) at `resolution` (req.resolution).  Which session: the reference's
// save_session writes data_loader's key poses and key clouds -- the PRIOR session including everything addNewPrior joined --
// and refuses to run outside relo mode; save_map writes the current session's cloudKeyPoses6D / surfCloudKeyFrames.  Both
// lay out the same directory.  Every key cloud goes through VoxelGrid(resolution) into PCDs/<i>.pcd; globalMap.pcd is ALL key
// clouds, transformed by their key poses, concatenated in key order and filtered at `resolution` (:590-619).
// resolution == 0: the reference never assigns surfCloudDS / globalMapCloudDS then and writes empty clouds; here the key
// clouds are written as they are and the global map is filtered at mappingSurfLeafSize (the filter's standing leaf).
bool mapOptimization::saveSession(const std::string& dir, float resolution) {
    if (!ctx_) { error_ = "no device context"; return false; }
    if (resolution < 0.f || resolution != resolution) { error_ = "saveSession: bad resolution"; return false; }
    std::error_code ec;
    std::filesystem::create_directories(dir + "/PCDs", ec);
    if (ec) { error_ = "cannot create " + dir + "/PCDs: " + ec.message(); return false; }
    const bool prior = mode == ModeType::RELO && !prior_.KeyPoses6D.empty();
    const std::vector<PointTypePose>& keyPoses = prior ? prior_.KeyPoses6D : cloudKeyPoses6D;
    const int id0 = prior ? kPriorKeyframeBase : 0;
    if (!saveSessionGraph(dir + "/singlesession_posegraph.g2o", keyPoses, error_)) return false;
    if (!savePosesPCDBinary(dir + "/transformations.pcd", keyPoses, error_)) return false;
    if (!prior) {                                     // save_map also writes the key positions (:437)
        std::vector<liloc_point> traj;
        for (const PointType& p : cloudKeyPoses3D) traj.push_back(liloc_point{p.x, p.y, p.z, p.intensity});
        if (!savePCDBinary(dir + "/trajectory.pcd", traj.data(), (int)traj.size(), error_)) return false;
    }
    std::vector<liloc_point> buf, ds;
    std::vector<int> ids; std::vector<float> poses;
    long long total = 0;
    for (int k = 0; k < (int)keyPoses.size(); ++k) {
        int n = 0;
        if (liloc_keyframe_get(ctx_, LILOC_SURF, id0 + k, nullptr, 0, &n) < 0) { error_ = liloc_last_error(ctx_); return false; }
        buf.resize((size_t)(n > 0 ? n : 1));
        if (liloc_keyframe_get(ctx_, LILOC_SURF, id0 + k, buf.data(), n, &n) < 0) { error_ = liloc_last_error(ctx_); return false; }
        const liloc_point* out = buf.data();
        int m = n;
        if (resolution != 0.f && n > 0) {
            ds.resize((size_t)n);
            if (liloc_voxelgrid(ctx_, buf.data(), n, resolution, ds.data(), n, &m) < 0) { error_ = liloc_last_error(ctx_); return false; }
            out = ds.data();
        }
        if (!savePCDBinary(dir + "/PCDs/" + std::to_string(k) + ".pcd", out, m, error_)) return false;
        const PointTypePose& p = keyPoses[k];
        const float pose6[6] = {p.roll, p.pitch, p.yaw, p.x, p.y, p.z};
        ids.push_back(id0 + k); poses.insert(poses.end(), pose6, pose6 + 6);
        total += n;
    }
    const float leaf = resolution != 0.f ? resolution : mappingSurfLeafSize;
    std::vector<liloc_point> gm;
    int M = 0;
    if (!ids.empty() && total > 0) {
        if (liloc_globalmap_build(ctx_, ids.data(), poses.data(), (int)ids.size(), leaf, nullptr, 0, &M) < 0) { error_ = liloc_last_error(ctx_); return false; }
        gm.resize((size_t)(M > 0 ? M : 1));
        if (M > 0 && liloc_globalmap_build(ctx_, ids.data(), poses.data(), (int)ids.size(), leaf, gm.data(), M, &M) < 0) { error_ = liloc_last_error(ctx_); return false; }
    }
    return savePCDBinary(dir + "/globalMap.pcd", gm.data(), M, error_);
}

void mapOptimization::initialposeHandler(float x, float y, float yaw) {        // :755-769: only x, y and yaw are taken
    transformTobeMappedInit[0] = 0.f; transformTobeMappedInit[1] = 0.f; transformTobeMappedInit[2] = yaw;
    transformTobeMappedInit[3] = x;   transformTobeMappedInit[4] = y;   transformTobeMappedInit[5] = 0.f;
    systemInitialized = true;
}

int mapOptimization::searchNearestSubMap(float x, float y, float z) const {
    int best = -1; float bd = 0.f;
    const PointType q{x, y, z, 0.f};
    for (int i = 0; i < (int)prior_.SubMapCenteriod.size(); ++i) {
        const float d = sqDist(q, prior_.SubMapCenteriod[i]);
        if (best < 0 || d < bd) { bd = d; best = i; }
    }
    return best;
}

// kdtreeSearchVertex_->radiusSearch(pose, 5.0), ascending distance, first three (dataLoader.hpp:366-389).  The reference
// searches subMapVertexCloudVec_[map_id].  For the submaps of the LOADED session that is one and the same cloud:
// generateSubMaps pushes the same, never cleared `vertexCloud` pointer for every submap (:311-325), so each entry is the whole
// loaded trajectory -- near a submap border the vertices of the neighbouring submap are found too.  A submap cut by
// addNewPrior has a vertex cloud of its own (anchorOptimize.hpp:343-361).  Vertices this session added within the last 10 key
// poses are skipped (:376-380; the test is `intensity > prior_size`, so the very first added pose is not).
std::vector<int> mapOptimization::searchVertexes(int submap, float x, float y, float z) const {
    std::vector<std::pair<float, int>> hits;
    const PointType q{x, y, z, 0.f};
    auto consider = [&](int id) {
        const PointTypePose& p = prior_.KeyPoses6D[id];
        const float d = sqDist(q, PointType{p.x, p.y, p.z, 0.f});
        if (d < 25.0f) hits.push_back({d, id});
    };
    const std::vector<int>& own = prior_.submap_vertexes[submap];
    if (own.empty()) for (int id = 0; id < prior_.prior_size; ++id) consider(id);
    else for (int id : own) consider(id);
    std::sort(hits.begin(), hits.end());
    const int curNew = (int)prior_.KeyPoses6D.size();
    std::vector<int> out;
    int count = 0;
    for (const auto& h : hits) {
        if (h.second > prior_.prior_size && std::abs(curNew - h.second) <= 10) continue;
        if (++count > 3) break;
        out.push_back(h.second);
    }
    return out;
}

// AnchorOptimization::addNewPrior, cloud side (anchorOptimize.hpp:272-380).  Called when fewer than two prior vertices are
// within reach (:387-390): the frame's pose and its FULL scan (curCloud_ = laserCloudSurfLast) join the prior session --
// at once after a frame that did match (first_add), otherwise only after the saveFrame-style motion gate against the last
// added pose -- and after every submap_size additions past the first a new submap is cut from them: transformed,
// concatenated, voxel-filtered at 0.2 (this one is: :363-367) and gridded for NDT.  The GTSAM factors of the function
// stay out of scope.
bool mapOptimization::addNewPrior() {
    PriorSession& ps = prior_;
    const PointTypePose cur{transformTobeMapped[3], transformTobeMapped[4], transformTobeMapped[5], (float)ps.KeyPoses6D.size(),
                            transformTobeMapped[0], transformTobeMapped[1], transformTobeMapped[2], timeLaserInfoCur};
    auto append = [&]() -> bool {
        const int id = (int)ps.KeyPoses6D.size();
        if (liloc_keyframe_put_from_raw_scan(ctx_, kPriorKeyframeBase + id) < 0) { error_ = liloc_last_error(ctx_); return false; }
        ps.KeyPoses6D.push_back(cur);
        ps.increNodePtIds.push_back(id);
        ++ps.count;
        ps.first_add = false;
        ps.lastPose = cur;
        rep_.prior_added = id;
        return true;
    };
    if (ps.first_add) return append();
    const Affine3f transStart = getTransformation(ps.lastPose.x, ps.lastPose.y, ps.lastPose.z, ps.lastPose.roll, ps.lastPose.pitch, ps.lastPose.yaw);
    const Affine3f transBetween = transStart.inverse() * trans2Affine3f(transformTobeMapped);
    float x, y, z, roll, pitch, yaw;
    getTranslationAndEulerAngles(transBetween, x, y, z, roll, pitch, yaw);
    if (std::abs(roll) < surroundingkeyframeAddingAngleThreshold && std::abs(pitch) < surroundingkeyframeAddingAngleThreshold &&
        std::abs(yaw) < surroundingkeyframeAddingAngleThreshold && std::sqrt(x * x + y * y + z * z) < surroundingkeyframeAddingDistThreshold)
        return true;
    if (!append()) return false;
    if (ps.count % ps.submap_size == 0) {                                        // :341-377
        std::vector<int> ids, vertexes; std::vector<float> poses;
        float xx = 0.f, xy = 0.f, xz = 0.f;
        for (int i = ps.end; i < (int)ps.KeyPoses6D.size(); ++i) {
            const PointTypePose& p = ps.KeyPoses6D[i];
            ids.push_back(kPriorKeyframeBase + i); vertexes.push_back(i);
            const float pose6[6] = {p.roll, p.pitch, p.yaw, p.x, p.y, p.z};
            poses.insert(poses.end(), pose6, pose6 + 6);
            xx += p.x; xy += p.y; xz += p.z;
        }
        const int sid = (int)ps.SubMapCenteriod.size();
        int leaves = 0, pts = 0;
        if (liloc_ndt_target_put_keyframes_filtered(ctx_, sid, ids.data(), poses.data(), (int)ids.size(), 0.2f, ndtResolution, &leaves, &pts) < 0) {
            error_ = liloc_last_error(ctx_); return false;
        }
        const float n = (float)ps.submap_size;                                   // divided by submap_size, not by the count (:355-358)
        ps.SubMapCenteriod.push_back(PointType{xx / n, xy / n, xz / n, (float)sid});
        ps.submap_first.push_back(ps.end); ps.submap_count.push_back((int)ids.size()); ps.submap_points.push_back(pts);
        ps.submap_vertexes.push_back(vertexes);
        ps.end = (int)ps.KeyPoses6D.size();
        ps.count = 0;
        rep_.prior_new_submap = sid;
    }
    return true;
}

// registration->setInputTarget(submap) / setInputSource(laserCloudSurfLast) / align x n_align / getFinalTransformation
int mapOptimization::ndtAlign(int submap, const float guess6[6], int n_align, float out6[6], int* iterations, int* converged) {
    int rc = liloc_ndt_source_put(ctx_, /*source_id=*/0, cloudInfo.cloud_deskewed, cloudInfo.cloud_size, cloudInfo.cloud_on_device);
    if (rc < 0) return rc;
    liloc_ndt_opts o{};
    o.step_size = 0.1; o.outlier_ratio = 0.55; o.trans_eps = ndtEpsilon; o.max_iterations = 35;          // pclomp defaults + :178
    o.search = regMethod == "DIRECT7" ? LILOC_NDT_DIRECT7 : (regMethod == "KDTREE" ? LILOC_NDT_KDTREE : LILOC_NDT_DIRECT1);   // :181-199
    o.n_align = n_align;
    liloc_ndt_job job{};
    job.target_id = submap; job.source_id = 0;
    eulerToMatrix4(guess6[0], guess6[1], guess6[2], guess6[3], guess6[4], guess6[5], job.guess);         // :276-280
    liloc_ndt_out out{};
    rc = liloc_ndt_align_batch(ctx_, &job, 1, &o, &out);
    if (rc < 0) return rc;
    matrix4ToEuler(out.T, out6[0], out6[1], out6[2], out6[3], out6[4], out6[5]);                         // :288-289
    if (iterations) *iterations = out.iterations;
    if (converged) *converged = out.converged;
    return 0;
}

int mapOptimization::commInit(int rank, int world, const void* id128) {
    if (!ctx_) { error_ = "no device context"; return LILOC_ERR_STATE; }
    const int rc = liloc_comm_init(ctx_, rank, world, id128);
    if (rc < 0) error_ = liloc_last_error(ctx_);
    return rc;
}

// Hypothesis h of n_yaw * n_xy around `center6`: yaw offset 2 pi (h / n_xy) / n_yaw; position k = h % n_xy on a sunflower
// spiral (r = radius sqrt(k / n_xy), angle k x the golden angle; k = 0 is the centre itself).  Deterministic, so every
// rank (and the tests) builds the same list.
void mapOptimization::hypothesisPose(const float center6[6], int n_yaw, int n_xy, float radius, int h, float out6[6]) {
    const int iy = h / n_xy, k = h % n_xy;
    const double golden = 2.399963229728653;
    const double r = (double)radius * std::sqrt((double)k / (double)n_xy), a = golden * (double)k;
    for (int i = 0; i < 6; ++i) out6[i] = center6[i];
    out6[2] = (float)((double)center6[2] + 6.283185307179586 * (double)iy / (double)n_yaw);
    out6[3] = (float)((double)center6[3] + r * std::cos(a));
    out6[4] = (float)((double)center6[4] + r * std::sin(a));
}

bool mapOptimization::relocalizeInit() {
    if (!ctx_) { error_ = "no device context"; return false; }
    if (prior_.SubMapCenteriod.empty()) { error_ = "relo mode: no prior session loaded"; return false; }
    const int H = relo_n_yaw_ * relo_n_xy_;
    if (H <= 1) {
        const int submap = searchNearestSubMap(transformTobeMappedInit[3], transformTobeMappedInit[4], transformTobeMappedInit[5]);   // :267-270
        float out6[6];
        int it = 0, conv = 0;
        if (ndtAlign(submap, transformTobeMappedInit, /*n_align=*/2, out6, &it, &conv) < 0) { error_ = liloc_last_error(ctx_); return false; }   // :273-287
        std::memcpy(transformTobeMappedInit, out6, sizeof(out6));                   // :291-296
        systemInitialized = true; poseInitialized = true;                          // :298-299
        rep_.relo_initialized = 1; rep_.relo_submap = submap; rep_.relo_iterations = it; rep_.relo_converged = conv;
        std::memcpy(rep_.relo_pose, out6, sizeof(out6));
        return true;
    }
    // every hypothesis = the reference's init (:267-290) from another starting pose; sharded over the ranks inside the C ABI
    if (liloc_ndt_source_put(ctx_, /*source_id=*/0, cloudInfo.cloud_deskewed, cloudInfo.cloud_size, cloudInfo.cloud_on_device) < 0) { error_ = liloc_last_error(ctx_); return false; }
    std::vector<liloc_ndt_job> items((size_t)H);
    for (int h = 0; h < H; ++h) {
        float g6[6];
        hypothesisPose(transformTobeMappedInit, relo_n_yaw_, relo_n_xy_, relo_radius_, h, g6);
        items[h].target_id = searchNearestSubMap(g6[3], g6[4], g6[5]);
        items[h].source_id = 0;
        eulerToMatrix4(g6[0], g6[1], g6[2], g6[3], g6[4], g6[5], items[h].guess);
    }
    liloc_ndt_opts o{};
    o.step_size = 0.1; o.outlier_ratio = 0.55; o.trans_eps = ndtEpsilon; o.max_iterations = 35;
    o.search = regMethod == "DIRECT7" ? LILOC_NDT_DIRECT7 : (regMethod == "KDTREE" ? LILOC_NDT_KDTREE : LILOC_NDT_DIRECT1);
    o.n_align = 2;
    std::vector<float> rows((size_t)H * 8);
    if (liloc_ndt_align_batch_sharded(ctx_, items.data(), H, &o, LILOC_SHARD_ROUND_ROBIN, rows.data(), nullptr) < 0) { error_ = liloc_last_error(ctx_); return false; }
    int best = 0;
    for (int h = 1; h < H; ++h) if (rows[(size_t)h * 8 + 6] > rows[(size_t)best * 8 + 6]) best = h;      // trans_probability: higher is better
    std::memcpy(transformTobeMappedInit, &rows[(size_t)best * 8], 6 * sizeof(float));
    systemInitialized = true; poseInitialized = true;
    rep_.relo_initialized = 1; rep_.relo_submap = items[best].target_id; rep_.relo_iterations = (int)rows[(size_t)best * 8 + 7]; rep_.relo_converged = 1;
    rep_.relo_hypothesis = best; rep_.relo_score = rows[(size_t)best * 8 + 6];
    std::memcpy(rep_.relo_pose, &rows[(size_t)best * 8], 6 * sizeof(float));
    return true;
}

void mapOptimization::saveRELOKeyFramesAndFactor() {
    if (rep_.status < 0) return;
    // optimize->addOdomFactors() -> addScanMatchingFactor() (anchorOptimize.hpp:382-407): one align of the full scan
    // against the submap nearest to the current pose, started from the current pose.  The reference feeds the result
    // into its ISAM2 graph (out of scope here); it is reported in the frame report instead.
    if (!prior_.SubMapCenteriod.empty() && !rep_.relo_initialized) {
        const int submap = searchNearestSubMap(transformTobeMapped[3], transformTobeMapped[4], transformTobeMapped[5]);
        const std::vector<int> vertexes = searchVertexes(submap, transformTobeMapped[3], transformTobeMapped[4], transformTobeMapped[5]);
        rep_.relo_n_vertices = (int)vertexes.size();
        for (size_t v = 0; v < vertexes.size(); ++v) rep_.relo_vertex[v] = vertexes[v];
        if (vertexes.size() <= 1) {                // `usingVertexes_->size() <= 1` => addNewPrior() instead of matching (:387-390)
            if (liloc_ndt_source_put(ctx_, /*source_id=*/0, cloudInfo.cloud_deskewed, cloudInfo.cloud_size, cloudInfo.cloud_on_device) < 0 || !addNewPrior()) {
                if (error_.empty()) error_ = liloc_last_error(ctx_);
                rep_.status = LILOC_ERR_CUDA; return;
            }
        } else {
            prior_.first_add = true;               // :392
            float out6[6]; int it = 0, conv = 0;
            if (ndtAlign(submap, transformTobeMapped, 1, out6, &it, &conv) < 0) { error_ = liloc_last_error(ctx_); rep_.status = LILOC_ERR_CUDA; return; }
            rep_.relo_submap = submap; rep_.relo_iterations = it; rep_.relo_converged = conv;
            std::memcpy(rep_.relo_pose, out6, sizeof(out6));
            // per prior vertex: getICPFitnessScore(its key cloud at its prior pose, the scan at the matched pose), capped
            // at 2 (:427-446).  Both clouds are already on the device.
            for (size_t v = 0; v < vertexes.size(); ++v) {
                rep_.relo_vertex[v] = vertexes[v];
                // vertices this session added carry no fitness score (:431-434)
                if (std::find(prior_.increNodePtIds.begin(), prior_.increNodePtIds.end(), vertexes[v]) != prior_.increNodePtIds.end()) { rep_.relo_fitness[v] = -1.f; continue; }
                const PointTypePose& p = prior_.KeyPoses6D[vertexes[v]];
                const float pose_i[6] = {p.roll, p.pitch, p.yaw, p.x, p.y, p.z};
                float score = 0.f;
                if (liloc_icp_fitness_resident(ctx_, kPriorKeyframeBase + vertexes[v], pose_i, /*ndt source*/ 0, out6, &score) < 0) {
                    error_ = liloc_last_error(ctx_); rep_.status = LILOC_ERR_CUDA; return;
                }
                if (score > 2) score = 2;
                rep_.relo_fitness[v] = score;
            }
        }
    }
    // every processed frame becomes a keyframe (the saveFrame() gate is commented out, :1338-1339)
    const int id = (int)cloudKeyPoses3D.size();
    PointType p3{transformTobeMapped[3], transformTobeMapped[4], transformTobeMapped[5], (float)id};
    PointTypePose p6{p3.x, p3.y, p3.z, p3.intensity, transformTobeMapped[0], transformTobeMapped[1], transformTobeMapped[2], timeLaserInfoCur};
    if (liloc_keyframe_put_from_scan(ctx_, id) < 0) { error_ = liloc_last_error(ctx_); return; }       // :1381-1385
    cloudKeyPoses3D.push_back(p3);
    cloudKeyPoses6D.push_back(p6);
    rep_.keyframe_id = id;
}

}  // namespace liloc_host
