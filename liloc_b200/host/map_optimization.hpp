// liloc_b200/host/map_optimization.hpp -- ROS-free host side of `class mapOptimization`
// (src/mapOptmization.cpp:18): same member names, same per-frame call order, same parameters; every
// PCL / Eigen / OpenCV call on the scan-to-map path is replaced by a call into the C ABI
// (include/liloc_gpu.h).  What stays on the host is exactly what SURVEY.md 8(a) marks "stays on host":
// updateInitialGuess (a10), extractNearby keyframe selection (a11), transformUpdate (a9), saveFrame.
//
// Out of scope here: ROS publishers/services, GTSAM.  In lio mode the
// reference's ISAM2 graph holds one prior and a chain of odometry BetweenFactors with no loop
// closures (correctPoses() is commented out, :346), whose unique optimum is the chain of measured
// poses, so saveLIOKeyFramesAndFactor stores transformTobeMapped itself as the key pose.
#pragma once
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/liloc_gpu.h"
#include "affine.hpp"
#include "param_server.hpp"

namespace liloc_host {

struct PointType { float x, y, z, intensity; };                                     // pcl::PointXYZI
struct PointTypePose { float x, y, z, intensity, roll, pitch, yaw; double time; };  // PointXYZIRPYT

// The fields of msg/cloud_info.msg that mapOptimization reads.
struct CloudInfo {
    double stamp = 0.0;                          // header.stamp
    int imuAvailable = 0, odomAvailable = 0;
    float imuRollInit = 0, imuPitchInit = 0, imuYawInit = 0;
    float initialGuessX = 0, initialGuessY = 0, initialGuessZ = 0, initialGuessRoll = 0, initialGuessPitch = 0, initialGuessYaw = 0;
    const liloc_point* cloud_deskewed = nullptr; // host pointer, or device pointer when cloud_on_device
    int cloud_size = 0;
    int cloud_on_device = 0;
};

struct FrameReport {                             // what the reference would publish / log for one frame
    int processed = 0;                           // 0: dropped by the timeInterval gate (:257)
    int status = 0;                              // liloc return code of the registration (LILOC_SKIPPED ...)
    float pose[6] = {0, 0, 0, 0, 0, 0};          // transformTobeMapped after the frame
    float guess[6] = {0, 0, 0, 0, 0, 0};         // transformTobeMapped after updateInitialGuess (LM start point)
    int is_degenerate = 0;
    int iters = 0, converged = 0, n_sel = 0;
    int q_all = 0, map_points = 0, n_raw = 0, n_selected_keyframes = 0;
    int keyframe_id = -1;                        // id if this frame became a keyframe
    double reg_ms = 0.0;                         // reg_time (:307-318): init guess + map + downsample + scan2map
    // relo mode
    int relo_initialized = 0;                    // this frame ran the relocalisation init (:261-305)
    int relo_submap = -1;                        // prior submap the scan was matched against (searchNearestSubMapAndVertex)
    int relo_iterations = 0, relo_converged = 0;
    float relo_pose[6] = {0, 0, 0, 0, 0, 0};     // NDT result as [roll, pitch, yaw, x, y, z] (the scan-matching factor's poseTo)
    int relo_n_vertices = 0;                     // prior vertices within 5 m of the pose (usingVertexes_, at most 3)
    int relo_vertex[3] = {-1, -1, -1};           // their prior keyframe ids
    float relo_fitness[3] = {0, 0, 0};           // getICPFitnessScore of each vertex' key cloud against the matched scan, capped at 2
                                                 // (-1: a vertex this session added itself, no score, anchorOptimize.hpp:431-434)
    int relo_hypothesis = -1;                    // multi-hypothesis init: index of the winning hypothesis
    float relo_score = 0.f;                      // ... and its trans_probability (score / points)
    int prior_added = -1;                        // addNewPrior appended this frame to the prior session under this key-pose id
    int prior_new_submap = -1;                   // ... and cut this new submap from the last submap_size additions
};

// What mapOptimization keeps of the prior session (dataManager::Session, include/dataManager/dataLoader.hpp:128-391):
// key poses and the submaps cut every `submap_size` keyframes.  The key clouds and the submap grids live on the device.
struct PriorSession {
    std::vector<PointTypePose> KeyPoses6D;       // prior key poses (g2o vertices)
    std::vector<PointType> SubMapCenteriod;      // mean key position of each submap (:321-325)
    std::vector<int> submap_first, submap_count; // keyframe range of each submap
    std::vector<int> submap_points;              // points of each submap cloud
    int submap_size = 10;
    // ---- growth of the prior map during the session: AnchorOptimization::addNewPrior (anchorOptimize.hpp:272-380) ----
    int prior_size = 0;                          // key poses of the session as loaded (dataLoader.hpp:186)
    std::vector<std::vector<int>> submap_vertexes;   // per submap: key-pose ids of ITS vertex cloud; empty = the loader's shared cloud
                                                     // (every loaded submap sees all prior_size loaded key poses, dataLoader.hpp:311-325)
    bool first_add = true;                       // member first_add (:51): the next addNewPrior adds unconditionally
    PointTypePose lastPose{};                    // static lastPose of addNewPrior
    int end = 0;                                 // static end: first key pose of the submap being collected
    int count = 0;                               // static count: key poses added since the last cut
    std::vector<int> increNodePtIds;             // key poses added by this session (increNodePtIds_)
};

class mapOptimization : public ParamServer {
public:
    // ---- state with the reference's names ----
    std::vector<PointType> cloudKeyPoses3D;
    std::vector<PointTypePose> cloudKeyPoses6D;
    std::map<int, PointTypePose> laserCloudMapContainer;   // keyframe id -> pose its cached transform was made with (:69)
    float transformTobeMapped[6] = {0, 0, 0, 0, 0, 0};
    float transformTobeMappedInit[6] = {0, 0, 0, 0, 0, 0};
    bool isDegenerate = false;
    float matP[36] = {0};
    bool systemInitialized = false, poseInitialized = false;
    int laserCloudSurfFromMapDSNum = 0, laserCloudSurfLastDSNum = 0;
    Affine3f incrementalOdometryAffineFront = Affine3f::Identity(), incrementalOdometryAffineBack = Affine3f::Identity();
    CloudInfo cloudInfo;
    double timeLaserInfoCur = 0.0;
    std::vector<double> reg_time;

    explicit mapOptimization(const std::string& yaml_path, const char* mode_override, int device);
    ~mapOptimization();
    mapOptimization(const mapOptimization&) = delete;
    mapOptimization& operator=(const mapOptimization&) = delete;

    void reset();                               // forget the session: key poses, cache, device keyframes
    bool ok() const { return error_.empty(); }
    const std::string& error() const { return error_; }
    liloc_ctx* context() { return ctx_; }

    // ---- the reference's per-frame methods (src/mapOptmization.cpp:246-353) ----
    FrameReport laserCloudInfoHandler(const CloudInfo& msgIn);
    void updateInitialGuess();                 // :831-900
    void extractSurroundingKeyFrames();        // :963-968
    void extractNearby();                      // :902-934
    void extractNearbyPlain(const PointType& last);   // its pose-only part, step by step like the reference
    bool extractNearbyFast(const PointType& last);    // the same result with one sort and windowed 1-NN searches
    void extractCloud(const std::vector<PointType>& cloudToExtract);   // :936-961
    void downsampleCurrentScan();              // :970-975
    void scan2MapOptimization();               // :1183-1207
    void transformUpdate();                    // :1209-1234
    float constraintTransformation(float value, float limit);          // :1236-1243
    bool saveFrame();                          // :1245-1264
    void saveLIOKeyFramesAndFactor();          // :1391-1461
    // the cloud publishGlobalMap (:698-753, 1 Hz visualisation thread) would publish: key poses within
    // globalMapVisualizationSearchRadius thinned at ...PoseDensity, their clouds transformed + down-sampled at ...LeafSize
    bool publishGlobalMap(std::vector<liloc_point>& globalMapKeyFramesDS);

    // ---- relo mode: relocalisation against a prior session (the NDT side of :171-206, :261-305, :755-769 and of
    //      AnchorOptimization::addScanMatchingFactor, anchorOptimize.hpp:382-421).  The GTSAM graphs are out of scope
    //      (SURVEY.md 2, rows 12-13): the registration result is reported, not fused.
    // Prior key clouds (body frame, concatenated; offsets has n + 1 entries) + key poses -> device keyframes + submaps.
    bool loadPriorSession(const PointTypePose* keyPoses, int n, const liloc_point* clouds, const long long* offsets);
    // The same from a session directory in the reference's on-disk format (singlesession_posegraph.g2o + PCDs/<k>.pcd,
    // dataLoader.hpp:207-304), and the writer of that format for the current session (save_session / save_map,
    // :420-635: pose graph + one PCD per keyframe; nothing is deleted, unlike the reference's `rm -r`).
    bool loadPriorSessionDir(const std::string& session_dir);
    bool saveSession(const std::string& session_dir, float resolution = 0.f);
    void initialposeHandler(float x, float y, float yaw);                       // :755-769
    int searchNearestSubMap(float x, float y, float z) const;                   // dataLoader.hpp:350-357 (1-NN on the centroids)
    std::vector<int> searchVertexes(int submap, float x, float y, float z) const;   // dataLoader.hpp:366-389 (radius 5 m, at most 3)
    bool relocalizeInit();                                                      // :261-305
    // Multi-hypothesis relocalisation (north_star; not in the reference, whose /initialpose fixes only x, y and yaw, :764-769):
    // n_yaw yaw offsets over the full circle x n_xy positions on a spiral of `radius` around the given pose, every
    // hypothesis through the reference's 2 x align against the submap nearest to it, the best trans_probability wins.
    // The hypotheses are sharded over the ranks of the context's communicator (liloc_comm_init through commInit):
    // liloc_ndt_align_batch_sharded -- every rank ends up with the same winner.  n_yaw * n_xy == 1 is the reference's init.
    void setReloHypotheses(int n_yaw, int n_xy, float radius) { relo_n_yaw_ = n_yaw < 1 ? 1 : n_yaw; relo_n_xy_ = n_xy < 1 ? 1 : n_xy; relo_radius_ = radius; }
    int commInit(int rank, int world, const void* id128);
    static void hypothesisPose(const float center6[6], int n_yaw, int n_xy, float radius, int h, float out6[6]);
    void saveRELOKeyFramesAndFactor();                                          // :1336-1389 (+ anchorOptimize.hpp:382-407)
    const PriorSession& prior() const { return prior_; }
    // AnchorOptimization::addNewPrior, cloud side (anchorOptimize.hpp:272-380): the current frame joins the prior map
    bool addNewPrior();
    static constexpr int kPriorKeyframeBase = 1 << 24;                          // device keyframe ids of the prior session

    Affine3f trans2Affine3f(const float t[6]) { return getTransformation(t[3], t[4], t[5], t[0], t[1], t[2]); }   // :404-406

    liloc_s2m_opts s2m_opts{20, 2, 50, 30, 0, 10, 1, {0}};
    bool nearby_plain_only = false;             // tests: always take extractNearbyPlain (and never the cached result)
    const std::vector<int>& lastSelection() const { return sel_ids_; }

private:
    liloc_ctx* ctx_ = nullptr;
    std::string error_;
    // function-static state of the reference, kept as members
    double timeLastProcessing_ = -1;
    Affine3f lastImuTransformation_ = Affine3f::Identity();
    bool lastImuPreTransAvailable_ = false;
    Affine3f lastImuPreTransformation_ = Affine3f::Identity();
    // staged inputs of the fused device call
    std::vector<int> sel_ids_;
    std::vector<float> sel_poses_;
    bool have_selection_ = false;
    // extractNearby's pose-only part, valid while the key poses are unchanged
    std::vector<PointType> nearby_ds_;
    size_t nearby_n_ = 0;
    PointType nearby_last_{0, 0, 0, 0}, nearby_first_{0, 0, 0, 0};
    std::vector<std::pair<float, int>> by_x_;    // key poses sorted by (x, index), extended as poses are added
    size_t by_x_n_ = 0;
    struct HitScratch { int key; float d; int i; };
    std::vector<HitScratch> hit_scratch_;
    FrameReport rep_;
    PriorSession prior_;
    int relo_n_yaw_ = 1, relo_n_xy_ = 1;
    float relo_radius_ = 2.0f;
    int ndtAlign(int submap, const float guess6[6], int n_align, float out6[6], int* iterations, int* converged);
};

// pcl::VoxelGrid on a handful of key positions (extractNearby :916-917); host-side because it is part of
// keyframe *selection* (SURVEY 8a a11), not of the point-cloud path.
std::vector<PointType> voxelGridSmall(const std::vector<PointType>& in, float leaf);

}  // namespace liloc_host
