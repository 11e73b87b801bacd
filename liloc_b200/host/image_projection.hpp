// liloc_b200/host/image_projection.hpp -- ROS-free host side of `class ImageProjection` (src/imageProjection.cpp:63):
// same member names and the same call order per sweep (cachePointCloud -> deskewInfo -> projectPointCloud ->
// publishClouds -> resetParameters).  What stays on the host is the message bookkeeping: the IMU / odometry queues, the
// gyro integration of imuDeskewInfo (a ~60-entry serial prefix sum in double) and the start-of-sweep odometry of
// odomDeskewInfo.  The per-point work of projectPointCloud -- filters, compaction, de-skew -- is one device launch
// (liloc_deskew, csrc/deskew.cuh), and the cloud it produces stays on the device: the CloudInfo handed on carries
// cloud_on_device = LILOC_SCAN_DESKEWED, which mapOptimization passes straight to the frame call.
#pragma once
#include <deque>
#include <string>
#include <utility>
#include <vector>

#include "../../include/liloc_gpu.h"
#include "affine.hpp"
#include "map_optimization.hpp"
#include "param_server.hpp"

namespace liloc_host {

struct ImuMsg {                                    // sensor_msgs::Imu
    double stamp;
    double angular_velocity[3], linear_acceleration[3];
    double orientation[4];                         // x, y, z, w
};
struct OdomMsg {                                   // nav_msgs::Odometry (odomTopic + "_incremental")
    double stamp;
    double position[3];
    double orientation[4];                         // x, y, z, w
    double covariance0;                            // pose.covariance[0]: the preintegration's reset counter
};
struct LivoxPoint { float x, y, z; unsigned char reflectivity, tag, line; unsigned offset_time; };   // livox_ros_driver::CustomPoint

const int queueLength = 2000;                      // :61

class ImageProjection : public ParamServer {
public:
    // `ctx`: the device context the de-skewed cloud is left in (mapOptimization::context() when both run in one process).
    // ctx == nullptr: host logic only (CPU tests) -- everything but the device call of projectPointCloud runs, and the
    // inputs that call would have received stay readable (lastSweep*).
    ImageProjection(const std::string& yaml_path, liloc_ctx* ctx);

    void imuHandler(const ImuMsg& imuMsg);                                   // :162-173 (imuConverter, include/utility.h:306-336)
    void odometryHandler(const OdomMsg& odometryMsg);                        // :175-178
    // :180-192.  Points in the common PointXYZIRT layout (ring / time as liloc_ring_time); has_time: the message has a
    // "time" or "t" field (:400-410).  With have_ring_time_channel == false ring / time are ignored and the relative time
    // comes from the azimuth (:279-326).  Returns true when a CloudInfo was produced (the reference publishes it).
    bool cloudHandler(double stamp, const liloc_point* pts, const liloc_ring_time* ring_time, int n, bool has_time, CloudInfo* out);
    // :194-206 / :208-268: Livox CustomMsg -> common layout (tag / line / NaN / IsNear filters, sort by time)
    bool cloudHandlerLivox(double stamp, const LivoxPoint* pts, int n, CloudInfo* out);

    bool ok() const { return error_.empty(); }
    const std::string& error() const { return error_; }
    int lastCloudSize() const { return last_size_; }                          // fullCloud->size() of the last published sweep, -1 unknown
    // what projectPointCloud handed (or would have handed) to liloc_deskew for the last published sweep
    const std::vector<double>& lastImu() const { return last_imu_; }         // 4 x (lastImuPointerCur() + 1): time | rotX | rotY | rotZ
    int lastImuPointerCur() const { return last_imu_pointer_; }
    const std::vector<liloc_point>& lastSweepPoints() const { return last_pts_; }
    const std::vector<liloc_ring_time>& lastSweepRingTime() const { return last_rt_; }

    // the sensor-specific time fields mapped onto PointXYZIRT::time (:332-379)
    static float ousterTime(unsigned t_ns) { return t_ns * 1e-9f; }
    static float mulranTime(unsigned t) { return static_cast<float>(t); }
    static float robosenseTime(double timestamp, double first_timestamp) { return (float)(timestamp - first_timestamp); }

    // ---- state with the reference's names ----
    std::deque<ImuMsg> imuQueue;
    std::deque<OdomMsg> odomQueue;
    std::vector<double> imuTime, imuRotX, imuRotY, imuRotZ;                  // queueLength entries
    int imuPointerCur = 0;
    int deskewFlag = 0;
    bool odomDeskewFlag = false;
    float odomIncreX = 0, odomIncreY = 0, odomIncreZ = 0;
    CloudInfo cloudInfo;
    double timeScanCur = 0, timeScanEnd = 0;

private:
    struct Sweep { double stamp; std::vector<liloc_point> pts; std::vector<liloc_ring_time> rt; bool has_time; };
    std::deque<Sweep> cloudQueue;
    std::deque<std::pair<double, std::vector<LivoxPoint>>> cloudQueueLivox;
    Sweep current_;
    liloc_ctx* ctx_;
    std::string error_;
    int last_size_ = -1;
    std::vector<double> last_imu_;
    int last_imu_pointer_ = -1;
    std::vector<liloc_point> last_pts_;
    std::vector<liloc_ring_time> last_rt_;
    double extRot_[3][3], extQRPY_[4];                                        // include/utility.h:275-278
    // cachePointCloudLivox's function statics (:209-211)
    bool livox_first_scan_ = true;
    double livox_last_time_ = 0.0, livox_first_scan_time_ = 0.0;

    void resetParameters();                                                   // :143-158
    bool cachePointCloud(Sweep&& msg);                                        // :270-423
    bool cachePointCloudLivox(double stamp, const LivoxPoint* pts, int n);    // :208-268
    bool deskewInfo();                                                        // :425-439
    void imuDeskewInfo();                                                     // :441-496
    void odomDeskewInfo();                                                    // :498-571
    bool projectPointCloud();                                                 // :645-673 -> liloc_deskew
    void publishClouds(CloudInfo* out);                                       // :675-679
};

}  // namespace liloc_host
