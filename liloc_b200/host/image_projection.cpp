// liloc_b200/host/image_projection.cpp -- see image_projection.hpp.  Reference: src/imageProjection.cpp.
#include "image_projection.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace liloc_host {

namespace {

// Eigen::Quaterniond(Matrix3d): the standard branch on the trace (Eigen/src/Geometry/Quaternion.h quaternionbase_assign_impl)
void quatFromMatrix(const double m[3][3], double q[4] /* x y z w */) {
    double t = m[0][0] + m[1][1] + m[2][2];
    if (t > 0.0) {
        t = std::sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (m[2][1] - m[1][2]) * t; q[1] = (m[0][2] - m[2][0]) * t; q[2] = (m[1][0] - m[0][1]) * t;
    } else {
        int i = 0;
        if (m[1][1] > m[0][0]) i = 1;
        if (m[2][2] > m[i][i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (m[k][j] - m[j][k]) * t;
        q[j] = (m[j][i] + m[i][j]) * t;
        q[k] = (m[k][i] + m[i][k]) * t;
    }
}

void quatMul(const double a[4], const double b[4], double o[4]) {       // x y z w
    o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    o[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    o[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
}

void rpyOf(const double q[4], double& roll, double& pitch, double& yaw) {   // tf::Matrix3x3(q).getRPY
    quatToRPY(Quat{q[0], q[1], q[2], q[3]}, roll, pitch, yaw);
}

// Relative time of every return of a cloud that has no time channel, from its azimuth (src/imageProjection.cpp:279-326).
// The sensor turns clockwise, so -atan2(y, x) grows along the sweep; it is unwrapped against the first return's azimuth
// until the sweep has passed the half turn and against the last return's afterwards.  A sweep lasts 0.1 s.  Types as in
// the reference: the azimuth of a return is a float, the two reference angles and every comparison are double.
void azimuthTimes(const std::vector<liloc_point>& cloud, std::vector<liloc_ring_time>& out) {
    auto azimuth = [](const liloc_point& p) -> float { return -std::atan2(p.y, p.x); };
    const double two_pi = 2 * M_PI;
    const double first = azimuth(cloud.front());
    double last = azimuth(cloud.back()) + two_pi;
    if (last - first > 3 * M_PI) last -= two_pi;
    else if (last - first < M_PI) last += two_pi;
    const double span = last - first;
    bool past_half = false;
    for (size_t i = 0; i < cloud.size(); ++i) {
        float a = azimuth(cloud[i]);
        if (past_half) {
            a += two_pi;
            if (a < last - M_PI * 3 / 2) a += two_pi;
            else if (a > last + M_PI / 2) a -= two_pi;
        } else {
            if (a < first - M_PI / 2) a += two_pi;
            else if (a > first + M_PI * 3 / 2) a -= two_pi;
            past_half = a - first > M_PI;
        }
        const float fraction = (a - first) / span;
        out[i].ring = 0;                                                  // PointXYZIRT::ring stays value-initialised
        out[i].time = 0.1 * fraction;
    }
}

}  // namespace

ImageProjection::ImageProjection(const std::string& yaml_path, liloc_ctx* ctx) : ParamServer(yaml_path), ctx_(ctx) {
    if (!param_error.empty()) error_ = param_error;
    // include/utility.h:275-278 (identity when the yaml has no extrinsics)
    double extRPY[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            extRot_[i][j] = extRotV.size() == 9 ? extRotV[3 * i + j] : (i == j ? 1.0 : 0.0);
            extRPY[i][j] = extRPYV.size() == 9 ? extRPYV[3 * i + j] : (i == j ? 1.0 : 0.0);
        }
    double q[4];
    quatFromMatrix(extRPY, q);
    const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];   // Quaterniond::inverse(): conjugate / squaredNorm
    extQRPY_[0] = -q[0] / n2; extQRPY_[1] = -q[1] / n2; extQRPY_[2] = -q[2] / n2; extQRPY_[3] = q[3] / n2;
    imuTime.assign(queueLength, 0.0); imuRotX.assign(queueLength, 0.0); imuRotY.assign(queueLength, 0.0); imuRotZ.assign(queueLength, 0.0);
    resetParameters();
}

void ImageProjection::resetParameters() {
    imuPointerCur = 0;
    odomDeskewFlag = false;
    std::fill(imuTime.begin(), imuTime.end(), 0.0); std::fill(imuRotX.begin(), imuRotX.end(), 0.0);
    std::fill(imuRotY.begin(), imuRotY.end(), 0.0); std::fill(imuRotZ.begin(), imuRotZ.end(), 0.0);
}

void ImageProjection::imuHandler(const ImuMsg& imuMsg) {
    ImuMsg o = imuMsg;                                                    // imuConverter
    for (int i = 0; i < 3; ++i) {
        o.linear_acceleration[i] = extRot_[i][0] * imuMsg.linear_acceleration[0] + extRot_[i][1] * imuMsg.linear_acceleration[1] + extRot_[i][2] * imuMsg.linear_acceleration[2];
        o.angular_velocity[i] = extRot_[i][0] * imuMsg.angular_velocity[0] + extRot_[i][1] * imuMsg.angular_velocity[1] + extRot_[i][2] * imuMsg.angular_velocity[2];
    }
    if (imuType) {
        quatMul(imuMsg.orientation, extQRPY_, o.orientation);
        const double* q = o.orientation;
        if (std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]) < 0.1) { error_ = "Invalid quaternion, please use a 9-axis IMU!"; return; }
    }
    if (correct) for (int i = 0; i < 3; ++i) o.linear_acceleration[i] *= 9.80665;      // kAccScale (:165-169)
    imuQueue.push_back(o);
}

void ImageProjection::odometryHandler(const OdomMsg& odometryMsg) { odomQueue.push_back(odometryMsg); }

bool ImageProjection::cloudHandler(double stamp, const liloc_point* pts, const liloc_ring_time* ring_time, int n, bool has_time, CloudInfo* out) {
    if (!ok()) return false;
    if (n <= 0 || !pts || (have_ring_time_channel && !ring_time)) { error_ = "cloudHandler: empty cloud"; return false; }
    Sweep s;
    s.stamp = stamp; s.has_time = has_time;
    s.pts.assign(pts, pts + n);
    if (ring_time) s.rt.assign(ring_time, ring_time + n); else s.rt.assign((size_t)n, liloc_ring_time{0, 0.f});
    if (!cachePointCloud(std::move(s))) return false;
    if (!deskewInfo()) return false;
    if (!projectPointCloud()) return false;
    publishClouds(out);
    resetParameters();
    return true;
}

bool ImageProjection::cloudHandlerLivox(double stamp, const LivoxPoint* pts, int n, CloudInfo* out) {
    if (!ok()) return false;
    if (!cachePointCloudLivox(stamp, pts, n)) return false;
    if (!deskewInfo()) return false;
    if (!projectPointCloud()) return false;
    publishClouds(out);
    resetParameters();
    return true;
}

bool ImageProjection::cachePointCloudLivox(double stamp, const LivoxPoint* pts, int n) {       // :208-268
    cloudQueueLivox.emplace_back(stamp, std::vector<LivoxPoint>(pts, pts + (n > 0 ? n : 0)));
    if (cloudQueueLivox.size() <= 2) return false;
    const std::pair<double, std::vector<LivoxPoint>> msg = std::move(cloudQueueLivox.front());
    cloudQueueLivox.pop_front();
    const double cur_time = msg.first;
    if (cur_time < livox_last_time_) { cloudQueueLivox.clear(); livox_last_time_ = cur_time; }  // "Livox Cloud Loop ."
    if (livox_first_scan_) { livox_first_scan_time_ = cur_time; livox_first_scan_ = false; }
    const double time_offset = (cur_time - livox_first_scan_time_) * (double)(1000);
    auto bad = [](float v) { return std::isinf(v) || std::isnan(v); };
    std::vector<liloc_point> P; std::vector<liloc_ring_time> R;
    for (size_t i = 1; i < msg.second.size(); ++i) {
        const LivoxPoint& p = msg.second[i]; const LivoxPoint& q = msg.second[i - 1];
        if (!(p.line < N_SCAN)) continue;
        if (!((p.tag & 0x30) == 0x10 || (p.tag & 0x30) == 0x00)) continue;
        if (bad(p.x) || bad(p.y) || bad(p.z)) continue;
        if (std::fabs(p.x - q.x) < 1e-7 || std::fabs(p.y - q.y) < 1e-7 || std::fabs(p.z - q.z) < 1e-7) continue;     // IsNear (include/utility.h:480-483)
        P.push_back(liloc_point{p.x, p.y, p.z, 0.f});
        R.push_back(liloc_ring_time{(int)p.line, (float)(time_offset + p.offset_time * 1e-6)});                       // ms
    }
    if (P.empty()) { error_ = "cachePointCloudLivox: no valid point"; return false; }
    std::vector<int> order(P.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return R[a].time < R[b].time; });                // std::sort upstream; stable = canonical
    current_.stamp = msg.first; current_.has_time = true;
    current_.pts.resize(order.size()); current_.rt.resize(order.size());
    for (size_t i = 0; i < order.size(); ++i) { current_.pts[i] = P[order[i]]; current_.rt[i] = R[order[i]]; }
    timeScanCur = current_.stamp;
    timeScanEnd = timeScanCur + current_.rt.back().time / (double)(1000);
    livox_first_scan_ = true;
    livox_last_time_ = cur_time;
    deskewFlag = 1;
    return true;
}

bool ImageProjection::cachePointCloud(Sweep&& msg) {
    cloudQueue.push_back(std::move(msg));
    if (cloudQueue.size() <= 2) return false;
    current_ = std::move(cloudQueue.front());
    cloudQueue.pop_front();
    if (!have_ring_time_channel) {                                        // :279-326: time from the azimuth, ring unknown (0)
        azimuthTimes(current_.pts, current_.rt);
        deskewFlag = 1;
    } else if (deskewFlag == 0) {                                         // :400-410
        deskewFlag = current_.has_time ? 1 : -1;
    }
    timeScanCur = current_.stamp;
    timeScanEnd = timeScanCur + current_.rt.back().time;
    return true;
}

bool ImageProjection::deskewInfo() {
    if (imuQueue.empty() || imuQueue.front().stamp > timeScanCur || imuQueue.back().stamp < timeScanEnd) return false;   // "Waiting for IMU data ..."
    imuDeskewInfo();
    odomDeskewInfo();
    return true;
}

void ImageProjection::imuDeskewInfo() {
    cloudInfo.imuAvailable = 0;
    while (!imuQueue.empty()) {
        if (imuQueue.front().stamp < timeScanCur - 0.01) imuQueue.pop_front();
        else break;
    }
    if (imuQueue.empty()) return;
    imuPointerCur = 0;
    for (int i = 0; i < (int)imuQueue.size(); ++i) {
        const ImuMsg& thisImuMsg = imuQueue[i];
        const double currentImuTime = thisImuMsg.stamp;
        if (imuType && currentImuTime <= timeScanCur) {                   // imuRPY2rosRPY
            double r, p, y;
            rpyOf(thisImuMsg.orientation, r, p, y);
            cloudInfo.imuRollInit = (float)r; cloudInfo.imuPitchInit = (float)p; cloudInfo.imuYawInit = (float)y;
        }
        if (currentImuTime > timeScanEnd + 0.01) break;
        if (imuPointerCur >= queueLength) break;                          // the reference would write past the arrays
        if (imuPointerCur == 0) {
            imuRotX[0] = 0; imuRotY[0] = 0; imuRotZ[0] = 0;
            imuTime[0] = currentImuTime;
            ++imuPointerCur;
            continue;
        }
        const double timeDiff = currentImuTime - imuTime[imuPointerCur - 1];
        imuRotX[imuPointerCur] = imuRotX[imuPointerCur - 1] + thisImuMsg.angular_velocity[0] * timeDiff;
        imuRotY[imuPointerCur] = imuRotY[imuPointerCur - 1] + thisImuMsg.angular_velocity[1] * timeDiff;
        imuRotZ[imuPointerCur] = imuRotZ[imuPointerCur - 1] + thisImuMsg.angular_velocity[2] * timeDiff;
        imuTime[imuPointerCur] = currentImuTime;
        ++imuPointerCur;
    }
    --imuPointerCur;
    if (imuPointerCur <= 0) return;
    cloudInfo.imuAvailable = 1;
}

void ImageProjection::odomDeskewInfo() {
    cloudInfo.odomAvailable = 0;
    const float sync_diff_time = (imuRate >= 300) ? 0.01f : 0.20f;
    while (!odomQueue.empty()) {
        if (odomQueue.front().stamp < timeScanCur - sync_diff_time) odomQueue.pop_front();
        else break;
    }
    if (odomQueue.empty()) return;
    if (odomQueue.front().stamp > timeScanCur) return;
    OdomMsg startOdomMsg = odomQueue.front();
    for (int i = 0; i < (int)odomQueue.size(); ++i) {
        startOdomMsg = odomQueue[i];
        if (startOdomMsg.stamp < timeScanCur) continue;
        else break;
    }
    double roll, pitch, yaw;
    rpyOf(startOdomMsg.orientation, roll, pitch, yaw);
    cloudInfo.initialGuessX = (float)startOdomMsg.position[0];
    cloudInfo.initialGuessY = (float)startOdomMsg.position[1];
    cloudInfo.initialGuessZ = (float)startOdomMsg.position[2];
    cloudInfo.initialGuessRoll = (float)roll; cloudInfo.initialGuessPitch = (float)pitch; cloudInfo.initialGuessYaw = (float)yaw;
    cloudInfo.odomAvailable = 1;
    odomDeskewFlag = false;
    if (odomQueue.back().stamp < timeScanEnd) return;
    OdomMsg endOdomMsg = odomQueue.front();
    for (int i = 0; i < (int)odomQueue.size(); ++i) {
        endOdomMsg = odomQueue[i];
        if (endOdomMsg.stamp < timeScanEnd) continue;
        else break;
    }
    if ((int)std::round(startOdomMsg.covariance0) != (int)std::round(endOdomMsg.covariance0)) return;
    const Affine3f transBegin = getTransformation((float)startOdomMsg.position[0], (float)startOdomMsg.position[1], (float)startOdomMsg.position[2],
                                                  (float)roll, (float)pitch, (float)yaw);
    rpyOf(endOdomMsg.orientation, roll, pitch, yaw);
    const Affine3f transEnd = getTransformation((float)endOdomMsg.position[0], (float)endOdomMsg.position[1], (float)endOdomMsg.position[2],
                                                (float)roll, (float)pitch, (float)yaw);
    const Affine3f transBt = transBegin.inverse() * transEnd;
    float rollIncre, pitchIncre, yawIncre;
    getTranslationAndEulerAngles(transBt, odomIncreX, odomIncreY, odomIncreZ, rollIncre, pitchIncre, yawIncre);
    odomDeskewFlag = true;            // consumed by findPosition only, whose body is commented out upstream (:599-613)
}

bool ImageProjection::projectPointCloud() {
    liloc_deskew_params p{};
    p.time_scan_cur = timeScanCur;
    p.imu_time = imuTime.data(); p.imu_rot_x = imuRotX.data(); p.imu_rot_y = imuRotY.data(); p.imu_rot_z = imuRotZ.data();
    p.imu_pointer_cur = imuPointerCur;
    p.deskew_flag = deskewFlag;
    p.imu_available = cloudInfo.imuAvailable;
    p.lidar_min_range = lidarMinRange; p.lidar_max_range = lidarMaxRange;
    p.n_scan = N_SCAN; p.downsample_rate = downsampleRate; p.point_filter_num = point_filter_num;
    // keep what this sweep's device call receives (diagnostics; the CPU tests of the host logic read it)
    const int m = imuPointerCur >= 0 ? imuPointerCur + 1 : 0;
    last_imu_pointer_ = imuPointerCur;
    last_imu_.resize((size_t)4 * m);
    for (int k = 0; k < m; ++k) { last_imu_[k] = imuTime[k]; last_imu_[m + k] = imuRotX[k]; last_imu_[2 * m + k] = imuRotY[k]; last_imu_[3 * m + k] = imuRotZ[k]; }
    last_pts_ = current_.pts; last_rt_ = current_.rt;
    last_size_ = -1;
    if (!ctx_) return true;                                                   // host logic only
    const int rc = liloc_deskew(ctx_, current_.pts.data(), current_.rt.data(), (int)current_.pts.size(), &p, nullptr);
    if (rc < 0) { error_ = liloc_last_error(ctx_); return false; }
    return true;
}

void ImageProjection::publishClouds(CloudInfo* out) {
    cloudInfo.stamp = current_.stamp;                                       // cloudHeader.stamp
    cloudInfo.cloud_deskewed = nullptr;
    cloudInfo.cloud_size = 0;
    cloudInfo.cloud_on_device = LILOC_SCAN_DESKEWED;                        // the cloud stays in the device context
    if (out) *out = cloudInfo;
}

}  // namespace liloc_host
