// oracle/deskew_oracle.cpp -- CPU restatement of imageProjection's de-skew and point filters.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as liloc_oracle.cpp: only tests/,
// __graft_entry__.smoke() and bench.py's CPU legs may load it).
//
// PARITY UNPINNED: the reference ships no tests or fixtures for this step; ROS / PCL / Eigen are absent here.
// Restated from src/imageProjection.cpp:
//   imuDeskewInfo   :441-496   gyro integration into imuTime / imuRotX/Y/Z (double)
//   findRotation    :573-597   first IMU sample later than the point, linear interpolation in double -> float
//   findPosition    :599-613   always zero (the positional de-skew is commented out upstream)
//   deskewPoint     :615-643   transBt = transStartInverse * transFinal, point = transBt * point
//   projectPointCloud :645-673 range / ring / ring-stride / index-stride filters, push_back order
// and from the third-party pieces it delegates to:
//   pcl::getTransformation              SURVEY A.3 (sin/cos in double rounded to float: the repo-wide canonical choice)
//   Eigen::Affine3f::inverse()          Eigen 3.3 Transform.h: linear().inverse() (3x3 cofactor formula, Inverse.h
//                                       compute_inverse<.,.,3>), translation = (-inv) * t
//   Eigen::Affine3f * Eigen::Affine3f   Transform.h transform_transform_product_impl: linear = L1*L2,
//                                       translation = L1*t2 + t1; 3-term dot products reduce as a0 + (a1 + a2)
//                                       (Redux.h redux_novec_unroller splits [0,1) | [1,3))
// float32 IEEE, no FMA contraction (-ffp-contract=off).
#include <cmath>
#include <cstdint>
#include <vector>

#include "affine_eigen.h"

namespace {

struct Pt { float x, y, z, intensity; };
struct RingTime { int ring; float time; };

using namespace orc_affine;

struct Imu { const double* time; const double* rx; const double* ry; const double* rz; int last; };

// findRotation :573-597
void find_rotation(const Imu& u, double pointTime, float* rx, float* ry, float* rz) {
    int front = 0;
    while (front < u.last) {
        if (pointTime < u.time[front]) break;
        ++front;
    }
    if (pointTime > u.time[front] || front == 0) {
        *rx = (float)u.rx[front]; *ry = (float)u.ry[front]; *rz = (float)u.rz[front];
    } else {
        const int back = front - 1;
        const double rf = (pointTime - u.time[back]) / (u.time[front] - u.time[back]);
        const double rb = (u.time[front] - pointTime) / (u.time[front] - u.time[back]);
        *rx = (float)(u.rx[front] * rf + u.rx[back] * rb);
        *ry = (float)(u.ry[front] * rf + u.ry[back] * rb);
        *rz = (float)(u.rz[front] * rf + u.rz[back] * rb);
    }
}

}  // namespace

extern "C" {

struct orc_deskew_params {
    double time_scan_cur;
    const double* imu_time; const double* imu_rot_x; const double* imu_rot_y; const double* imu_rot_z;
    int imu_pointer_cur;          // index of the last valid entry (after the `--imuPointerCur` of :490)
    int deskew_flag;              // -1: no time channel
    int imu_available;
    float lidar_min_range, lidar_max_range;
    int n_scan, downsample_rate, point_filter_num;
};

// imuDeskewInfo :441-496 for an already-trimmed queue (stamps + angular velocity): fills the four arrays (capacity cap),
// returns imuPointerCur after the final decrement (<= 0: imuAvailable stays false).
int orc_imu_integrate(const double* stamp, const double* gyro_xyz, int n, double time_scan_end, int cap,
                      double* imu_time, double* rot_x, double* rot_y, double* rot_z) {
    int cur = 0;
    for (int i = 0; i < n && cur < cap; ++i) {
        const double t = stamp[i];
        if (t > time_scan_end + 0.01) break;
        if (cur == 0) { rot_x[0] = rot_y[0] = rot_z[0] = 0; imu_time[0] = t; ++cur; continue; }
        const double dt = t - imu_time[cur - 1];
        rot_x[cur] = rot_x[cur - 1] + gyro_xyz[3 * i + 0] * dt;
        rot_y[cur] = rot_y[cur - 1] + gyro_xyz[3 * i + 1] * dt;
        rot_z[cur] = rot_z[cur - 1] + gyro_xyz[3 * i + 2] * dt;
        imu_time[cur] = t;
        ++cur;
    }
    return cur - 1;
}

// projectPointCloud :645-673 (+ deskewPoint :615-643).  Returns the size of fullCloud.
int orc_deskew(const Pt* pts, const RingTime* rt, int n, const orc_deskew_params* p, Pt* out) {
    const Imu u{p->imu_time, p->imu_rot_x, p->imu_rot_y, p->imu_rot_z, p->imu_pointer_cur};
    bool firstPointFlag = true;
    Aff transStartInverse{};
    int m = 0;
    for (int i = 0; i < n; ++i) {
        const Pt pt = pts[i];
        const float range = std::sqrt(pt.x * pt.x + pt.y * pt.y + pt.z * pt.z);
        if (range < p->lidar_min_range || range > p->lidar_max_range) continue;
        const int row = rt[i].ring;
        if (row < 0 || row >= p->n_scan) continue;
        if (row % p->downsample_rate != 0) continue;
        if (i % p->point_filter_num != 0) continue;
        if (p->deskew_flag == -1 || !p->imu_available) { out[m++] = pt; continue; }
        const double pointTime = p->time_scan_cur + (double)rt[i].time;
        float rx, ry, rz;
        find_rotation(u, pointTime, &rx, &ry, &rz);
        const float px = 0.f, py = 0.f, pz = 0.f;                               // findPosition :599-613
        if (firstPointFlag) {
            transStartInverse = affine_inverse(get_transformation(px, py, pz, rx, ry, rz));
            firstPointFlag = false;
        }
        const Aff transFinal = get_transformation(px, py, pz, rx, ry, rz);
        const Aff bt = affine_mul(transStartInverse, transFinal);
        Pt o;
        o.x = bt.l[0][0] * pt.x + bt.l[0][1] * pt.y + bt.l[0][2] * pt.z + bt.t[0];
        o.y = bt.l[1][0] * pt.x + bt.l[1][1] * pt.y + bt.l[1][2] * pt.z + bt.t[1];
        o.z = bt.l[2][0] * pt.x + bt.l[2][1] * pt.y + bt.l[2][2] * pt.z + bt.t[2];
        o.intensity = pt.intensity;
        out[m++] = o;
    }
    return m;
}

// the no-ring / no-time channel clouds (:279-326): relative time from the azimuth of every point (0.1 s sweep)
void orc_orientation_times(const Pt* pts, int n, float* time_out) {
    if (n <= 0) return;
    bool halfPassed = false;
    const double startOrientation = -std::atan2(pts[0].y, pts[0].x);
    double endOrientation = -std::atan2(pts[n - 1].y, pts[n - 1].x) + 2 * M_PI;
    if (endOrientation - startOrientation > 3 * M_PI) endOrientation -= 2 * M_PI;
    else if (endOrientation - startOrientation < M_PI) endOrientation += 2 * M_PI;
    const double orientationDiff = endOrientation - startOrientation;
    for (int i = 0; i < n; ++i) {
        float ori = -std::atan2(pts[i].y, pts[i].x);
        if (!halfPassed) {
            if (ori < startOrientation - M_PI / 2) ori += 2 * M_PI;
            else if (ori > startOrientation + M_PI * 3 / 2) ori -= 2 * M_PI;
            if (ori - startOrientation > M_PI) halfPassed = true;
        } else {
            ori += 2 * M_PI;
            if (ori < endOrientation - M_PI * 3 / 2) ori += 2 * M_PI;
            else if (ori > endOrientation + M_PI / 2) ori -= 2 * M_PI;
        }
        const float relTime = (ori - startOrientation) / orientationDiff;
        time_out[i] = 0.1 * relTime;
    }
}

}  // extern "C"
