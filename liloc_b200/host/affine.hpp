// liloc_b200/host/affine.hpp -- the few Eigen::Affine3f / PCL / tf helpers the host class needs
// (Eigen, PCL and tf are not available here).  Conventions: SURVEY.md A.3.
#pragma once
#include <cmath>

namespace liloc_host {

struct Affine3f {
    float m[3][4];      // top 3x4 of the homogeneous matrix, row-major

    static Affine3f Identity() { Affine3f a{}; a.m[0][0] = a.m[1][1] = a.m[2][2] = 1.f; return a; }

    // Eigen evaluates a fixed-size 3-term dot product as a0 + (a1 + a2) (Redux.h, redux_novec_unroller splits [0,1) | [1,3))
    static float sum3(float a0, float a1, float a2) { return a0 + (a1 + a2); }

    // Eigen::Transform * Transform (transform_transform_product_impl): linear = L1 L2, translation = L1 t2 + t1
    Affine3f operator*(const Affine3f& b) const {
        Affine3f r{};
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) r.m[i][j] = sum3(m[i][0] * b.m[0][j], m[i][1] * b.m[1][j], m[i][2] * b.m[2][j]);
            r.m[i][3] = sum3(m[i][0] * b.m[0][3], m[i][1] * b.m[1][3], m[i][2] * b.m[2][3]) + m[i][3];
        }
        return r;
    }

    // Eigen::Transform<float,3,Affine>::inverse(): the 3x3 cofactor inverse of the linear part (InverseImpl.h compute_inverse
    // <., ., 3>: cofactors of column 0 give the determinant), translation = (-inverse) * t
    float cofactor(int i, int j) const {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        return m[i1][j1] * m[i2][j2] - m[i1][j2] * m[i2][j1];
    }
    Affine3f inverse() const {
        const float c0[3] = {cofactor(0, 0), cofactor(1, 0), cofactor(2, 0)};
        const float det = sum3(c0[0] * m[0][0], c0[1] * m[1][0], c0[2] * m[2][0]);
        const float invdet = 1.0f / det;
        Affine3f r{};
        for (int k = 0; k < 3; ++k) r.m[0][k] = c0[k] * invdet;
        for (int k = 0; k < 3; ++k) r.m[1][k] = cofactor(k, 1) * invdet;
        for (int k = 0; k < 3; ++k) r.m[2][k] = cofactor(k, 2) * invdet;
        for (int i = 0; i < 3; ++i) r.m[i][3] = sum3((-r.m[i][0]) * m[0][3], (-r.m[i][1]) * m[1][3], (-r.m[i][2]) * m[2][3]);
        return r;
    }
};

// pcl::getTransformation(x, y, z, roll, pitch, yaw) -- same evaluation as the device (common.cuh pose_to_T):
// sin/cos in double, rounded to float.
inline Affine3f getTransformation(float x, float y, float z, float roll, float pitch, float yaw) {
    const float A = (float)std::cos((double)yaw), B = (float)std::sin((double)yaw);
    const float C = (float)std::cos((double)pitch), D = (float)std::sin((double)pitch);
    const float E = (float)std::cos((double)roll), F = (float)std::sin((double)roll);
    const float DE = D * E, DF = D * F;
    Affine3f t{};
    t.m[0][0] = A * C; t.m[0][1] = A * DF - B * E; t.m[0][2] = B * F + A * DE; t.m[0][3] = x;
    t.m[1][0] = B * C; t.m[1][1] = A * E + B * DF; t.m[1][2] = B * DE - A * F; t.m[1][3] = y;
    t.m[2][0] = -D;    t.m[2][1] = C * F;          t.m[2][2] = C * E;          t.m[2][3] = z;
    return t;
}

// pcl::getTranslationAndEulerAngles
inline void getTranslationAndEulerAngles(const Affine3f& t, float& x, float& y, float& z, float& roll, float& pitch, float& yaw) {
    x = t.m[0][3]; y = t.m[1][3]; z = t.m[2][3];
    roll = std::atan2(t.m[2][1], t.m[2][2]);
    pitch = std::asin(-t.m[2][0]);
    yaw = std::atan2(t.m[1][0], t.m[0][0]);
}

// eulerToRotation (include/math_tools.h:229-237): Rz(yaw) * Ry(pitch) * Rx(roll) as the product of three float rotation
// matrices (Eigen::AngleAxisf::toRotationMatrix: cosf / sinf of each angle), with the translation in the last column:
// the init_guess of the NDT call sites (src/mapOptmization.cpp:276-280, anchorOptimize.hpp:397-401).  Row-major 4x4.
inline void eulerToMatrix4(float roll, float pitch, float yaw, float x, float y, float z, float* T16) {
    const float cr = std::cos(roll), sr = std::sin(roll), cp = std::cos(pitch), sp = std::sin(pitch), cy = std::cos(yaw), sy = std::sin(yaw);
    const float Rz[3][3] = {{cy, -sy, 0}, {sy, cy, 0}, {0, 0, 1}}, Ry[3][3] = {{cp, 0, sp}, {0, 1, 0}, {-sp, 0, cp}}, Rx[3][3] = {{1, 0, 0}, {0, cr, -sr}, {0, sr, cr}};
    float M[3][3], R[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M[i][j] = Rz[i][0] * Ry[0][j] + Rz[i][1] * Ry[1][j] + Rz[i][2] * Ry[2][j];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R[i][j] = M[i][0] * Rx[0][j] + M[i][1] * Rx[1][j] + M[i][2] * Rx[2][j];
    const float t[3] = {x, y, z};
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T16[4 * i + j] = R[i][j]; T16[4 * i + 3] = t[i]; }
    T16[12] = T16[13] = T16[14] = 0.f; T16[15] = 1.f;
}

// RotMtoEuler<float> (include/math_tools.h:92-111) on the rotation block of a row-major 4x4, plus the translation.
inline void matrix4ToEuler(const float* T16, float& roll, float& pitch, float& yaw, float& x, float& y, float& z) {
    const float r00 = T16[0], r10 = T16[4], r20 = T16[8], r21 = T16[9], r22 = T16[10], r11 = T16[5], r12 = T16[6];
    const float sy = std::sqrt(r00 * r00 + r10 * r10);
    if (!(sy < 1e-6)) { roll = std::atan2(r21, r22); pitch = std::atan2(-r20, sy); yaw = std::atan2(r10, r00); }
    else { roll = std::atan2(-r12, r11); pitch = std::atan2(-r20, sy); yaw = 0.f; }
    x = T16[3]; y = T16[7]; z = T16[11];
}

// tf::Quaternion::setRPY / slerp and tf::Matrix3x3::getRPY, double precision, as used by transformUpdate
struct Quat { double x, y, z, w; };
inline Quat quatFromRPY(double roll, double pitch, double yaw) {
    const double hr = roll * 0.5, hp = pitch * 0.5, hy = yaw * 0.5;
    const double cr = std::cos(hr), sr = std::sin(hr), cp = std::cos(hp), sp = std::sin(hp), cy = std::cos(hy), sy = std::sin(hy);
    return Quat{sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy};
}
inline Quat slerp(const Quat& a, const Quat& b, double t) {
    const double dot = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    const double len2 = std::sqrt((a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w) * (b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w));
    const double theta = std::acos(std::fmin(1.0, std::fmax(-1.0, dot / len2)));
    if (theta == 0.0) return a;
    const double d = 1.0 / std::sin(theta), s0 = std::sin((1.0 - t) * theta), s1 = std::sin(t * theta);
    if (dot < 0) return Quat{(a.x * s0 - b.x * s1) * d, (a.y * s0 - b.y * s1) * d, (a.z * s0 - b.z * s1) * d, (a.w * s0 - b.w * s1) * d};
    return Quat{(a.x * s0 + b.x * s1) * d, (a.y * s0 + b.y * s1) * d, (a.z * s0 + b.z * s1) * d, (a.w * s0 + b.w * s1) * d};
}
inline void quatToRPY(const Quat& q, double& roll, double& pitch, double& yaw) {
    const double n = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w, s = 2.0 / n;
    const double r00 = 1 - s * (q.y * q.y + q.z * q.z), r10 = s * (q.x * q.y + q.w * q.z);
    const double r20 = s * (q.x * q.z - q.w * q.y), r21 = s * (q.y * q.z + q.w * q.x), r22 = 1 - s * (q.x * q.x + q.y * q.y);
    pitch = std::asin(std::fmin(1.0, std::fmax(-1.0, -r20)));
    roll = std::atan2(r21, r22);
    yaw = std::atan2(r10, r00);
}

}  // namespace liloc_host
