// oracle/affine_eigen.h -- Eigen::Affine3f / pcl::getTransformation algebra as the reference's compiled code evaluates it
// (TEST INFRASTRUCTURE, shared by deskew_oracle.cpp and host_oracle.cpp; the host class mirrors the same order in
// liloc_b200/host/affine.hpp, independently written):
//   pcl::getTransformation              SURVEY A.3 (sin/cos in double rounded to float: the repo-wide canonical choice)
//   Eigen::Affine3f::inverse()          Eigen 3.3 Transform.h: linear().inverse() (3x3 cofactor formula, Inverse.h
//                                       compute_inverse<.,.,3>), translation = (-inv) * t
//   Eigen::Affine3f * Eigen::Affine3f   Transform.h transform_transform_product_impl: linear = L1*L2,
//                                       translation = L1*t2 + t1; 3-term dot products reduce as a0 + (a1 + a2)
//                                       (Redux.h redux_novec_unroller splits [0,1) | [1,3))
#pragma once
#include <cmath>

namespace orc_affine {

struct Aff { float l[3][3]; float t[3]; };

// Eigen's reduction order for a fixed-size 3-vector sum
inline float sum3(float a0, float a1, float a2) { return a0 + (a1 + a2); }

inline Aff get_transformation(float x, float y, float z, float roll, float pitch, float yaw) {
    const float A = (float)std::cos((double)yaw), B = (float)std::sin((double)yaw);
    const float C = (float)std::cos((double)pitch), D = (float)std::sin((double)pitch);
    const float E = (float)std::cos((double)roll), F = (float)std::sin((double)roll);
    const float DE = D * E, DF = D * F;
    Aff t;
    t.l[0][0] = A * C; t.l[0][1] = A * DF - B * E; t.l[0][2] = B * F + A * DE; t.t[0] = x;
    t.l[1][0] = B * C; t.l[1][1] = A * E + B * DF; t.l[1][2] = B * DE - A * F; t.t[1] = y;
    t.l[2][0] = -D;    t.l[2][1] = C * F;          t.l[2][2] = C * E;          t.t[2] = z;
    return t;
}

// cofactor_3x3<i,j> of Eigen/src/LU/InverseImpl.h
inline float cofactor(const float m[3][3], int i, int j) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m[i1][j1] * m[i2][j2] - m[i1][j2] * m[i2][j1];
}

inline Aff affine_inverse(const Aff& a) {
    float c0[3];
    for (int k = 0; k < 3; ++k) c0[k] = cofactor(a.l, k, 0);
    const float det = sum3(c0[0] * a.l[0][0], c0[1] * a.l[1][0], c0[2] * a.l[2][0]);
    const float invdet = 1.0f / det;
    Aff r;
    for (int k = 0; k < 3; ++k) r.l[0][k] = c0[k] * invdet;
    for (int k = 0; k < 3; ++k) r.l[1][k] = cofactor(a.l, k, 1) * invdet;
    for (int k = 0; k < 3; ++k) r.l[2][k] = cofactor(a.l, k, 2) * invdet;
    for (int i = 0; i < 3; ++i) r.t[i] = sum3((-r.l[i][0]) * a.t[0], (-r.l[i][1]) * a.t[1], (-r.l[i][2]) * a.t[2]);
    return r;
}

inline Aff affine_mul(const Aff& a, const Aff& b) {
    Aff r;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) r.l[i][j] = sum3(a.l[i][0] * b.l[0][j], a.l[i][1] * b.l[1][j], a.l[i][2] * b.l[2][j]);
        r.t[i] = sum3(a.l[i][0] * b.t[0], a.l[i][1] * b.t[1], a.l[i][2] * b.t[2]) + a.t[i];
    }
    return r;
}

// pcl::getTranslationAndEulerAngles (float)
inline void get_translation_and_euler(const Aff& t, float* x, float* y, float* z, float* roll, float* pitch, float* yaw) {
    *x = t.t[0]; *y = t.t[1]; *z = t.t[2];
    *roll = std::atan2(t.l[2][1], t.l[2][2]);
    *pitch = std::asin(-t.l[2][0]);
    *yaw = std::atan2(t.l[1][0], t.l[0][0]);
}

}  // namespace orc_affine
