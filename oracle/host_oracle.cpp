// oracle/host_oracle.cpp -- CPU restatement of the host-side pose algebra around the registration (TEST INFRASTRUCTURE).
//
// PARITY UNPINNED (the reference ships no tests; Eigen / PCL are absent here).  Restated from src/mapOptmization.cpp:
//   updateInitialGuess, odometry-increment branch   :865-883   transIncre = lastImuPreTransformation.inverse() * transBack,
//                                                              transFinal = trans2Affine3f(transformTobeMapped) * transIncre
//   saveFrame                                       :1245-1264 transBetween = transStart.inverse() * transFinal, thresholds
// in float32 with Eigen's evaluation order (affine_eigen.h).  oracle/ref_pipeline.py drives the CPU lio pipeline with these;
// the product's host class (liloc_b200/host/map_optimization.cpp) is compared against it differentially
// (tests/test_host_logic.py, bench.py parity_checked_frames).
#include "affine_eigen.h"

using namespace orc_affine;

extern "C" {

// pose6 / odom6 = [roll, pitch, yaw, x, y, z].  last_odom6 = the previous frame's initialGuess* (lastImuPreTransformation).
void orc_update_initial_guess(const float* pose6, const float* last_odom6, const float* odom6, float* out6) {
    const Aff transBack = get_transformation(odom6[3], odom6[4], odom6[5], odom6[0], odom6[1], odom6[2]);
    const Aff last = get_transformation(last_odom6[3], last_odom6[4], last_odom6[5], last_odom6[0], last_odom6[1], last_odom6[2]);
    const Aff transIncre = affine_mul(affine_inverse(last), transBack);
    const Aff transTobe = get_transformation(pose6[3], pose6[4], pose6[5], pose6[0], pose6[1], pose6[2]);
    const Aff transFinal = affine_mul(transTobe, transIncre);
    get_translation_and_euler(transFinal, &out6[3], &out6[4], &out6[5], &out6[0], &out6[1], &out6[2]);
}

// saveFrame() given the last key pose: 1 = the frame becomes a keyframe
int orc_save_frame(const float* last_key6, const float* pose6, float angle_threshold, float dist_threshold) {
    const Aff transStart = get_transformation(last_key6[3], last_key6[4], last_key6[5], last_key6[0], last_key6[1], last_key6[2]);
    const Aff transFinal = get_transformation(pose6[3], pose6[4], pose6[5], pose6[0], pose6[1], pose6[2]);
    const Aff transBetween = affine_mul(affine_inverse(transStart), transFinal);
    float x, y, z, roll, pitch, yaw;
    get_translation_and_euler(transBetween, &x, &y, &z, &roll, &pitch, &yaw);
    if (std::abs(roll) < angle_threshold && std::abs(pitch) < angle_threshold && std::abs(yaw) < angle_threshold &&
        std::sqrt(x * x + y * y + z * z) < dist_threshold)
        return 0;
    return 1;
}

}  // extern "C"
