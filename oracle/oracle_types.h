// oracle/oracle_types.h -- plain-C structs of the CPU oracle (test infrastructure only; see
// liloc_oracle.cpp header).  Mirrored field-for-field by ctypes in tests/oracle_lib.py.
#ifndef LILOC_ORACLE_TYPES_H
#define LILOC_ORACLE_TYPES_H

#ifdef __cplusplus
extern "C" {
#endif

// {x, y, z, intensity}: the 16-byte device layout of pcl::PointXYZI (include/utility.h:83)
typedef struct { float x, y, z, w; } orc_point;

// members isDegenerate / matP of class mapOptimization (src/mapOptmization.cpp:90-91)
typedef struct { int is_degenerate; float matP[36]; } orc_lm_state;

typedef struct {
    int max_iters;     // 20  (src/mapOptmization.cpp:1190)
    int stride;        // 2   (:985)
    int min_sel;       // 50  (:1080)
    int min_points;    // 30  (:1187, gate is Q_all > min_points)
    int knn;           // 0 brute force, 1 kd-tree
    int threads;       // numberOfCores (:984)
} orc_s2m_opts;

typedef struct {
    int iters;         // loop iterations executed
    int converged;     // LMOptimization returned true
    int n_sel;         // correspondences in the last executed iteration
    int is_degenerate;
    int too_few;       // last iteration had < min_sel correspondences
    int skipped;       // Q_all <= min_points
} orc_s2m_result;

#ifdef __cplusplus
}
#endif
#endif
