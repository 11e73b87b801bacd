"""ctypes binding of the CPU parity oracle (oracle/libliloc_oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "oracle", "libliloc_oracle.so")


class LmState(C.Structure):
    _fields_ = [("is_degenerate", C.c_int), ("matP", C.c_float * 36)]


class S2mOpts(C.Structure):
    _fields_ = [("max_iters", C.c_int), ("stride", C.c_int), ("min_sel", C.c_int), ("min_points", C.c_int),
                ("knn", C.c_int), ("threads", C.c_int)]

    def __init__(self, max_iters=20, stride=2, min_sel=50, min_points=30, knn=1, threads=1):
        super().__init__(max_iters, stride, min_sel, min_points, knn, threads)


class S2mResult(C.Structure):
    _fields_ = [("iters", C.c_int), ("converged", C.c_int), ("n_sel", C.c_int), ("is_degenerate", C.c_int),
                ("too_few", C.c_int), ("skipped", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} missing: run `make -C oracle` or `python -m liloc_b200._build`")
    L = C.CDLL(LIB_PATH)
    P = C.c_void_p
    L.orc_num_threads.restype = C.c_int
    L.orc_voxelgrid.restype = C.c_int
    L.orc_voxelgrid.argtypes = [P, C.c_int, C.c_float, P]
    L.orc_pose_to_T.argtypes = [P, P]
    L.orc_transform_cloud.argtypes = [P, C.c_int, P, P, C.c_int]
    L.orc_apply_T.argtypes = [P, C.c_int, P, P]
    L.orc_knn5_brute.argtypes = [P, C.c_int, P, C.c_int, P, P, C.c_int]
    L.orc_knn5_kdtree.argtypes = [P, C.c_int, P, C.c_int, P, P, C.c_int]
    L.orc_plane_lsq.argtypes = [P, P]
    L.orc_solve6.argtypes = [P, P, P]
    L.orc_eigen6.argtypes = [P, P, P]
    L.orc_degeneracy.restype = C.c_int
    L.orc_degeneracy.argtypes = [P, P]
    L.orc_jacobian_row.argtypes = [P, P, P, P, P]
    L.orc_surf_pass.restype = C.c_int
    L.orc_surf_pass.argtypes = [P, C.c_int, P, C.c_int, C.c_int, P, C.c_int, C.c_int, P, P, P, P, P]
    L.orc_lm_step.restype = C.c_int
    L.orc_lm_step.argtypes = [P, P, C.c_int, C.c_int, C.c_int, P, C.POINTER(LmState), P, P]
    L.orc_scan2map.restype = C.c_int
    L.orc_scan2map.argtypes = [P, C.c_int, P, C.c_int, P, C.POINTER(S2mOpts), C.POINTER(LmState), C.POINTER(S2mResult)]
    L.orc_localmap_build.restype = C.c_int
    L.orc_localmap_build.argtypes = [P, P, P, C.c_int, C.c_float, C.c_int, P, P]
    L.orc_frame.restype = C.c_int
    L.orc_frame.argtypes = [P, P, P, C.c_int, C.c_float, P, C.c_int, C.c_float, P, C.POINTER(S2mOpts),
                            C.POINTER(LmState), C.POINTER(S2mResult), P, P, P]
    L.orc_icp_fitness.restype = C.c_float
    L.orc_icp_fitness.argtypes = [P, C.c_int, P, P, C.c_int, P, C.c_int, C.c_int]
    L.orc_eigen3.argtypes = [P, P, P]
    L.orc_line_residual.restype = C.c_int
    L.orc_line_residual.argtypes = [P, P, P]
    L.orc_corner_pass.restype = C.c_int
    L.orc_corner_pass.argtypes = [P, C.c_int, P, C.c_int, C.c_int, P, C.c_int, C.c_int, P, P, P]
    L.orc_scan2map_features.restype = C.c_int
    L.orc_scan2map_features.argtypes = [P, C.c_int, P, C.c_int, P, C.c_int, P, C.c_int, C.c_int, C.c_int, P,
                                        C.POINTER(S2mOpts), C.POINTER(LmState), C.POINTER(S2mResult)]
    L.orc_imu_integrate.restype = C.c_int
    L.orc_imu_integrate.argtypes = [P, P, C.c_int, C.c_double, C.c_int, P, P, P, P]
    L.orc_deskew.restype = C.c_int
    L.orc_deskew.argtypes = [P, P, C.c_int, C.POINTER(DeskewParams), P]
    L.orc_orientation_times.argtypes = [P, C.c_int, P]
    L.orc_update_initial_guess.argtypes = [P, P, P, P]
    L.orc_save_frame.restype = C.c_int
    L.orc_save_frame.argtypes = [P, P, C.c_float, C.c_float]
    return L


class DeskewParams(C.Structure):
    _fields_ = [("time_scan_cur", C.c_double),
                ("imu_time", C.c_void_p), ("imu_rot_x", C.c_void_p), ("imu_rot_y", C.c_void_p), ("imu_rot_z", C.c_void_p),
                ("imu_pointer_cur", C.c_int), ("deskew_flag", C.c_int), ("imu_available", C.c_int),
                ("lidar_min_range", C.c_float), ("lidar_max_range", C.c_float),
                ("n_scan", C.c_int), ("downsample_rate", C.c_int), ("point_filter_num", C.c_int)]


RING_TIME = np.dtype([("ring", np.int32), ("time", np.float32)])

lib = _load()


def _c(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 4
    return a


def _p(a):
    return C.c_void_p(a.ctypes.data)


def num_threads() -> int:
    return lib.orc_num_threads()


def update_initial_guess(pose6, last_odom6, odom6) -> np.ndarray:
    """updateInitialGuess, odometry-increment branch (src/mapOptmization.cpp:865-883), float32 / Eigen order."""
    a, b, c = (np.ascontiguousarray(v, np.float32).reshape(6) for v in (pose6, last_odom6, odom6))
    out = np.zeros(6, np.float32)
    lib.orc_update_initial_guess(_p(a), _p(b), _p(c), _p(out))
    return out


def save_frame(last_key6, pose6, angle_threshold, dist_threshold) -> bool:
    """saveFrame (src/mapOptmization.cpp:1245-1264)."""
    a, b = (np.ascontiguousarray(v, np.float32).reshape(6) for v in (last_key6, pose6))
    return bool(lib.orc_save_frame(_p(a), _p(b), float(angle_threshold), float(dist_threshold)))


def pose_to_T(pose6) -> np.ndarray:
    p = np.ascontiguousarray(pose6, np.float32).reshape(6)
    T = np.empty(12, np.float32)
    lib.orc_pose_to_T(_p(p), _p(T))
    return T.reshape(3, 4)


def transform_cloud(pts, pose6, threads=1) -> np.ndarray:
    a = _c(pts)
    p = np.ascontiguousarray(pose6, np.float32).reshape(6)
    out = np.empty_like(a)
    lib.orc_transform_cloud(_p(a), len(a), _p(p), _p(out), threads)
    return out


def voxelgrid(pts, leaf) -> np.ndarray:
    a = _c(pts)
    out = np.empty_like(a)
    m = lib.orc_voxelgrid(_p(a), len(a), leaf, _p(out))
    return out[:m].copy()


def knn5(map_pts, queries, kdtree=False, threads=1):
    m, q = _c(map_pts), _c(queries)
    idx = np.empty((len(q), 5), np.int32)
    d2 = np.empty((len(q), 5), np.float32)
    f = lib.orc_knn5_kdtree if kdtree else lib.orc_knn5_brute
    f(_p(m), len(m), _p(q), len(q), _p(idx), _p(d2), threads)
    return idx, d2


def plane_lsq(pts5x3) -> np.ndarray:
    a = np.ascontiguousarray(pts5x3, np.float32).reshape(15)
    x = np.empty(3, np.float32)
    lib.orc_plane_lsq(_p(a), _p(x))
    return x


def solve6(A, b) -> np.ndarray:
    A = np.ascontiguousarray(A, np.float32).reshape(36)
    b = np.ascontiguousarray(b, np.float32).reshape(6)
    x = np.empty(6, np.float32)
    lib.orc_solve6(_p(A), _p(b), _p(x))
    return x


def eigen6(A):
    A = np.ascontiguousarray(A, np.float32).reshape(36)
    w = np.empty(6, np.float32)
    V = np.empty(36, np.float32)
    lib.orc_eigen6(_p(A), _p(w), _p(V))
    return w, V.reshape(6, 6)


def degeneracy(A):
    A = np.ascontiguousarray(A, np.float32).reshape(36)
    P = np.empty(36, np.float32)
    d = lib.orc_degeneracy(_p(A), _p(P))
    return bool(d), P.reshape(6, 6)


def jacobian_row(pose6, ori, coef):
    p = np.ascontiguousarray(pose6, np.float32).reshape(6)
    o = np.ascontiguousarray(ori, np.float32).reshape(4)
    c = np.ascontiguousarray(coef, np.float32).reshape(4)
    row = np.empty(6, np.float32)
    rhs = np.empty(1, np.float32)
    lib.orc_jacobian_row(_p(p), _p(o), _p(c), _p(row), _p(rhs))
    return row, float(rhs[0])


def surf_pass(map_pts, scan_ds, pose6, stride=2, kdtree=False, threads=1, want_knn=False):
    """Returns (coeff[Q_all,4], flag[Q_all], ori_sel, coef_sel, sel_index[, idx5, d2])."""
    m, s = _c(map_pts), _c(scan_ds)
    p = np.ascontiguousarray(pose6, np.float32).reshape(6)
    Q = len(s)
    ori = np.empty((Q, 4), np.float32)
    coef = np.empty((Q, 4), np.float32)
    sel = np.empty(Q, np.int32)
    idx5 = np.full((Q, 5), -1, np.int32) if want_knn else None
    d2 = np.full((Q, 5), np.float32(3.4e38), np.float32) if want_knn else None
    n = lib.orc_surf_pass(_p(m), len(m), _p(s), Q, stride, _p(p), 1 if kdtree else 0, threads, _p(ori), _p(coef), _p(sel),
                          _p(idx5) if want_knn else None, _p(d2) if want_knn else None)
    full = np.zeros((Q, 4), np.float32)
    flag = np.zeros(Q, np.uint8)
    full[sel[:n]] = coef[:n]
    flag[sel[:n]] = 1
    out = (full, flag, ori[:n].copy(), coef[:n].copy(), sel[:n].copy())
    return out + ((idx5, d2) if want_knn else ())


def lm_step(ori, coef, iter_count, pose6, state: LmState | None = None, min_sel=50):
    """LMOptimization (src/mapOptmization.cpp:1061-1181) on explicit correspondences.  Returns (rc, pose, AtA, AtB, state):
    rc -1 fewer than min_sel correspondences (pose untouched), 0 updated, 1 updated and converged."""
    o, c = _c(ori), _c(coef)
    p = np.array(pose6, np.float32).reshape(6).copy()
    st = state if state is not None else LmState()
    AtA, AtB = np.zeros(36, np.float32), np.zeros(6, np.float32)
    rc = lib.orc_lm_step(_p(o), _p(c), len(o), int(iter_count), int(min_sel), _p(p), C.byref(st), _p(AtA), _p(AtB))
    return rc, p, AtA.reshape(6, 6), AtB, st


def scan2map(map_pts, scan_ds, pose6, opts: S2mOpts | None = None, state: LmState | None = None):
    m, s = _c(map_pts), _c(scan_ds)
    p = np.array(pose6, np.float32).reshape(6).copy()
    o = opts or S2mOpts()
    st = state if state is not None else LmState()
    res = S2mResult()
    rc = lib.orc_scan2map(_p(m), len(m), _p(s), len(s), _p(p), C.byref(o), C.byref(st), C.byref(res))
    return p, res, rc, st


# ---- upstream LIO-SAM corner superset (not in the reference; oracle/liloc_oracle.cpp "corner path") ----
def eigen3(A):
    A = np.ascontiguousarray(A, np.float32).reshape(9)
    w, V = np.empty(3, np.float32), np.empty(9, np.float32)
    lib.orc_eigen3(_p(A), _p(w), _p(V))
    return w, V.reshape(3, 3)


def line_residual(pts5x3, sel3):
    a = np.ascontiguousarray(pts5x3, np.float32).reshape(15)
    q = np.ascontiguousarray(sel3, np.float32).reshape(3)
    c = np.zeros(4, np.float32)
    keep = lib.orc_line_residual(_p(a), _p(q), _p(c))
    return c, keep


def corner_pass(map_pts, scan_ds, pose6, stride=1, kdtree=False, threads=1):
    """Returns (coeff[Q_all,4], flag[Q_all]) of cornerOptimization at a fixed pose."""
    m, s = _c(map_pts), _c(scan_ds)
    p = np.ascontiguousarray(pose6, np.float32).reshape(6)
    Q = len(s)
    ori, coef, sel = np.empty((Q, 4), np.float32), np.empty((Q, 4), np.float32), np.empty(Q, np.int32)
    n = lib.orc_corner_pass(_p(m), len(m), _p(s), Q, stride, _p(p), 1 if kdtree else 0, threads, _p(ori), _p(coef), _p(sel))
    full, flag = np.zeros((Q, 4), np.float32), np.zeros(Q, np.uint8)
    full[sel[:n]] = coef[:n]
    flag[sel[:n]] = 1
    return full, flag


def upstream_opts(knn=1, threads=1) -> S2mOpts:
    """Upstream LIO-SAM loop shape: 30 iterations, surf stride 1, gate surf > 100 (corner > 10 is an argument)."""
    return S2mOpts(max_iters=30, stride=1, min_sel=50, min_points=100, knn=knn, threads=threads)


def scan2map_features(map_s, scan_s, map_c, scan_c, pose6, opts: S2mOpts | None = None, state: LmState | None = None,
                      min_corner=10, stride_corner=1):
    ms, ss, mc, sc = _c(map_s), _c(scan_s), _c(map_c), _c(scan_c)
    p = np.array(pose6, np.float32).reshape(6).copy()
    o = opts or upstream_opts()
    st = state if state is not None else LmState()
    res = S2mResult()
    rc = lib.orc_scan2map_features(_p(ms), len(ms), _p(ss), len(ss), _p(mc), len(mc), _p(sc), len(sc),
                                   min_corner, stride_corner, _p(p), C.byref(o), C.byref(st), C.byref(res))
    return p, res, rc, st


def icp_fitness(cloud1, pose6_1, cloud2, pose6_2, kdtree=True, threads=1) -> float:
    """getICPFitnessScore (anchorOptimize.hpp:465-481): mean squared 1-NN distance of cloud 1 (transformed) in cloud 2."""
    a, b = _c(cloud1), _c(cloud2)
    p1 = np.ascontiguousarray(pose6_1, np.float32).reshape(6)
    p2 = np.ascontiguousarray(pose6_2, np.float32).reshape(6)
    return float(lib.orc_icp_fitness(_p(a), len(a), _p(p1), _p(b), len(b), _p(p2), 1 if kdtree else 0, threads))


def imu_integrate(stamps, gyro, time_scan_end, cap=2000):
    """imuDeskewInfo (src/imageProjection.cpp:441-496) -> (imu_time (m,), imu_rot (m, 3)); m - 1 = imuPointerCur."""
    st = np.ascontiguousarray(stamps, np.float64)
    g = np.ascontiguousarray(gyro, np.float64).reshape(-1, 3)
    t = np.zeros(cap, np.float64); r = [np.zeros(cap, np.float64) for _ in range(3)]
    last = lib.orc_imu_integrate(_p(st), _p(g), len(st), float(time_scan_end), cap, _p(t), _p(r[0]), _p(r[1]), _p(r[2]))
    m = last + 1
    return t[:m].copy(), np.stack([a[:m] for a in r], axis=1)


def deskew(pts, ring, time, imu_time=None, imu_rot=None, time_scan_cur=0.0, deskew_flag=1, imu_available=True,
           min_range=1.0, max_range=1000.0, n_scan=16, downsample_rate=1, point_filter_num=3) -> np.ndarray:
    """projectPointCloud + deskewPoint (src/imageProjection.cpp:615-673) -> fullCloud."""
    p = _c(pts)
    rt = np.empty(len(p), RING_TIME)
    rt["ring"], rt["time"] = np.asarray(ring, np.int32), np.asarray(time, np.float32)
    dp = DeskewParams()
    dp.time_scan_cur = float(time_scan_cur)
    keep = []
    if imu_time is not None and len(imu_time):
        t = np.ascontiguousarray(imu_time, np.float64)
        r = [np.ascontiguousarray(np.asarray(imu_rot, np.float64)[:, k]) for k in range(3)]
        keep = [t, *r]
        dp.imu_time, dp.imu_rot_x, dp.imu_rot_y, dp.imu_rot_z = (a.ctypes.data for a in keep)
        dp.imu_pointer_cur = len(t) - 1
    else:
        dp.imu_pointer_cur = -1
        imu_available = False
    dp.deskew_flag, dp.imu_available = int(deskew_flag), int(bool(imu_available))
    dp.lidar_min_range, dp.lidar_max_range = float(min_range), float(max_range)
    dp.n_scan, dp.downsample_rate, dp.point_filter_num = int(n_scan), int(downsample_rate), int(point_filter_num)
    out = np.empty((max(len(p), 1), 4), np.float32)
    m = lib.orc_deskew(_p(p), _p(rt), len(p), C.byref(dp), _p(out))
    return out[:m].copy()


def orientation_times(pts) -> np.ndarray:
    """relative time of every point of a cloud without ring / time channels (src/imageProjection.cpp:279-326)"""
    p = _c(pts)
    out = np.zeros(len(p), np.float32)
    lib.orc_orientation_times(_p(p), len(p), _p(out))
    return out


def localmap_build(kf_clouds, poses6, leaf, threads=1):
    pts = np.ascontiguousarray(np.concatenate([_c(c) for c in kf_clouds], axis=0)) if kf_clouds else np.zeros((0, 4), np.float32)
    offs = np.zeros(len(kf_clouds) + 1, np.int64)
    offs[1:] = np.cumsum([len(c) for c in kf_clouds])
    poses = np.ascontiguousarray(poses6, np.float32).reshape(-1, 6)
    out = np.empty((max(len(pts), 1), 4), np.float32)
    M = lib.orc_localmap_build(_p(pts), _p(offs), _p(poses), len(kf_clouds), leaf, threads, _p(out), None)
    return out[:M].copy()


def frame(kf_pts, kf_offs, poses6, map_leaf, scan_raw, scan_leaf, pose6, opts: S2mOpts, state: LmState | None = None):
    """Whole reference frame on the CPU with timing. kf_pts: concatenated (N,4); kf_offs: int64 (K+1)."""
    pts = _c(kf_pts)
    offs = np.ascontiguousarray(kf_offs, np.int64)
    poses = np.ascontiguousarray(poses6, np.float32).reshape(-1, 6)
    s = _c(scan_raw)
    p = np.array(pose6, np.float32).reshape(6).copy()
    st = state if state is not None else LmState()
    res = S2mResult()
    M, Q = C.c_int(), C.c_int()
    ms = np.zeros(4, np.float64)
    rc = lib.orc_frame(_p(pts), _p(offs), _p(poses), len(offs) - 1, map_leaf, _p(s), len(s), scan_leaf, _p(p),
                       C.byref(opts), C.byref(st), C.byref(res), C.byref(M), C.byref(Q), _p(ms))
    return p, res, rc, M.value, Q.value, ms


# ---- NDT oracle (oracle/ndt_oracle.cpp) -----------------------------------------------------------------
class NdtOpts(C.Structure):
    _fields_ = [("step_size", C.c_double), ("outlier_ratio", C.c_double), ("trans_eps", C.c_double),
                ("resolution", C.c_float), ("max_iterations", C.c_int), ("search", C.c_int), ("threads", C.c_int)]

    def __init__(self, resolution=1.0, trans_eps=0.01, search=0, threads=1, step_size=0.1, outlier_ratio=0.55, max_iterations=35):
        super().__init__(step_size, outlier_ratio, float(np.float32(trans_eps)), resolution, max_iterations, search, threads)


class NdtResult(C.Structure):
    _fields_ = [("iterations", C.c_int), ("converged", C.c_int), ("sweeps", C.c_int), ("score", C.c_double)]


lib.orc_ndt_target_create.restype = C.c_void_p
lib.orc_ndt_target_create.argtypes = [C.c_void_p, C.c_int, C.c_float]
lib.orc_ndt_target_destroy.argtypes = [C.c_void_p]
lib.orc_ndt_target_leaves.restype = C.c_int
lib.orc_ndt_target_leaves.argtypes = [C.c_void_p]
lib.orc_ndt_target_get.restype = C.c_int
lib.orc_ndt_target_get.argtypes = [C.c_void_p] * 4 + [C.c_int]
lib.orc_ndt_target_keys.restype = C.c_int
lib.orc_ndt_target_keys.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
lib.orc_ndt_pose_to_matrix.argtypes = [C.c_void_p, C.c_void_p]
lib.orc_ndt_matrix_to_pose.argtypes = [C.c_void_p, C.c_void_p]
lib.orc_ndt_derivatives.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(NdtOpts), C.c_void_p]
lib.orc_ndt_newton.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
lib.orc_ndt_newton_pinv.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
lib.orc_ndt_derivatives_sequential.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(NdtOpts), C.c_void_p]
lib.orc_ndt_trial_value.restype = C.c_double
lib.orc_ndt_trial_value.argtypes = [C.c_double] * 9
lib.orc_ndt_align.restype = C.c_int
lib.orc_ndt_align.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(NdtOpts), C.c_void_p, C.POINTER(NdtResult)]


class NdtTarget:
    def __init__(self, pts, resolution):
        self._pts = _c(pts)
        self._h = lib.orc_ndt_target_create(_p(self._pts), len(self._pts), resolution)

    def __del__(self):
        if getattr(self, "_h", None):
            lib.orc_ndt_target_destroy(self._h)
            self._h = None

    def leaves(self):
        n = lib.orc_ndt_target_leaves(self._h)
        means, icovs, counts = np.zeros((n, 3)), np.zeros((n, 9), np.float32), np.zeros(n, np.int32)
        lib.orc_ndt_target_get(self._h, _p(means), _p(icovs), _p(counts), n)
        return means, icovs, counts

    def keys(self):
        n = lib.orc_ndt_target_leaves(self._h)
        k = np.zeros(n, np.int32)
        lib.orc_ndt_target_keys(self._h, _p(k), n)
        return k


def ndt_pose_to_matrix(p6):
    p = np.ascontiguousarray(p6, np.float64).reshape(6)
    T = np.zeros(12, np.float32)
    lib.orc_ndt_pose_to_matrix(_p(p), _p(T))
    return T.reshape(3, 4)


def ndt_matrix_to_pose(T12):
    T = np.ascontiguousarray(T12, np.float32).reshape(12)
    p = np.zeros(6)
    lib.orc_ndt_matrix_to_pose(_p(T), _p(p))
    return p


def ndt_derivatives(target: NdtTarget, src, p6, opts: NdtOpts):
    s = _c(src)
    p = np.ascontiguousarray(p6, np.float64).reshape(6)
    out = np.zeros(28)
    lib.orc_ndt_derivatives(target._h, _p(s), len(s), _p(p), C.byref(opts), _p(out))
    return out


def ndt_derivatives_sequential(target: NdtTarget, src, p6, opts: NdtOpts):
    """The same 28 sums in plain ascending point order (cross-check of the canonical tile order)."""
    s = _c(src)
    p = np.ascontiguousarray(p6, np.float64).reshape(6)
    out = np.empty(28, np.float64)
    lib.orc_ndt_derivatives_sequential(target._h, _p(s), len(s), _p(p), C.byref(opts), _p(out))
    return out


def ndt_align(target: NdtTarget, src, guess12, opts: NdtOpts):
    """guess12 / result: (3,4) float32 = top rows of the 4x4 Eigen::Matrix4f of registration->align(out, guess)."""
    s = _c(src)
    g = np.ascontiguousarray(guess12, np.float32).reshape(12)
    out = np.zeros(12, np.float32)
    res = NdtResult()
    lib.orc_ndt_align(target._h, _p(s), len(s), _p(g), C.byref(opts), _p(out), C.byref(res))
    return out.reshape(3, 4), res


def ndt_newton(H, g):
    """JacobiSVD(H).solve(-g) for a symmetric 6x6 H (upper triangle is read)."""
    H = np.asarray(H, np.float64); g = np.ascontiguousarray(g, np.float64)
    up = np.ascontiguousarray([H[i, j] for i in range(6) for j in range(i, 6)], np.float64)
    dp = np.zeros(6, np.float64)
    lib.orc_ndt_newton(_p(up), _p(g), _p(dp))
    return dp


def ndt_newton_pinv(H, g):
    """The pure pseudo-inverse solve (Jacobi eigen-decomposition), the fallback branch of ndt_newton."""
    H = np.asarray(H, np.float64); g = np.ascontiguousarray(g, np.float64)
    up = np.ascontiguousarray([H[i, j] for i in range(6) for j in range(i, 6)], np.float64)
    dp = np.zeros(6, np.float64)
    lib.orc_ndt_newton_pinv(_p(up), _p(g), _p(dp))
    return dp


def ndt_trial_value(a_l, f_l, g_l, a_u, f_u, g_u, a_t, f_t, g_t):
    return lib.orc_ndt_trial_value(a_l, f_l, g_l, a_u, f_u, g_u, a_t, f_t, g_t)
