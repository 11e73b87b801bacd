"""ctypes binding of libliloc_host.so: the ROS-free `mapOptimization` host class (liloc_b200/host/)
driven through its per-frame callback, exactly as the reference's ROS spinner drives it."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
lib_path = os.path.join(_HERE, "libliloc_host.so")
CONFIG_DIR = os.path.join(os.path.dirname(_HERE), "config")


class CloudInfo(C.Structure):
    """msg/cloud_info.msg fields read by mapOptimization."""
    _fields_ = [("stamp", C.c_double), ("imuAvailable", C.c_int), ("odomAvailable", C.c_int),
                ("imuRollInit", C.c_float), ("imuPitchInit", C.c_float), ("imuYawInit", C.c_float),
                ("initialGuessX", C.c_float), ("initialGuessY", C.c_float), ("initialGuessZ", C.c_float),
                ("initialGuessRoll", C.c_float), ("initialGuessPitch", C.c_float), ("initialGuessYaw", C.c_float),
                ("cloud_deskewed", C.c_void_p), ("cloud_size", C.c_int), ("cloud_on_device", C.c_int)]


class Report(C.Structure):
    _fields_ = [("processed", C.c_int), ("status", C.c_int), ("pose", C.c_float * 6), ("guess", C.c_float * 6),
                ("is_degenerate", C.c_int), ("iters", C.c_int), ("converged", C.c_int), ("n_sel", C.c_int),
                ("q_all", C.c_int), ("map_points", C.c_int), ("n_raw", C.c_int), ("n_selected_keyframes", C.c_int),
                ("keyframe_id", C.c_int), ("reg_ms", C.c_double),
                ("relo_initialized", C.c_int), ("relo_submap", C.c_int), ("relo_iterations", C.c_int), ("relo_converged", C.c_int),
                ("relo_pose", C.c_float * 6), ("relo_n_vertices", C.c_int), ("relo_vertex", C.c_int * 3), ("relo_fitness", C.c_float * 3),
                ("prior_added", C.c_int), ("prior_new_submap", C.c_int), ("relo_hypothesis", C.c_int), ("relo_score", C.c_float)]


def _load():
    if not os.path.exists(lib_path):
        raise ImportError(f"{lib_path} missing: run `python -m liloc_b200._build`")
    L = C.CDLL(lib_path)
    P = C.c_void_p
    L.liloc_host_create.restype = P
    L.liloc_host_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    L.liloc_host_destroy.argtypes = [P]
    L.liloc_host_error.restype = C.c_char_p
    L.liloc_host_error.argtypes = [P]
    L.liloc_host_ctx.restype = P
    L.liloc_host_ctx.argtypes = [P]
    L.liloc_host_handle.restype = C.c_int
    L.liloc_host_handle.argtypes = [P, C.POINTER(CloudInfo), C.POINTER(Report)]
    L.liloc_host_run.restype = C.c_int
    L.liloc_host_run.argtypes = [P, C.POINTER(CloudInfo), C.c_int, C.POINTER(Report)]
    L.liloc_host_reset.argtypes = [P]
    L.liloc_host_global_map.restype = C.c_int
    L.liloc_host_global_map.argtypes = [P, P, C.c_int]
    L.liloc_host_load_prior.restype = C.c_int
    L.liloc_host_load_prior.argtypes = [P, P, C.c_int, P, P]
    L.liloc_host_initialpose.argtypes = [P, C.c_float, C.c_float, C.c_float]
    L.liloc_host_set_relo_hypotheses.argtypes = [P, C.c_int, C.c_int, C.c_float]
    L.liloc_host_set_share.argtypes = [P, C.c_int]; L.liloc_host_set_share.restype = C.c_int
    L.liloc_host_comm_init.restype = C.c_int
    L.liloc_host_comm_init.argtypes = [P, C.c_int, C.c_int, P]
    L.liloc_host_hypothesis_pose.argtypes = [P, C.c_int, C.c_int, C.c_float, C.c_int, P]
    L.liloc_host_load_prior_dir.restype = C.c_int
    L.liloc_host_load_prior_dir.argtypes = [P, C.c_char_p]
    L.liloc_host_save_session.restype = C.c_int
    L.liloc_host_save_session.argtypes = [P, C.c_char_p]
    L.liloc_host_save_session_res.restype = C.c_int
    L.liloc_host_save_session_res.argtypes = [P, C.c_char_p, C.c_float]
    L.liloc_host_save_poses_pcd.restype = C.c_int
    L.liloc_host_save_poses_pcd.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.c_int]
    L.liloc_host_load_pcd.restype = C.c_int
    L.liloc_host_load_pcd.argtypes = [C.c_char_p, P, C.c_int, C.c_char_p, C.c_int]
    L.liloc_host_save_pcd.restype = C.c_int
    L.liloc_host_save_pcd.argtypes = [C.c_char_p, P, C.c_int]
    L.liloc_host_load_g2o.restype = C.c_int
    L.liloc_host_load_g2o.argtypes = [C.c_char_p, P, C.c_int]
    L.liloc_host_save_g2o.restype = C.c_int
    L.liloc_host_save_g2o.argtypes = [C.c_char_p, P, C.c_int]
    L.liloc_host_submap_info.restype = C.c_int
    L.liloc_host_submap_info.argtypes = [P, C.c_int, P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.liloc_host_set_s2m_opts.argtypes = [P, C.c_int, C.c_int, C.c_int, C.c_int]
    L.liloc_host_num_keyframes.restype = C.c_int
    L.liloc_host_num_keyframes.argtypes = [P]
    L.liloc_host_keyposes.restype = C.c_int
    L.liloc_host_keyposes.argtypes = [P, P, C.c_int]
    L.liloc_host_last_selection.restype = C.c_int
    L.liloc_host_last_selection.argtypes = [P, P, C.c_int]
    L.liloc_host_param.restype = C.c_int
    L.liloc_host_param.argtypes = [P, C.c_char_p, C.POINTER(C.c_double)]
    L.liloc_host_yaml_param.restype = C.c_int
    L.liloc_host_yaml_param.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_double)]
    L.liloc_host_debug_add_keypose.argtypes = [P, P, C.c_double]
    L.liloc_host_debug_extract_nearby.restype = C.c_int
    L.liloc_host_debug_extract_nearby.argtypes = [P, C.c_double, P, C.c_int]
    L.liloc_host_debug_nearby_plain.argtypes = [P, C.c_int]
    L.liloc_host_debug_set_pose.argtypes = [P, P]
    L.liloc_host_debug_get_pose.argtypes = [P, P]
    L.liloc_host_debug_update_initial_guess.argtypes = [P, C.POINTER(CloudInfo)]
    L.liloc_host_debug_euler_to_matrix.argtypes = [P, P]
    L.liloc_host_debug_save_frame.restype = C.c_int
    L.liloc_host_debug_save_frame.argtypes = [P]
    L.liloc_host_debug_transform_update.argtypes = [P, C.POINTER(CloudInfo)]
    L.liloc_host_debug_voxelgrid_small.restype = C.c_int
    L.liloc_host_debug_voxelgrid_small.argtypes = [P, C.c_int, C.c_float, P]
    L.liloc_host_ip_create.restype = P
    L.liloc_host_ip_create.argtypes = [C.c_char_p, P]
    L.liloc_host_ip_destroy.argtypes = [P]
    L.liloc_host_ip_error.restype = C.c_char_p
    L.liloc_host_ip_error.argtypes = [P]
    L.liloc_host_ip_imu.argtypes = [P, P]
    L.liloc_host_ip_odom.argtypes = [P, P]
    L.liloc_host_ip_cloud.restype = C.c_int
    L.liloc_host_ip_cloud.argtypes = [P, C.c_double, P, P, C.c_int, C.c_int, C.POINTER(CloudInfo)]
    L.liloc_host_ip_cloud_livox.restype = C.c_int
    L.liloc_host_ip_cloud_livox.argtypes = [P, C.c_double, P, C.c_int, C.POINTER(CloudInfo)]
    L.liloc_host_ip_times.argtypes = [P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.liloc_host_ip_last_sweep.restype = C.c_int
    L.liloc_host_ip_last_sweep.argtypes = [P, P, C.c_int, C.POINTER(C.c_int), P, P, C.c_int]
    return L


_lib = None
PRIOR_KEYFRAME_BASE = 1 << 24      # mapOptimization::kPriorKeyframeBase: device keyframe ids of the prior session's key clouds


def hypothesis_pose(center6, n_yaw: int, n_xy: int, radius: float, h: int) -> np.ndarray:
    """mapOptimization::hypothesisPose: hypothesis h of the multi-hypothesis relocalisation init."""
    c = np.ascontiguousarray(center6, np.float32).reshape(6)
    out = np.zeros(6, np.float32)
    lib().liloc_host_hypothesis_pose(c.ctypes.data, int(n_yaw), int(n_xy), float(radius), int(h), out.ctypes.data)
    return out


def euler_to_matrix(pose6) -> np.ndarray:
    """init_guess of the NDT call sites exactly as the host class builds it (eulerToRotation in float32,
    src/mapOptmization.cpp:276-280): (4, 4) float32.  Parity tests hand the oracle these very bits."""
    p = np.ascontiguousarray(pose6, np.float32).reshape(6)
    T = np.zeros(16, np.float32)
    lib().liloc_host_debug_euler_to_matrix(p.ctypes.data, T.ctypes.data)
    return T.reshape(4, 4)


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def load_pcd(path: str) -> np.ndarray:
    """pcl::io::loadPCDFile<PointXYZI> as the host reads session files (liloc_b200/host/session_io.cpp)."""
    err = C.create_string_buffer(512)
    n = lib().liloc_host_load_pcd(path.encode(), None, 0, err, 512)
    if n < 0:
        raise IOError(err.value.decode())
    out = np.empty((max(n, 1), 4), np.float32)
    lib().liloc_host_load_pcd(path.encode(), out.ctypes.data, len(out), err, 512)
    return out[:n]


def save_pcd(path: str, pts) -> None:
    p = np.ascontiguousarray(pts, np.float32).reshape(-1, 4)
    if lib().liloc_host_save_pcd(path.encode(), p.ctypes.data, len(p)) != 0:
        raise IOError(path)


def load_g2o(path: str) -> np.ndarray:
    """Key poses (n, 7) [roll, pitch, yaw, x, y, z, node id] of a singlesession_posegraph.g2o."""
    n = lib().liloc_host_load_g2o(path.encode(), None, 0)
    if n < 0:
        raise IOError(path)
    out = np.zeros((n, 7), np.float64)
    lib().liloc_host_load_g2o(path.encode(), out.ctypes.data, n)
    return out


def save_g2o(path: str, keyposes7) -> None:
    kp = np.ascontiguousarray(keyposes7, np.float64).reshape(-1, 7)
    if lib().liloc_host_save_g2o(path.encode(), kp.ctypes.data, len(kp)) != 0:
        raise IOError(path)


def yaml_param(yaml: str, key: str) -> float:
    v = C.c_double()
    rc = lib().liloc_host_yaml_param(config_path(yaml).encode(), key.encode(), C.byref(v))
    if rc != 0:
        raise KeyError(f"{yaml}:{key} ({rc})")
    return v.value


def config_path(name: str) -> str:
    return name if os.path.isabs(name) else os.path.join(CONFIG_DIR, name)


class MapOptimization:
    """Handle on one host `mapOptimization` object (one GPU context inside)."""

    def __init__(self, yaml: str, mode: str | None = None, device: int = 0):
        L = lib()
        self._h = L.liloc_host_create(config_path(yaml).encode(), (mode or "").encode(), device)
        if not self._h:
            raise RuntimeError("mapOptimization: " + (L.liloc_host_error(None) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            lib().liloc_host_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def param(self, key: str) -> float:
        v = C.c_double()
        if lib().liloc_host_param(self._h, key.encode(), C.byref(v)) != 0:
            raise KeyError(key)
        return v.value

    def reset(self):
        lib().liloc_host_reset(self._h)

    def set_s2m_opts(self, max_iters=20, stride=2, min_sel=50, min_points=30):
        lib().liloc_host_set_s2m_opts(self._h, max_iters, stride, min_sel, min_points)

    def handle(self, info: CloudInfo) -> Report:
        r = Report()
        rc = lib().liloc_host_handle(self._h, C.byref(info), C.byref(r))
        if rc < 0:
            raise RuntimeError(f"laserCloudInfoHandler failed ({rc}): " + (lib().liloc_host_error(self._h) or b"").decode())
        return r

    def run(self, infos) -> list:
        """infos: ctypes array of CloudInfo.  Processes them back to back natively."""
        n = len(infos)
        reps = (Report * n)()
        rc = lib().liloc_host_run(self._h, infos, n, reps)
        if rc < 0:
            raise RuntimeError(f"liloc_host_run failed ({rc}): " + (lib().liloc_host_error(self._h) or b"").decode())
        return reps

    def set_share(self, share: int) -> int:
        """Throughput mode: this object is one of `share` objects processed concurrently on its GPU, one host thread each
        (liloc_set_share: 1 / share of the blocks per launch so that the frames of the sequences are co-resident)."""
        return int(lib().liloc_host_set_share(self._h, int(share)))

    # ---- relo mode ----
    def set_relo_hypotheses(self, n_yaw: int, n_xy: int, radius: float = 2.0) -> None:
        """Multi-hypothesis relocalisation init (n_yaw * n_xy hypotheses, sharded over the ranks of comm_init)."""
        lib().liloc_host_set_relo_hypotheses(self._h, int(n_yaw), int(n_xy), float(radius))

    def comm_init(self, rank: int, world: int, unique_id: bytes | None) -> None:
        buf = C.create_string_buffer(unique_id or b"", 128)
        rc = lib().liloc_host_comm_init(self._h, int(rank), int(world), buf if world > 1 else None)
        if rc < 0:
            raise RuntimeError("comm_init: " + (lib().liloc_host_error(self._h) or b"").decode())

    def load_prior_session(self, keyposes7, clouds) -> int:
        """Prior session: key poses (n, 7) [roll, pitch, yaw, x, y, z, time] + their key clouds (body frame).
        Returns the number of submaps (dataLoader.hpp generateSubMaps, every 10 keyframes)."""
        kp = np.ascontiguousarray(keyposes7, np.float64).reshape(-1, 7)
        flat = np.ascontiguousarray(np.concatenate(clouds), np.float32)
        offs = np.zeros(len(clouds) + 1, np.int64)
        offs[1:] = np.cumsum([len(c) for c in clouds])
        n = lib().liloc_host_load_prior(self._h, kp.ctypes.data, len(kp), flat.ctypes.data, offs.ctypes.data)
        if n < 0:
            raise RuntimeError("loadPriorSession: " + (lib().liloc_host_error(self._h) or b"").decode())
        return n

    def load_prior_session_dir(self, session_dir: str) -> int:
        n = lib().liloc_host_load_prior_dir(self._h, session_dir.encode())
        if n < 0:
            raise RuntimeError("loadPriorSessionDir: " + (lib().liloc_host_error(self._h) or b"").decode())
        return n

    def save_session(self, session_dir: str, resolution: float = 0.0) -> None:
        """save_map (lio) / save_session (relo: the prior session with what this session joined to it) at req.resolution."""
        if lib().liloc_host_save_session_res(self._h, session_dir.encode(), float(resolution)) != 0:
            raise RuntimeError("saveSession: " + (lib().liloc_host_error(self._h) or b"").decode())

    def initialpose(self, x: float, y: float, yaw: float):
        """/initialpose (src/mapOptmization.cpp:755-769)."""
        lib().liloc_host_initialpose(self._h, float(x), float(y), float(yaw))

    def submap_info(self, s: int) -> dict:
        c = np.zeros(3, np.float32)
        first, count, pts = C.c_int(), C.c_int(), C.c_int()
        if lib().liloc_host_submap_info(self._h, s, c.ctypes.data, C.byref(first), C.byref(count), C.byref(pts)) != 0:
            raise IndexError(s)
        return {"centroid": c, "first": first.value, "count": count.value, "points": pts.value}

    def global_map(self) -> np.ndarray:
        """The cloud publishGlobalMap (:698-753) would publish."""
        n = lib().liloc_host_global_map(self._h, None, 0)
        if n < 0:
            raise RuntimeError("publishGlobalMap: " + (lib().liloc_host_error(self._h) or b"").decode())
        out = np.empty((max(n, 1), 4), np.float32)
        lib().liloc_host_global_map(self._h, out.ctypes.data, len(out))
        return out[:n]

    def num_keyframes(self) -> int:
        return lib().liloc_host_num_keyframes(self._h)

    def keyposes(self) -> np.ndarray:
        n = self.num_keyframes()
        out = np.zeros((max(n, 1), 7), np.float64)
        lib().liloc_host_keyposes(self._h, out.ctypes.data, n)
        return out[:n]

    def last_selection(self) -> np.ndarray:
        ids = np.zeros(4096, np.int32)
        n = lib().liloc_host_last_selection(self._h, ids.ctypes.data, len(ids))
        return ids[:n].copy()

    # ---- host-logic hooks (CPU tests) ----
    def debug_add_keypose(self, pose6, time: float):
        p = np.ascontiguousarray(pose6, np.float32)
        lib().liloc_host_debug_add_keypose(self._h, p.ctypes.data, float(time))

    def debug_extract_nearby(self, stamp: float) -> np.ndarray:
        ids = np.zeros(8192, np.int32)
        n = lib().liloc_host_debug_extract_nearby(self._h, float(stamp), ids.ctypes.data, len(ids))
        return ids[:n].copy()

    def debug_nearby_plain(self, on: bool):
        lib().liloc_host_debug_nearby_plain(self._h, int(on))

    def debug_set_pose(self, pose6):
        p = np.ascontiguousarray(pose6, np.float32)
        lib().liloc_host_debug_set_pose(self._h, p.ctypes.data)

    def debug_get_pose(self) -> np.ndarray:
        p = np.zeros(6, np.float32)
        lib().liloc_host_debug_get_pose(self._h, p.ctypes.data)
        return p

    def debug_update_initial_guess(self, info: CloudInfo) -> np.ndarray:
        lib().liloc_host_debug_update_initial_guess(self._h, C.byref(info))
        return self.debug_get_pose()

    def debug_save_frame(self) -> bool:
        return bool(lib().liloc_host_debug_save_frame(self._h))

    def debug_transform_update(self, info: CloudInfo) -> np.ndarray:
        lib().liloc_host_debug_transform_update(self._h, C.byref(info))
        return self.debug_get_pose()

    def ctx_handle(self) -> int:
        return lib().liloc_host_ctx(self._h)


LIVOX_POINT = np.dtype([("x", np.float32), ("y", np.float32), ("z", np.float32), ("reflectivity", np.uint8), ("tag", np.uint8),
                        ("line", np.uint8), ("_pad", np.uint8), ("offset_time", np.uint32)])
RING_TIME = np.dtype([("ring", np.int32), ("time", np.float32)])


class ImageProjection:
    """Handle on one host `ImageProjection` (src/imageProjection.cpp:63) that shares the device context of a
    MapOptimization: the CloudInfo it produces refers to the de-skewed sweep left on the device."""

    def __init__(self, yaml: str, map_optimization: MapOptimization | None):
        """map_optimization = None: host logic only (no device; projectPointCloud's device call is skipped)."""
        L = lib()
        self._mo = map_optimization
        self._h = L.liloc_host_ip_create(config_path(yaml).encode(), map_optimization._h if map_optimization is not None else None)
        if not self._h:
            raise RuntimeError("ImageProjection: " + (L.liloc_host_ip_error(None) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            lib().liloc_host_ip_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def imu(self, stamp, gyro, acc=(0.0, 0.0, 9.8), quat=(0.0, 0.0, 0.0, 1.0)) -> None:
        m = np.array([stamp, *gyro, *acc, *quat], np.float64)
        lib().liloc_host_ip_imu(self._h, m.ctypes.data)

    def odom(self, stamp, position, quat=(0.0, 0.0, 0.0, 1.0), covariance0=0.0) -> None:
        m = np.array([stamp, *position, *quat, covariance0], np.float64)
        lib().liloc_host_ip_odom(self._h, m.ctypes.data)

    def _ret(self, rc, info):
        if rc < 0:
            raise RuntimeError("ImageProjection: " + (lib().liloc_host_ip_error(self._h) or b"").decode())
        return info if rc == 1 else None

    def cloud(self, stamp, pts, ring=None, time=None, has_time=True):
        """cloudHandler: returns the CloudInfo the reference would publish, or None (queue filling / waiting for IMU)."""
        p = np.ascontiguousarray(pts, np.float32)
        rt = None
        if ring is not None:
            rt = np.empty(len(p), RING_TIME)
            rt["ring"], rt["time"] = np.asarray(ring, np.int32), np.asarray(time, np.float32)
        info = CloudInfo()
        rc = lib().liloc_host_ip_cloud(self._h, float(stamp), p.ctypes.data, rt.ctypes.data if rt is not None else None, len(p),
                                       1 if has_time else 0, C.byref(info))
        return self._ret(rc, info)

    def cloud_livox(self, stamp, points):
        p = np.ascontiguousarray(points, LIVOX_POINT)
        info = CloudInfo()
        rc = lib().liloc_host_ip_cloud_livox(self._h, float(stamp), p.ctypes.data, len(p), C.byref(info))
        return self._ret(rc, info)

    def last_sweep(self):
        """(imu_time (m,), imu_rot (m, 3), pts (n, 4), ring (n,), time (n,)) of the last published sweep's de-skew call."""
        cap = 2000
        imu = np.zeros((4, cap), np.float64)
        last = C.c_int()
        n = lib().liloc_host_ip_last_sweep(self._h, imu.ctypes.data, cap, C.byref(last), None, None, 0)
        pts = np.zeros((n, 4), np.float32); rt = np.zeros(n, RING_TIME)
        lib().liloc_host_ip_last_sweep(self._h, imu.ctypes.data, cap, C.byref(last), pts.ctypes.data, rt.ctypes.data, n)
        m = last.value + 1 if last.value >= 0 else 0
        return imu[0, :m].copy(), imu[1:, :m].T.copy(), pts, rt["ring"].copy(), rt["time"].copy()

    def times(self):
        a, b, f = C.c_double(), C.c_double(), C.c_int()
        lib().liloc_host_ip_times(self._h, C.byref(a), C.byref(b), C.byref(f))
        return a.value, b.value, f.value


def make_infos(scans, odom_poses6, stamps, device_ptrs=None):
    """Build a ctypes array of CloudInfo.  scans: list of (n,4) float32 host arrays (kept alive by the
    caller); odom_poses6: (F,6) [roll,pitch,yaw,x,y,z] 'IMU pre-integration' poses (initialGuess*);
    device_ptrs: optional list of (ptr, n) for device-resident clouds."""
    F = len(stamps)
    arr = (CloudInfo * F)()
    for i in range(F):
        ci = arr[i]
        ci.stamp = float(stamps[i])
        ci.imuAvailable, ci.odomAvailable = 0, 1
        r, p, y, x, yy, z = [float(v) for v in odom_poses6[i]]
        ci.imuRollInit, ci.imuPitchInit, ci.imuYawInit = r, p, y
        ci.initialGuessX, ci.initialGuessY, ci.initialGuessZ = x, yy, z
        ci.initialGuessRoll, ci.initialGuessPitch, ci.initialGuessYaw = r, p, y
        if device_ptrs is not None:
            ci.cloud_deskewed, ci.cloud_size, ci.cloud_on_device = int(device_ptrs[i][0]), int(device_ptrs[i][1]), 1
        else:
            ci.cloud_deskewed, ci.cloud_size, ci.cloud_on_device = scans[i].ctypes.data, len(scans[i]), 0
    return arr
