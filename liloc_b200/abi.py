"""ctypes binding of include/liloc_gpu.h (the same stub a maintainer of the reference would write
in C++ -- see INTEGRATION.md).  Every call goes through the C ABI; nothing here computes."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
lib_path = os.path.join(_HERE, "libliloc_gpu.so")

LILOC_OK = 0
LILOC_SKIPPED = 1

POINT_DTYPE = np.dtype(np.float32)   # clouds are (n, 4) float32 arrays: x, y, z, intensity


class LilocError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"liloc error {code}: {msg}")
        self.code = code


class _Config(C.Structure):
    _fields_ = [("device", C.c_int), ("max_scan_points", C.c_int), ("max_raw_points", C.c_int), ("map_cache", C.c_int),
                ("share", C.c_int), ("reserved", C.c_int * 3)]


class S2mOpts(C.Structure):
    """liloc_s2m_opts; defaults are the reference's constants (src/mapOptmization.cpp:1190,985,1080,1187)."""
    _fields_ = [("max_iters", C.c_int), ("stride", C.c_int), ("min_sel", C.c_int), ("min_points", C.c_int),
                ("enable_corner", C.c_int), ("min_points_corner", C.c_int), ("stride_corner", C.c_int),
                ("reserved", C.c_int * 1)]

    def __init__(self, max_iters: int = 20, stride: int = 2, min_sel: int = 50, min_points: int = 30,
                 enable_corner: int = 0, min_points_corner: int = 10, stride_corner: int = 1):
        super().__init__(max_iters, stride, min_sel, min_points, enable_corner, min_points_corner, stride_corner)

    @classmethod
    def upstream(cls) -> "S2mOpts":
        """Upstream LIO-SAM loop shape named by the north_star: corner + surf, stride 1, 30 iterations,
        gate corner > 10 and surf > 100."""
        return cls(max_iters=30, stride=1, min_sel=50, min_points=100, enable_corner=1, min_points_corner=10, stride_corner=1)


class FrameIn(C.Structure):
    _fields_ = [("kf_ids", C.c_void_p), ("kf_poses6", C.c_void_p), ("K", C.c_int), ("map_leaf", C.c_float * 2),
                ("scan", C.c_void_p * 2), ("n_scan", C.c_int * 2), ("scan_leaf", C.c_float * 2), ("scan_on_device", C.c_int)]


class S2mExtra(C.Structure):
    _fields_ = [("q_corner", C.c_int), ("map_points_corner", C.c_int), ("tiles_corner", C.c_int), ("tiles_surf", C.c_int),
                ("sort_redos", C.c_int), ("reserved", C.c_int * 3)]


class S2mResult(C.Structure):
    _fields_ = [("iters", C.c_int), ("converged", C.c_int), ("n_sel", C.c_int), ("is_degenerate", C.c_int),
                ("too_few", C.c_int), ("skipped", C.c_int), ("q_all", C.c_int), ("map_points", C.c_int),
                ("matP", C.c_float * 36), ("AtA0", C.c_float * 36)]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k in ("iters", "converged", "n_sel", "is_degenerate", "too_few",
                                               "skipped", "q_all", "map_points")}


class Timing(C.Structure):
    _fields_ = [(k, C.c_float) for k in ("assemble_ms", "map_voxel_ms", "grid_ms", "scan_ms", "s2m_ms", "total_ms")]


class Stats(C.Structure):
    _fields_ = [(k, C.c_longlong) for k in ("frames", "launches", "s2m_launches", "s2m_iters", "h2d_bytes", "d2h_bytes")] + \
               [(k, C.c_double) for k in ("assemble_ms", "map_voxel_ms", "grid_ms", "scan_ms", "s2m_ms", "total_ms",
                                          "alg_bytes_map", "alg_bytes_scan", "alg_bytes_s2m")]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_}


def stats_of(ctx_handle) -> dict:
    st = Stats()
    lib.liloc_stats_get(C.c_void_p(ctx_handle), C.byref(st))
    return st.as_dict()


def last_deskew_ms_of(ctx_handle) -> float:
    ms = C.c_float()
    lib.liloc_last_deskew_ms(C.c_void_p(ctx_handle), C.byref(ms))
    return ms.value


class DeskewParams(C.Structure):
    """liloc_deskew_params: what imuDeskewInfo / cachePointCloud leave for projectPointCloud (src/imageProjection.cpp)."""
    _fields_ = [("time_scan_cur", C.c_double),
                ("imu_time", C.c_void_p), ("imu_rot_x", C.c_void_p), ("imu_rot_y", C.c_void_p), ("imu_rot_z", C.c_void_p),
                ("imu_pointer_cur", C.c_int), ("deskew_flag", C.c_int), ("imu_available", C.c_int),
                ("lidar_min_range", C.c_float), ("lidar_max_range", C.c_float),
                ("n_scan", C.c_int), ("downsample_rate", C.c_int), ("point_filter_num", C.c_int), ("reserved", C.c_int * 2)]


RING_TIME = np.dtype([("ring", np.int32), ("time", np.float32)])
SCAN_HOST, SCAN_DEVICE, SCAN_DESKEWED = 0, 1, 2
DESKEW_AHEAD = 1


class NdtOpts(C.Structure):
    """liloc_ndt_opts; defaults = ndt_omp defaults + LiLoc's yaml (ndtEpsilon 0.01, regMethod DIRECT1)."""
    _fields_ = [("step_size", C.c_double), ("outlier_ratio", C.c_double), ("trans_eps", C.c_double),
                ("max_iterations", C.c_int), ("search", C.c_int), ("n_align", C.c_int), ("reserved", C.c_int)]

    def __init__(self, trans_eps=0.01, search=0, n_align=1, step_size=0.1, outlier_ratio=0.55, max_iterations=35):
        super().__init__(step_size, outlier_ratio, float(np.float32(trans_eps)), max_iterations, search, n_align, 0)


class NdtStats(C.Structure):
    _fields_ = [(k, C.c_longlong) for k in ("batches", "jobs", "launches", "blocks", "tickets", "updates", "job_sweeps", "max_job_sweeps")] + \
               [(k, C.c_double) for k in ("kernel_ms", "alg_bytes", "block_sweep_ms", "block_update_ms", "block_wait_ms")] + \
               [("update_cycles", C.c_double * 4)]


class NdtJob(C.Structure):
    _fields_ = [("target_id", C.c_int), ("source_id", C.c_int), ("guess", C.c_float * 16)]


class NdtOut(C.Structure):
    _fields_ = [("T", C.c_float * 16), ("score", C.c_double), ("iterations", C.c_int), ("converged", C.c_int),
                ("sweeps", C.c_int), ("n_source", C.c_int)]


_NDT_JOB_DT = np.dtype([("target_id", np.int32), ("source_id", np.int32), ("guess", np.float32, 16)])
_NDT_OUT_DT = np.dtype([("T", np.float32, 16), ("score", np.float64), ("iterations", np.int32), ("converged", np.int32),
                        ("sweeps", np.int32), ("n_source", np.int32)], align=True)
assert _NDT_JOB_DT.itemsize == C.sizeof(NdtJob) and _NDT_OUT_DT.itemsize == C.sizeof(NdtOut)


def _load() -> C.CDLL:
    if not os.path.exists(lib_path):
        raise ImportError(
            f"{lib_path} is missing: build it with `python -m liloc_b200._build` (nvcc, sm_100a). "
            "liloc_b200 has no CPU fallback.")
    L = C.CDLL(lib_path)
    P = C.c_void_p
    ip, fp = C.POINTER(C.c_int), C.POINTER(C.c_float)
    sig = {
        "liloc_create": (C.c_int, [C.POINTER(_Config), C.POINTER(P)]),
        "liloc_destroy": (None, [P]),
        "liloc_last_error": (C.c_char_p, [P]),
        "liloc_version": (C.c_char_p, []),
        "liloc_keyframe_put": (C.c_int, [P, C.c_int, P, C.c_int]),
        "liloc_keyframe_put_from_scan": (C.c_int, [P, C.c_int]),
        "liloc_keyframe_size": (C.c_int, [P, C.c_int]),
        "liloc_deskew": (C.c_int, [P, P, P, C.c_int, C.POINTER(DeskewParams), ip]),
        "liloc_deskewed_get": (C.c_int, [P, P, C.c_int, ip]),
        "liloc_deskew_advance": (C.c_int, [P]),
        "liloc_last_deskew_ms": (C.c_int, [P, fp]),
        "liloc_last_deskew_marks": (C.c_int, [P, P]),
        "liloc_keyframe_clear": (C.c_int, [P]),
        "liloc_localmap_build": (C.c_int, [P, P, P, C.c_int, C.c_float, ip, ip]),
        "liloc_localmap_set": (C.c_int, [P, P, C.c_int, C.c_float, ip]),
        "liloc_localmap_get": (C.c_int, [P, P, C.c_int, ip]),
        "liloc_scan_put": (C.c_int, [P, P, C.c_int, C.c_float, ip]),
        "liloc_scan_put_dev": (C.c_int, [P, P, C.c_int, C.c_float, ip]),
        "liloc_scan_ds_get": (C.c_int, [P, P, C.c_int, ip]),
        "liloc_scan2map": (C.c_int, [P, P, C.POINTER(S2mOpts), C.POINTER(S2mResult)]),
        "liloc_frame": (C.c_int, [P, P, P, C.c_int, C.c_float, P, C.c_int, C.c_int, C.c_float, P,
                                  C.POINTER(S2mOpts), C.POINTER(S2mResult)]),
        "liloc_voxelgrid": (C.c_int, [P, P, C.c_int, C.c_float, P, C.c_int, ip]),
        "liloc_knn5": (C.c_int, [P, P, C.c_int, P, P, P]),
        "liloc_surf_pass": (C.c_int, [P, P, C.c_int, P, P]),
        "liloc_pose_to_T": (C.c_int, [P, P, P]),
        "liloc_last_timing": (C.c_int, [P, C.POINTER(Timing)]),
        "liloc_set_timing": (C.c_int, [P, C.c_int]),
        "liloc_kernel_launches": (C.c_longlong, [P]),
        "liloc_last_trace": (C.c_int, [P, P, C.c_int]),
        "liloc_debug_knn5_bounded": (C.c_int, [P, C.c_int, P, C.c_int, P, P, P, P]),
        "liloc_ndt_target_put": (C.c_int, [P, C.c_int, P, C.c_int, C.c_float, ip]),
        "liloc_ndt_source_put": (C.c_int, [P, C.c_int, P, C.c_int, C.c_int]),
        "liloc_ndt_align_batch": (C.c_int, [P, C.POINTER(NdtJob), C.c_int, C.POINTER(NdtOpts), C.POINTER(NdtOut)]),
        "liloc_ndt_clear": (C.c_int, [P]),
        "liloc_ndt_target_get": (C.c_int, [P, C.c_int, P, P, P, P, C.c_int]),
        "liloc_ndt_derivatives": (C.c_int, [P, C.c_int, C.c_int, P, C.POINTER(NdtOpts), P]),
        "liloc_last_clocks": (C.c_int, [P, P, C.c_int]),
        "liloc_stats_get": (C.c_int, [P, C.POINTER(Stats)]),
        "liloc_stats_reset": (C.c_int, [P]),
        "liloc_stream": (C.c_void_p, [P]),
        "liloc_debug_lm6": (C.c_int, [P, P, P, C.c_int, P, P, P, P]),
        "liloc_debug_plane": (C.c_int, [P, P, C.c_int, P]),
        "liloc_frame_features": (C.c_int, [P, C.POINTER(FrameIn), P, C.POINTER(S2mOpts), C.POINTER(S2mResult)]),
        "liloc_last_s2m_extra": (C.c_int, [P, C.POINTER(S2mExtra)]),
        "liloc_scan_prefetch": (C.c_int, [P, P, C.c_int]),
        "liloc_keyframe_put_features": (C.c_int, [P, C.c_int, P, C.c_int, P, C.c_int]),
        "liloc_keyframe_size_class": (C.c_int, [P, C.c_int, C.c_int]),
        "liloc_localmap_build_class": (C.c_int, [P, C.c_int, P, P, C.c_int, C.c_float, ip, ip]),
        "liloc_localmap_set_class": (C.c_int, [P, C.c_int, P, C.c_int, C.c_float, ip]),
        "liloc_localmap_get_class": (C.c_int, [P, C.c_int, P, C.c_int, ip]),
        "liloc_scan_put_class": (C.c_int, [P, C.c_int, P, C.c_int, C.c_int, C.c_float, ip]),
        "liloc_scan_ds_get_class": (C.c_int, [P, C.c_int, P, C.c_int, ip]),
        "liloc_knn5_class": (C.c_int, [P, C.c_int, P, C.c_int, P, P, P]),
        "liloc_feature_pass": (C.c_int, [P, C.c_int, P, C.c_int, P, P]),
        "liloc_debug_eigen3": (C.c_int, [P, P, C.c_int, P, P]),
        "liloc_last_marks": (C.c_int, [P, P, C.c_int]),
        "liloc_set_map_cache": (C.c_longlong, [P, C.c_int]),
        "liloc_set_share": (C.c_int, [P, C.c_int]),
        "liloc_keyframe_get": (C.c_int, [P, C.c_int, C.c_int, P, C.c_int, ip]),
        "liloc_globalmap_build": (C.c_int, [P, P, P, C.c_int, C.c_float, P, C.c_int, ip]),
        "liloc_icp_fitness": (C.c_int, [P, P, C.c_int, P, P, C.c_int, P, fp]),
        "liloc_icp_fitness_resident": (C.c_int, [P, C.c_int, P, C.c_int, P, fp]),
        "liloc_ndt_target_put_keyframes": (C.c_int, [P, C.c_int, P, P, C.c_int, C.c_float, ip, ip]),
        "liloc_ndt_target_size": (C.c_int, [P, C.c_int]),
        "liloc_ndt_stats_get": (C.c_int, [P, C.POINTER(NdtStats)]),
        "liloc_ndt_stats_reset": (C.c_int, [P]),
        "liloc_debug_line": (C.c_int, [P, P, P, C.c_int, P, P]),
        "liloc_keyframe_put_from_raw_scan": (C.c_int, [P, C.c_int]),
        "liloc_ndt_target_put_keyframes_filtered": (C.c_int, [P, C.c_int, P, P, C.c_int, C.c_float, C.c_float, ip, ip]),
        "liloc_comm_unique_id": (C.c_int, [P]),
        "liloc_comm_init": (C.c_int, [P, C.c_int, C.c_int, P]),
        "liloc_comm_destroy": (C.c_int, [P]),
        "liloc_shard_indices": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, P, C.c_int]),
        "liloc_unshard_rows": (C.c_int, [P, C.c_int, C.c_int, C.c_int, C.c_int, P]),
        "liloc_comm_rank": (C.c_int, [P]),
        "liloc_comm_world": (C.c_int, [P]),
        "liloc_ndt_align_batch_sharded": (C.c_int, [P, C.POINTER(NdtJob), C.c_int, C.POINTER(NdtOpts), C.c_int, P, C.POINTER(NdtOut)]),
        "liloc_gather_results": (C.c_int, [P, P, C.c_int, C.c_int, P]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    return L


lib = _load()
EXPORTS = ("liloc_create liloc_destroy liloc_last_error liloc_version liloc_keyframe_put "
           "liloc_keyframe_put_from_scan liloc_keyframe_size liloc_keyframe_clear liloc_localmap_build "
           "liloc_localmap_set liloc_localmap_get liloc_scan_put liloc_scan_put_dev liloc_scan_ds_get "
           "liloc_scan2map liloc_frame liloc_voxelgrid liloc_knn5 liloc_surf_pass liloc_pose_to_T "
           "liloc_last_timing liloc_set_timing liloc_kernel_launches liloc_last_trace liloc_last_clocks liloc_ndt_target_put liloc_ndt_source_put liloc_ndt_align_batch liloc_ndt_clear liloc_ndt_target_get liloc_ndt_derivatives liloc_stats_get liloc_stats_reset liloc_stream liloc_debug_lm6 liloc_debug_plane "
           "liloc_frame_features liloc_last_s2m_extra liloc_keyframe_put_features liloc_keyframe_size_class "
           "liloc_localmap_build_class liloc_localmap_set_class liloc_localmap_get_class liloc_scan_put_class "
           "liloc_scan_ds_get_class liloc_knn5_class liloc_feature_pass liloc_debug_eigen3 liloc_debug_line liloc_last_marks liloc_ndt_stats_get liloc_ndt_stats_reset liloc_ndt_target_put_keyframes liloc_ndt_target_size liloc_set_map_cache liloc_icp_fitness liloc_icp_fitness_resident liloc_globalmap_build liloc_keyframe_get liloc_deskew liloc_deskewed_get liloc_last_deskew_ms liloc_last_deskew_marks liloc_deskew_advance "
           "liloc_keyframe_put_from_raw_scan liloc_ndt_target_put_keyframes_filtered liloc_comm_unique_id liloc_comm_init liloc_comm_destroy liloc_shard_indices liloc_unshard_rows liloc_comm_rank liloc_comm_world liloc_ndt_align_batch_sharded liloc_gather_results liloc_scan_prefetch liloc_set_share liloc_debug_knn5_bounded").split()


COMM_ID_BYTES = 128
SHARD_BLOCKS, SHARD_ROUND_ROBIN = 0, 1


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the C ABI (rank 0); the launcher hands the 128 bytes to every rank."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    rc = lib.liloc_comm_unique_id(buf)
    if rc != 0:
        raise LilocError(rc, (lib.liloc_last_error(None) or b"").decode())
    return buf.raw


def shard_indices(n_items: int, rank: int, world: int, round_robin: bool = False) -> np.ndarray:
    """liloc_shard_indices: the items of `rank`, in the order their rows travel."""
    out = np.zeros(max((n_items + world - 1) // world, 1), np.int32)
    n = lib.liloc_shard_indices(n_items, rank, world, SHARD_ROUND_ROBIN if round_robin else SHARD_BLOCKS, _ptr(out), len(out))
    if n < 0:
        raise LilocError(n, "liloc_shard_indices: bad arguments")
    return out[:n].astype(np.int64)


def unshard_rows(rows_all: np.ndarray, n_items: int, round_robin: bool = False) -> np.ndarray:
    """liloc_unshard_rows: gathered (world, per_rank, 8) rows -> (n_items, 8) in item order."""
    a = np.ascontiguousarray(rows_all, np.float32)
    world, per = a.shape[0], a.shape[1]
    out = np.zeros((n_items, 8), np.float32)
    rc = lib.liloc_unshard_rows(_ptr(a), n_items, world, per, SHARD_ROUND_ROBIN if round_robin else SHARD_BLOCKS, _ptr(out))
    if rc != 0:
        raise LilocError(rc, "liloc_unshard_rows: bad arguments")
    return out


def _cloud(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] != 4:
        raise ValueError("cloud must have shape (n, 4): x, y, z, intensity")
    return a


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


class Context:
    """One liloc_ctx (one GPU, one stream).  Mirrors the calls mapOptimization makes per frame."""

    def __init__(self, device: int = 0, max_scan_points: int = 32 * 1800, max_raw_points: int = 1 << 20, map_cache: bool = False,
                 share: int = 0):
        cfg = _Config(device, max_scan_points, max_raw_points, int(map_cache), int(share))
        h = C.c_void_p()
        rc = lib.liloc_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise LilocError(rc, (lib.liloc_last_error(None) or b"").decode())
        self._h = h
        self.device = device

    @classmethod
    def borrow(cls, handle: int, device: int = 0) -> "Context":
        """A view on a liloc_ctx owned by someone else (the host class): close() does not destroy it."""
        self = cls.__new__(cls)
        self._h, self.device, self._borrowed = C.c_void_p(handle), device, True
        return self

    def close(self) -> None:
        if getattr(self, "_h", None):
            if not getattr(self, "_borrowed", False):
                lib.liloc_destroy(self._h)
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

    def _chk(self, rc: int) -> int:
        if rc < 0:
            raise LilocError(rc, (lib.liloc_last_error(self._h) or b"").decode())
        return rc

    # ---- keyframes ----
    def keyframe_put(self, kf_id: int, pts) -> None:
        p = _cloud(pts)
        self._chk(lib.liloc_keyframe_put(self._h, kf_id, _ptr(p), len(p)))

    def keyframe_put_features(self, kf_id: int, surf, corner) -> None:
        a, b = _cloud(surf), _cloud(corner)
        self._chk(lib.liloc_keyframe_put_features(self._h, kf_id, _ptr(a), len(a), _ptr(b), len(b)))

    def keyframe_put_from_scan(self, kf_id: int) -> None:
        self._chk(lib.liloc_keyframe_put_from_scan(self._h, kf_id))

    def keyframe_size(self, kf_id: int, cls: int = 0) -> int:
        return lib.liloc_keyframe_size_class(self._h, cls, kf_id)

    def keyframe_get(self, kf_id: int, cls: int = 0) -> np.ndarray:
        n = C.c_int()
        self._chk(lib.liloc_keyframe_get(self._h, cls, kf_id, None, 0, C.byref(n)))
        out = np.empty((max(n.value, 1), 4), np.float32)
        self._chk(lib.liloc_keyframe_get(self._h, cls, kf_id, _ptr(out), len(out), C.byref(n)))
        return out[: n.value]

    def keyframe_clear(self) -> None:
        self._chk(lib.liloc_keyframe_clear(self._h))

    # ---- local map ----
    def localmap_build(self, kf_ids, poses6, leaf: float, cls: int = 0) -> tuple[int, int]:
        ids = np.ascontiguousarray(kf_ids, dtype=np.int32)
        poses = np.ascontiguousarray(poses6, dtype=np.float32).reshape(-1, 6)
        assert len(ids) == len(poses)
        M, n_raw = C.c_int(), C.c_int()
        self._chk(lib.liloc_localmap_build_class(self._h, cls, _ptr(ids), _ptr(poses), len(ids), leaf, C.byref(M), C.byref(n_raw)))
        return M.value, n_raw.value

    def localmap_set(self, pts_map, leaf: float, cls: int = 0) -> int:
        p = _cloud(pts_map)
        M = C.c_int()
        self._chk(lib.liloc_localmap_set_class(self._h, cls, _ptr(p), len(p), leaf, C.byref(M)))
        return M.value

    def localmap_get(self, cls: int = 0) -> np.ndarray:
        M = C.c_int()
        self._chk(lib.liloc_localmap_get_class(self._h, cls, None, 0, C.byref(M)))
        out = np.empty((M.value, 4), np.float32)
        self._chk(lib.liloc_localmap_get_class(self._h, cls, _ptr(out), M.value, C.byref(M)))
        return out

    # ---- scan ----
    def scan_put(self, pts_body, leaf: float, cls: int = 0) -> int:
        p = _cloud(pts_body)
        Q = C.c_int()
        self._chk(lib.liloc_scan_put_class(self._h, cls, _ptr(p), len(p), 0, leaf, C.byref(Q)))
        return Q.value

    def scan_put_dev(self, dev_ptr: int, n: int, leaf: float) -> int:
        Q = C.c_int()
        self._chk(lib.liloc_scan_put_dev(self._h, C.c_void_p(dev_ptr), n, leaf, C.byref(Q)))
        return Q.value

    # ---- de-skew (imageProjection::projectPointCloud) ----
    def deskew(self, pts, ring, time, imu_time=None, imu_rot=None, time_scan_cur: float = 0.0, deskew_flag: int = 1,
               imu_available: bool = True, min_range: float = 1.0, max_range: float = 1000.0, n_scan: int = 16,
               downsample_rate: int = 1, point_filter_num: int = 3, sync: bool = True, ahead: bool = False) -> int | None:
        """liloc_deskew.  imu_time: (m,) float64, imu_rot: (m, 3) float64 integrated angles (imuRotX/Y/Z); the last valid
        index is m - 1.  Returns the size of the de-skewed cloud (None with sync=False: it stays on the device)."""
        p = _cloud(pts)
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
        dp.reserved[0] = DESKEW_AHEAD if ahead else 0
        n = C.c_int()
        self._chk(lib.liloc_deskew(self._h, _ptr(p), _ptr(rt), len(p), C.byref(dp), C.byref(n) if sync else None))
        return n.value if sync else None

    def deskew_advance(self) -> None:
        """The sweep staged with ahead=True becomes the current one."""
        self._chk(lib.liloc_deskew_advance(self._h))

    def deskewed_get(self) -> np.ndarray:
        n = C.c_int()
        self._chk(lib.liloc_deskewed_get(self._h, None, 0, C.byref(n)))
        out = np.empty((n.value, 4), np.float32)
        self._chk(lib.liloc_deskewed_get(self._h, _ptr(out), n.value, C.byref(n)))
        return out

    def last_deskew_ms(self) -> float:
        ms = C.c_float()
        self._chk(lib.liloc_last_deskew_ms(self._h, C.byref(ms)))
        return ms.value

    def last_deskew_marks(self) -> np.ndarray:
        out = np.zeros(3, np.uint64)
        self._chk(lib.liloc_last_deskew_marks(self._h, _ptr(out)))
        return out

    def scan_put_deskewed(self, leaf: float) -> int:
        Q = C.c_int()
        self._chk(lib.liloc_scan_put_class(self._h, 0, None, 0, SCAN_DESKEWED, leaf, C.byref(Q)))
        return Q.value

    def scan_ds_get(self, cls: int = 0) -> np.ndarray:
        Q = C.c_int()
        self._chk(lib.liloc_scan_ds_get_class(self._h, cls, None, 0, C.byref(Q)))
        out = np.empty((Q.value, 4), np.float32)
        self._chk(lib.liloc_scan_ds_get_class(self._h, cls, _ptr(out), Q.value, C.byref(Q)))
        return out

    # ---- registration ----
    def scan2map(self, pose6, opts: S2mOpts | None = None) -> tuple[np.ndarray, S2mResult, int]:
        pose = np.array(pose6, dtype=np.float32).reshape(6).copy()
        res = S2mResult()
        o = opts or S2mOpts()
        rc = self._chk(lib.liloc_scan2map(self._h, _ptr(pose), C.byref(o), C.byref(res)))
        return pose, res, rc

    def frame(self, kf_ids, poses6, map_leaf: float, scan, scan_leaf: float, pose6, opts: S2mOpts | None = None,
              scan_dev_ptr: int | None = None, n_scan: int | None = None) -> tuple[np.ndarray, S2mResult, int]:
        ids = np.ascontiguousarray(kf_ids, dtype=np.int32)
        poses = np.ascontiguousarray(poses6, dtype=np.float32).reshape(-1, 6)
        pose = np.array(pose6, dtype=np.float32).reshape(6).copy()
        res = S2mResult()
        o = opts or S2mOpts()
        if scan is None and scan_dev_ptr is None:           # the de-skewed sweep the context holds
            sp, n, on_dev = None, 0, SCAN_DESKEWED
        elif scan_dev_ptr is not None:
            sp, n, on_dev = C.c_void_p(scan_dev_ptr), int(n_scan), 1
        else:
            s = _cloud(scan)
            sp, n, on_dev = _ptr(s), len(s), 0
        rc = self._chk(lib.liloc_frame(self._h, _ptr(ids), _ptr(poses), len(ids), map_leaf, sp, n, on_dev, scan_leaf,
                                       _ptr(pose), C.byref(o), C.byref(res)))
        return pose, res, rc

    def scan_prefetch(self, scan: np.ndarray) -> None:
        """liloc_scan_prefetch: `scan` (a C-contiguous float32 (n, 4) array the caller keeps alive and unchanged) is the NEXT
        frame's surf scan; pass the very same array to that frame() call."""
        assert scan.dtype == np.float32 and scan.flags.c_contiguous and scan.ndim == 2 and scan.shape[1] == 4
        self._chk(lib.liloc_scan_prefetch(self._h, scan.ctypes.data, len(scan)))

    def frame_features(self, kf_ids, poses6, map_leafs, scans, scan_leafs, pose6, opts: S2mOpts | None = None,
                       scan_dev_ptrs=None, n_scans=None) -> tuple[np.ndarray, S2mResult, int]:
        """liloc_frame_features: scans / leaf sizes are (surf, corner) pairs."""
        ids = np.ascontiguousarray(kf_ids, dtype=np.int32)
        poses = np.ascontiguousarray(poses6, dtype=np.float32).reshape(-1, 6)
        pose = np.array(pose6, dtype=np.float32).reshape(6).copy()
        res = S2mResult()
        o = opts or S2mOpts.upstream()
        fi = FrameIn()
        fi.kf_ids, fi.kf_poses6, fi.K = ids.ctypes.data, poses.ctypes.data, len(ids)
        keep = []
        for c in range(2):
            fi.map_leaf[c], fi.scan_leaf[c] = float(map_leafs[c]), float(scan_leafs[c])
            if scan_dev_ptrs is not None:
                fi.scan[c], fi.n_scan[c] = int(scan_dev_ptrs[c]), int(n_scans[c])
            else:
                sc = _cloud(scans[c]); keep.append(sc)
                fi.scan[c], fi.n_scan[c] = sc.ctypes.data if len(sc) else None, len(sc)
        fi.scan_on_device = 1 if scan_dev_ptrs is not None else 0
        rc = self._chk(lib.liloc_frame_features(self._h, C.byref(fi), _ptr(pose), C.byref(o), C.byref(res)))
        return pose, res, rc

    def last_s2m_extra(self) -> dict:
        e = S2mExtra()
        lib.liloc_last_s2m_extra(self._h, C.byref(e))
        return {k: getattr(e, k) for k in ("q_corner", "map_points_corner", "tiles_corner", "tiles_surf", "sort_redos")}

    # ---- kernel-level ----
    def voxelgrid(self, pts, leaf: float) -> np.ndarray:
        p = _cloud(pts)
        out = np.empty_like(p)
        m = C.c_int()
        self._chk(lib.liloc_voxelgrid(self._h, _ptr(p), len(p), leaf, _ptr(out), len(p), C.byref(m)))
        return out[: m.value].copy()

    def knn5(self, queries_map, cls: int = 0) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        q = _cloud(queries_map)
        idx = np.empty((len(q), 5), np.int32)
        d2 = np.empty((len(q), 5), np.float32)
        found = np.empty(len(q), np.int32)
        self._chk(lib.liloc_knn5_class(self._h, cls, _ptr(q), len(q), _ptr(idx), _ptr(d2), _ptr(found)))
        return idx, d2, found

    def knn5_bounded(self, queries_map, bounds, cls: int = 0) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        q = _cloud(queries_map)
        b = np.ascontiguousarray(bounds, np.float32).reshape(len(q))
        idx = np.empty((len(q), 5), np.int32)
        d2 = np.empty((len(q), 5), np.float32)
        found = np.empty(len(q), np.int32)
        self._chk(lib.liloc_debug_knn5_bounded(self._h, cls, _ptr(q), len(q), _ptr(b), _ptr(idx), _ptr(d2), _ptr(found)))
        return idx, d2, found

    def feature_pass(self, cls: int, pose6, stride: int, q_all: int) -> tuple[np.ndarray, np.ndarray]:
        pose = np.ascontiguousarray(pose6, dtype=np.float32).reshape(6)
        coeff = np.zeros((q_all, 4), np.float32)
        flag = np.zeros(q_all, np.uint8)
        self._chk(lib.liloc_feature_pass(self._h, cls, _ptr(pose), stride, _ptr(coeff), _ptr(flag)))
        return coeff, flag

    def debug_eigen3(self, A):
        A = np.ascontiguousarray(A, np.float32).reshape(-1, 9)
        w, V = np.empty((len(A), 3), np.float32), np.empty((len(A), 9), np.float32)
        self._chk(lib.liloc_debug_eigen3(self._h, _ptr(A), len(A), _ptr(w), _ptr(V)))
        return w, V.reshape(-1, 3, 3)

    def debug_line(self, pts, sel):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 15)
        sel = np.ascontiguousarray(sel, np.float32).reshape(-1, 3)
        coeff, keep = np.empty((len(pts), 4), np.float32), np.empty(len(pts), np.int32)
        self._chk(lib.liloc_debug_line(self._h, _ptr(pts), _ptr(sel), len(pts), _ptr(coeff), _ptr(keep)))
        return coeff, keep

    def surf_pass(self, pose6, stride: int, q_all: int) -> tuple[np.ndarray, np.ndarray]:
        pose = np.ascontiguousarray(pose6, dtype=np.float32).reshape(6)
        coeff = np.zeros((q_all, 4), np.float32)
        flag = np.zeros(q_all, np.uint8)
        self._chk(lib.liloc_surf_pass(self._h, _ptr(pose), stride, _ptr(coeff), _ptr(flag)))
        return coeff, flag

    def pose_to_T(self, pose6) -> np.ndarray:
        pose = np.ascontiguousarray(pose6, dtype=np.float32).reshape(6)
        T = np.empty(12, np.float32)
        self._chk(lib.liloc_pose_to_T(self._h, _ptr(pose), _ptr(T)))
        return T.reshape(3, 4)

    # ---- NDT (relo mode) ----
    def ndt_target_put(self, target_id: int, pts_map, resolution: float) -> int:
        p = _cloud(pts_map)
        n = C.c_int()
        self._chk(lib.liloc_ndt_target_put(self._h, target_id, _ptr(p), len(p), resolution, C.byref(n)))
        return n.value

    def ndt_target_put_keyframes(self, target_id: int, kf_ids, poses6, resolution: float) -> tuple[int, int]:
        ids = np.ascontiguousarray(kf_ids, dtype=np.int32)
        poses = np.ascontiguousarray(poses6, dtype=np.float32).reshape(-1, 6)
        nl, npts = C.c_int(), C.c_int()
        self._chk(lib.liloc_ndt_target_put_keyframes(self._h, target_id, _ptr(ids), _ptr(poses), len(ids), resolution, C.byref(nl), C.byref(npts)))
        return nl.value, npts.value

    def ndt_source_put(self, source_id: int, pts_body=None, dev_ptr: int | None = None, n: int | None = None) -> None:
        if pts_body is None and dev_ptr is None:            # the de-skewed sweep the context holds
            self._chk(lib.liloc_ndt_source_put(self._h, source_id, None, 0, SCAN_DESKEWED))
        elif dev_ptr is not None:
            self._chk(lib.liloc_ndt_source_put(self._h, source_id, C.c_void_p(dev_ptr), int(n), 1))
        else:
            p = _cloud(pts_body)
            self._chk(lib.liloc_ndt_source_put(self._h, source_id, _ptr(p), len(p), 0))

    def globalmap_build(self, kf_ids, poses6, leaf: float) -> np.ndarray:
        ids = np.ascontiguousarray(kf_ids, dtype=np.int32)
        poses = np.ascontiguousarray(poses6, dtype=np.float32).reshape(-1, 6)
        M = C.c_int()
        self._chk(lib.liloc_globalmap_build(self._h, _ptr(ids), _ptr(poses), len(ids), leaf, None, 0, C.byref(M)))
        out = np.empty((max(M.value, 1), 4), np.float32)
        self._chk(lib.liloc_globalmap_build(self._h, _ptr(ids), _ptr(poses), len(ids), leaf, _ptr(out), len(out), C.byref(M)))
        return out[: M.value].copy()

    def icp_fitness(self, cloud1, pose6_1, cloud2, pose6_2) -> float:
        a, b = _cloud(cloud1), _cloud(cloud2)
        p1 = None if pose6_1 is None else np.ascontiguousarray(pose6_1, np.float32).reshape(6)
        p2 = None if pose6_2 is None else np.ascontiguousarray(pose6_2, np.float32).reshape(6)
        out = C.c_float()
        self._chk(lib.liloc_icp_fitness(self._h, _ptr(a), len(a), None if p1 is None else _ptr(p1), _ptr(b), len(b),
                                        None if p2 is None else _ptr(p2), C.byref(out)))
        return out.value

    def icp_fitness_resident(self, kf_id: int, pose6_kf, source_id: int, pose6_src) -> float:
        p1 = np.ascontiguousarray(pose6_kf, np.float32).reshape(6)
        p2 = np.ascontiguousarray(pose6_src, np.float32).reshape(6)
        out = C.c_float()
        self._chk(lib.liloc_icp_fitness_resident(self._h, kf_id, _ptr(p1), source_id, _ptr(p2), C.byref(out)))
        return out.value

    def ndt_stats(self, reset: bool = False) -> dict:
        st = NdtStats()
        lib.liloc_ndt_stats_get(self._h, C.byref(st))
        if reset:
            lib.liloc_ndt_stats_reset(self._h)
        return {k: (list(getattr(st, k)) if k == "update_cycles" else getattr(st, k)) for k, _ in NdtStats._fields_}

    def ndt_clear(self) -> None:
        self._chk(lib.liloc_ndt_clear(self._h))

    def ndt_align_batch(self, target_ids, source_ids, guesses, opts: NdtOpts | None = None):
        """guesses: (J, 4, 4) or (J, 3, 4) float32 row-major.  Returns (T (J,4,4), score, iterations, converged, sweeps)."""
        g = np.asarray(guesses, np.float32)
        J = len(g)
        jobs = np.zeros(J, _NDT_JOB_DT)                      # same layout as liloc_ndt_job (checked at import)
        jobs["target_id"] = np.asarray(target_ids, np.int32)
        jobs["source_id"] = np.asarray(source_ids, np.int32)
        m = np.tile(np.eye(4, dtype=np.float32), (J, 1, 1))
        m[:, : g.shape[1], :] = g
        jobs["guess"] = m.reshape(J, 16)
        outs = np.zeros(J, _NDT_OUT_DT)
        o = opts or NdtOpts()
        self._chk(lib.liloc_ndt_align_batch(self._h, C.cast(jobs.ctypes.data, C.POINTER(NdtJob)), J, C.byref(o),
                                            C.cast(outs.ctypes.data, C.POINTER(NdtOut))))
        return (outs["T"].reshape(J, 4, 4).copy(), outs["score"].copy(), outs["iterations"].copy(), outs["converged"].copy(), outs["sweeps"].copy())

    @staticmethod
    def _pack_jobs(target_ids, source_ids, guesses):
        g = np.asarray(guesses, np.float32)
        J = len(g)
        jobs = np.zeros(J, _NDT_JOB_DT)
        jobs["target_id"] = np.asarray(target_ids, np.int32)
        jobs["source_id"] = np.asarray(source_ids, np.int32)
        m = np.tile(np.eye(4, dtype=np.float32), (J, 1, 1))
        m[:, : g.shape[1], :] = g
        jobs["guess"] = m.reshape(J, 16)
        return jobs

    # ---- multi-GPU (liloc_comm_*, liloc_ndt_align_batch_sharded, liloc_gather_results) ----
    def comm_init(self, rank: int, world: int, unique_id: bytes | None) -> None:
        buf = C.create_string_buffer(unique_id or b"", COMM_ID_BYTES)
        self._chk(lib.liloc_comm_init(self._h, rank, world, buf if world > 1 else None))

    def comm_destroy(self) -> None:
        self._chk(lib.liloc_comm_destroy(self._h))

    def comm_world(self) -> tuple[int, int]:
        return lib.liloc_comm_rank(self._h), lib.liloc_comm_world(self._h)

    def ndt_align_batch_sharded(self, target_ids, source_ids, guesses, opts: NdtOpts | None = None, round_robin: bool = False, want_local: bool = False):
        """The FULL item list on every rank; this rank runs its share, one ncclAllGather of 8 floats per item.
        Returns rows (n_items, 8) = {roll, pitch, yaw, x, y, z, score / n_source, iterations} (+ this rank's liloc_ndt_out records)."""
        jobs = self._pack_jobs(target_ids, source_ids, guesses)
        n = len(jobs)
        rows = np.zeros((n, 8), np.float32)
        o = opts or NdtOpts()
        rank, world = self.comm_world()
        outs = np.zeros((n + world - 1) // world, _NDT_OUT_DT) if want_local else None
        self._chk(lib.liloc_ndt_align_batch_sharded(self._h, C.cast(jobs.ctypes.data, C.POINTER(NdtJob)), n, C.byref(o), SHARD_ROUND_ROBIN if round_robin else SHARD_BLOCKS,
                                                    _ptr(rows), C.cast(outs.ctypes.data, C.POINTER(NdtOut)) if want_local else None))
        return (rows, outs) if want_local else rows

    def gather_results(self, rows_local: np.ndarray, per_rank: int) -> np.ndarray:
        """All-gather of this rank's (n_local <= per_rank, 8) rows -> (world, per_rank, 8) on every rank."""
        r = np.ascontiguousarray(rows_local, np.float32).reshape(-1, 8)
        _, world = self.comm_world()
        out = np.zeros((world, per_rank, 8), np.float32)
        self._chk(lib.liloc_gather_results(self._h, _ptr(r) if len(r) else None, len(r), per_rank, _ptr(out)))
        return out

    def ndt_target_get(self, target_id: int):
        n = lib.liloc_ndt_target_get(self._h, target_id, None, None, None, None, 0)
        means, icovs = np.zeros((max(n, 1), 3)), np.zeros((max(n, 1), 9), np.float32)
        counts, keys = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32)
        self._chk(lib.liloc_ndt_target_get(self._h, target_id, _ptr(means), _ptr(icovs), _ptr(counts), _ptr(keys), n))
        o = np.argsort(keys[:n], kind="stable")
        return means[:n][o], icovs[:n][o], counts[:n][o], keys[:n][o]

    def ndt_derivatives(self, target_id: int, source_id: int, p6, opts: NdtOpts | None = None) -> np.ndarray:
        p = np.ascontiguousarray(p6, np.float64).reshape(6)
        out = np.zeros(28)
        o = opts or NdtOpts()
        self._chk(lib.liloc_ndt_derivatives(self._h, target_id, source_id, _ptr(p), C.byref(o), _ptr(out)))
        return out

    # ---- instrumentation ----
    def set_timing(self, on: bool) -> None:
        lib.liloc_set_timing(self._h, int(on))

    def last_timing(self) -> dict:
        t = Timing()
        lib.liloc_last_timing(self._h, C.byref(t))
        return {k: getattr(t, k) for k, _ in Timing._fields_}

    def debug_lm6(self, A, b):
        A = np.ascontiguousarray(A, np.float32).reshape(-1, 36)
        b = np.ascontiguousarray(b, np.float32).reshape(-1, 6)
        n = len(A)
        x, w = np.empty((n, 6), np.float32), np.empty((n, 6), np.float32)
        P, deg = np.empty((n, 36), np.float32), np.empty(n, np.int32)
        self._chk(lib.liloc_debug_lm6(self._h, _ptr(A), _ptr(b), n, _ptr(x), _ptr(w), _ptr(P), _ptr(deg)))
        return x, w, P.reshape(n, 6, 6), deg

    def debug_plane(self, pts):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 15)
        x = np.empty((len(pts), 3), np.float32)
        self._chk(lib.liloc_debug_plane(self._h, _ptr(pts), len(pts), _ptr(x)))
        return x

    def last_trace(self, iters: int = 32) -> np.ndarray:
        out = np.zeros((32, 16), np.float32)
        n = lib.liloc_last_trace(self._h, _ptr(out), iters)
        return out[:n]

    def set_share(self, share: int) -> int:
        """liloc_set_share: this context is one of `share` contexts driving the GPU concurrently; returns the previous value."""
        return int(lib.liloc_set_share(self._h, int(share)))

    def set_map_cache(self, on: bool) -> int:
        """Local-map memoisation on / off; returns the number of frames that have reused their map so far."""
        return int(lib.liloc_set_map_cache(self._h, int(on)))

    def last_marks(self) -> np.ndarray:
        """%globaltimer marks of the last timed frame (see include/liloc_gpu.h liloc_last_marks)."""
        out = np.zeros(2 * 16 + 2 + 3 * 32, np.uint64)           # [130] = end of the registration kernel
        n = lib.liloc_last_marks(self._h, _ptr(out), len(out))
        return out[:n]

    def last_clocks(self, iters: int = 32) -> np.ndarray:
        out = np.zeros((32, 8), np.int32)
        n = lib.liloc_last_clocks(self._h, _ptr(out), iters)
        return out[:n]

    def stats(self) -> dict:
        return stats_of(self._h.value)

    def stats_reset(self) -> None:
        lib.liloc_stats_reset(self._h)

    def kernel_launches(self) -> int:
        return int(lib.liloc_kernel_launches(self._h))
