"""Seeded synthetic worlds, sensors, trajectories and scans.

Inputs only -- nothing here is on the measured path.  The ray caster is native
(``libliloc_synth.so``, built from ``synth.cpp``); worlds and trajectories are numpy.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(os.path.dirname(_HERE), "libliloc_synth.so")


class _Sensor(C.Structure):
    _fields_ = [("pattern", C.c_int), ("n_rings", C.c_int), ("n_az", C.c_int),
                ("el_min_deg", C.c_float), ("el_max_deg", C.c_float),
                ("range_min", C.c_float), ("range_max", C.c_float), ("noise_sigma", C.c_float)]


_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise ImportError(f"{_LIB_PATH} missing: run `python -m liloc_b200._build`")
        _lib = C.CDLL(_LIB_PATH)
        _lib.synth_scan.restype = C.c_int
        _lib.synth_scan.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(_Sensor), C.c_uint64, C.c_void_p]
    return _lib


@dataclass(frozen=True)
class Sensor:
    name: str
    pattern: int            # 0 spinning, 1 random
    n_rings: int
    n_az: int
    el_min_deg: float
    el_max_deg: float
    range_min: float
    range_max: float
    noise_sigma: float = 0.02

    @property
    def max_points(self) -> int:
        return self.n_rings * self.n_az if self.pattern == 0 else self.n_az


# config/lio_sam_default.yaml: N_SCAN 16, Horizon_SCAN 1800, range 2..100 m
VLP16 = Sensor("vlp16", 0, 16, 1800, -15.0, 15.0, 2.0, 100.0)
# config/nclt.yaml: N_SCAN 32, Horizon_SCAN 1800, range 1..100 m
HDL32 = Sensor("hdl32e", 0, 32, 1800, -30.67, 10.67, 1.0, 100.0)
# config/lio_sam_mid360.yaml: non-repetitive pattern approximated by 20 000 uniform rays / frame
MID360 = Sensor("mid360", 1, 1, 20000, -7.0, 52.0, 0.5, 100.0)


def make_world(size_xy=(60.0, 40.0), height=12.0, n_pillars=12, seed=0) -> np.ndarray:
    """Box room centred on the origin with the floor at z=0, plus box pillars.  (n_boxes, 6) float32."""
    rng = np.random.default_rng(seed)
    sx, sy = size_xy
    boxes = [[-sx / 2, -sy / 2, 0.0, sx / 2, sy / 2, height]]
    for _ in range(n_pillars):
        w, d = rng.uniform(0.6, 2.5, 2)
        h = rng.uniform(2.0, height)
        cx = rng.uniform(-sx / 2 + 3, sx / 2 - 3)
        cy = rng.uniform(-sy / 2 + 3, sy / 2 - 3)
        if abs(cy) < 2.5:                       # keep a corridor for the trajectory
            cy = np.sign(cy if cy != 0 else 1.0) * (2.5 + abs(cy))
        boxes.append([cx - w / 2, cy - d / 2, 0.0, cx + w / 2, cy + d / 2, h])
    return np.asarray(boxes, dtype=np.float32)


def trajectory_line(n: int, step=1.0, x0=-25.0, height=1.8) -> np.ndarray:
    """advance along +x with a 0.5*sin(k/8) lateral weave and 0.02 rad/step yaw drift.
    Returns (n, 6) float64 [roll, pitch, yaw, x, y, z]."""
    k = np.arange(n, dtype=np.float64)
    pose = np.zeros((n, 6))
    pose[:, 0] = 0.01 * np.sin(k / 5.0)
    pose[:, 1] = 0.01 * np.cos(k / 7.0)
    pose[:, 2] = 0.02 * k * np.cos(k / 20.0) * 0.3
    pose[:, 3] = x0 + step * k
    pose[:, 4] = 0.5 * np.sin(k / 8.0)
    pose[:, 5] = height + 0.02 * np.sin(k / 3.0)
    return pose


def trajectory_loop(n: int, radius_xy=(35.0, 20.0), laps=1.0, height=1.8) -> np.ndarray:
    """Closed elliptical loop (cfg 2: 1000 frames at 10 Hz along a loop inside a 120 x 80 m world)."""
    s = np.linspace(0.0, 2 * np.pi * laps, n, endpoint=False)
    pose = np.zeros((n, 6))
    pose[:, 3] = radius_xy[0] * np.cos(s)
    pose[:, 4] = radius_xy[1] * np.sin(s)
    dx = -radius_xy[0] * np.sin(s)
    dy = radius_xy[1] * np.cos(s)
    pose[:, 2] = np.arctan2(dy, dx)
    pose[:, 0] = 0.01 * np.sin(7 * s)
    pose[:, 1] = 0.01 * np.cos(5 * s)
    pose[:, 5] = height + 0.03 * np.sin(11 * s)
    return pose


def scan(world: np.ndarray, pose6, sensor: Sensor, seed: int) -> np.ndarray:
    """One body-frame scan (n, 4) float32 = cloud_deskewed."""
    lib = _load()
    w = np.ascontiguousarray(world, dtype=np.float32)
    p = np.ascontiguousarray(pose6, dtype=np.float32).reshape(6)
    out = np.empty((sensor.max_points, 4), np.float32)
    s = _Sensor(sensor.pattern, sensor.n_rings, sensor.n_az, sensor.el_min_deg, sensor.el_max_deg,
                sensor.range_min, sensor.range_max, sensor.noise_sigma)
    n = lib.synth_scan(w.ctypes.data, len(w), p.ctypes.data, C.byref(s), C.c_uint64(seed), out.ctypes.data)
    return out[:n].copy()


def split_features(world: np.ndarray, pose6, scan_body: np.ndarray, tol: float = 0.12):
    """Split a body-frame scan into (corner, surf) feature clouds the way an upstream LIO-SAM front end would
    hand them to mapOptimization: a return is a *corner* feature when it lies within `tol` of a vertical edge
    of the world (room corners, pillar edges), everything else is a *surf* feature.  (The reference has no
    feature extraction -- SURVEY.md 0.2 D2 -- so this exists only for the north_star's corner superset.)"""
    p = np.asarray(pose6, np.float64)
    R = rot_zyx(p[0], p[1], p[2])
    w = scan_body[:, :3].astype(np.float64) @ R.T + p[3:6]
    best = np.full(len(w), np.inf)
    for b in np.asarray(world, np.float64):
        for ex in (b[0], b[3]):
            for ey in (b[1], b[4]):
                d = np.hypot(w[:, 0] - ex, w[:, 1] - ey)
                d = np.where((w[:, 2] >= b[2] - tol) & (w[:, 2] <= b[5] + tol), d, np.inf)
                best = np.minimum(best, d)
    is_corner = best < tol
    return np.ascontiguousarray(scan_body[is_corner]), np.ascontiguousarray(scan_body[~is_corner])


def transform_cloud(pts: np.ndarray, pose6) -> np.ndarray:
    """Body-frame cloud -> map frame by the pose [roll, pitch, yaw, x, y, z] (float32 result; used to build synthetic
    prior-session submaps, i.e. inputs -- the product's own transform runs on the device)."""
    p = np.asarray(pose6, np.float64)
    R = rot_zyx(p[0], p[1], p[2]).astype(np.float32)
    out = np.empty_like(pts, dtype=np.float32)
    out[:, :3] = pts[:, :3].astype(np.float32) @ R.T + p[3:6].astype(np.float32)
    out[:, 3] = pts[:, 3]
    return out


def raw_sweep(scan_body: np.ndarray, sensor: Sensor, gyro=(0.05, -0.03, 0.4), imu_rate: float = 500.0, sweep_s: float = 0.1,
              time_scan_cur: float = 1000.0, seed: int = 0, outliers: float = 0.02) -> dict:
    """A raw (skewed) sweep as the LiDAR driver would publish it, built from a motion-free body-frame scan: every point
    gets a ring (elevation bin) and a relative time (azimuth fraction of the sweep; points are ordered by time), and
    is rotated by the inverse of the motion the sensor made since the start of the sweep under the body rate `gyro`
    (rad/s).  A few points are made invalid (ring out of range, range below the minimum) so that the filters of
    projectPointCloud have something to drop.  Also returns the IMU samples covering the sweep.  Inputs only: the
    de-skew itself runs on the device."""
    rng = np.random.default_rng(seed)
    p = np.asarray(scan_body, np.float64)
    az = np.arctan2(p[:, 1], p[:, 0])
    t = (az + np.pi) / (2 * np.pi) * sweep_s
    order = np.argsort(t, kind="stable")
    p, t = p[order], t[order]
    el = np.degrees(np.arctan2(p[:, 2], np.hypot(p[:, 0], p[:, 1])))
    ring = np.clip(np.round((el - sensor.el_min_deg) / max(sensor.el_max_deg - sensor.el_min_deg, 1e-9) * (sensor.n_rings - 1)), 0,
                   sensor.n_rings - 1).astype(np.int32)
    w = np.asarray(gyro, np.float64)
    n_imu = int(np.ceil((sweep_s + 0.03) * imu_rate)) + 1
    stamps = time_scan_cur - 0.008 + np.arange(n_imu) / imu_rate
    gyro_s = np.tile(w, (n_imu, 1)) + rng.normal(0, 1e-3, (n_imu, 3))
    # the motion the reference assumes: integrated angles -> Rz Ry Rx; raw = R(t)^-1 R(0) p
    ang0 = w * 0.008
    R0 = rot_zyx(*ang0)
    raw = np.empty((len(p), 4), np.float32)
    ang = ang0[None, :] + w[None, :] * t[:, None]
    Rt = np.moveaxis(rot_zyx(ang[:, 0], ang[:, 1], ang[:, 2]), -1, 0)        # (n, 3, 3)
    raw[:, :3] = np.einsum("nji,nj->ni", Rt, p[:, :3] @ R0.T)               # R(t)^T (R0 p)
    raw[:, 3] = p[:, 3]
    bad = rng.random(len(p)) < outliers
    ring = np.where(bad & (rng.random(len(p)) < 0.5), sensor.n_rings + 3, ring).astype(np.int32)
    near = bad & (ring < sensor.n_rings)
    raw[near, :3] *= (0.3 / np.maximum(np.linalg.norm(raw[near, :3], axis=1), 1e-6))[:, None].astype(np.float32)
    return {"pts": raw, "ring": ring, "time": t.astype(np.float32), "imu_stamps": stamps, "imu_gyro": gyro_s,
            "time_scan_cur": time_scan_cur, "time_scan_end": time_scan_cur + float(np.float32(t[-1])) if len(t) else time_scan_cur,
            "truth": p.astype(np.float32)}


def imu_integrate(stamps, gyro, time_scan_end: float, queue_length: int = 2000):
    """imuDeskewInfo (src/imageProjection.cpp:441-496) for an already-trimmed IMU queue: integrated angles per sample
    (double, sequential).  Returns (imu_time (m,), imu_rot (m, 3)); imuPointerCur = m - 1."""
    tt, rr = [], []
    for i in range(len(stamps)):
        t = float(stamps[i])
        if t > time_scan_end + 0.01 or len(tt) >= queue_length:
            break
        if not tt:
            tt.append(t); rr.append((0.0, 0.0, 0.0)); continue
        dt = t - tt[-1]
        rr.append(tuple(rr[-1][k] + float(gyro[i][k]) * dt for k in range(3)))
        tt.append(t)
    return np.asarray(tt, np.float64), np.asarray(rr, np.float64).reshape(-1, 3)


def perturb(pose6, seed: int, dt=0.10, dr=0.02) -> np.ndarray:
    """Initial guess = ground truth + U(+-dt m, +-dr rad) per axis."""
    rng = np.random.default_rng(seed)
    p = np.asarray(pose6, dtype=np.float64).copy()
    p[:3] += rng.uniform(-dr, dr, 3)
    p[3:] += rng.uniform(-dt, dt, 3)
    return p.astype(np.float32)


def rot_zyx(roll, pitch, yaw) -> np.ndarray:
    cr, sr, cp, sp, cy, sy = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, sy * sr + cy * sp * cr],
                     [sy * cp, cy * cr + sy * sp * sr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


@dataclass
class Sequence:
    """A synthetic lio sequence: what three ROS nodes would deliver to mapOptimization frame by frame."""
    sensor: Sensor
    world: np.ndarray
    truth: np.ndarray          # (F, 6) float64 ground-truth sensor poses
    odom: np.ndarray           # (F, 6) float32 'IMU pre-integration' poses (cloud_info.initialGuess*)
    stamps: np.ndarray         # (F,) seconds
    scans: list                # F body-frame clouds (n_i, 4) float32


def make_sequence(sensor: Sensor, n_frames: int, seed: int, world_size=(120.0, 80.0), n_pillars=40,
                  loop_radius=(45.0, 25.0), dt=0.2, odom_sigma=(0.01, 0.002), start_frac=0.0, n_total=None) -> Sequence:
    """cfg 2 of SURVEY 8d: frames along a closed loop inside a box world with pillars; `dt` is the stamp
    spacing (use the yaml's timeInterval so that every frame passes the rate gate, :257)."""
    world = make_world(size_xy=world_size, height=12.0, n_pillars=n_pillars, seed=seed)
    # keep pillars off the path
    keep = [0]
    for b in range(1, len(world)):
        cx, cy = 0.5 * (world[b, 0] + world[b, 3]), 0.5 * (world[b, 1] + world[b, 4])
        rr = (cx / loop_radius[0]) ** 2 + (cy / loop_radius[1]) ** 2
        if abs(np.sqrt(rr) - 1.0) > 0.12:
            keep.append(b)
    world = world[keep]
    n_total = n_frames if n_total is None else max(n_total, n_frames)   # trajectory spacing of an n_total-frame loop
    truth = trajectory_loop(n_total, radius_xy=loop_radius)
    if start_frac:
        truth = np.roll(truth, -int(start_frac * n_total), axis=0)
    truth = truth[:n_frames]
    rng = np.random.default_rng(seed + 77)
    odom = truth.copy()
    odom[:, 3:] += rng.normal(0, odom_sigma[0], (n_frames, 3))
    odom[:, :3] += rng.normal(0, odom_sigma[1], (n_frames, 3))
    scans = [scan(world, truth[k], sensor, seed * 100003 + k) for k in range(n_frames)]
    return Sequence(sensor, world, truth, odom.astype(np.float32), np.arange(n_frames) * dt + 100.0, scans)
