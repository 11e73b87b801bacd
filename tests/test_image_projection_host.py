"""CPU tests of the host side of class ImageProjection (liloc_b200/host/image_projection.cpp; src/imageProjection.cpp):
message queues, imuDeskewInfo, odomDeskewInfo, the azimuth timing of clouds without ring / time channels and the Livox
conversion, against the oracle's restatement and independent numpy / scipy evaluations.  No device: the de-skew launch
itself is covered by tests/test_gpu_deskew.py."""
import os

import numpy as np
import pytest

YAML = "lio_sam_default.yaml"          # velodyne, N_SCAN 16, 6-axis IMU, extrinsicRot != identity, imuRate 500


def _quat_from_rpy(r, p, y):          # tf::Quaternion::setRPY
    cr, sr, cp, sp, cy, sy = np.cos(r / 2), np.sin(r / 2), np.cos(p / 2), np.sin(p / 2), np.cos(y / 2), np.sin(y / 2)
    return (sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy)


def _ext_rot(yaml_path):
    import yaml as pyyaml
    cfg = pyyaml.safe_load(open(yaml_path))
    return np.array(cfg["Sensors"]["extrinsicRot"], np.float64).reshape(3, 3), np.array(cfg["Sensors"]["extrinsicRPY"], np.float64).reshape(3, 3)


def _cloud(n, seed, t_end=0.0995):
    rng = np.random.default_rng(seed)
    pts = (rng.normal(size=(n, 4)) * np.array([20, 20, 2, 30])).astype(np.float32)
    ring = rng.integers(0, 16, n).astype(np.int32)
    tm = np.sort(rng.uniform(0, t_end, n)).astype(np.float32)
    return pts, ring, tm


def _temp_yaml(tmp_path, **sensor_overrides):
    import yaml as pyyaml
    from liloc_b200 import host_api
    cfg = pyyaml.safe_load(open(host_api.config_path(YAML)))
    cfg["Sensors"].update(sensor_overrides)
    p = tmp_path / "ip.yaml"
    with open(p, "w") as f:
        for sec, kv in cfg.items():
            f.write(f"{sec}:\n")
            for k, v in kv.items():
                f.write(f"  {k}: {'true' if v is True else 'false' if v is False else v}\n")
    return str(p)


def test_queue_imu_integration_and_odometry(orc):
    from liloc_b200 import host_api
    ip = host_api.ImageProjection(YAML, None)
    R, _ = _ext_rot(host_api.config_path(YAML))
    rng = np.random.default_rng(1)
    stamps = 100.0 + 0.2 * np.arange(5)
    clouds = [_cloud(3000 + 100 * k, 10 + k) for k in range(5)]
    gyro = {}
    for k in range(5):
        t_imu = stamps[k] - 0.008 + np.arange(66) * 0.002
        g = rng.normal(0, 0.3, (66, 3))
        gyro[k] = (t_imu, g)
        for t, gg in zip(t_imu, g):
            ip.imu(t, gg)
        ip.odom(stamps[k] - 0.005, (1.0 + k, 2.0, 0.5), _quat_from_rpy(0.01, -0.02, 0.3 + 0.1 * k), covariance0=3.0)
        ip.odom(stamps[k], (1.5 + k, 2.0, 0.5), _quat_from_rpy(0.01, -0.02, 0.3 + 0.1 * k), covariance0=3.0)
        pts, ring, tm = clouds[k]
        info = ip.cloud(stamps[k], pts, ring, tm, has_time=True)
        if k < 2:
            assert info is None                                    # cachePointCloud keeps two sweeps back (:271-273)
            continue
        j = k - 2                                                  # the sweep that was just processed
        assert info is not None and info.stamp == stamps[j] and info.cloud_on_device == 2
        assert info.imuAvailable == 1 and info.odomAvailable == 1
        # odomDeskewInfo (:498-571): the first odometry message at or after the sweep's start
        assert abs(info.initialGuessX - (1.5 + j)) < 1e-6 and abs(info.initialGuessY - 2.0) < 1e-6
        assert abs(info.initialGuessRoll - 0.01) < 1e-6 and abs(info.initialGuessPitch + 0.02) < 1e-6 and abs(info.initialGuessYaw - (0.3 + 0.1 * j)) < 1e-6
        t_cur, t_end, flag = ip.times()
        pj, rj, tj = clouds[j]
        assert t_cur == stamps[j] and flag == 1 and t_end == stamps[j] + float(tj[-1])
        # imuDeskewInfo (:441-496) on the gyro rotated by the extrinsics (imuConverter, include/utility.h:306-320)
        imu_t, imu_r, lp, lr, lt = ip.last_sweep()
        ts, g = gyro[j]
        g_rot = np.array([[R[i, 0] * v[0] + R[i, 1] * v[1] + R[i, 2] * v[2] for i in range(3)] for v in g])
        et, er = orc.imu_integrate(ts, g_rot, t_end)
        assert np.array_equal(imu_t, et) and np.array_equal(imu_r, er)
        assert imu_t[0] <= t_cur and imu_t[-1] <= t_end + 0.01 and len(imu_t) > 50
        # the sweep handed to the de-skew call is the queued cloud, untouched
        assert np.array_equal(lp, pj) and np.array_equal(lr, rj) and np.array_equal(lt, tj)
    ip.close()


def test_waits_for_imu_and_odometry_gates(orc):
    from liloc_b200 import host_api
    ip = host_api.ImageProjection(YAML, None)
    clouds = [_cloud(500, 30 + k) for k in range(4)]
    # no IMU at all: nothing is published ("Waiting for IMU data ...", :429-432)
    for k in range(3):
        assert ip.cloud(10.0 + 0.2 * k, *clouds[k]) is None
    # IMU that starts after the sweep: still waiting
    for t in 10.05 + np.arange(100) * 0.002:
        ip.imu(t, (0, 0, 0.1))
    assert ip.cloud(10.6, *clouds[3]) is None
    ip.close()
    # IMU covers the sweep, odometry starts after it: published with odomAvailable = 0 (:509-510)
    ip = host_api.ImageProjection(YAML, None)
    for t in 19.99 + np.arange(400) * 0.002:
        ip.imu(t, (0, 0, 0.1))
    ip.odom(20.05, (1, 2, 3))
    infos = [ip.cloud(20.0 + 0.2 * k, *clouds[k]) for k in range(3)]
    assert infos[0] is None and infos[1] is None and infos[2] is not None
    assert infos[2].imuAvailable == 1 and infos[2].odomAvailable == 0
    # a single IMU sample inside the window: imuPointerCur == 0 -> imuAvailable stays false (:492-495)
    ip.close()


def test_nine_axis_imu_attitude(orc):
    """imuType 1: imuRollInit / PitchInit / YawInit from the last orientation at or before the sweep's start, after the
    extrinsic quaternion (imuConverter :321-328, imuRPY2rosRPY :376-388) -- against scipy."""
    from scipy.spatial.transform import Rotation as Rot
    import tempfile, pathlib
    from liloc_b200 import host_api
    tmp = pathlib.Path(tempfile.mkdtemp())
    y = _temp_yaml(tmp, imuType=1)
    _, ext_rpy = _ext_rot(host_api.config_path(YAML))
    ip = host_api.ImageProjection(y, None)
    q_in = Rot.from_euler("xyz", [0.1, -0.05, 0.7])
    for t in 29.99 + np.arange(400) * 0.002:
        ip.imu(t, (0, 0, 0.0), quat=tuple(q_in.as_quat()))
    clouds = [_cloud(400, 50 + k) for k in range(3)]
    info = [ip.cloud(30.0 + 0.2 * k, *clouds[k]) for k in range(3)][2]
    assert info is not None
    q_ext_inv = Rot.from_matrix(ext_rpy).inv()
    want = (q_in * q_ext_inv).as_euler("xyz")          # roll, pitch, yaw of q_from * extQRPY
    got = np.array([info.imuRollInit, info.imuPitchInit, info.imuYawInit])
    assert np.abs(got - want).max() < 1e-6, (got, want)
    ip.close()


def test_clouds_without_ring_and_time_channels(orc, tmp_path):
    from liloc_b200 import host_api
    y = _temp_yaml(tmp_path, have_ring_time_channel=False)
    ip = host_api.ImageProjection(y, None)
    for t in 39.99 + np.arange(400) * 0.002:
        ip.imu(t, (0.01, 0.0, 0.2))
    n = 3600
    az = -np.linspace(0.03, 2 * np.pi - 0.03, n)
    clouds = []
    for k in range(3):
        rad = 10 + 5 * np.sin(3 * az + k)
        clouds.append(np.stack([rad * np.cos(az), rad * np.sin(az), 0.1 * np.cos(az), np.ones(n)], axis=1).astype(np.float32))
    info = [ip.cloud(40.0 + 0.2 * k, clouds[k], None, None, has_time=False) for k in range(3)][2]
    assert info is not None
    _, _, lp, lr, lt = ip.last_sweep()
    assert np.array_equal(lp, clouds[0]) and np.all(lr == 0)
    assert np.array_equal(lt, orc.orientation_times(clouds[0]))
    t_cur, t_end, flag = ip.times()
    assert flag == 1 and t_end == t_cur + float(lt[-1])
    ip.close()


def test_livox_conversion(orc, tmp_path):
    from liloc_b200 import host_api
    ip = host_api.ImageProjection("lio_sam_mid360.yaml", None)
    rng = np.random.default_rng(7)
    msgs = []
    for f in range(3):
        n = 4000
        lp = np.zeros(n, host_api.LIVOX_POINT)
        xyz = rng.normal(size=(n, 3)) * np.array([10, 10, 2])
        lp["x"], lp["y"], lp["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
        lp["line"] = rng.integers(0, 10, n)
        lp["tag"] = rng.choice([0x00, 0x10, 0x20, 0x30, 0x01], n)
        lp["offset_time"] = rng.permutation(n).astype(np.uint32) * 20000
        lp["y"][7] = lp["y"][6]
        lp["x"][100] = np.nan
        stamp = 60.0 + 0.1 * f
        msgs.append((stamp, lp))
        for k in range(60):
            ip.imu(stamp - 0.008 + k * 0.002, (0.0, 0.01, 0.3))
        assert (ip.cloud_livox(stamp, lp) is None) == (f < 2)
    stamp, lp = msgs[0]
    keep = np.ones(len(lp), bool); keep[0] = False
    keep &= (lp["line"] < 8) & (((lp["tag"] & 0x30) == 0x10) | ((lp["tag"] & 0x30) == 0x00))
    keep &= ~(np.isnan(lp["x"]) | np.isnan(lp["y"]) | np.isnan(lp["z"]))
    near = np.zeros(len(lp), bool)
    with np.errstate(invalid="ignore"):
        for c in "xyz":
            near[1:] |= np.abs(lp[c][1:] - lp[c][:-1]) < 1e-7
    keep &= ~near
    tms = (lp["offset_time"].astype(np.float64) * 1e-6).astype(np.float32)[keep]
    order = np.argsort(tms, kind="stable")
    want_pts = np.stack([lp["x"][keep], lp["y"][keep], lp["z"][keep], np.zeros(keep.sum(), np.float32)], axis=1)[order]
    _, _, gp, gr, gt = ip.last_sweep()
    assert np.array_equal(gp, want_pts) and np.array_equal(gr, lp["line"][keep][order].astype(np.int32)) and np.array_equal(gt, tms[order])
    t_cur, t_end, flag = ip.times()
    assert t_cur == stamp and flag == 1 and t_end == stamp + float(tms[order][-1]) / 1000.0
    ip.close()
