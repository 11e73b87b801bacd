"""CPU tests of the host side (liloc_b200/host): the ROS-free `mapOptimization` class keeps the
reference's parameters and host-resident logic.  No device work here:
the objects are created with device=-1."""
import json
import os

import numpy as np
import pytest

from liloc_b200 import host_api

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

NUMERIC = {
    "System/numberOfCores": "numberOfCores", "Sensors/N_SCAN": "N_SCAN", "Sensors/Horizon_SCAN": "Horizon_SCAN",
    "Sensors/downsampleRate": "downsampleRate", "Sensors/point_filter_num": "point_filter_num",
    "Sensors/lidarMinRange": "lidarMinRange", "Sensors/lidarMaxRange": "lidarMaxRange", "Sensors/imuType": "imuType",
    "Sensors/imuRate": "imuRate", "Sensors/imuAccNoise": "imuAccNoise", "Sensors/imuGyrNoise": "imuGyrNoise",
    "Sensors/imuAccBiasN": "imuAccBiasN", "Sensors/imuGyrBiasN": "imuGyrBiasN", "Sensors/imuGravity": "imuGravity",
    "Sensors/imuRPYWeight": "imuRPYWeight", "Mapping/z_tollerance": "z_tollerance",
    "Mapping/rotation_tollerance": "rotation_tollerance", "Mapping/ndtResolution": "ndtResolution",
    "Mapping/ndtEpsilon": "ndtEpsilon", "Mapping/timeInterval": "timeInterval",
    "Mapping/mappingProcessInterval": "mappingProcessInterval", "Mapping/mappingSurfLeafSize": "mappingSurfLeafSize",
    "Mapping/surroundingKeyframeMapLeafSize": "surroundingKeyframeMapLeafSize",
    "Mapping/surroundingkeyframeAddingDistThreshold": "surroundingkeyframeAddingDistThreshold",
    "Mapping/surroundingkeyframeAddingAngleThreshold": "surroundingkeyframeAddingAngleThreshold",
    "Mapping/surroundingKeyframeDensity": "surroundingKeyframeDensity",
    "Mapping/surroundingKeyframeSearchRadius": "surroundingKeyframeSearchRadius",
    "Mapping/globalMapVisualizationSearchRadius": "globalMapVisualizationSearchRadius",
    "Mapping/globalMapVisualizationPoseDensity": "globalMapVisualizationPoseDensity",
    "Mapping/globalMapVisualizationLeafSize": "globalMapVisualizationLeafSize",
}


@pytest.mark.parametrize("name", ["lio_sam_default.yaml", "nclt.yaml", "m2dgr.yaml", "lio_sam_mid360.yaml"])
def test_yaml_parameters_equal_the_reference_files(name):
    """config/<name> parsed by our ParamServer == the reference's config/<name> (fixture made from
    /root/reference/config by tests/golden/make_golden.py)."""
    ref = json.load(open(os.path.join(GOLD, "reference_params.json")))[name]
    for key, member in NUMERIC.items():
        assert host_api.yaml_param(name, member) == pytest.approx(float(ref[key]), rel=1e-6), key
    assert host_api.yaml_param(name, "mode") == (0 if ref["System/mode"] == "lio" else 1)
    assert host_api.yaml_param(name, "regMethod.DIRECT1") == (1 if ref["Mapping/regMethod"] == "DIRECT1" else 0)
    assert host_api.yaml_param(name, "have_ring_time_channel") == float(bool(ref["Sensors/have_ring_time_channel"]))
    assert host_api.yaml_param(name, "correct") == float(bool(ref["Sensors/correct"]))
    sensors = {"velodyne": 0, "ouster": 1, "livox": 2, "robosense": 3, "mulran": 4}
    assert host_api.yaml_param(name, "sensor") == sensors[ref["Sensors/sensor"]]
    for lst, n in (("extrinsicRot", 9), ("extrinsicRPY", 9), ("extrinsicTrans", 3)):
        for i in range(n):
            assert host_api.yaml_param(name, f"{lst}.{i}") == pytest.approx(float(ref[f"Sensors/{lst}"][i]))


def test_parameter_defaults_and_multiline_lists(tmp_path):
    """Defaults of include/utility.h:195-304 apply to missing keys; multi-line inline lists parse."""
    y = tmp_path / "p.yaml"
    y.write_text('System:\n  mode: "lio"\nSensors:\n  sensor: ouster   # comment\n  extrinsicRot: [ 1, 0, 0,\n                  0, 1, 0,\n                  0, 0, 1 ]\nMapping:\n  timeInterval: 0.3\n')
    p = lambda k: host_api.yaml_param(str(y), k)
    assert p("numberOfCores") == 4 and p("N_SCAN") == 16 and p("Horizon_SCAN") == 1800 and p("point_filter_num") == 3
    assert p("lidarMinRange") == 1.0 and p("lidarMaxRange") == 1000.0 and p("imuRPYWeight") == pytest.approx(0.01)
    assert p("mappingSurfLeafSize") == pytest.approx(0.2) and p("surroundingKeyframeMapLeafSize") == pytest.approx(0.4)
    assert p("surroundingKeyframeSearchRadius") == 50.0 and p("surroundingKeyframeDensity") == 1.0
    assert p("timeInterval") == pytest.approx(0.3) and p("ndtEpsilon") == pytest.approx(0.01) and p("sensor") == 1
    assert p("z_tollerance") > 1e38 and p("extrinsicRot.8") == 1.0 and p("extrinsicRot.1") == 0.0
    y.write_text('System:\n  mode: "slam"\nSensors:\n  sensor: velodyne\n')
    with pytest.raises(KeyError):
        host_api.yaml_param(str(y), "N_SCAN")              # invalid mode -> the reference shuts down (:207-211)


def _host(yaml="lio_sam_default.yaml", mode="lio"):
    return host_api.MapOptimization(yaml, mode, device=-1)


def _rot(r, p, y):
    from liloc_b200.synth import rot_zyx
    return rot_zyx(r, p, y)


def test_voxelgrid_small_equals_oracle(orc):
    rng = np.random.default_rng(0)
    pts = (rng.normal(size=(300, 4)) * np.array([20, 20, 0.5, 100])).astype(np.float32)
    out = np.empty_like(pts)
    n = host_api.lib().liloc_host_debug_voxelgrid_small(pts.ctypes.data, len(pts), 2.0, out.ctypes.data)
    assert np.array_equal(out[:n], orc.voxelgrid(pts, 2.0))


def _extract_nearby_numpy(orc, keyposes, times, stamp, R, density):
    """numpy restatement of extractNearby + the selection part of extractCloud (src/mapOptmization.cpp:902-952)."""
    P = keyposes[:, 3:6].astype(np.float32)
    last = P[-1]
    d2 = ((last[0] - P[:, 0]) ** 2 + (last[1] - P[:, 1]) ** 2) + (last[2] - P[:, 2]) ** 2
    hit = np.flatnonzero(d2 < np.float32(R) * np.float32(R))
    hit = hit[np.lexsort((hit, d2[hit]))]
    cloud = np.c_[P[hit], hit.astype(np.float32)].astype(np.float32)
    ds = orc.voxelgrid(cloud, density)
    ids = []
    for c in ds:
        dd = ((c[0] - P[:, 0]) ** 2 + (c[1] - P[:, 1]) ** 2) + (c[2] - P[:, 2]) ** 2
        ids.append((int(np.argmin(dd)), c[:3]))
    sel = [(i, P[i]) for i, _ in ids]
    sel_pos = [c for _, c in ids]
    for i in range(len(P) - 1, -1, -1):
        if stamp - times[i] < 5.0:
            sel.append((i, P[i])); sel_pos.append(P[i])
        else:
            break
    out = []
    for (i, _), pos in zip(sel, sel_pos):
        if np.sqrt(np.float32((pos[0] - last[0]) ** 2 + (pos[1] - last[1]) ** 2 + (pos[2] - last[2]) ** 2)) > R:
            continue
        out.append(i)
    return np.asarray(out, np.int32)


@pytest.mark.parametrize("yaml,n,step", [("lio_sam_default.yaml", 60, 0.8), ("nclt.yaml", 120, 1.0), ("lio_sam_mid360.yaml", 150, 0.5)])
def test_extract_nearby_selection(orc, synth, yaml, n, step):
    traj = synth.trajectory_loop(n, radius_xy=(step * n / 6.3, step * n / 9.0)).astype(np.float32)
    times = np.arange(n) * 0.7
    with _host(yaml) as m:
        for k in range(n):
            m.debug_add_keypose(traj[k], times[k])
        stamp = times[-1] + 0.2
        got = m.debug_extract_nearby(stamp)
        R, dens = m.param("surroundingKeyframeSearchRadius"), m.param("surroundingKeyframeDensity")
    want = _extract_nearby_numpy(orc, traj, times, stamp, R, dens)
    assert np.array_equal(got, want)
    assert len(got) > 3 and len(set(got.tolist())) < len(got)          # the "last 5 s" append duplicates ids (:924-931)


def test_update_initial_guess_odometry_increment(synth):
    """odomAvailable branch (:865-883): new guess = previous estimate * (lastOdom^-1 * odom)."""
    with _host() as m:
        m.debug_add_keypose(np.zeros(6, np.float32), 0.0)
        est = np.array([0.01, -0.02, 0.5, 3.0, -1.0, 0.2], np.float32)
        odom0 = np.array([0.0, 0.01, 0.45, 2.9, -1.1, 0.25], np.float32)
        odom1 = np.array([0.005, 0.012, 0.52, 3.6, -0.7, 0.22], np.float32)

        def info(o):
            ci = host_api.CloudInfo()
            ci.odomAvailable = 1
            ci.initialGuessRoll, ci.initialGuessPitch, ci.initialGuessYaw = [float(v) for v in o[:3]]
            ci.initialGuessX, ci.initialGuessY, ci.initialGuessZ = [float(v) for v in o[3:]]
            return ci
        m.debug_set_pose(est)
        assert np.array_equal(m.debug_update_initial_guess(info(odom0)), est)        # first odom message only primes (:868-871)
        got = m.debug_update_initial_guess(info(odom1))

    def T(p):
        M = np.eye(4); M[:3, :3] = _rot(*p[:3].astype(np.float64)); M[:3, 3] = p[3:]; return M
    Tf = T(est) @ np.linalg.inv(T(odom0)) @ T(odom1)
    want = np.array([np.arctan2(Tf[2, 1], Tf[2, 2]), np.arcsin(-Tf[2, 0]), np.arctan2(Tf[1, 0], Tf[0, 0]), *Tf[:3, 3]])
    assert np.allclose(got, want, atol=2e-6)


def test_first_frame_guess_uses_imu_attitude():
    with _host() as m:
        ci = host_api.CloudInfo()
        ci.imuRollInit, ci.imuPitchInit, ci.imuYawInit = 0.1, -0.2, 1.5
        got = m.debug_update_initial_guess(ci)                                        # :835-842
    assert np.allclose(got, [0.1, -0.2, 1.5, 0, 0, 0])


def test_save_frame_thresholds():
    with _host("nclt.yaml") as m:                     # dist 1.0 m, angle 0.2 rad
        assert m.debug_save_frame()                   # no keyframes yet -> true (:1246-1247)
        m.debug_add_keypose(np.array([0, 0, 0.3, 5, 5, 1], np.float32), 0.0)
        m.debug_set_pose(np.array([0, 0, 0.35, 5.5, 5.2, 1.0], np.float32))
        assert not m.debug_save_frame()
        m.debug_set_pose(np.array([0, 0, 0.35, 5.9, 5.6, 1.0], np.float32))
        assert m.debug_save_frame()                   # moved > 1 m
        m.debug_set_pose(np.array([0, 0, 0.55, 5.1, 5.0, 1.0], np.float32))
        assert m.debug_save_frame()                   # yawed > 0.2 rad


def test_transform_update_clamps_and_imu_slerp(tmp_path):
    y = tmp_path / "p.yaml"
    y.write_text('System:\n  mode: "lio"\nSensors:\n  sensor: velodyne\n  imuType: 1\n  imuRPYWeight: 0.25\nMapping:\n  z_tollerance: 0.5\n  rotation_tollerance: 0.3\n')
    with host_api.MapOptimization(str(y), "lio", device=-1) as m:
        ci = host_api.CloudInfo()
        ci.imuAvailable = 1
        ci.imuRollInit, ci.imuPitchInit = 0.2, -0.1
        m.debug_set_pose(np.array([0.1, 0.1, 1.0, 1, 2, 3.0], np.float32))
        got = m.debug_transform_update(ci)
        # slerp between pure-roll (pure-pitch) quaternions is linear in the angle (:1217-1225)
        assert got[0] == pytest.approx(0.1 + 0.25 * (0.2 - 0.1), abs=1e-6)
        assert got[1] == pytest.approx(0.1 + 0.25 * (-0.1 - 0.1), abs=1e-6)
        assert got[5] == pytest.approx(0.5) and got[2] == pytest.approx(1.0)           # z clamped (:1231), yaw untouched
        m.debug_set_pose(np.array([0.9, -0.9, 0, 0, 0, -3.0], np.float32))
        ci.imuAvailable = 0
        got = m.debug_transform_update(ci)
        assert np.allclose(got, [0.3, -0.3, 0, 0, 0, -0.5])


def test_time_interval_gate_without_device():
    """Frames closer than timeInterval to the last processed one are dropped (:257)."""
    with _host("nclt.yaml") as m:                     # timeInterval 0.2
        ci = host_api.CloudInfo()
        ci.stamp = 10.0
        with pytest.raises(RuntimeError):             # no device context: the device call is refused loudly, no CPU fallback
            m.handle(ci)


def test_extract_nearby_fast_equals_plain():
    """extractNearbyFast (one sort, windowed 1-NN on an x-sorted index) against the step-by-step version on random key-pose
    sets: clustered, with exact duplicates (ties between poses and between distances), on voxel boundaries, growing one pose
    at a time (the incremental x index), and with radii that cut the set."""
    import tempfile, pathlib
    rng = np.random.default_rng(17)
    n_cases = 0
    for case in range(12):
        dens = [1.0, 2.0, 0.5][case % 3]
        radius = [50.0, 8.0, 3.0, 1000.0][case % 4]
        tmp = pathlib.Path(tempfile.mkdtemp())
        import yaml as pyyaml
        from liloc_b200 import host_api
        cfg = pyyaml.safe_load(open(host_api.config_path("nclt.yaml")))
        cfg["Mapping"]["surroundingKeyframeDensity"] = dens
        cfg["Mapping"]["surroundingKeyframeSearchRadius"] = radius
        y = tmp / "c.yaml"
        with open(y, "w") as f:
            for sec, kv in cfg.items():
                f.write(f"{sec}:\n")
                for k, v in kv.items():
                    f.write(f"  {k}: {'true' if v is True else 'false' if v is False else v}\n")
        n = int(rng.integers(1, 260))
        kind = case % 4
        if kind == 0:
            P = rng.normal(0, 6, (n, 3))
        elif kind == 1:
            P = np.round(rng.normal(0, 4, (n, 3)) * 2) / 2                      # a lattice: duplicates, voxel boundaries, ties
        elif kind == 2:
            P = np.cumsum(rng.normal(0, 0.4, (n, 3)), axis=0)                   # a trajectory
        else:
            P = np.repeat(rng.normal(0, 3, (max(n // 4, 1), 3)), 4, axis=0)[:n]  # every pose four times
        P[:, 2] *= 0.2
        P = P.astype(np.float32)
        with host_api.MapOptimization(str(y), "lio", device=-1) as fast, host_api.MapOptimization(str(y), "lio", device=-1) as plain:
            plain.debug_nearby_plain(True)
            for k in range(len(P)):
                pose = np.array([0, 0, 0, P[k, 0], P[k, 1], P[k, 2]], np.float32)
                fast.debug_add_keypose(pose, 0.1 * k); plain.debug_add_keypose(pose, 0.1 * k)
                if k % 7 == 0 or k == len(P) - 1:
                    a, b = fast.debug_extract_nearby(0.1 * k + 0.05), plain.debug_extract_nearby(0.1 * k + 0.05)
                    assert np.array_equal(a, b), (case, k, a, b)
                    n_cases += 1
    assert n_cases > 150


def test_host_class_against_ref_pipeline_differentially(orc, synth):
    """The C++ host class's updateInitialGuess (:865-883), extractNearby (:902-934) and saveFrame (:1245-1264) against the
    CPU pipeline the bench times (oracle/ref_pipeline.py RefLio, float32 Eigen-order restatements in oracle/host_oracle.cpp):
    the same guess bits, the same keyframe selection and the same keyframe decisions over a winding 400-frame trajectory
    (no device: the registration result is replaced by a known pose near the guess on both sides)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import ref_pipeline as rp
    rng = np.random.default_rng(5)
    for yaml in ("nclt.yaml", "lio_sam_default.yaml"):
        P = rp.Params(yaml)
        lio = rp.RefLio(P, threads=1)
        n = 400
        truth = synth.trajectory_loop(n, radius_xy=(30.0, 18.0)).astype(np.float64)
        odom = truth + np.cumsum(rng.normal(0, [1e-4, 1e-4, 3e-4, 2e-3, 2e-3, 5e-4], (n, 6)), axis=0)     # drifting odometry
        n_kf = n_sel_checked = 0
        with _host(yaml) as m:
            for f in range(n):
                stamp = 100.0 + f * (P.time_interval + 0.01)
                ci = host_api.CloudInfo()
                ci.stamp = stamp; ci.odomAvailable = 1
                ci.initialGuessRoll, ci.initialGuessPitch, ci.initialGuessYaw = [float(np.float32(v)) for v in odom[f, :3]]
                ci.initialGuessX, ci.initialGuessY, ci.initialGuessZ = [float(np.float32(v)) for v in odom[f, 3:]]
                # --- updateInitialGuess
                got = m.debug_update_initial_guess(ci)
                if lio.key_poses:
                    o32 = np.array([ci.initialGuessRoll, ci.initialGuessPitch, ci.initialGuessYaw, ci.initialGuessX, ci.initialGuessY, ci.initialGuessZ], np.float32)
                    if lio.last_odom is None:
                        lio.last_odom = o32
                    else:
                        lio.pose = orc.update_initial_guess(lio.pose, lio.last_odom, o32)
                        lio.last_odom = o32
                    assert np.array_equal(got, lio.pose), (yaml, f, got, lio.pose)
                    # --- extractNearby
                    sel = lio._select(stamp)
                    assert np.array_equal(m.debug_extract_nearby(stamp), np.asarray(sel, np.int32)), (yaml, f)
                    n_sel_checked += 1
                # --- "registration": both sides land on the same pose near the truth
                reg = (truth[f] + rng.normal(0, [1e-4, 1e-4, 2e-4, 5e-3, 5e-3, 1e-3])).astype(np.float32)
                m.debug_set_pose(reg); lio.pose = reg.copy()
                # --- saveFrame
                save = m.debug_save_frame()
                want = True if not lio.key_poses else orc.save_frame(lio.key_poses[-1], lio.pose, P.add_angle, P.add_dist)
                assert save == want, (yaml, f)
                if save:
                    m.debug_add_keypose(reg, stamp)
                    lio.key_poses.append(reg.copy()); lio.key_times.append(stamp); lio.key_clouds.append(None)
                    n_kf += 1
        assert n_kf > 40 and n_sel_checked > 350, (n_kf, n_sel_checked)
