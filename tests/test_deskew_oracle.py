"""The de-skew restatement (oracle/deskew_oracle.cpp; src/imageProjection.cpp:441-673) against independent numpy
re-derivations and the properties the step must have.  CPU only."""
import numpy as np
import pytest

from liloc_b200 import synth
import oracle_lib as orc


def _sweep(sensor=synth.VLP16, seed=3, **kw):
    world = synth.make_world(seed=seed)
    pose = np.array([0.0, 0.0, 0.3, -5.0, 2.0, 1.8], np.float32)
    scan = synth.scan(world, pose, sensor, seed=seed)
    return synth.raw_sweep(scan, sensor, seed=seed, **kw)


def test_imu_integration_matches_numpy():
    sw = _sweep()
    t, r = orc.imu_integrate(sw["imu_stamps"], sw["imu_gyro"], sw["time_scan_end"])
    t2, r2 = synth.imu_integrate(sw["imu_stamps"], sw["imu_gyro"], sw["time_scan_end"])
    assert np.array_equal(t, t2) and np.array_equal(r, r2)
    assert t[0] <= sw["time_scan_cur"] and t[-1] <= sw["time_scan_end"] + 0.01
    assert np.all(r[0] == 0)


def _keep_mask(sw, min_range, max_range, n_scan, ds, pf):
    p = sw["pts"]
    rng = np.sqrt((p[:, 0] * p[:, 0] + p[:, 1] * p[:, 1] + p[:, 2] * p[:, 2]).astype(np.float32)).astype(np.float32)
    ring = sw["ring"]
    idx = np.arange(len(p))
    return (rng >= min_range) & (rng <= max_range) & (ring >= 0) & (ring < n_scan) & (ring % ds == 0) & (idx % pf == 0)


@pytest.mark.parametrize("ds,pf", [(1, 3), (2, 1), (1, 1), (4, 5)])
def test_filters_and_order_without_deskew(ds, pf):
    sw = _sweep()
    out = orc.deskew(sw["pts"], sw["ring"], sw["time"], deskew_flag=-1, min_range=2.0, max_range=60.0, n_scan=16,
                     downsample_rate=ds, point_filter_num=pf)
    m = _keep_mask(sw, 2.0, 60.0, 16, ds, pf)
    assert 0 < m.sum() < len(m)
    assert np.array_equal(out, sw["pts"][m])          # untouched points, push_back order


def test_imu_unavailable_returns_points_unchanged():
    sw = _sweep()
    t, r = orc.imu_integrate(sw["imu_stamps"], sw["imu_gyro"], sw["time_scan_end"])
    out = orc.deskew(sw["pts"], sw["ring"], sw["time"], t, r, sw["time_scan_cur"], imu_available=False, point_filter_num=1)
    assert np.array_equal(out, sw["pts"][_keep_mask(sw, 1.0, 1000.0, 16, 1, 1)])


def test_deskew_recovers_the_motion_free_cloud():
    """raw = R(t)^-1 R(0) p  =>  de-skewed = R(t_first)^-1 R(t) raw = R(t_first)^-1 R(0) p"""
    sw = _sweep(gyro=(0.2, -0.1, 1.5), outliers=0.0)
    sw["imu_gyro"][:] = (0.2, -0.1, 1.5)               # noise-free rates: linear interpolation is then exact
    t, r = orc.imu_integrate(sw["imu_stamps"], sw["imu_gyro"], sw["time_scan_end"])
    out = orc.deskew(sw["pts"], sw["ring"], sw["time"], t, r, sw["time_scan_cur"], min_range=0.0, point_filter_num=1)
    assert len(out) == len(sw["pts"])
    w = np.array((0.2, -0.1, 1.5))
    a_first = w * (0.008 + float(sw["time"][0]))
    expect = (synth.rot_zyx(*a_first).T @ synth.rot_zyx(*(w * 0.008)) @ sw["truth"][:, :3].astype(np.float64).T).T
    err = np.abs(out[:, :3] - expect).max()
    skew = np.abs(sw["pts"][:, :3] - sw["truth"][:, :3]).max()
    assert skew > 0.5                                   # the sweep really was distorted (1.5 rad/s over 0.1 s at tens of metres)
    assert err < 2e-4, err                              # float32 transform of points up to ~60 m
    assert np.array_equal(out[:, 3], sw["pts"][:, 3])   # intensity copied


def test_first_surviving_point_defines_the_start_frame():
    sw = _sweep()
    t, r = orc.imu_integrate(sw["imu_stamps"], sw["imu_gyro"], sw["time_scan_end"])
    out = orc.deskew(sw["pts"], sw["ring"], sw["time"], t, r, sw["time_scan_cur"], point_filter_num=3)
    m = _keep_mask(sw, 1.0, 1000.0, 16, 1, 3)
    first = np.flatnonzero(m)[0]
    # transBt of the first survivor is inverse(T) * T: the identity up to float rounding
    assert np.abs(out[0, :3] - sw["pts"][first, :3]).max() < 1e-4


def _f32(x):
    return np.float32(x)


def _np_get_transformation(roll, pitch, yaw):
    A, B = _f32(np.cos(np.float64(yaw))), _f32(np.sin(np.float64(yaw)))
    Cc, D = _f32(np.cos(np.float64(pitch))), _f32(np.sin(np.float64(pitch)))
    E, F = _f32(np.cos(np.float64(roll))), _f32(np.sin(np.float64(roll)))
    DE, DF = D * E, D * F
    return np.array([[A * Cc, A * DF - B * E, B * F + A * DE], [B * Cc, A * E + B * DF, B * DE - A * F], [-D, Cc * F, Cc * E]], np.float32)


def _sum3(a, b, c):
    return _f32(a + _f32(b + c))


def _np_inverse(m):
    def cof(i, j):
        i1, i2, j1, j2 = (i + 1) % 3, (i + 2) % 3, (j + 1) % 3, (j + 2) % 3
        return _f32(_f32(m[i1, j1] * m[i2, j2]) - _f32(m[i1, j2] * m[i2, j1]))
    c0 = [cof(k, 0) for k in range(3)]
    det = _sum3(c0[0] * m[0, 0], c0[1] * m[1, 0], c0[2] * m[2, 0])
    inv = _f32(_f32(1.0) / det)
    r = np.empty((3, 3), np.float32)
    for k in range(3):
        r[0, k] = c0[k] * inv
        r[1, k] = cof(k, 1) * inv
        r[2, k] = cof(k, 2) * inv
    return r


def test_arithmetic_against_numpy_float32_restatement():
    """An independent float32 numpy evaluation of deskewPoint for a few points (Eigen's evaluation order), bit for bit."""
    sw = _sweep(seed=5)
    t, r = orc.imu_integrate(sw["imu_stamps"], sw["imu_gyro"], sw["time_scan_end"])
    sel = np.r_[0:6, 500:506, len(sw["pts"]) - 6:len(sw["pts"])]
    pts, ring, tim = sw["pts"][sel], np.zeros(len(sel), np.int32), sw["time"][sel]
    out = orc.deskew(pts, ring, tim, t, r, sw["time_scan_cur"], min_range=0.0, point_filter_num=1)
    assert len(out) == len(sel)

    def rot_at(rel):
        pt = np.float64(sw["time_scan_cur"]) + np.float64(rel)
        front = 0
        while front < len(t) - 1 and not (pt < t[front]):
            front += 1
        if pt > t[front] or front == 0:
            return [_f32(r[front, k]) for k in range(3)]
        back = front - 1
        rf = (pt - t[back]) / (t[front] - t[back]); rb = (t[front] - pt) / (t[front] - t[back])
        return [_f32(r[front, k] * rf + r[back, k] * rb) for k in range(3)]

    with np.errstate(all="ignore"):
        start_inv = _np_inverse(_np_get_transformation(*rot_at(tim[0])))
        for i in range(len(sel)):
            Tf = _np_get_transformation(*rot_at(tim[i]))
            bt = np.empty((3, 3), np.float32)
            for a in range(3):
                for b in range(3):
                    bt[a, b] = _sum3(start_inv[a, 0] * Tf[0, b], start_inv[a, 1] * Tf[1, b], start_inv[a, 2] * Tf[2, b])
            x, y, z = pts[i, 0], pts[i, 1], pts[i, 2]
            exp = [_f32(_f32(_f32(bt[a, 0] * x) + _f32(bt[a, 1] * y)) + _f32(bt[a, 2] * z)) for a in range(3)]
            # + transBt(a, 3): the translation is a signed zero, which leaves a non-zero sum unchanged
            assert [float(v) for v in exp] == [float(v) for v in out[i, :3]], i


def test_rotation_lookup_edges():
    pts = np.array([[10, 0, 0, 1], [0, 10, 0, 2], [0, 0, 10, 3]], np.float32)
    t = np.array([100.0, 100.01, 100.02]); r = np.array([[0, 0, 0], [0, 0, 0.1], [0, 0, 0.2]])
    # point 0 before the first IMU sample (front == 0: sample 0), point 1 in between, point 2 after the last (last sample)
    out = orc.deskew(pts, [0, 0, 0], [-0.5, 0.015, 0.9], t, r, 100.0, min_range=0.0, point_filter_num=1)
    c, s = np.cos(0.15), np.sin(0.15)
    assert np.allclose(out[0, :3], [10, 0, 0], atol=1e-6)
    assert np.allclose(out[1, :3], [-10 * s, 10 * c, 0], atol=1e-5)
    assert np.allclose(out[2, :3], [0, 0, 10], atol=1e-6)
    # a single IMU sample: every point takes it (imuPointerCur == 0), i.e. no relative rotation at all
    out1 = orc.deskew(pts, [0, 0, 0], [0.0, 0.05, 0.1], t[:1], r[:1], 100.0, min_range=0.0, point_filter_num=1)
    assert np.allclose(out1[:, :3], pts[:, :3], atol=1e-6)


def test_empty_and_fully_filtered():
    e = np.zeros((0, 4), np.float32)
    assert len(orc.deskew(e, [], [], deskew_flag=-1)) == 0
    pts = np.array([[0.1, 0, 0, 1], [0.2, 0, 0, 1]], np.float32)
    assert len(orc.deskew(pts, [0, 1], [0, 0.1], deskew_flag=-1, min_range=1.0)) == 0


def test_orientation_times_of_a_spinning_sweep():
    """no ring / time channel (:279-326): azimuth -> relative time in [0, 0.1], increasing along a clockwise sweep"""
    n = 3600
    az = -np.linspace(0.05, 2 * np.pi - 0.05, n)           # Velodyne spins clockwise: ori = -atan2(y, x) increases
    pts = np.stack([20 * np.cos(az), 20 * np.sin(az), np.zeros(n), np.ones(n)], axis=1).astype(np.float32)
    tm = orc.orientation_times(pts)
    assert tm.min() >= 0 and tm.max() <= 0.1 + 1e-6
    assert np.all(np.diff(tm) > -1e-6)
    assert abs(tm[-1] - 0.1) < 1e-3 and tm[0] == 0


def _golden():
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "deskew_np32.npz"))
    kw = dict(min_range=float(g["min_range"]), max_range=float(g["max_range"]), n_scan=int(g["n_scan"]),
              downsample_rate=int(g["downsample_rate"]), point_filter_num=int(g["point_filter_num"]))
    return g, kw


def test_golden_fixture_from_the_numpy_restatement():
    """tests/golden/deskew_np32.npz (made by make_golden_deskew.py without the oracle): filters, compaction order, rotation
    look-up incl. points later than the last IMU sample, every float32 operation -- bit for bit."""
    g, kw = _golden()
    out = orc.deskew(g["pts"], g["ring"], g["time"], g["imu_time"], g["imu_rot"], float(g["time_scan_cur"]), **kw)
    assert out.shape == g["out"].shape and len(out) > 50
    assert np.array_equal(out.view(np.uint32), g["out"].view(np.uint32))
    # the integration of the fixture's IMU samples is reproduced too (imuDeskewInfo)
    dt = np.diff(g["imu_time"])
    assert np.all(dt > 0) and np.all(g["imu_rot"][0] == 0)


def test_zero_motion_is_the_identity_bit_for_bit():
    """All integrated angles zero: getTransformation is the exact identity, its cofactor inverse and the product too, and
    1*x + 0*y + 0*z + 0 reproduces every coordinate exactly."""
    sw = _sweep(seed=9)
    t, r = orc.imu_integrate(sw["imu_stamps"], np.zeros_like(sw["imu_gyro"]), sw["time_scan_end"])
    assert np.all(r == 0)
    out = orc.deskew(sw["pts"], sw["ring"], sw["time"], t, r, sw["time_scan_cur"], point_filter_num=2)
    assert np.array_equal(out, sw["pts"][_keep_mask(sw, 1.0, 1000.0, 16, 1, 2)])


def test_filter_parameters_property():
    """hypothesis: for any ranges / strides the survivors are exactly the numpy mask, in order (deskew switched off)."""
    from hypothesis import given, settings, strategies as st
    sw = _sweep(seed=10)

    @settings(max_examples=40, deadline=None)
    @given(st.floats(0.0, 30.0), st.floats(5.0, 120.0), st.integers(1, 40), st.integers(1, 5), st.integers(1, 9))
    def check(minr, maxr, n_scan, ds, pf):
        out = orc.deskew(sw["pts"], sw["ring"], sw["time"], deskew_flag=-1, min_range=minr, max_range=maxr, n_scan=n_scan,
                         downsample_rate=ds, point_filter_num=pf)
        m = _keep_mask(sw, np.float32(minr), np.float32(maxr), n_scan, ds, pf)
        assert np.array_equal(out, sw["pts"][m])
    check()


def test_rigid_motion_preserves_ranges():
    """de-skewing is a rotation about the sensor origin (findPosition is zero): ranges are preserved to float accuracy"""
    sw = _sweep(seed=11, gyro=(0.4, 0.3, -1.0))
    t, r = orc.imu_integrate(sw["imu_stamps"], sw["imu_gyro"], sw["time_scan_end"])
    out = orc.deskew(sw["pts"], sw["ring"], sw["time"], t, r, sw["time_scan_cur"], min_range=0.0, point_filter_num=1)
    keep = _keep_mask(sw, 0.0, 1000.0, 16, 1, 1)
    r_in = np.linalg.norm(sw["pts"][keep, :3].astype(np.float64), axis=1)
    r_out = np.linalg.norm(out[:, :3].astype(np.float64), axis=1)
    assert np.abs(r_in - r_out).max() < 2e-5 * max(1.0, r_in.max())
