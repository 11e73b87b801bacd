"""GPU parity of the de-skew step (src/imageProjection.cpp:573-673) through the C ABI: the
surviving points, their order and every coordinate are BIT-EXACT against oracle/deskew_oracle.cpp; the de-skewed
sweep then feeds the frame without leaving the device and the frame's result is identical to the one obtained from
a host copy of the same cloud."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sweep(synth, sensor, seed, **kw):
    world = synth.make_world(seed=seed)
    pose = np.array([0.01, -0.02, 0.3, -5.0, 2.0, 1.8], np.float32)
    return synth.raw_sweep(synth.scan(world, pose, sensor, seed=seed), sensor, seed=seed, **kw)


def _both(gpu_ctx, orc, sw, t, r, **kw):
    ref = orc.deskew(sw["pts"], sw["ring"], sw["time"], t, r, sw["time_scan_cur"], **kw)
    n = gpu_ctx.deskew(sw["pts"], sw["ring"], sw["time"], t, r, sw["time_scan_cur"], **kw)
    got = gpu_ctx.deskewed_get()
    assert n == len(ref) == len(got)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    return ref


@pytest.mark.parametrize("sensor_name,n_scan", [("VLP16", 16), ("HDL32", 32), ("MID360", 8)])
def test_deskew_bit_exact(gpu_ctx, orc, synth, sensor_name, n_scan):
    sensor = getattr(synth, sensor_name)
    sw = _sweep(synth, sensor, seed=21, gyro=(0.3, -0.2, 1.2))
    t, r = orc.imu_integrate(sw["imu_stamps"], sw["imu_gyro"], sw["time_scan_end"])
    ref = _both(gpu_ctx, orc, sw, t, r, n_scan=n_scan, point_filter_num=3, min_range=sensor.range_min, max_range=80.0)
    assert 0 < len(ref) < len(sw["pts"])


@pytest.mark.parametrize("ds,pf,minr,maxr", [(1, 1, 1.0, 1000.0), (2, 1, 2.0, 40.0), (1, 5, 0.0, 15.0), (4, 2, 5.0, 6.0), (3, 7, 1.0, 1000.0)])
def test_filters_and_compaction_order(gpu_ctx, orc, synth, ds, pf, minr, maxr):
    sw = _sweep(synth, synth.HDL32, seed=22, outliers=0.2)
    t, r = orc.imu_integrate(sw["imu_stamps"], sw["imu_gyro"], sw["time_scan_end"])
    _both(gpu_ctx, orc, sw, t, r, n_scan=32, downsample_rate=ds, point_filter_num=pf, min_range=minr, max_range=maxr)


def test_without_time_channel_or_imu_points_pass_unchanged(gpu_ctx, orc, synth):
    sw = _sweep(synth, synth.VLP16, seed=23)
    t, r = orc.imu_integrate(sw["imu_stamps"], sw["imu_gyro"], sw["time_scan_end"])
    a = _both(gpu_ctx, orc, sw, t, r, deskew_flag=-1, point_filter_num=2)
    b = _both(gpu_ctx, orc, sw, t, r, imu_available=False, point_filter_num=2)
    c = _both(gpu_ctx, orc, sw, None, None, point_filter_num=2)
    assert np.array_equal(a, b) and np.array_equal(a, c)
    p = sw["pts"]
    rng = np.sqrt(p[:, 0] * p[:, 0] + p[:, 1] * p[:, 1] + p[:, 2] * p[:, 2])
    keep = (rng >= 1.0) & (rng <= 1000.0) & (sw["ring"] >= 0) & (sw["ring"] < 16) & (np.arange(len(p)) % 2 == 0)
    assert np.array_equal(a, p[keep])


def test_edges(gpu_ctx, orc, synth):
    # empty sweep, everything filtered, a single IMU sample, the first survivor deep inside the sweep, 2000 IMU samples
    e = {"pts": np.zeros((0, 4), np.float32), "ring": np.zeros(0, np.int32), "time": np.zeros(0, np.float32), "time_scan_cur": 5.0}
    assert gpu_ctx.deskew(e["pts"], e["ring"], e["time"], None, None) == 0
    assert len(gpu_ctx.deskewed_get()) == 0
    sw = _sweep(synth, synth.VLP16, seed=24)
    t, r = orc.imu_integrate(sw["imu_stamps"], sw["imu_gyro"], sw["time_scan_end"])
    assert len(_both(gpu_ctx, orc, sw, t, r, min_range=500.0)) == 0
    _both(gpu_ctx, orc, sw, t[:1], r[:1], point_filter_num=1)
    late = dict(sw); late["ring"] = sw["ring"].copy(); late["ring"][:len(sw["ring"]) // 2 + 77] = 99      # first half dropped
    _both(gpu_ctx, orc, late, t, r, point_filter_num=3)
    rng = np.random.default_rng(0)
    tt = sw["time_scan_cur"] - 0.005 + np.sort(rng.uniform(0, 0.12, 2000)); tt[5] = tt[4]                 # a repeated stamp too
    rr = np.cumsum(rng.normal(0, 1e-3, (2000, 3)), axis=0)
    _both(gpu_ctx, orc, sw, tt, rr, point_filter_num=1)
    with pytest.raises(Exception):
        gpu_ctx.deskew(sw["pts"], sw["ring"], sw["time"], np.zeros(2001), np.zeros((2001, 3)))
    with pytest.raises(Exception):
        gpu_ctx.deskew(sw["pts"], sw["ring"], sw["time"], t, r, point_filter_num=0)


def test_deskewed_sweep_feeds_the_scan_pipeline_on_device(gpu_ctx, orc, synth):
    sw = _sweep(synth, synth.HDL32, seed=25)
    t, r = orc.imu_integrate(sw["imu_stamps"], sw["imu_gyro"], sw["time_scan_end"])
    ref = orc.deskew(sw["pts"], sw["ring"], sw["time"], t, r, sw["time_scan_cur"], n_scan=32)
    # count left on the device: the voxel pipeline reads it from there
    assert gpu_ctx.deskew(sw["pts"], sw["ring"], sw["time"], t, r, sw["time_scan_cur"], n_scan=32, sync=False) is None
    Q = gpu_ctx.scan_put_deskewed(0.4)
    ds = orc.voxelgrid(ref, 0.4)
    assert Q == len(ds)
    assert np.array_equal(gpu_ctx.scan_ds_get(), ds)
    # and again with the count already known on the host
    assert gpu_ctx.deskew(sw["pts"], sw["ring"], sw["time"], t, r, sw["time_scan_cur"], n_scan=32) == len(ref)
    assert gpu_ctx.scan_put_deskewed(0.2) == len(orc.voxelgrid(ref, 0.2))


def test_frame_from_raw_sweep_equals_frame_from_host_cloud(gpu_ctx, orc, synth, scene_small):
    s = scene_small
    gpu_ctx.keyframe_clear()
    for k, c in enumerate(s.kf_clouds):
        gpu_ctx.keyframe_put(k, c)
    ids = list(range(len(s.kf_clouds)))
    sw = synth.raw_sweep(s.scan_raw, s.sensor, gyro=(0.1, 0.05, -0.6), seed=26)
    t, r = orc.imu_integrate(sw["imu_stamps"], sw["imu_gyro"], sw["time_scan_end"])
    kw = dict(n_scan=16, point_filter_num=1, min_range=1.0)
    cloud = orc.deskew(sw["pts"], sw["ring"], sw["time"], t, r, sw["time_scan_cur"], **kw)
    pose_h, res_h, rc_h = gpu_ctx.frame(ids, s.kf_poses, s.map_leaf, cloud, s.scan_leaf, s.guess)
    gpu_ctx.deskew(sw["pts"], sw["ring"], sw["time"], t, r, sw["time_scan_cur"], sync=False, **kw)
    pose_d, res_d, rc_d = gpu_ctx.frame(ids, s.kf_poses, s.map_leaf, None, s.scan_leaf, s.guess)
    assert rc_h == rc_d == 0
    assert np.array_equal(pose_h, pose_d)
    assert (res_h.iters, res_h.n_sel, res_h.q_all, res_h.map_points) == (res_d.iters, res_d.n_sel, res_d.q_all, res_d.map_points)
    assert res_d.converged
    assert np.abs(pose_d[3:] - s.truth[3:]).max() < 0.05
    # the frame after: a new raw sweep reuses the buffers
    gpu_ctx.deskew(sw["pts"], sw["ring"], sw["time"], t, r, sw["time_scan_cur"], sync=False, **kw)
    pose_d2, _, _ = gpu_ctx.frame(ids, s.kf_poses, s.map_leaf, None, s.scan_leaf, s.guess)
    assert np.array_equal(pose_d2, pose_d)
    gpu_ctx.keyframe_put_from_scan(100)
    assert gpu_ctx.keyframe_size(100) == res_d.q_all


def test_ndt_source_from_the_deskewed_sweep(gpu_ctx, orc, synth, scene_small):
    from liloc_b200.abi import NdtOpts
    s = scene_small
    sw = synth.raw_sweep(s.scan_raw, s.sensor, gyro=(0.0, 0.0, 0.5), seed=27)
    t, r = orc.imu_integrate(sw["imu_stamps"], sw["imu_gyro"], sw["time_scan_end"])
    cloud = orc.deskew(sw["pts"], sw["ring"], sw["time"], t, r, sw["time_scan_cur"], n_scan=16, point_filter_num=2)
    submap = np.concatenate([synth.transform_cloud(c, p) for c, p in zip(s.kf_clouds, s.kf_poses)])
    gpu_ctx.ndt_clear()
    gpu_ctx.ndt_target_put(0, submap, 1.0)
    T = np.eye(4, dtype=np.float32); T[:3, :3] = synth.rot_zyx(*s.guess[:3]); T[:3, 3] = s.guess[3:]
    gpu_ctx.ndt_source_put(0, cloud)
    out_h = gpu_ctx.ndt_align_batch([0], [0], T[None], NdtOpts())
    gpu_ctx.deskew(sw["pts"], sw["ring"], sw["time"], t, r, sw["time_scan_cur"], n_scan=16, point_filter_num=2, sync=False)
    gpu_ctx.ndt_source_put(1)
    out_d = gpu_ctx.ndt_align_batch([0], [1], T[None], NdtOpts())
    assert np.array_equal(out_h[0], out_d[0]) and out_h[2][0] == out_d[2][0] and out_h[4][0] == out_d[4][0]
    gpu_ctx.ndt_clear()
