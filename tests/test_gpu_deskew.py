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


def _quat_from_rpy(r, p, y):        # tf::Quaternion::setRPY
    cr, sr, cp, sp, cy, sy = np.cos(r / 2), np.sin(r / 2), np.cos(p / 2), np.sin(p / 2), np.cos(y / 2), np.sin(y / 2)
    return (sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy)


def test_image_projection_node_in_front_of_map_optimization(orc, synth):
    """The two host classes in one process: ImageProjection (IMU / odometry queues, gyro integration on the host,
    de-skew on the device) hands mapOptimization a CloudInfo whose cloud never left the device; the frames must equal,
    bit for bit, the frames of a second mapOptimization fed the oracle's de-skewed cloud from host memory."""
    from liloc_b200 import host_api
    from liloc_b200.abi import Context
    yaml = "lio_sam_default.yaml"
    F = 9
    seq = synth.make_sequence(synth.VLP16, F, seed=6, world_size=(60.0, 40.0), n_pillars=10, loop_radius=(20.0, 11.0), dt=0.21)
    with host_api.MapOptimization(yaml, "lio", device=0) as m_raw, host_api.MapOptimization(yaml, "lio", device=0) as m_ref:
        ip = host_api.ImageProjection(yaml, m_raw)
        ctx = Context.borrow(m_raw.ctx_handle())
        R = np.array([0.9994142, 0.03275957, 0.00990251, -0.03266505, 0.99942062, -0.00956125, -0.01021, 0.00923218, 0.99990526]).reshape(3, 3)
        sweeps, produced = [], 0
        for f in range(F):
            sw = synth.raw_sweep(seq.scans[f], seq.sensor, gyro=(0.05, -0.02, 0.3 + 0.05 * f), time_scan_cur=float(seq.stamps[f]), seed=100 + f)
            sweeps.append(sw)
            for st, g in zip(sw["imu_stamps"], sw["imu_gyro"]):
                ip.imu(st, g)
            r, p, y, x, yy, z = [float(v) for v in seq.odom[f]]
            ip.odom(seq.stamps[f] - 0.005, (x, yy, z), _quat_from_rpy(r, p, y))
            ip.odom(seq.stamps[f], (x, yy, z), _quat_from_rpy(r, p, y))
            info = ip.cloud(seq.stamps[f], sw["pts"], sw["ring"], sw["time"], has_time=True)
            if f < 2:
                assert info is None                      # cachePointCloud holds two sweeps back (:271-273)
                continue
            k = f - 2                                    # the sweep that was processed
            s = sweeps[k]
            assert info is not None and info.cloud_on_device == 2 and info.stamp == seq.stamps[k]
            assert info.imuAvailable == 1 and info.odomAvailable == 1
            assert abs(info.initialGuessX - seq.odom[k][3]) < 1e-6 and abs(info.initialGuessYaw - seq.odom[k][2]) < 1e-6
            t_cur, t_end, flag = ip.times()
            assert (t_cur, flag) == (seq.stamps[k], 1) and t_end == s["time_scan_end"]
            # the oracle's view of the same sweep: extrinsic rotation of the gyro (imuConverter), integration, de-skew
            g_rot = np.array([[R[i, 0] * g[0] + R[i, 1] * g[1] + R[i, 2] * g[2] for i in range(3)] for g in s["imu_gyro"]])
            it, ir = orc.imu_integrate(s["imu_stamps"], g_rot, s["time_scan_end"])
            ref = orc.deskew(s["pts"], s["ring"], s["time"], it, ir, s["time_scan_cur"], n_scan=16, point_filter_num=1, min_range=2.0, max_range=100.0)
            got = ctx.deskewed_get()
            assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
            rep_raw = m_raw.handle(info)
            info_h = host_api.CloudInfo.from_buffer_copy(info)
            info_h.cloud_deskewed, info_h.cloud_size, info_h.cloud_on_device = ref.ctypes.data, len(ref), 0
            rep_ref = m_ref.handle(info_h)
            assert rep_raw.processed == rep_ref.processed == 1 and rep_raw.status == rep_ref.status == 0
            assert list(rep_raw.pose) == list(rep_ref.pose) and list(rep_raw.guess) == list(rep_ref.guess)
            assert (rep_raw.q_all, rep_raw.map_points, rep_raw.iters, rep_raw.n_sel, rep_raw.keyframe_id) == \
                   (rep_ref.q_all, rep_ref.map_points, rep_ref.iters, rep_ref.n_sel, rep_ref.keyframe_id)
            produced += 1
        assert produced == F - 2 and m_raw.num_keyframes() == m_ref.num_keyframes() >= 3
        # a sweep whose IMU data has not arrived yet is not published ("Waiting for IMU data ...", :429-432)
        sw = synth.raw_sweep(seq.scans[0], seq.sensor, time_scan_cur=float(seq.stamps[-1]) + 5.0, seed=7)
        assert ip.cloud(float(seq.stamps[-1]) + 5.0, sw["pts"], sw["ring"], sw["time"]) is not None     # sweep F-2: its IMU data is there
        assert ip.cloud(float(seq.stamps[-1]) + 5.2, sw["pts"], sw["ring"], sw["time"]) is not None     # sweep F-1
        assert ip.cloud(float(seq.stamps[-1]) + 5.4, sw["pts"], sw["ring"], sw["time"]) is None         # the sweep at +5.0: no IMU data
        ctx.close(); ip.close()


def test_livox_custom_msg_path(orc, synth):
    """cloudHandlerLivox (:194-268): tag / line / IsNear filters, time in ms, sort by time, then the common path."""
    from liloc_b200 import host_api
    from liloc_b200.abi import Context
    yaml = "lio_sam_mid360.yaml"
    world = synth.make_world(seed=8)
    rng = np.random.default_rng(8)
    with host_api.MapOptimization(yaml, "lio", device=0) as m:
        ip = host_api.ImageProjection(yaml, m)
        ctx = Context.borrow(m.ctx_handle())
        msgs = []
        for f in range(3):
            sc = synth.scan(world, np.array([0, 0, 0.1 * f, -5.0 + f, 2.0, 1.8], np.float32), synth.MID360, seed=80 + f)
            lp = np.zeros(len(sc), host_api.LIVOX_POINT)
            lp["x"], lp["y"], lp["z"] = sc[:, 0], sc[:, 1], sc[:, 2]
            lp["line"] = rng.integers(0, 10, len(sc))                      # N_SCAN 8: lines 8, 9 are dropped
            lp["tag"] = rng.choice([0x00, 0x10, 0x20, 0x30, 0x01], len(sc))
            lp["offset_time"] = rng.permutation(len(sc)).astype(np.uint32) * 4000    # ns, unsorted
            lp["x"][5] = lp["x"][4]                                          # IsNear with the previous point
            stamp = 50.0 + 0.1 * f
            msgs.append((stamp, lp))
            for k in range(40):
                ip.imu(stamp - 0.008 + k * 0.005, (0.01, 0.02, 0.2))
            info = ip.cloud_livox(stamp, lp)
            assert (info is None) == (f < 2)
        stamp, lp = msgs[0]
        keep = np.ones(len(lp), bool); keep[0] = False
        keep &= (lp["line"] < 8) & (((lp["tag"] & 0x30) == 0x10) | ((lp["tag"] & 0x30) == 0x00))
        near = np.zeros(len(lp), bool)
        for c in "xyz":
            near[1:] |= np.abs(lp[c][1:] - lp[c][:-1]) < 1e-7
        keep &= ~near
        tms = (0.0 + lp["offset_time"].astype(np.float64) * 1e-6).astype(np.float32)[keep]
        order = np.argsort(tms, kind="stable")
        pts = np.stack([lp["x"][keep], lp["y"][keep], lp["z"][keep], np.zeros(keep.sum(), np.float32)], axis=1)[order]
        ring, tms = lp["line"][keep][order].astype(np.int32), tms[order]
        t_cur, t_end, flag = ip.times()
        assert t_cur == stamp and flag == 1 and t_end == stamp + float(tms[-1]) / 1000.0
        st = np.array([stamp - 0.008 + k * 0.005 for k in range(40)] + [msgs[1][0] - 0.008 + k * 0.005 for k in range(40)] + [msgs[2][0] - 0.008 + k * 0.005 for k in range(40)])
        it, ir = orc.imu_integrate(st, np.tile((0.01, 0.02, 0.2), (len(st), 1)), t_end)
        ref = orc.deskew(pts, ring, tms, it, ir, stamp, n_scan=8, point_filter_num=1, min_range=1.0, max_range=100.0)
        got = ctx.deskewed_get()
        assert len(ref) > 1000 and np.array_equal(got.view(np.uint32), ref.view(np.uint32))
        ctx.close(); ip.close()


def test_sweep_staged_ahead_of_the_current_frame(gpu_ctx, orc, synth, scene_small):
    """LILOC_DESKEW_AHEAD: the next sweep is uploaded and de-skewed (on a stream of its own) while LILOC_SCAN_DESKEWED still
    refers to the current one; liloc_deskew_advance swaps.  Frames pipelined this way equal the sequential ones."""
    s = scene_small
    gpu_ctx.keyframe_clear()
    for k, c in enumerate(s.kf_clouds):
        gpu_ctx.keyframe_put(k, c)
    ids = list(range(len(s.kf_clouds)))
    sweeps = []
    for f in range(4):
        sw = synth.raw_sweep(s.scan_raw, s.sensor, gyro=(0.05 * f, 0.02, 0.3 - 0.2 * f), seed=40 + f)
        t, r = orc.imu_integrate(sw["imu_stamps"], sw["imu_gyro"], sw["time_scan_end"])
        sweeps.append((sw, t, r))
    kw = dict(n_scan=16, point_filter_num=1, min_range=1.0)

    def dk(i, **extra):
        sw, t, r = sweeps[i]
        return gpu_ctx.deskew(sw["pts"], sw["ring"], sw["time"], t, r, sw["time_scan_cur"], sync=False, **kw, **extra)
    seq = []
    for i in range(4):                                   # sequential: de-skew, then the frame
        dk(i)
        seq.append(gpu_ctx.frame(ids, s.kf_poses, s.map_leaf, None, s.scan_leaf, s.guess))
    with pytest.raises(Exception):
        gpu_ctx.deskew_advance()                         # nothing staged
    dk(0)
    for i in range(4):                                   # pipelined: sweep i + 1 is staged while frame i runs
        if i + 1 < 4:
            dk(i + 1, ahead=True)
        pose, res, rc = gpu_ctx.frame(ids, s.kf_poses, s.map_leaf, None, s.scan_leaf, s.guess)
        assert rc == 0 and np.array_equal(pose, seq[i][0]) and (res.iters, res.n_sel, res.q_all) == (seq[i][1].iters, seq[i][1].n_sel, seq[i][1].q_all), i
        sw, t, r = sweeps[i]
        ref = orc.deskew(sw["pts"], sw["ring"], sw["time"], t, r, sw["time_scan_cur"], **kw)
        assert np.array_equal(gpu_ctx.deskewed_get().view(np.uint32), ref.view(np.uint32))      # still the current sweep
        if i + 1 < 4:
            gpu_ctx.deskew_advance()
    assert len({tuple(p[0]) for p in seq}) > 1           # the four sweeps really differ


def test_large_sweep_several_tiles_per_block(gpu_ctx, orc):
    """A 128-beam sweep (262 144 returns): more 256-point tiles than resident blocks, so every block walks several tiles
    and the tile offsets cross block boundaries."""
    rng = np.random.default_rng(61)
    n = 262144
    pts = (rng.normal(size=(n, 4)) * np.array([30.0, 30.0, 3.0, 40.0])).astype(np.float32)
    ring = rng.integers(-2, 132, n).astype(np.int32)
    tm = np.sort(rng.uniform(0.0, 0.1, n)).astype(np.float32)
    imu_t = 700.0 - 0.004 + np.arange(60) * 0.002
    imu_r = np.cumsum(rng.normal(0, 2e-3, (60, 3)), axis=0); imu_r[0] = 0
    for pf, ds in ((1, 1), (3, 2)):
        kw = dict(n_scan=128, downsample_rate=ds, point_filter_num=pf, min_range=2.0, max_range=90.0)
        ref = orc.deskew(pts, ring, tm, imu_t, imu_r, 700.0, **kw)
        n_out = gpu_ctx.deskew(pts, ring, tm, imu_t, imu_r, 700.0, **kw)
        got = gpu_ctx.deskewed_get()
        assert n_out == len(ref) and 1000 < len(ref) < n
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    # and staged ahead (one block per SM: twice as many tiles per block)
    gpu_ctx.deskew(pts, ring, tm, imu_t, imu_r, 700.0, ahead=True, sync=False, **kw)
    gpu_ctx.deskew_advance()
    assert np.array_equal(gpu_ctx.deskewed_get().view(np.uint32), ref.view(np.uint32))


def test_golden_fixture(gpu_ctx):
    """The device against tests/golden/deskew_np32.npz, the oracle-independent numpy float32 restatement."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "deskew_np32.npz"))
    n = gpu_ctx.deskew(g["pts"], g["ring"], g["time"], g["imu_time"], g["imu_rot"], float(g["time_scan_cur"]),
                       min_range=float(g["min_range"]), max_range=float(g["max_range"]), n_scan=int(g["n_scan"]),
                       downsample_rate=int(g["downsample_rate"]), point_filter_num=int(g["point_filter_num"]))
    got = gpu_ctx.deskewed_get()
    assert n == len(g["out"]) == len(got)
    assert np.array_equal(got.view(np.uint32), g["out"].view(np.uint32))
