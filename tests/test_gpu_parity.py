"""GPU parity tests (pytest -m gpu, run on the B200 box): every call goes through the C ABI
(include/liloc_gpu.h -> libliloc_gpu.so) and is compared with the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star): integer / index / per-point float work is BIT-EXACT (both sides evaluate
the same IEEE float32 sequence without FMA); final poses agree within 1e-4 m and 1e-5 rad, and
iters / n_sel / is_degenerate are compared exactly.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

POSE_TOL_M = 1e-4      # north_star tolerance, metres
POSE_TOL_RAD = 1e-5    # north_star tolerance, radians


def _assert_pose_close(got, ref):
    assert np.abs(got[3:] - ref[3:]).max() <= POSE_TOL_M, (got, ref)
    assert np.abs(got[:3] - ref[:3]).max() <= POSE_TOL_RAD, (got, ref)


def test_pose_to_T_bit_exact(gpu_ctx, orc):
    rng = np.random.default_rng(0)
    for _ in range(200):
        pose = np.concatenate([rng.uniform(-3.2, 3.2, 3), rng.uniform(-100, 100, 3)]).astype(np.float32)
        assert np.array_equal(gpu_ctx.pose_to_T(pose), orc.pose_to_T(pose))


def test_small_dense_algebra_bit_exact(gpu_ctx, orc):
    """6x6 QR solve, Jacobi eigen and degeneracy projector, and the 5x3 plane fit, evaluated on the
    device, against the oracle on the golden systems: same float32 operation sequence -> same bits."""
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(gold, "lm6x6_cv2.npz"))
    x, w, P, deg = gpu_ctx.debug_lm6(g["A"], g["b"])
    for i, (A, b) in enumerate(zip(g["A"], g["b"])):
        assert np.array_equal(x[i], orc.solve6(A, b))
        assert np.array_equal(w[i], orc.eigen6(A)[0])
        d, Pi = orc.degeneracy(A)
        assert bool(deg[i]) == d and np.array_equal(P[i], Pi)
    pg = np.load(os.path.join(gold, "plane5x3_lstsq.npz"))["P"]
    rng = np.random.default_rng(3)
    extra = np.concatenate([pg, np.repeat(np.array([[[1, 2, 3]] * 5], np.float32), 2, 0),      # rank 1
                            (rng.normal(size=(500, 5, 3)) * np.array([30, 30, 3]) + np.array([50, -80, 2])).astype(np.float32)])
    xs = gpu_ctx.debug_plane(extra)
    for i, pts in enumerate(extra):
        assert np.array_equal(xs[i], orc.plane_lsq(pts)), i


@pytest.mark.parametrize("n,leaf,seed", [(1, 0.2, 0), (2, 0.2, 1), (31, 0.5, 2), (2049, 0.4, 3), (60000, 0.2, 4), (500000, 0.4, 5)])
def test_voxelgrid_bit_exact_random(gpu_ctx, orc, n, leaf, seed):
    rng = np.random.default_rng(seed)
    pts = (rng.normal(size=(n, 4)) * np.array([15, 9, 1.5, 60])).astype(np.float32)
    got = gpu_ctx.voxelgrid(pts, leaf)
    ref = orc.voxelgrid(pts, leaf)
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)


def test_voxelgrid_bit_exact_scans(gpu_ctx, orc, synth):
    w = synth.make_world(seed=2)
    tr = synth.trajectory_line(3)
    for sensor, leaf in ((synth.VLP16, 0.2), (synth.HDL32, 0.4), (synth.MID360, 0.2)):
        s = synth.scan(w, tr[1], sensor, 5)
        got, ref = gpu_ctx.voxelgrid(s, leaf), orc.voxelgrid(s, leaf)
        assert got.shape == ref.shape and np.array_equal(got, ref)


def test_voxelgrid_edge_cases(gpu_ctx, orc):
    assert len(gpu_ctx.voxelgrid(np.zeros((0, 4), np.float32), 0.2)) == 0
    # many points in one voxel: long sequential float sums must keep the source order
    rng = np.random.default_rng(9)
    pts = (rng.uniform(0.01, 0.19, (5000, 4))).astype(np.float32)
    assert np.array_equal(gpu_ctx.voxelgrid(pts, 0.2), orc.voxelgrid(pts, 0.2))
    # duplicates
    pts = np.repeat(rng.normal(size=(50, 4)).astype(np.float32), 7, axis=0)
    assert np.array_equal(gpu_ctx.voxelgrid(pts, 0.3), orc.voxelgrid(pts, 0.3))
    # PCL overflow guard: leaf too small -> input returned unchanged
    pts = np.array([[0, 0, 0, 1], [3000, 3000, 3000, 2], [1, 2, 3, 4]], np.float32)
    assert np.array_equal(gpu_ctx.voxelgrid(pts, 0.001), pts)
    # negative coordinates / voxel boundaries
    g = (np.arange(-40, 40) * 0.25).astype(np.float32)
    pts = np.stack(np.meshgrid(g, g[:20], g[:5], indexing="ij"), -1).reshape(-1, 3)
    pts = np.c_[pts, np.arange(len(pts))].astype(np.float32)
    assert np.array_equal(gpu_ctx.voxelgrid(pts, 0.5), orc.voxelgrid(pts, 0.5))


@pytest.mark.parametrize("n", [255, 256, 257, 1023, 1024, 1025, 148 * 1024 - 1, 148 * 1024, 148 * 1024 + 1, 300001])
def test_voxelgrid_sort_tile_boundaries(gpu_ctx, orc, n):
    """Sizes around the tile / chunk boundaries of the in-kernel radix sort (256-thread rounds, 1024-item tiles,
    148 sorting blocks)."""
    rng = np.random.default_rng(n)
    pts = (rng.normal(size=(n, 4)) * np.array([25, 14, 2, 50])).astype(np.float32)
    assert np.array_equal(gpu_ctx.voxelgrid(pts, 0.4), orc.voxelgrid(pts, 0.4))


def test_voxelgrid_sort_key_widths_and_long_runs(gpu_ctx, orc):
    rng = np.random.default_rng(77)
    # 1..4 radix passes: the key width is decided on the device from the bounding box
    for extent, leaf in ((3.0, 0.5), (40.0, 0.4), (400.0, 0.4), (900.0, 0.1)):
        pts = rng.uniform(-extent, extent, (40000, 4)).astype(np.float32)
        pts[:, 2] *= 0.05
        got, ref = gpu_ctx.voxelgrid(pts, leaf), orc.voxelgrid(pts, leaf)
        assert got.shape == ref.shape and np.array_equal(got, ref), (extent, leaf)
    # one voxel holding 300 k points (a run across ~300 tiles and every sorting block) next to sparse ones
    pts = np.concatenate([rng.uniform(0.01, 0.19, (300000, 4)), rng.uniform(-30, 30, (5000, 4))]).astype(np.float32)
    pts = pts[rng.permutation(len(pts))]
    assert np.array_equal(gpu_ctx.voxelgrid(pts, 0.2), orc.voxelgrid(pts, 0.2))
    # a skewed digit distribution: everything on one line along x (only the low key bits vary), then along z
    for axis in (0, 2):
        pts = np.zeros((60000, 4), np.float32)
        pts[:, axis] = rng.uniform(-200, 200, 60000)
        pts[:, 3] = np.arange(60000)
        assert np.array_equal(gpu_ctx.voxelgrid(pts, 0.3), orc.voxelgrid(pts, 0.3))


def test_localmap_build_bit_exact(gpu_ctx, orc, scene_hdl):
    s = scene_hdl
    gpu_ctx.keyframe_clear()
    for k, c in enumerate(s.kf_clouds):
        gpu_ctx.keyframe_put(k, c)
    # selection order with a duplicate, like extractNearby's "last 5 s" append (src/mapOptmization.cpp:924-931)
    ids = list(range(len(s.kf_clouds)))[::2] + [len(s.kf_clouds) - 1, len(s.kf_clouds) - 2, 0]
    M, n_raw = gpu_ctx.localmap_build(ids, s.kf_poses[ids], s.map_leaf)
    ref = orc.localmap_build([s.kf_clouds[i] for i in ids], s.kf_poses[ids], s.map_leaf)
    assert n_raw == sum(len(s.kf_clouds[i]) for i in ids)
    got = gpu_ctx.localmap_get()
    assert M == len(ref) == len(got)
    assert np.array_equal(got, ref)


def _knn_compare(gpu_ctx, orc, map_ds, queries):
    idx, d2, found = gpu_ctx.knn5(queries)
    ridx, rd2 = orc.knn5(map_ds, queries, kdtree=False, threads=8)
    accept = (ridx[:, 4] >= 0) & (rd2[:, 4] < 1.0)          # the reference's gate, src/mapOptmization.cpp:1002
    assert np.array_equal(found == 5, accept)
    assert np.array_equal(idx[accept], ridx[accept])           # index sets AND order, ties by index
    assert np.array_equal(d2[accept], rd2[accept])
    # rejected queries: whatever was found below 1.0 is a prefix of the true neighbour list
    for r in np.flatnonzero(~accept)[:200]:
        f = found[r]
        assert np.array_equal(idx[r, :f], ridx[r, :f]) and np.all(idx[r, f:] == -1)
        assert f == np.count_nonzero(rd2[r] < 1.0)
    return int(accept.sum())


def test_knn5_bit_exact_scene(gpu_ctx, orc, scene_hdl):
    s = scene_hdl
    raw = np.concatenate([orc.transform_cloud(c, p) for c, p in zip(s.kf_clouds, s.kf_poses)])
    M = gpu_ctx.localmap_set(raw, s.map_leaf)
    map_ds = gpu_ctx.localmap_get()
    assert M == len(map_ds) and np.array_equal(map_ds, orc.voxelgrid(raw, s.map_leaf))
    ds = orc.voxelgrid(s.scan_raw, s.scan_leaf)
    q = orc.transform_cloud(ds, s.guess)
    n_acc = _knn_compare(gpu_ctx, orc, map_ds, q)
    assert n_acc > 0.5 * len(q)
    # queries far outside the map, and on the bounding box
    far = q.copy(); far[:, 0] += 1000
    assert _knn_compare(gpu_ctx, orc, map_ds, far) == 0
    edge = map_ds[:500].copy(); edge[:, :3] += np.float32(0.3)
    _knn_compare(gpu_ctx, orc, map_ds, edge)


def test_knn5_exact_ties_by_index(gpu_ctx, orc):
    # lattice with spacing 0.25: many exactly equal distances; leaf tiny so VoxelGrid keeps every point
    g = (np.arange(-12, 13) * 0.25).astype(np.float32)
    pts = np.stack(np.meshgrid(g, g, g[:6], indexing="ij"), -1).reshape(-1, 3)
    rng = np.random.default_rng(4)
    pts = pts[rng.permutation(len(pts))]
    cloud = np.c_[pts, np.zeros(len(pts))].astype(np.float32)
    M = gpu_ctx.localmap_set(cloud, 0.1)
    map_ds = gpu_ctx.localmap_get()
    assert M == len(cloud)
    q = np.c_[rng.integers(-12, 13, (400, 3)) * 0.25 + 0.125 * rng.integers(0, 2, (400, 3)), np.zeros(400)].astype(np.float32)
    q[:, 2] = (rng.integers(-12, -6, 400) * 0.25 + 0.125 * rng.integers(0, 2, 400)).astype(np.float32)
    assert _knn_compare(gpu_ctx, orc, map_ds, q) > 300


def test_knn5_pruned_by_a_bound_is_the_unbounded_search(gpu_ctx, orc, scene_hdl):
    """The registration kernel skips search cells whose nearest face lies beyond a bound on the fifth neighbour's distance
    (knn_grid.cuh).  With any valid bound -- down to the fifth distance itself, where a cell's lower bound can EQUAL the bound
    and ties are decided by the point index -- the neighbours must be those of the unbounded search, bit for bit."""
    s = scene_hdl
    raw = np.concatenate([orc.transform_cloud(c, p) for c, p in zip(s.kf_clouds, s.kf_poses)])
    gpu_ctx.localmap_set(raw, s.map_leaf)
    rng = np.random.default_rng(11)
    q = orc.transform_cloud(orc.voxelgrid(s.scan_raw, s.scan_leaf), s.guess)
    # lattice map with exact ties as a second case
    g = (np.arange(-12, 13) * 0.25).astype(np.float32)
    lat = np.stack(np.meshgrid(g, g, g[:6], indexing="ij"), -1).reshape(-1, 3)
    lat = np.c_[lat[rng.permutation(len(lat))], np.zeros(len(lat))].astype(np.float32)
    ql = np.c_[rng.integers(-12, 13, (400, 3)) * 0.25 + 0.125 * rng.integers(0, 2, (400, 3)), np.zeros(400)].astype(np.float32)
    ql[:, 2] = (rng.integers(-12, -6, 400) * 0.25 + 0.125 * rng.integers(0, 2, 400)).astype(np.float32)
    for cloud, leaf, queries in ((None, None, q), (lat, 0.1, ql)):
        if cloud is not None:
            gpu_ctx.localmap_set(cloud, leaf)
        idx, d2, found = gpu_ctx.knn5(queries)
        full = found == 5
        assert full.sum() > 100
        d5 = np.where(full, d2[:, 4], np.float32(1.0)).astype(np.float32)
        for bounds in (d5,                                                   # the tightest valid bound
                       np.nextafter(d5, np.float32(np.inf)),
                       d5 * np.float32(1.5),
                       np.where(rng.random(len(d5)) < 0.5, d5, np.float32(3.4e38))):
            b = np.where(full, bounds, np.float32(3.4e38)).astype(np.float32)
            idx_b, d2_b, found_b = gpu_ctx.knn5_bounded(queries, b)
            assert np.array_equal(found_b, found)
            assert np.array_equal(idx_b, idx) and np.array_equal(d2_b.view(np.uint32), d2.view(np.uint32))


def test_knn5_tiny_maps(gpu_ctx, orc):
    for n in (1, 4, 5, 6):
        rng = np.random.default_rng(n)
        cloud = np.c_[rng.uniform(-0.3, 0.3, (n, 3)), np.zeros(n)].astype(np.float32)
        gpu_ctx.localmap_set(cloud, 0.01)
        map_ds = gpu_ctx.localmap_get()
        q = np.c_[rng.uniform(-0.4, 0.4, (40, 3)), np.zeros(40)].astype(np.float32)
        _knn_compare(gpu_ctx, orc, map_ds, q)


def _load_scene(gpu_ctx, orc, s):
    gpu_ctx.keyframe_clear()
    for k, c in enumerate(s.kf_clouds):
        gpu_ctx.keyframe_put(k, c)
    ids = list(range(len(s.kf_clouds)))
    M, _ = gpu_ctx.localmap_build(ids, s.kf_poses, s.map_leaf)
    Q = gpu_ctx.scan_put(s.scan_raw, s.scan_leaf)
    map_ds = orc.localmap_build(s.kf_clouds, s.kf_poses, s.map_leaf, threads=8)
    ds = orc.voxelgrid(s.scan_raw, s.scan_leaf)
    assert M == len(map_ds) and Q == len(ds)
    assert np.array_equal(gpu_ctx.scan_ds_get(), ds)
    return map_ds, ds


@pytest.mark.parametrize("scene_name", ["scene_small", "scene_hdl"])
def test_surf_pass_bit_exact(gpu_ctx, orc, request, scene_name):
    s = request.getfixturevalue(scene_name)
    map_ds, ds = _load_scene(gpu_ctx, orc, s)
    for stride in (2, 1):
        coeff, flag = gpu_ctx.surf_pass(s.guess, stride, len(ds))
        rcoeff, rflag, *_ = orc.surf_pass(map_ds, ds, s.guess, stride=stride, kdtree=True, threads=8)
        assert np.array_equal(flag, rflag)
        assert flag.sum() > 100
        assert np.array_equal(coeff[flag == 1], rcoeff[rflag == 1])


@pytest.mark.parametrize("scene_name", ["scene_small", "scene_hdl"])
def test_scan2map_parity(gpu_ctx, orc, request, scene_name):
    s = request.getfixturevalue(scene_name)
    map_ds, ds = _load_scene(gpu_ctx, orc, s)
    pose, res, rc = gpu_ctx.scan2map(s.guess)
    rpose, rres, rrc, rst = orc.scan2map(map_ds, ds, s.guess, orc.S2mOpts(knn=1, threads=8))
    assert rc == rrc == 0
    assert (res.iters, res.converged, res.n_sel, res.is_degenerate, res.too_few) == \
           (rres.iters, rres.converged, rres.n_sel, rres.is_degenerate, rres.too_few)
    _assert_pose_close(pose, rpose)
    assert res.q_all == len(ds) and res.map_points == len(map_ds)
    assert np.abs(pose[3:] - s.truth[3:]).max() < 0.03


def test_scan2map_many_seeds(gpu_ctx, orc, synth):
    """Pose tolerance + exact iteration / correspondence counts over a spread of seeds and options."""
    from conftest import Scene
    mism = []
    for seed in range(20, 32):
        sensor = (synth.VLP16, synth.HDL32)[seed % 2]
        s = Scene(synth, orc, sensor, n_kf=5 + seed % 4, scan_leaf=(0.2, 0.4)[seed % 2], map_leaf=(0.5, 0.4)[seed % 2], seed=seed)
        map_ds, ds = _load_scene(gpu_ctx, orc, s)
        from liloc_b200 import S2mOpts
        for max_iters, stride in ((20, 2), (30, 1)):
            pose, res, rc = gpu_ctx.scan2map(s.guess, S2mOpts(max_iters=max_iters, stride=stride))
            rpose, rres, rrc, _ = orc.scan2map(map_ds, ds, s.guess, orc.S2mOpts(max_iters=max_iters, stride=stride, knn=1, threads=8))
            assert rc == rrc
            _assert_pose_close(pose, rpose)
            if (res.iters, res.n_sel, res.is_degenerate, res.converged) != (rres.iters, rres.n_sel, rres.is_degenerate, rres.converged):
                mism.append((seed, max_iters, stride, res.as_dict(), rres.as_dict()))
    assert not mism, mism


def test_scan2map_gates(gpu_ctx, orc, scene_small):
    from liloc_b200 import LILOC_SKIPPED
    s = scene_small
    map_ds, ds = _load_scene(gpu_ctx, orc, s)
    # fewer than 50 correspondences (:1080): pose unchanged, loop stops being useful
    far = s.guess.copy(); far[3] += 500
    pose, res, rc = gpu_ctx.scan2map(far)
    assert rc == 0 and res.too_few == 1 and res.converged == 0 and np.array_equal(pose, far)
    # fewer than 31 down-sampled points (:1187): skipped
    Q = gpu_ctx.scan_put(s.scan_raw[:25], s.scan_leaf)
    assert Q <= 25
    pose, res, rc = gpu_ctx.scan2map(s.guess)
    assert rc == LILOC_SKIPPED and res.skipped == 1 and np.array_equal(pose, s.guess)


def test_scan2map_degenerate_scene(gpu_ctx, orc, synth):
    world = np.array([[-500, -500, 0, 500, 500, 400]], np.float32)
    traj = synth.trajectory_line(5)
    sensor = synth.Sensor("down", 0, 16, 900, -40.0, -8.0, 0.5, 60.0)
    kfs = [orc.voxelgrid(synth.scan(world, traj[k], sensor, 40 + k), 0.4) for k in range(4)]
    scan_raw = synth.scan(world, traj[4], sensor, 49)
    guess = synth.perturb(traj[4], 3, dt=0.05, dr=0.01)
    gpu_ctx.keyframe_clear()
    for k, c in enumerate(kfs):
        gpu_ctx.keyframe_put(k, c)
    gpu_ctx.localmap_build(list(range(4)), traj[:4].astype(np.float32), 0.5)
    gpu_ctx.scan_put(scan_raw, 0.4)
    pose, res, rc = gpu_ctx.scan2map(guess)
    m = orc.localmap_build(kfs, traj[:4], 0.5)
    ds = orc.voxelgrid(scan_raw, 0.4)
    rpose, rres, _, rst = orc.scan2map(m, ds, guess)
    assert res.is_degenerate == rres.is_degenerate == 1
    assert (res.iters, res.n_sel) == (rres.iters, rres.n_sel)
    _assert_pose_close(pose, rpose)
    assert np.allclose(np.array(res.matP), np.array(rst.matP), atol=1e-5)


def test_frame_sequence_matches_oracle(gpu_ctx, orc, synth):
    """liloc_frame over a short lio sequence: each frame's result becomes the next keyframe pose
    (saveLIOKeyFramesAndFactor, src/mapOptmization.cpp:1391-1461), exactly as the host class drives it."""
    world = synth.make_world(seed=21)
    traj = synth.trajectory_line(12, step=1.0, x0=-6)
    sensor, scan_leaf, map_leaf = synth.VLP16, 0.4, 0.5
    gpu_ctx.keyframe_clear()
    first = orc.voxelgrid(synth.scan(world, traj[0], sensor, 7000), scan_leaf)
    gpu_ctx.keyframe_put(0, first)
    kf_clouds, kf_poses = [first], [traj[0].astype(np.float32)]
    o_clouds, o_poses = [first], [traj[0].astype(np.float32)]
    for k in range(1, 12):
        raw = synth.scan(world, traj[k], sensor, 7000 + k)
        guess = synth.perturb(traj[k], 900 + k, dt=0.05, dr=0.01)
        ids = list(range(len(kf_clouds)))
        pose, res, rc = gpu_ctx.frame(ids, np.asarray(kf_poses), map_leaf, raw, scan_leaf, guess)
        m = orc.localmap_build(o_clouds, np.asarray(o_poses), map_leaf, threads=8)
        ds = orc.voxelgrid(raw, scan_leaf)
        rpose, rres, rrc, _ = orc.scan2map(m, ds, guess, orc.S2mOpts(knn=1, threads=8))
        assert rc == rrc == 0
        assert (res.iters, res.n_sel, res.converged) == (rres.iters, rres.n_sel, rres.converged)
        _assert_pose_close(pose, rpose)
        gpu_ctx.keyframe_put_from_scan(k)                 # device-to-device copy of the DS scan
        kf_clouds.append(ds); kf_poses.append(pose)
        o_clouds.append(ds); o_poses.append(pose)         # the oracle continues from the SAME key poses
        assert gpu_ctx.keyframe_size(k) == len(ds)
    assert gpu_ctx.kernel_launches() > 0


def test_map_cache_is_output_identical(gpu_ctx, orc, synth):
    """liloc_config.map_cache: frames whose keyframe selection, poses and leaf are unchanged reuse the local map and
    its search grid; every output (pose bits, counters, the map itself) must equal the uncached run."""
    world = synth.make_world(seed=23)
    traj = synth.trajectory_line(9, step=0.3, x0=-2)
    sensor, scan_leaf, map_leaf = synth.VLP16, 0.4, 0.5
    kfs = [orc.voxelgrid(synth.scan(world, traj[k], sensor, 2300 + k), scan_leaf) for k in range(4)]
    scans = [synth.scan(world, traj[k], sensor, 2300 + k) for k in range(4, 9)]

    def run(cache):
        gpu_ctx.keyframe_clear()
        gpu_ctx.set_map_cache(False)
        hits0 = gpu_ctx.set_map_cache(cache)
        for k, c in enumerate(kfs):
            gpu_ctx.keyframe_put(k, c)
        ids, poses = [0, 1, 2, 3], traj[:4].astype(np.float32)
        out = []
        for j, raw in enumerate(scans):
            if j == 3:                                   # selection changes once: the cached map must be dropped
                ids, poses = [1, 2, 3], traj[1:4].astype(np.float32)
            pose, res, rc = gpu_ctx.frame(ids, poses, map_leaf, raw, scan_leaf, synth.perturb(traj[4 + j], 70 + j, dt=0.05, dr=0.01))
            out.append((pose.tobytes(), res.iters, res.n_sel, res.converged, res.map_points, res.q_all, gpu_ctx.localmap_get().tobytes()))
        return out, gpu_ctx.set_map_cache(False) - hits0

    ref, h0 = run(False)
    got, h1 = run(True)
    assert h0 == 0 and h1 == 3                           # frames 1, 2 (same selection) and 4 (same as frame 3)
    assert got == ref


def test_scan_prefetch_is_output_identical(gpu_ctx, orc, synth):
    """liloc_scan_prefetch: the next frame's scan uploaded behind the current frame's launches.  Every output equals the
    run without hints -- also when a hint names a scan no frame consumes, when the hinted array is not the one the next frame
    passes, and when one array is hinted twice."""
    world = synth.make_world(seed=29)
    traj = synth.trajectory_line(10, step=0.4, x0=-2)
    sensor, scan_leaf, map_leaf = synth.VLP16, 0.4, 0.5
    kfs = [orc.voxelgrid(synth.scan(world, traj[k], sensor, 2900 + k), scan_leaf) for k in range(3)]
    scans = [np.ascontiguousarray(synth.scan(world, traj[k], sensor, 2900 + k), dtype=np.float32) for k in range(3, 10)]
    decoy = np.ascontiguousarray(scans[0][::-1])

    def run(hints):
        gpu_ctx.keyframe_clear()
        for k, c in enumerate(kfs):
            gpu_ctx.keyframe_put(k, c)
        ids, poses = [0, 1, 2], traj[:3].astype(np.float32)
        gpu_ctx.stats_reset()
        out = []
        for j, raw in enumerate(scans):
            if hints and j + 1 < len(scans):
                nxt = scans[j + 1]
                if j == 2:
                    nxt = decoy                              # a hint the next frame does not honour
                gpu_ctx.scan_prefetch(nxt)
                if j == 4:
                    gpu_ctx.scan_prefetch(nxt)               # twice: still one upload
            pose, res, rc = gpu_ctx.frame(ids, poses, map_leaf, raw, scan_leaf, synth.perturb(traj[3 + j], 90 + j, dt=0.05, dr=0.01))
            out.append((rc, pose.tobytes(), res.iters, res.n_sel, res.converged, res.map_points, res.q_all, gpu_ctx.scan_ds_get().tobytes()))
        return out, gpu_ctx.stats()

    ref, st0 = run(False)
    got, st1 = run(True)
    assert got == ref
    scan_bytes = sum(len(s) for s in scans) * 16
    extra = st1["h2d_bytes"] - st0["h2d_bytes"]
    assert st0["h2d_bytes"] >= scan_bytes and extra == len(decoy) * 16      # only the unconsumed hint cost a copy


def test_share_is_output_identical(gpu_ctx, orc, synth):
    """liloc_set_share (throughput mode: 1 / S of the blocks per launch): the launch size changes, no output does --
    VoxelGrid, local map, search grid and the registration have fixed summation orders."""
    world = synth.make_world(seed=31)
    traj = synth.trajectory_line(8, step=0.4, x0=-2)
    sensor, scan_leaf, map_leaf = synth.HDL32, 0.4, 0.4
    kfs = [orc.voxelgrid(synth.scan(world, traj[k], sensor, 3100 + k), scan_leaf) for k in range(4)]
    scans = [synth.scan(world, traj[k], sensor, 3100 + k) for k in range(4, 8)]

    def run(share):
        assert gpu_ctx.set_share(share) >= 1
        gpu_ctx.keyframe_clear()
        for k, c in enumerate(kfs):
            gpu_ctx.keyframe_put(k, c)
        ids, poses = [0, 1, 2, 3], traj[:4].astype(np.float32)
        out = []
        for j, raw in enumerate(scans):
            pose, res, rc = gpu_ctx.frame(ids, poses, map_leaf, raw, scan_leaf, synth.perturb(traj[4 + j], 40 + j, dt=0.05, dr=0.01))
            out.append((rc, pose.tobytes(), res.iters, res.n_sel, res.converged, res.map_points, res.q_all,
                        gpu_ctx.localmap_get().tobytes(), gpu_ctx.scan_ds_get().tobytes()))
        return out

    try:
        ref = run(1)
        for share in (3, 8, 64, 1000):                      # 1000: one block per launch
            assert run(share) == ref, share
    finally:
        assert gpu_ctx.set_share(1) == 1000


def test_errors_are_codes_not_exceptions_from_c(gpu_ctx):
    from liloc_b200 import LilocError
    with pytest.raises(LilocError) as e:
        gpu_ctx.localmap_build([123456], np.zeros((1, 6), np.float32), 0.4)
    assert e.value.code < 0 and "unknown keyframe" in str(e.value)


def test_non_finite_input_is_rejected_and_the_context_survives(gpu_ctx, orc, synth, scene_small):
    """A NaN / Inf coordinate in any cloud that passes the point-cloud pipeline (keyframe, scan, VoxelGrid input, NDT target)
    is an argument error, never an out-of-range write: it leaves the bounding box but would still get a voxel key and a search
    cell.  The map is put far from the origin (cell_coord(NaN) = -g0 is then out of range).  After each refused call the same
    context runs a good frame and reproduces the earlier result bit for bit.  The de-skew filters drop such points."""
    from liloc_b200 import LilocError, NdtOpts
    sc = scene_small
    shift = np.array([0, 0, 0, 700.0, -450.0, 30.0], np.float32)
    poses = sc.kf_poses + shift
    guess = (sc.guess + shift).astype(np.float32)
    K = len(sc.kf_clouds)

    def good_frame():
        gpu_ctx.keyframe_clear()
        for k, c in enumerate(sc.kf_clouds):
            gpu_ctx.keyframe_put(k, c)
        pose, res, rc = gpu_ctx.frame(list(range(K)), poses, sc.map_leaf, sc.scan_raw, sc.scan_leaf, guess)
        return pose, (res.iters, res.n_sel, res.converged, res.map_points, res.q_all), rc

    p0, r0, rc0 = good_frame()
    assert rc0 == 0 and r0[0] > 0
    for bad_value in (np.nan, np.inf, -np.inf):
        # (1) a bad point in a keyframe cloud
        bad_kf = sc.kf_clouds[2].copy(); bad_kf[len(bad_kf) // 2, 1] = bad_value
        gpu_ctx.keyframe_put(900, bad_kf)
        with pytest.raises(LilocError) as e:
            gpu_ctx.frame([0, 1, 900, 3], poses[[0, 1, 2, 3]], sc.map_leaf, sc.scan_raw, sc.scan_leaf, guess)
        assert e.value.code == -1 and "finite" in str(e.value)
        with pytest.raises(LilocError):                       # a failed frame leaves no usable map / scan behind
            gpu_ctx.scan2map(guess)
        assert good_frame() [1] == r0
        # (2) a bad point in the scan
        bad_scan = sc.scan_raw.copy(); bad_scan[7, 2] = bad_value
        with pytest.raises(LilocError) as e:
            gpu_ctx.frame(list(range(K)), poses, sc.map_leaf, bad_scan, sc.scan_leaf, guess)
        assert e.value.code == -1
        # (3) kernel-level VoxelGrid and an NDT target
        with pytest.raises(LilocError):
            gpu_ctx.voxelgrid(bad_scan, 0.4)
        with pytest.raises(LilocError):
            gpu_ctx.ndt_target_put(77, bad_scan, 1.0)
        with pytest.raises(LilocError):                       # the failed build left no half-made target behind
            gpu_ctx.ndt_align_batch([77], [77], [np.eye(4, dtype=np.float32)], NdtOpts())
        p1, r1, rc1 = good_frame()
        assert rc1 == 0 and r1 == r0 and np.array_equal(p1, p0)
    assert np.array_equal(gpu_ctx.voxelgrid(sc.scan_raw, 0.4), orc.voxelgrid(sc.scan_raw, 0.4))
    # (4) de-skew: the range filter refuses non-finite points
    sw = synth.raw_sweep(sc.scan_raw, synth.VLP16, seed=3)
    it, ir = orc.imu_integrate(sw["imu_stamps"], sw["imu_gyro"], sw["time_scan_end"])
    n_ok = gpu_ctx.deskew(sw["pts"], sw["ring"], sw["time"], it, ir, sw["time_scan_cur"], n_scan=16, point_filter_num=1)
    pts = sw["pts"].copy(); pts[5, 0] = np.nan; pts[6, 1] = np.inf; pts[9, 2] = -np.inf
    keep = np.ones(len(pts), bool); keep[[5, 6, 9]] = False
    n_bad = gpu_ctx.deskew(pts, sw["ring"], sw["time"], it, ir, sw["time_scan_cur"], n_scan=16, point_filter_num=1)
    ref = orc.deskew(sw["pts"], sw["ring"], sw["time"], it, ir, sw["time_scan_cur"], n_scan=16, point_filter_num=1)
    kept_ref = orc.deskew(sw["pts"][keep], sw["ring"][keep], sw["time"][keep], it, ir, sw["time_scan_cur"], n_scan=16, point_filter_num=1)
    assert n_ok == len(ref) and n_bad == len(kept_ref) and np.isfinite(gpu_ctx.deskewed_get()).all()
