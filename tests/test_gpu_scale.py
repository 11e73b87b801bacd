"""GPU parity at the sizes BASELINE.json names: configs[0] VLP-16 against a 50-keyframe local map
(reference loop shape and the upstream corner + surf superset, 30 iterations) and configs[2] Mid-360-shaped dense
scans against a local map of ~1 M points (130 keyframes, leaf 0.2).  Full-size comparisons against the oracle
(which needs a few seconds on the CPU for these), plus size-independent properties of the VoxelGrid output."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _assert_pose_close(got, ref):
    assert np.abs(got[3:] - ref[3:]).max() <= 1e-4, (got, ref)
    assert np.abs(got[:3] - ref[:3]).max() <= 1e-5, (got, ref)


def _voxelgrid_properties(cloud_in, out, leaf):
    """Count conservation; every output point is the mean of exactly the input points of one voxel; outputs are in
    ascending voxel index (independent numpy restatement, float64 sums)."""
    inv = np.float32(1.0) / np.float32(leaf)
    ijk = np.floor(cloud_in[:, :3] * inv).astype(np.int64)
    mn = ijk.min(0); ijk -= mn
    dims = ijk.max(0) + 1
    key = ijk[:, 0] + ijk[:, 1] * dims[0] + ijk[:, 2] * dims[0] * dims[1]
    uk, cnt = np.unique(key, return_counts=True)
    assert len(out) == len(uk)
    sums = np.zeros((len(uk), 3)); np.add.at(sums, np.searchsorted(uk, key), cloud_in[:, :3].astype(np.float64))
    # row i of `sums` belongs to the i-th voxel in ascending index: matching it row by row also proves the output order
    assert np.abs(sums / cnt[:, None] - out[:, :3]).max() < 1e-3


def test_cfg0_vlp16_50_keyframes_reference_and_superset(gpu_ctx, orc, synth):
    from liloc_b200 import S2mOpts
    world = synth.make_world(size_xy=(120.0, 80.0), n_pillars=40, seed=1000)
    K = 50
    traj = synth.trajectory_line(K + 1, step=1.0, x0=-K / 2)
    gpu_ctx.keyframe_clear()
    kfs, kfc = [], []
    for k in range(K):
        c, s = synth.split_features(world, traj[k], synth.scan(world, traj[k], synth.VLP16, 1000 + k), 0.15)
        kfs.append(orc.voxelgrid(np.concatenate([c, s]), 0.2)); kfc.append(orc.voxelgrid(c, 0.2))
    # (a) reference semantics: the whole scan is the surf set, stride 2, 20 iterations (lio_sam_default.yaml leaves)
    for k in range(K):
        gpu_ctx.keyframe_put(k, kfs[k])
    raw = synth.scan(world, traj[K], synth.VLP16, 1999)
    guess = synth.perturb(traj[K], 1000)
    ids, poses = list(range(K)), traj[:K].astype(np.float32)
    pose, res, rc = gpu_ctx.frame(ids, poses, 0.5, raw, 0.2, guess)
    m = orc.localmap_build(kfs, poses, 0.5, threads=8)
    ds = orc.voxelgrid(raw, 0.2)
    assert res.map_points == len(m) and res.q_all == len(ds)
    assert np.array_equal(gpu_ctx.localmap_get(), m)
    rpose, rres, rrc, _ = orc.scan2map(m, ds, guess, orc.S2mOpts(knn=1, threads=8))
    assert rc == rrc == 0 and (res.iters, res.n_sel, res.converged) == (rres.iters, rres.n_sel, rres.converged)
    _assert_pose_close(pose, rpose)
    # (b) north_star superset: corner + surf features, stride 1, 30 iterations
    gpu_ctx.keyframe_clear()
    kfs2 = []
    for k in range(K):
        c, s = synth.split_features(world, traj[k], synth.scan(world, traj[k], synth.VLP16, 1000 + k), 0.15)
        kfs2.append(orc.voxelgrid(s, 0.4))
        gpu_ctx.keyframe_put_features(k, kfs2[k], kfc[k])
    c, s = synth.split_features(world, traj[K], raw, 0.15)
    pose, res, rc = gpu_ctx.frame_features(ids, poses, (0.4, 0.2), (s, c), (0.4, 0.2), guess, S2mOpts.upstream())
    ms = orc.localmap_build(kfs2, poses, 0.4, threads=8); mc = orc.localmap_build(kfc, poses, 0.2, threads=8)
    sd, cd = orc.voxelgrid(s, 0.4), orc.voxelgrid(c, 0.2)
    ex = gpu_ctx.last_s2m_extra()
    assert (res.map_points, res.q_all, ex["map_points_corner"], ex["q_corner"]) == (len(ms), len(sd), len(mc), len(cd))
    rpose, rres, rrc, _ = orc.scan2map_features(ms, sd, mc, cd, guess, orc.upstream_opts(knn=1, threads=8))
    assert rc == rrc == 0 and (res.iters, res.n_sel, res.converged) == (rres.iters, rres.n_sel, rres.converged)
    _assert_pose_close(pose, rpose)


def test_cfg2_mid360_million_point_map(gpu_ctx, orc, synth):
    """~1 M-point local map (leaf 0.2, 130 keyframes 0.5 m apart): local map bit-exact, kNN / registration parity."""
    world = synth.make_world(size_xy=(200.0, 150.0), height=25.0, n_pillars=60, seed=3000)
    K = 130
    traj = synth.trajectory_line(K + 1, step=0.5, x0=-K / 4)
    sensor = synth.Sensor("mid360x", 1, 1, 40000, -7.0, 52.0, 0.5, 100.0)      # denser than one Mid-360 frame so that M reaches ~1 M
    kfs = [synth.scan(world, traj[k], sensor, 3000 + k) for k in range(K)]       # un-thinned keyframes (relo mode keeps every frame)
    gpu_ctx.keyframe_clear()
    for k in range(K):
        gpu_ctx.keyframe_put(k, kfs[k])
    raw = synth.scan(world, traj[K], synth.MID360, 3999)
    guess = synth.perturb(traj[K], 3000)
    ids, poses = list(range(K)), traj[:K].astype(np.float32)
    pose, res, rc = gpu_ctx.frame(ids, poses, 0.2, raw, 0.2, guess)
    n_raw = sum(len(c) for c in kfs)
    assert n_raw > 3_000_000 and res.map_points > 700_000, (n_raw, res.map_points)
    gm = gpu_ctx.localmap_get()
    raw_map = np.concatenate([orc.transform_cloud(kfs[k], poses[k], threads=8) for k in range(K)])
    _voxelgrid_properties(raw_map, gm, 0.2)
    m = orc.localmap_build(kfs, poses, 0.2, threads=8)
    assert np.array_equal(gm, m)                                                  # bit-exact at full size
    ds = orc.voxelgrid(raw, 0.2)
    assert np.array_equal(gpu_ctx.scan_ds_get(), ds)
    q = orc.transform_cloud(ds[::2], guess)
    idx, d2, found = gpu_ctx.knn5(q)
    ridx, rd2 = orc.knn5(m, q, kdtree=True, threads=8)
    full = found == 5
    assert full.sum() > 1000 and np.array_equal(idx[full], ridx[full]) and np.array_equal(d2[full], rd2[full])
    rpose, rres, rrc, _ = orc.scan2map(m, ds, guess, orc.S2mOpts(knn=1, threads=8))
    assert rc == rrc == 0 and (res.iters, res.n_sel, res.converged) == (rres.iters, rres.n_sel, rres.converged)
    _assert_pose_close(pose, rpose)
