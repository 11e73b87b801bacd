"""GPU parity tests of the corner + surf superset (upstream LIO-SAM cornerOptimization; the north_star names it,
the reference dropped it -- SURVEY.md 0.2 D1, Appendix B).  Same bars as test_gpu_parity.py: per-point work is
BIT-EXACT against the oracle, poses within 1e-4 m / 1e-5 rad, counters exact.  Everything goes through the C ABI.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SURF, CORNER = 0, 1
LEAF_C, LEAF_S, MAP_LEAF_C, MAP_LEAF_S = 0.2, 0.4, 0.2, 0.4     # upstream mappingCornerLeafSize / mappingSurfLeafSize


def _assert_pose_close(got, ref):
    assert np.abs(got[3:] - ref[3:]).max() <= 1e-4, (got, ref)
    assert np.abs(got[:3] - ref[:3]).max() <= 1e-5, (got, ref)


class FeatScene:
    def __init__(self, synth, orc, seed, n_kf, sensor=None, tol=0.12):
        sensor = sensor or synth.VLP16
        self.world = synth.make_world(seed=seed)
        self.traj = synth.trajectory_line(n_kf + 1, x0=-n_kf / 2)
        self.kfc, self.kfs = [], []
        for k in range(n_kf):
            c, s = synth.split_features(self.world, self.traj[k], synth.scan(self.world, self.traj[k], sensor, seed * 100 + k), tol)
            self.kfc.append(orc.voxelgrid(c, LEAF_C)); self.kfs.append(orc.voxelgrid(s, LEAF_S))
        self.poses = self.traj[:n_kf].astype(np.float32)
        self.c_raw, self.s_raw = synth.split_features(self.world, self.traj[n_kf], synth.scan(self.world, self.traj[n_kf], sensor, seed * 100 + 99), tol)
        self.guess = synth.perturb(self.traj[n_kf], seed)
        self.truth = self.traj[n_kf]


@pytest.fixture(scope="module")
def fscene(synth, orc):
    return FeatScene(synth, orc, seed=31, n_kf=8)


def _load(gpu_ctx, orc, s):
    gpu_ctx.keyframe_clear()
    for k in range(len(s.kfs)):
        gpu_ctx.keyframe_put_features(k, s.kfs[k], s.kfc[k])
    ids = list(range(len(s.kfs)))
    Ms, _ = gpu_ctx.localmap_build(ids, s.poses, MAP_LEAF_S, cls=SURF)
    Mc, _ = gpu_ctx.localmap_build(ids, s.poses, MAP_LEAF_C, cls=CORNER)
    Qs = gpu_ctx.scan_put(s.s_raw, LEAF_S, cls=SURF)
    Qc = gpu_ctx.scan_put(s.c_raw, LEAF_C, cls=CORNER)
    ms = orc.localmap_build(s.kfs, s.poses, MAP_LEAF_S, threads=8)
    mc = orc.localmap_build(s.kfc, s.poses, MAP_LEAF_C, threads=8)
    sd, cd = orc.voxelgrid(s.s_raw, LEAF_S), orc.voxelgrid(s.c_raw, LEAF_C)
    assert (Ms, Mc, Qs, Qc) == (len(ms), len(mc), len(sd), len(cd))
    assert np.array_equal(gpu_ctx.localmap_get(cls=CORNER), mc) and np.array_equal(gpu_ctx.scan_ds_get(cls=CORNER), cd)
    assert np.array_equal(gpu_ctx.localmap_get(cls=SURF), ms) and np.array_equal(gpu_ctx.scan_ds_get(cls=SURF), sd)
    return ms, mc, sd, cd


def test_eigen3_and_line_residual_bit_exact(gpu_ctx, orc):
    g = np.load(os.path.join(GOLD, "corner3x3_cv2.npz"))
    rng = np.random.default_rng(9)
    extra = []
    for _ in range(300):
        B = rng.normal(size=(5, 3)) * rng.uniform(0.01, 2.0, 3)
        extra.append((B.T @ B / 5).astype(np.float32))
    extra.append(np.zeros((3, 3), np.float32)); extra.append(np.eye(3, dtype=np.float32))
    A = np.concatenate([g["A"], np.stack(extra)])
    w, V = gpu_ctx.debug_eigen3(A)
    for i, Ai in enumerate(A):
        rw, rV = orc.eigen3(Ai)
        assert np.array_equal(w[i], rw) and np.array_equal(V[i], rV), i
    pts, sel = [], []
    for _ in range(600):
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        c = rng.uniform(-60, 60, 3)
        spread = rng.choice([0.01, 0.05, 0.3])
        pts.append((c + rng.uniform(-0.8, 0.8, (5, 1)) * d + rng.normal(0, spread, (5, 3))).astype(np.float32))
        sel.append((c + rng.normal(0, 0.3, 3)).astype(np.float32))
    coeff, keep = gpu_ctx.debug_line(np.stack(pts), np.stack(sel))
    n_keep = 0
    for i in range(len(pts)):
        rc, rk = orc.line_residual(pts[i], sel[i])
        assert keep[i] == rk, i
        assert np.array_equal(coeff[i], rc), (i, coeff[i], rc)
        n_keep += rk
    assert 100 < n_keep < 600


def test_corner_knn_and_pass_bit_exact(gpu_ctx, orc, fscene):
    ms, mc, sd, cd = _load(gpu_ctx, orc, fscene)
    q = orc.transform_cloud(cd, fscene.guess)
    idx, d2, found = gpu_ctx.knn5(q, cls=CORNER)
    ridx, rd2 = orc.knn5(mc, q, kdtree=True, threads=8)
    full = found == 5
    assert full.sum() > 50
    assert np.array_equal(idx[full], ridx[full]) and np.array_equal(d2[full], rd2[full])
    assert np.all(rd2[~full][:, 4] >= 1.0)
    for stride in (1, 2):
        coeff, flag = gpu_ctx.feature_pass(CORNER, fscene.guess, stride, len(cd))
        rcoeff, rflag = orc.corner_pass(mc, cd, fscene.guess, stride=stride, kdtree=True, threads=8)
        assert np.array_equal(flag, rflag) and flag.sum() > 30
        assert np.array_equal(coeff[flag == 1], rcoeff[rflag == 1])
    coeff, flag = gpu_ctx.feature_pass(SURF, fscene.guess, 1, len(sd))
    rcoeff, rflag, *_ = orc.surf_pass(ms, sd, fscene.guess, stride=1, kdtree=True, threads=8)
    assert np.array_equal(flag, rflag) and np.array_equal(coeff[flag == 1], rcoeff[rflag == 1])


def test_scan2map_features_parity(gpu_ctx, orc, fscene):
    from liloc_b200 import S2mOpts
    ms, mc, sd, cd = _load(gpu_ctx, orc, fscene)
    pose, res, rc = gpu_ctx.scan2map(fscene.guess, S2mOpts.upstream())
    rpose, rres, rrc, _ = orc.scan2map_features(ms, sd, mc, cd, fscene.guess, orc.upstream_opts(knn=1, threads=8))
    assert rc == rrc == 0
    assert (res.iters, res.converged, res.n_sel, res.is_degenerate, res.too_few) == \
           (rres.iters, rres.converged, rres.n_sel, rres.is_degenerate, rres.too_few)
    _assert_pose_close(pose, rpose)
    ex = gpu_ctx.last_s2m_extra()
    assert ex["q_corner"] == len(cd) and ex["map_points_corner"] == len(mc) and ex["tiles_corner"] > 0
    assert np.abs(pose[3:] - fscene.truth[3:]).max() < 0.03
    # corners off on the same context state: the reference's surf-only loop, unaffected by the corner class
    pose0, res0, _ = gpu_ctx.scan2map(fscene.guess, S2mOpts(max_iters=30, stride=1, min_points=100))
    rpose0, rres0, _, _ = orc.scan2map(ms, sd, fscene.guess, orc.S2mOpts(max_iters=30, stride=1, min_points=100, knn=1, threads=8))
    assert (res0.iters, res0.n_sel) == (rres0.iters, rres0.n_sel) and res0.n_sel < res.n_sel
    _assert_pose_close(pose0, rpose0)


def test_features_gates(gpu_ctx, orc, fscene):
    from liloc_b200 import S2mOpts, LILOC_SKIPPED
    _load(gpu_ctx, orc, fscene)
    Qc = gpu_ctx.scan_put(fscene.c_raw[:8], LEAF_C, cls=CORNER)
    assert Qc <= 10
    pose, res, rc = gpu_ctx.scan2map(fscene.guess, S2mOpts.upstream())
    assert rc == LILOC_SKIPPED and res.skipped == 1 and np.array_equal(pose, fscene.guess)
    pose, res, rc = gpu_ctx.scan2map(fscene.guess, S2mOpts(max_iters=30, stride=1, min_points=100))     # surf-only still runs
    assert rc == 0 and res.iters > 0


def test_frame_features_sequence_matches_oracle(gpu_ctx, orc, synth):
    """liloc_frame_features over a short sequence with keyframes appended device-to-device, both classes."""
    from liloc_b200 import S2mOpts
    world = synth.make_world(seed=41)
    traj = synth.trajectory_line(10, x0=-5)
    gpu_ctx.keyframe_clear()
    c0, s0 = synth.split_features(world, traj[0], synth.scan(world, traj[0], synth.VLP16, 4100))
    kfc, kfs, poses = [orc.voxelgrid(c0, LEAF_C)], [orc.voxelgrid(s0, LEAF_S)], [traj[0].astype(np.float32)]
    gpu_ctx.keyframe_put_features(0, kfs[0], kfc[0])
    for k in range(1, 10):
        c, s = synth.split_features(world, traj[k], synth.scan(world, traj[k], synth.VLP16, 4100 + k))
        guess = synth.perturb(traj[k], 500 + k, dt=0.05, dr=0.01)
        ids = list(range(len(kfs)))
        pose, res, rc = gpu_ctx.frame_features(ids, np.asarray(poses), (MAP_LEAF_S, MAP_LEAF_C), (s, c), (LEAF_S, LEAF_C), guess, S2mOpts.upstream())
        ms = orc.localmap_build(kfs, np.asarray(poses), MAP_LEAF_S, threads=8)
        mc = orc.localmap_build(kfc, np.asarray(poses), MAP_LEAF_C, threads=8)
        sd, cd = orc.voxelgrid(s, LEAF_S), orc.voxelgrid(c, LEAF_C)
        rpose, rres, rrc, _ = orc.scan2map_features(ms, sd, mc, cd, guess, orc.upstream_opts(knn=1, threads=8))
        assert rc == rrc == 0, (k, rc, rrc)
        assert (res.iters, res.n_sel, res.converged) == (rres.iters, rres.n_sel, rres.converged), (k, res.as_dict(), rres.as_dict())
        _assert_pose_close(pose, rpose)
        gpu_ctx.keyframe_put_from_scan(k)
        assert gpu_ctx.keyframe_size(k, SURF) == len(sd) and gpu_ctx.keyframe_size(k, CORNER) == len(cd)
        kfs.append(sd); kfc.append(cd); poses.append(pose)
