"""CPU tests of the oracle's corner path (upstream LIO-SAM cornerOptimization).

The reference has no corner code at all, so there is nothing of the reference to pin this
against; the checks are independent ones: cv2.eigen known answers (tests/golden/corner3x3_cv2.npz), the geometry
of the point-to-line residual, and end-to-end recovery of the ground-truth pose on a synthetic scene.
"""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_eigen3_golden_cv2(orc):
    g = np.load(os.path.join(GOLD, "corner3x3_cv2.npz"))
    for A, w_ref, V_ref in zip(g["A"], g["w"], g["V"]):
        w, V = orc.eigen3(A)
        scale = max(abs(float(w_ref[0])), 1e-6)
        assert np.all(np.diff(w) <= 0)                                  # descending
        assert np.abs(w - w_ref).max() <= 2e-6 * scale + 1e-9
        assert np.abs(V @ V.T - np.eye(3)).max() < 1e-5                 # rows orthonormal
        assert np.abs(V @ A @ V.T - np.diag(w)).max() <= 5e-6 * scale + 1e-9
        if w_ref[0] > 3 * w_ref[1]:                                     # principal direction, up to sign
            assert abs(abs(float(V[0] @ V_ref[0])) - 1.0) < 1e-4


def test_line_residual_is_point_to_line_distance(orc):
    rng = np.random.default_rng(5)
    kept = 0
    for _ in range(300):
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        c = rng.uniform(-40, 40, 3)
        pts = (c + rng.uniform(-0.8, 0.8, (5, 1)) * d + rng.normal(0, 0.005, (5, 3))).astype(np.float32)
        sel = (c + rng.uniform(-0.5, 0.5) * d + rng.normal(0, 0.15, 3)).astype(np.float32)
        coeff, keep = orc.line_residual(pts, sel)
        # independent fit in float64
        m = pts.astype(np.float64).mean(0)
        _, _, vt = np.linalg.svd(pts.astype(np.float64) - m)
        v = vt[0]
        r = sel.astype(np.float64) - m
        perp = r - (r @ v) * v
        dist = np.linalg.norm(perp)
        s = 1 - 0.9 * dist
        assert keep == (1 if s > 0.1 else 0) or abs(s - 0.1) < 1e-3
        if keep and dist > 1e-3:
            kept += 1
            assert abs(coeff[3] - s * dist) < 2e-3
            n = coeff[:3] / np.linalg.norm(coeff[:3])
            assert np.linalg.norm(n - perp / dist) < 2e-2               # unit vector from the line to the point
            assert abs(np.linalg.norm(coeff[:3]) - s) < 2e-3
    assert kept > 150


def test_line_residual_rejects_non_lines(orc):
    rng = np.random.default_rng(6)
    rej = 0
    for _ in range(100):
        pts = (rng.uniform(-40, 40, 3) + rng.normal(0, 0.3, (5, 3))).astype(np.float32)     # isotropic blob
        _, keep = orc.line_residual(pts, pts.mean(0) + 0.1)
        w = np.linalg.eigvalsh(np.cov(pts.astype(np.float64).T, bias=True))[::-1]
        if not (w[0] > 3 * w[1] * 1.01 or w[0] < 3 * w[1] * 0.99):
            continue
        assert keep == 0 or w[0] > 3 * w[1]
        rej += (keep == 0)
    assert rej > 20


def _feature_scene(synth, orc, seed=1, n_kf=10, tol=0.12):
    world = synth.make_world(seed=seed)
    traj = synth.trajectory_line(n_kf + 1)
    kfc, kfs = [], []
    for k in range(n_kf):
        c, s = synth.split_features(world, traj[k], synth.scan(world, traj[k], synth.VLP16, 10 + k), tol)
        kfc.append(orc.voxelgrid(c, 0.2)); kfs.append(orc.voxelgrid(s, 0.4))
    c, s = synth.split_features(world, traj[n_kf], synth.scan(world, traj[n_kf], synth.VLP16, 999), tol)
    return world, traj, kfc, kfs, c, s


def test_scan2map_features_recovers_ground_truth(orc, synth):
    world, traj, kfc, kfs, c, s = _feature_scene(synth, orc)
    mc = orc.localmap_build(kfc, traj[:10], 0.2)
    ms = orc.localmap_build(kfs, traj[:10], 0.4)
    cd, sd = orc.voxelgrid(c, 0.2), orc.voxelgrid(s, 0.4)
    assert len(cd) > 100 and len(sd) > 1000
    guess = synth.perturb(traj[10], 5)
    coef, flag = orc.corner_pass(mc, cd, guess)
    assert flag.sum() > 0.5 * len(cd) and np.isfinite(coef).all()
    for knn in (0, 1):
        pose, res, rc, _ = orc.scan2map_features(ms, sd, mc, cd, guess, orc.upstream_opts(knn=knn, threads=4))
        assert rc == 0 and res.converged == 1 and res.iters <= 30
        assert res.n_sel > len(sd) // 2 + flag.sum() // 2
        assert np.abs(pose[3:] - traj[10][3:]).max() < 0.03 and np.abs(pose[:3] - traj[10][:3]).max() < 5e-3
        if knn == 0:
            first = (pose.copy(), res.as_dict())
    assert np.array_equal(first[0], pose) and first[1] == res.as_dict()      # kd-tree == brute force


def test_scan2map_features_gate(orc, synth):
    world, traj, kfc, kfs, c, s = _feature_scene(synth, orc, n_kf=4)
    mc = orc.localmap_build(kfc, traj[:4], 0.2)
    ms = orc.localmap_build(kfs, traj[:4], 0.4)
    cd, sd = orc.voxelgrid(c, 0.2), orc.voxelgrid(s, 0.4)
    guess = synth.perturb(traj[4], 5)
    pose, res, rc, _ = orc.scan2map_features(ms, sd, mc, cd[:10], guess)        # corner gate: > 10
    assert rc == 1 and res.skipped == 1 and np.array_equal(pose, guess)
    pose, res, rc, _ = orc.scan2map_features(ms, sd[:100], mc, cd, guess)       # surf gate: > 100
    assert rc == 1 and res.skipped == 1
