"""CPU tests: the oracle (oracle/liloc_oracle.cpp) against independent known answers.

The reference ships no tests or golden vectors for this path ("parity unpinned"),
so the oracle is pinned here against libraries available in the container: cv2 for the 6x6 algebra
(the library the reference itself calls), scipy.cKDTree for exact kNN, numpy for the plane fit and a
finite-difference check of the Jacobian.  Fixtures: tests/golden/*.npz (made by make_golden.py).
"""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_pose_to_T_matches_rz_ry_rx(orc, synth):
    rng = np.random.default_rng(0)
    for _ in range(50):
        pose = np.concatenate([rng.uniform(-3.1, 3.1, 3), rng.uniform(-50, 50, 3)]).astype(np.float32)
        T = orc.pose_to_T(pose)
        R = synth.rot_zyx(*pose[:3].astype(np.float64))
        assert np.allclose(T[:, :3], R, atol=3e-7)
        assert np.array_equal(T[:, 3], pose[3:])


def test_transform_cloud_left_to_right_float32(orc):
    rng = np.random.default_rng(1)
    pts = rng.uniform(-40, 40, (1000, 4)).astype(np.float32)
    pose = np.array([0.02, -0.01, 1.3, 4.0, -7.0, 1.8], np.float32)
    T = orc.pose_to_T(pose)
    out = orc.transform_cloud(pts, pose, threads=4)
    f = np.float32
    for i in range(0, 1000, 97):
        x, y, z = pts[i, :3]
        for r in range(3):
            e = f(f(f(T[r, 0] * x) + f(T[r, 1] * y)) + f(T[r, 2] * z)) + T[r, 3]
            assert out[i, r] == f(e)
        assert out[i, 3] == pts[i, 3]


def _voxel_reference(pts, leaf):
    """numpy restatement of SURVEY A.1 used to cross-check the C++ oracle."""
    f = np.float32
    inv = f(1.0) / f(leaf)
    mn = pts[:, :3].min(0); mx = pts[:, :3].max(0)
    minb = np.floor(mn * inv).astype(np.int64); maxb = np.floor(mx * inv).astype(np.int64)
    div = maxb - minb + 1
    ijk = (np.floor(pts[:, :3] * inv) - minb.astype(np.float32)).astype(np.int64)
    key = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    order = np.argsort(key, kind="stable")
    out = []
    ks = key[order]
    starts = np.flatnonzero(np.r_[True, ks[1:] != ks[:-1]])
    ends = np.r_[starts[1:], len(ks)]
    for s, e in zip(starts, ends):
        acc = np.zeros(4, np.float32)
        for j in order[s:e]:
            acc = (acc + pts[j]).astype(np.float32)
        out.append(acc / f(e - s))
    return np.asarray(out, np.float32)


@pytest.mark.parametrize("n,leaf,seed", [(1, 0.2, 0), (17, 0.5, 1), (5000, 0.4, 2), (20000, 0.2, 3)])
def test_voxelgrid_matches_numpy_restatement(orc, n, leaf, seed):
    rng = np.random.default_rng(seed)
    pts = (rng.normal(size=(n, 4)) * np.array([8, 5, 1, 50])).astype(np.float32)
    got = orc.voxelgrid(pts, leaf)
    ref = _voxel_reference(pts, leaf)
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)


def test_voxelgrid_properties(orc, synth):
    w = synth.make_world(seed=5)
    s = synth.scan(w, synth.trajectory_line(1)[0], synth.VLP16, 9)
    out = orc.voxelgrid(s, 0.2)
    assert 0 < len(out) < len(s)
    # centroids stay inside the input bounding box; empty input -> empty output
    assert np.all(out[:, :3].min(0) >= s[:, :3].min(0)) and np.all(out[:, :3].max(0) <= s[:, :3].max(0))
    assert len(orc.voxelgrid(np.zeros((0, 4), np.float32), 0.2)) == 0
    # idempotence up to voxel membership: filtering the output again keeps the count
    assert len(orc.voxelgrid(out, 0.2)) <= len(out)


def test_voxelgrid_overflow_guard_returns_input(orc):
    pts = np.array([[0, 0, 0, 1], [3000, 3000, 3000, 2], [1, 2, 3, 4]], np.float32)
    out = orc.voxelgrid(pts, 0.001)          # (3e6)^3 voxels > INT32_MAX
    assert np.array_equal(out, pts)


def test_knn5_golden_ckdtree(orc):
    g = np.load(os.path.join(GOLD, "knn5_ckdtree.npz"))
    m = np.c_[g["map"], np.zeros(len(g["map"]), np.float32)]
    q = np.c_[g["q"], np.zeros(len(g["q"]), np.float32)]
    for kd in (False, True):
        idx, d2 = orc.knn5(m, q, kdtree=kd, threads=2)
        assert np.array_equal(idx, g["idx"])
        assert np.allclose(np.sqrt(d2.astype(np.float64)), g["dist"], rtol=1e-5, atol=1e-6)


def test_knn5_ties_broken_by_index_and_kdtree_equals_brute(orc):
    # integer lattice: masses of exact float ties
    g = np.arange(-6, 7, dtype=np.float32)
    X, Y, Z = np.meshgrid(g, g, g[:5], indexing="ij")
    m = np.c_[X.ravel(), Y.ravel(), Z.ravel(), np.zeros(X.size, np.float32)].astype(np.float32)
    rng = np.random.default_rng(3)
    m = m[rng.permutation(len(m))]
    q = np.c_[rng.integers(-6, 7, (200, 3)).astype(np.float32) + np.float32(0.5) * rng.integers(0, 2, (200, 3)).astype(np.float32),
              np.zeros(200, np.float32)].astype(np.float32)
    ib, db = orc.knn5(m, q, kdtree=False)
    ik, dk = orc.knn5(m, q, kdtree=True)
    assert np.array_equal(ib, ik) and np.array_equal(db, dk)
    # (d2, index) ascending inside every answer
    for r in range(len(q)):
        pairs = list(zip(db[r], ib[r]))
        assert pairs == sorted(pairs)
    # python brute force with the same total order
    for r in range(0, 200, 17):
        d = ((q[r, 0] - m[:, 0]) ** 2 + (q[r, 1] - m[:, 1]) ** 2) + (q[r, 2] - m[:, 2]) ** 2
        order = np.lexsort((np.arange(len(m)), d))[:5]
        assert np.array_equal(order.astype(np.int32), ib[r])


def test_knn5_fewer_than_five_points(orc):
    m = np.array([[0, 0, 0, 0], [1, 0, 0, 0], [0, 1, 0, 0]], np.float32)
    q = np.array([[0.1, 0.1, 0, 0]], np.float32)
    for kd in (False, True):
        idx, d2 = orc.knn5(m, q, kdtree=kd)
        assert list(idx[0][:3]) == [0, 1, 2] and list(idx[0][3:]) == [-1, -1]


def test_plane_lsq_golden_lstsq(orc):
    g = np.load(os.path.join(GOLD, "plane5x3_lstsq.npz"))
    for P, X in zip(g["P"], g["X"]):
        x = orc.plane_lsq(P)
        n = X / np.linalg.norm(X)
        assert np.allclose(x / np.linalg.norm(x), n, atol=2e-3)
        # the normalised plane must pass through the points like the float64 solution does
        assert np.allclose(x.astype(np.float64) / np.linalg.norm(X), X / np.linalg.norm(X), atol=5e-3)


def test_plane_lsq_rank_deficient_is_finite(orc):
    P = np.array([[1, 2, 3]] * 5, np.float32)          # rank 1: Eigen zeroes the non-pivot unknowns
    x = orc.plane_lsq(P)
    assert np.all(np.isfinite(x))
    assert np.allclose(P @ x, -1, atol=1e-5)


def test_solve6_and_eigen6_golden_cv2(orc):
    g = np.load(os.path.join(GOLD, "lm6x6_cv2.npz"))
    for A, b, x_cv, w_cv, V_cv in zip(g["A"], g["b"], g["x"], g["w"], g["V"]):
        x = orc.solve6(A, b)
        x64 = np.linalg.solve(A.astype(np.float64), b.astype(np.float64))
        scale = np.abs(x64).max() + 1e-12
        cond = np.linalg.cond(A.astype(np.float64))
        tol = max(5e-5, 3e-7 * cond)
        assert np.abs(x - x64).max() / scale < tol
        assert np.abs(x_cv - x64).max() / scale < tol          # cv2 agrees with the same reference
        w, V = orc.eigen6(A)
        assert np.all(np.diff(w) <= 0)                          # descending, like cv::eigen
        assert np.allclose(w, w_cv, rtol=2e-5, atol=2e-5 * abs(w_cv).max())
        assert np.allclose(V @ V.T, np.eye(6), atol=1e-5)       # eigenvectors are the ROWS
        assert np.allclose(V @ A.astype(np.float64) @ V.T, np.diag(w), atol=3e-5 * abs(w).max())
        # same degeneracy decision and projector as the cv2 eigen-decomposition (src/mapOptmization.cpp:1141-1153)
        deg, P = orc.degeneracy(A)
        V2 = V_cv.copy(); deg_cv = False
        for i in range(5, -1, -1):
            if w_cv[i] < 100: V2[i] = 0; deg_cv = True
            else: break
        P_cv = np.linalg.inv(V_cv.astype(np.float64)) @ V2
        assert deg == deg_cv
        assert np.allclose(P, P_cv, atol=2e-4)
    assert any(orc.degeneracy(A)[0] for A in g["A"]) and not all(orc.degeneracy(A)[0] for A in g["A"])


def test_jacobian_matches_finite_differences(orc, synth):
    """Row = d/d(pose) of  c . (R(rpy) p + t)  with the reference's LOAM relabelling (src/mapOptmization.cpp:1106-1123)."""
    rng = np.random.default_rng(7)
    for _ in range(20):
        pose = np.concatenate([rng.uniform(-0.5, 0.5, 2), rng.uniform(-3, 3, 1), rng.uniform(-20, 20, 3)])
        p = rng.uniform(-30, 30, 3); n = rng.normal(size=3); n /= np.linalg.norm(n)
        coef = np.r_[n, 0.123]
        row, rhs = orc.jacobian_row(pose.astype(np.float32), np.r_[p, 0].astype(np.float32), coef.astype(np.float32))
        assert rhs == -np.float32(0.123)

        def resid(q):
            return n @ (synth.rot_zyx(*q[:3]) @ p + q[3:])
        num = np.zeros(6)
        for k in range(6):
            e = np.zeros(6); e[k] = 1e-6
            num[k] = (resid(pose + e) - resid(pose - e)) / 2e-6
        # Quirk preserved from the reference (and upstream LIO-SAM), src/mapOptmization.cpp:1110: the
        # pitch column uses srx*sry*srz where the true derivative has srx*cry*srz.
        sy_, sp_, cp_, sr_ = np.sin(pose[2]), np.sin(pose[1]), np.cos(pose[1]), np.sin(pose[0])
        num[1] -= (sy_ * cp_ * sr_ - sy_ * sp_ * sr_) * p[1] * n[1]
        assert np.allclose(row, num, rtol=2e-3, atol=2e-4)


def test_scan2map_recovers_ground_truth(orc, scene_hdl):
    s = scene_hdl
    m = orc.localmap_build(s.kf_clouds, s.kf_poses, s.map_leaf, threads=4)
    ds = orc.voxelgrid(s.scan_raw, s.scan_leaf)
    pose, res, rc, st = orc.scan2map(m, ds, s.guess, orc.S2mOpts(knn=1, threads=4))
    assert rc == 0 and res.converged == 1 and 1 <= res.iters <= 20 and res.n_sel > 1000
    assert np.abs(pose[3:] - s.truth[3:]).max() < 0.02
    assert np.abs(pose[:3] - s.truth[:3]).max() < 0.003
    pose_b, res_b, _, _ = orc.scan2map(m, ds, s.guess, orc.S2mOpts(knn=0, threads=4))
    assert np.array_equal(pose, pose_b) and res.as_dict() == res_b.as_dict()      # kd-tree == brute force
    assert res.is_degenerate == 0


def test_scan2map_gates(orc, scene_small):
    s = scene_small
    m = orc.localmap_build(s.kf_clouds, s.kf_poses, s.map_leaf)
    ds = orc.voxelgrid(s.scan_raw, s.scan_leaf)
    # :1187 fewer than 31 points -> skipped, pose untouched
    pose, res, rc, _ = orc.scan2map(m, ds[:30], s.guess)
    assert rc == 1 and res.skipped == 1 and np.array_equal(pose, s.guess)
    # :1080 fewer than 50 correspondences -> no update
    far = s.guess.copy(); far[3] += 500
    pose, res, rc, _ = orc.scan2map(m, ds, far)
    assert rc == 0 and res.too_few == 1 and res.converged == 0 and np.array_equal(pose, far)


def test_scan2map_degenerate_scene_sets_flag(orc, synth):
    """Only a ground plane: x, y and yaw are unobservable -> isDegenerate (src/mapOptmization.cpp:1141-1152)."""
    world = np.array([[-500, -500, 0, 500, 500, 400]], np.float32)
    traj = synth.trajectory_line(5)
    sensor = synth.Sensor("down", 0, 16, 900, -40.0, -8.0, 0.5, 60.0)
    kfs = [orc.voxelgrid(synth.scan(world, traj[k], sensor, 40 + k), 0.4) for k in range(4)]
    m = orc.localmap_build(kfs, traj[:4], 0.5)
    ds = orc.voxelgrid(synth.scan(world, traj[4], sensor, 49), 0.4)
    pose, res, rc, st = orc.scan2map(m, ds, synth.perturb(traj[4], 3, dt=0.05, dr=0.01))
    assert rc == 0 and res.is_degenerate == 1 and st.is_degenerate == 1
    P = np.array(st.matP, np.float32).reshape(6, 6)
    assert np.linalg.matrix_rank(P.astype(np.float64), tol=1e-3) < 6
    assert abs(pose[5] - traj[4][5]) < 0.02            # z is still observable


def test_icp_fitness_matches_ckdtree(orc):
    """getICPFitnessScore restatement (anchorOptimize.hpp:465-481) against scipy: unbounded exact 1-NN, mean squared distance."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(31)
    a = (rng.normal(size=(2000, 4)) * [12, 9, 1.5, 1]).astype(np.float32)
    b = (rng.normal(size=(2500, 4)) * [12, 9, 1.5, 1]).astype(np.float32)
    b[:50, :3] += 300.0                                   # far clutter: must not matter
    p1 = np.array([0.02, -0.01, 0.4, 1.5, -2.0, 0.2], np.float32)
    p2 = np.array([0.0, 0.01, -0.1, 0.3, 0.1, 0.0], np.float32)
    at, bt = orc.transform_cloud(a, p1), orc.transform_cloud(b, p2)
    d, _ = cKDTree(bt[:, :3].astype(np.float64)).query(at[:, :3].astype(np.float64))
    want = float((d ** 2).mean())
    for kd in (True, False):
        got = orc.icp_fitness(a, p1, b, p2, kdtree=kd, threads=4)
        assert abs(got - want) <= 2e-5 * want
    assert orc.icp_fitness(a, p1, b, p2, kdtree=True) == orc.icp_fitness(a, p1, b, p2, kdtree=False)
    assert orc.icp_fitness(a, p1, a, p1) == 0.0
