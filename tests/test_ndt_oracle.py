"""CPU tests of the NDT oracle (oracle/ndt_oracle.cpp).  ndt_omp is un-vendored and unpinned in the reference
(README.md:39,47), and nothing in the reference tests it: parity for this path is UNPINNED.  The oracle is
checked here for internal consistency against independent numpy restatements of the published algorithm
(Magnusson 2009): voxel Gaussians, the score function and its derivatives, and end-to-end convergence."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def ndt_scene(synth, orc):
    world = synth.make_world(seed=3)
    traj = synth.trajectory_line(12, step=1.0, x0=-5)
    sub = np.concatenate([orc.transform_cloud(synth.scan(world, traj[k], synth.VLP16, 30 + k), traj[k].astype(np.float32)) for k in range(10)])
    src = synth.scan(world, traj[10], synth.VLP16, 77)
    return sub, src, traj[10]


def guess_matrix(synth, pose):
    """init_guess = [eulerToRotation(r,p,y) | t]  (src/mapOptmization.cpp:276-280, include/math_tools.h:229)"""
    T = np.zeros((3, 4), np.float32)
    T[:, :3] = synth.rot_zyx(*pose[:3]); T[:, 3] = pose[3:]
    return T


def test_voxel_gaussians_match_numpy(orc, ndt_scene):
    sub, _, _ = ndt_scene
    res = np.float32(1.0)
    tg = orc.NdtTarget(sub, res)
    means, icovs, counts = tg.leaves()
    inv = np.float32(1.0) / res
    ijk = np.floor(sub[:, :3] * inv).astype(np.int64)
    ijk -= ijk.min(0)
    dims = ijk.max(0) + 1
    key = ijk[:, 0] + ijk[:, 1] * dims[0] + ijk[:, 2] * dims[0] * dims[1]
    order = np.argsort(key, kind="stable")
    ks = key[order]
    starts = np.flatnonzero(np.r_[True, ks[1:] != ks[:-1]])
    ends = np.r_[starts[1:], len(ks)]
    li = 0
    for s, e in zip(starts, ends):
        if e - s < 6:
            continue
        P = sub[order[s:e], :3].astype(np.float64)
        n = len(P)
        mean = P.mean(0)
        cov = np.cov(P.T, bias=True) * ((n - 1.0) / n)                  # ndt_omp's single-pass formula and its (n-1)/n factor
        w, V = np.linalg.eigh(cov)
        if w[0] < 0 or w[1] < 0 or w[2] <= 0:
            continue
        w = np.maximum(w, 0.01 * w[2]) if w[0] < 0.01 * w[2] else w
        cov2 = (V * w) @ V.T
        icov = np.linalg.inv(cov2)
        if not np.all(np.isfinite(icov)):
            continue
        assert counts[li] == n
        assert np.allclose(means[li], mean, atol=1e-9)
        assert np.allclose(icovs[li].reshape(3, 3), icov, rtol=2e-4, atol=1e-3 * abs(icov).max())
        li += 1
    assert li == len(means) > 1000


def test_derivatives_match_finite_differences(orc, synth, ndt_scene):
    """Analytic gradient / Hessian (Magnusson eq. 6.12, 6.13) against finite differences of the score
    -d1 * exp(-d2/2 * q^T C q), q = T(p) x - mu, evaluated in float64 numpy with the point-to-voxel assignment
    FROZEN at the base pose (the real score is only piecewise smooth: points hop between voxels)."""
    sub, src, truth = ndt_scene
    src = src[::3]
    res = np.float32(1.0)
    tg = orc.NdtTarget(sub, res)
    means, icovs, counts = tg.leaves()
    keys = tg.keys()
    o = orc.NdtOpts(resolution=1.0, threads=4)
    p = orc.ndt_matrix_to_pose(guess_matrix(synth, truth + np.array([0.004, -0.003, 0.01, 0.05, -0.04, 0.02])))
    base = orc.ndt_derivatives(tg, src, p, o)
    grad = base[1:7]
    H = np.zeros((6, 6)); k = 7
    for i in range(6):
        for j in range(i, 6):
            H[i, j] = H[j, i] = base[k]; k += 1
    # frozen assignment at the base pose (DIRECT1: the voxel containing the transformed point)
    inv = np.float32(1.0) / res
    minb = np.floor(sub[:, :3].min(0) * inv).astype(np.int64); maxb = np.floor(sub[:, :3].max(0) * inv).astype(np.int64)
    div = maxb - minb + 1
    T0 = orc.ndt_pose_to_matrix(p)
    xt = (src[:, :3] @ T0[:, :3].T + T0[:, 3]).astype(np.float32)
    ijk = np.floor(xt * inv).astype(np.int64)
    inside = np.all((ijk >= minb) & (ijk <= maxb), axis=1)
    key = (ijk - minb) @ np.array([1, div[0], div[0] * div[1]])
    leaf_of = {int(kk): i for i, kk in enumerate(keys)}
    idx = np.array([leaf_of.get(int(kk), -1) if ins else -1 for kk, ins in zip(key, inside)])
    use = idx >= 0
    X = src[use, :3].astype(np.float64); MU = means[idx[use]]; C = icovs[idx[use]].reshape(-1, 3, 3).astype(np.float64)
    c1, c2 = 10 * (1 - 0.55), 0.55 / 1.0 ** 3
    d3 = -np.log(c2); d1 = -np.log(c1 + c2) - d3; d2 = -2 * np.log((-np.log(c1 * np.exp(-0.5) + c2) - d3) / d1)

    def score(q):
        ca, sa, cb, sb, cc, sc = np.cos(q[3]), np.sin(q[3]), np.cos(q[4]), np.sin(q[4]), np.cos(q[5]), np.sin(q[5])
        R = np.array([[1, 0, 0], [0, ca, -sa], [0, sa, ca]]) @ np.array([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]]) @ np.array([[cc, -sc, 0], [sc, cc, 0], [0, 0, 1]])
        d = X @ R.T + q[:3] - MU
        return np.sum(-d1 * np.exp(-0.5 * d2 * np.einsum("ni,nij,nj->n", d, C, d)))
    assert abs(score(p) - base[0]) < 1e-3 * abs(base[0])
    h = 1e-5
    num_g = np.array([(score(p + h * e) - score(p - h * e)) / (2 * h) for e in np.eye(6)])
    assert np.allclose(grad, num_g, rtol=2e-3, atol=2e-3 * np.abs(grad).max())
    num_H = np.zeros((6, 6))
    hh = 1e-4
    for i in range(6):
        for j in range(6):
            ei, ej = np.eye(6)[i] * hh, np.eye(6)[j] * hh
            num_H[i, j] = (score(p + ei + ej) - score(p + ei - ej) - score(p - ei + ej) + score(p - ei - ej)) / (4 * hh * hh)
    assert np.allclose(H, num_H, rtol=5e-3, atol=5e-3 * np.abs(H).max())


def test_pose_matrix_round_trip(orc):
    rng = np.random.default_rng(0)
    for _ in range(200):
        p = np.r_[rng.uniform(-50, 50, 3), rng.uniform(-3.1, 3.1, 3)]
        T = orc.ndt_pose_to_matrix(p)
        R = T[:, :3].astype(np.float64)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-6) and np.linalg.det(R) > 0.999
        p2 = orc.ndt_matrix_to_pose(T)
        assert np.allclose(orc.ndt_pose_to_matrix(p2), T, atol=3e-7)           # Eigen eulerAngles(0,1,2) picks another branch, same matrix
        ca, sa, cb, sb, cc, sc = np.cos(p[3]), np.sin(p[3]), np.cos(p[4]), np.sin(p[4]), np.cos(p[5]), np.sin(p[5])
        Rx = np.array([[1, 0, 0], [0, ca, -sa], [0, sa, ca]]); Ry = np.array([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]]); Rz = np.array([[cc, -sc, 0], [sc, cc, 0], [0, 0, 1]])
        assert np.allclose(R, Rx @ Ry @ Rz, atol=1e-6)                            # p = [t, eulerAngles(0,1,2)]: R = Rx Ry Rz


@pytest.mark.parametrize("search", [0, 1, 2])
def test_align_recovers_truth(orc, synth, ndt_scene, search):
    sub, src, truth = ndt_scene
    tg = orc.NdtTarget(sub, 1.0)
    o = orc.NdtOpts(resolution=1.0, trans_eps=0.01, search=search, threads=8)
    g = truth.copy(); g[3] += 0.3; g[4] -= 0.15; g[2] += 0.05
    T, res = orc.ndt_align(tg, src, guess_matrix(synth, g), o)
    assert res.converged == 1 and 1 <= res.iterations <= 36 and res.sweeps >= res.iterations
    assert np.abs(T[:, 3] - truth[3:]).max() < 0.02
    assert np.abs(T[:, :3] - synth.rot_zyx(*truth[:3])).max() < 2e-3
    T2, res2 = orc.ndt_align(tg, src, T, o)                                      # the reference aligns twice at relo init (:282-290)
    assert np.abs(T2[:, 3] - truth[3:]).max() < 0.02


def test_align_without_overlap_terminates(orc, synth, ndt_scene):
    sub, src, truth = ndt_scene
    tg = orc.NdtTarget(sub, 1.0)
    g = truth.copy(); g[3] += 500
    T, res = orc.ndt_align(tg, src, guess_matrix(synth, g), orc.NdtOpts(resolution=1.0, threads=4))
    assert res.sweeps == 1 and res.iterations == 0              # zero gradient -> delta_p_norm == 0 -> return (computeTransformation)
    assert np.allclose(T, guess_matrix(synth, g))


def test_kdtree_neighbourhood_matches_scipy_radius_search(orc, synth, ndt_scene):
    """regMethod KDTREE (src/mapOptmization.cpp:195-199): the neighbourhood of a transformed point is every leaf whose
    centroid is closer than `resolution` -- pinned against scipy's ball query over the oracle's own leaf centroids."""
    from scipy.spatial import cKDTree
    sub, src, truth = ndt_scene
    src = src[::2]
    tg = orc.NdtTarget(sub, 1.0)
    means, icovs, counts = tg.leaves()
    p = orc.ndt_matrix_to_pose(guess_matrix(synth, truth + np.array([0.004, -0.003, 0.01, 0.05, -0.04, 0.02])))
    got = orc.ndt_derivatives(tg, src, p, orc.NdtOpts(resolution=1.0, search=2, threads=4))
    T0 = orc.ndt_pose_to_matrix(p)
    xt = (src[:, :3] @ T0[:, :3].T + T0[:, 3]).astype(np.float32)
    tree = cKDTree(means.astype(np.float32).astype(np.float64))
    nb = tree.query_ball_point(xt.astype(np.float64), r=1.0)
    c1, c2 = 10 * (1 - 0.55), 0.55 / 1.0 ** 3
    d3 = -np.log(c2); d1 = -np.log(c1 + c2) - d3; d2 = -2 * np.log((-np.log(c1 * np.exp(-0.5) + c2) - d3) / d1)
    score, pairs = 0.0, 0
    for i, ids in enumerate(nb):
        for l in ids:
            d = xt[i].astype(np.float64) - means[l]
            score += -d1 * np.exp(-0.5 * d2 * d @ icovs[l].reshape(3, 3).astype(np.float64) @ d)
            pairs += 1
    assert pairs > 2 * len(src)                                     # several Gaussians per point, unlike DIRECT1
    assert abs(got[0] - score) < 1e-4 * abs(score)
    s1 = orc.ndt_derivatives(tg, src, p, orc.NdtOpts(resolution=1.0, search=0, threads=4))[0]
    s7 = orc.ndt_derivatives(tg, src, p, orc.NdtOpts(resolution=1.0, search=1, threads=4))[0]
    assert got[0] > s7 > s1 > 0                                     # more neighbours, more score (-d1 > 0)


def test_newton_direction_is_the_minimum_norm_solution(orc):
    """JacobiSVD(H).solve(-g) (computeTransformation): H^-1 (-g) when H is regular, the pseudo-inverse solution when a
    direction is unobservable (planar scenes: zero rows / columns) -- against numpy."""
    rng = np.random.default_rng(3)
    for _ in range(20):
        A = rng.normal(size=(6, 6)); H = A @ A.T + 0.1 * np.eye(6); g = rng.normal(size=6)
        assert np.allclose(orc.ndt_newton(H, g), np.linalg.solve(H, -g), rtol=1e-9, atol=1e-12)
    for _ in range(20):                                             # rank 4: two unobservable directions
        B = rng.normal(size=(6, 4)); H = B @ B.T; g = B @ rng.normal(size=4)
        want = np.linalg.pinv(H, rcond=1e-12) @ -g
        got = orc.ndt_newton(H, g)
        assert np.allclose(got, want, rtol=1e-6, atol=1e-9), (got, want)
    # indefinite Hessians (far from the optimum) are solved as they are, like the SVD does
    H = np.diag([3.0, -2.0, 1.0, 5.0, -0.5, 4.0]); g = np.arange(1.0, 7.0)
    assert np.allclose(orc.ndt_newton(H, g), -g / np.diag(H))


def test_newton_paths_agree(orc):
    """The canonical Newton solve (elimination when well conditioned, pseudo-inverse otherwise; shared with the device)
    against the pure pseudo-inverse (JacobiSVD semantics): same direction to ~1e-10 on Hessians shaped like the NDT ones
    (translation block ~1e3, rotation block ~1e5, indefinite far from the optimum)."""
    rng = np.random.default_rng(11)
    S = np.diag([30.0, 30.0, 30.0, 300.0, 300.0, 300.0])
    for k in range(50):
        A = rng.normal(size=(6, 6)); H = S @ (A @ A.T - (1.5 if k % 3 == 0 else 0.0) * np.eye(6)) @ S; g = S @ rng.normal(size=6) * 10
        a = orc.ndt_newton(H, g); b = orc.ndt_newton_pinv(H, g)
        assert np.allclose(a, b, rtol=1e-8, atol=1e-12 * np.abs(b).max()), (k, a, b)
    # singular: both drop the null space (the elimination refuses the tiny pivot)
    B = rng.normal(size=(6, 3)); H = B @ B.T; g = B @ rng.normal(size=3)
    assert np.allclose(orc.ndt_newton(H, g), orc.ndt_newton_pinv(H, g), rtol=1e-9, atol=1e-12)


def test_canonical_summation_order(orc, synth, ndt_scene):
    """The 28 sums are added in ONE fixed order (1024-point tiles, 8 warps of 4 x 32 points, hits dealt to 32 accumulators,
    butterfly, warps, tiles -- oracle/ndt_oracle.cpp derivative_sweep): bit-identical for every thread count, and equal to the
    plain ascending-point-order sums up to double rounding."""
    sub, src, truth = ndt_scene
    tg = orc.NdtTarget(sub, 1.0)
    rng = np.random.default_rng(2)
    for search in (0, 1, 2):
        for n in (len(src), 1024, 1025, 37):
            p = orc.ndt_matrix_to_pose(guess_matrix(synth, truth + np.r_[rng.uniform(-0.03, 0.03, 3), rng.uniform(-0.5, 0.5, 3)]))
            ref = orc.ndt_derivatives(tg, src[:n], p, orc.NdtOpts(resolution=1.0, search=search, threads=1))
            for th in (2, 3, 8):
                assert np.array_equal(ref, orc.ndt_derivatives(tg, src[:n], p, orc.NdtOpts(resolution=1.0, search=search, threads=th)))
            seq = orc.ndt_derivatives_sequential(tg, src[:n], p, orc.NdtOpts(resolution=1.0, search=search))
            assert np.allclose(ref, seq, rtol=1e-12, atol=1e-12 * np.abs(seq).max())


def test_more_thuente_trial_values(orc):
    """trialValueSelectionMT (More & Thuente 1994, section 4): case 1 picks the cubic minimiser when it is closer to a_l
    than the quadratic one, else their mean; the cubic / quadratic / secant minimisers are checked on polynomials whose
    minimisers are known in closed form."""
    # case 1 (f_t > f_l): interpolate f(a) = (a - 1)^2 through a_l = 0 and a_t = 3 -> both minimisers are a = 1
    f = lambda a: (a - 1.0) ** 2
    df = lambda a: 2.0 * (a - 1.0)
    a = orc.ndt_trial_value(0.0, f(0.0), df(0.0), 5.0, f(5.0), df(5.0), 3.0, f(3.0), df(3.0))
    assert abs(a - 1.0) < 1e-12
    # case 1 with a genuine cubic f(a) = a^3 - 3a (local minimum at a = 1): a_l = 0, a_t = 2
    f = lambda a: a ** 3 - 3 * a
    df = lambda a: 3 * a * a - 3
    a_c = 1.0                                                       # cubic through (0, 2) with both slopes reproduces f
    a_q = 0.0 - 0.5 * (0.0 - 2.0) * df(0.0) / (df(0.0) - (f(0.0) - f(2.0)) / (0.0 - 2.0))
    want = a_c if abs(a_c - 0.0) < abs(a_q - 0.0) else 0.5 * (a_q + a_c)
    assert abs(orc.ndt_trial_value(0.0, f(0.0), df(0.0), 4.0, f(4.0), df(4.0), 2.0, f(2.0), df(2.0)) - want) < 1e-12
    # case 2 (f_t <= f_l, slopes of opposite sign): cubic vs secant minimiser, the one farther from a_t ... on a parabola both are exact
    f = lambda a: (a - 2.0) ** 2
    df = lambda a: 2.0 * (a - 2.0)
    assert abs(orc.ndt_trial_value(0.0, f(0.0), df(0.0), 9.0, f(9.0), df(9.0), 3.0, f(3.0), df(3.0)) - 2.0) < 1e-12
