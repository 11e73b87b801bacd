"""GPU parity tests of the relo-mode NDT path (C ABI liloc_ndt_*) against the oracle (oracle/ndt_oracle.cpp).
Per-point terms follow the same float32 operation sequence on both sides and the 28 double sums are added in ONE
canonical order on both sides (oracle/ndt_oracle.cpp derivative_sweep <-> ndt.cuh ndt_sweep_tile), so the sums are
bit-identical, the Newton / More-Thuente iteration takes the same path (iterations and sweeps exact) and final
poses agree far inside the north_star tolerance (1e-4 m, 1e-5 rad)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from test_ndt_oracle import guess_matrix, ndt_scene  # noqa: E402,F401


def test_target_leaves_bit_exact(gpu_ctx, orc, ndt_scene):
    sub, _, _ = ndt_scene
    for res in (1.0, 0.8, 2.0):
        n = gpu_ctx.ndt_target_put(1, sub, res)
        means, icovs, counts, keys = gpu_ctx.ndt_target_get(1)
        rm, ri, rc = orc.NdtTarget(sub, np.float32(res)).leaves()
        assert n == len(rm) == len(means)
        assert np.array_equal(counts, rc)
        assert np.array_equal(means, rm)
        assert np.array_equal(icovs, ri)


def test_derivatives_match_oracle(gpu_ctx, orc, synth, ndt_scene):
    from liloc_b200 import NdtOpts
    sub, src, truth = ndt_scene
    gpu_ctx.ndt_target_put(2, sub, 1.0)
    gpu_ctx.ndt_source_put(2, src)
    tg = orc.NdtTarget(sub, 1.0)
    rng = np.random.default_rng(1)
    exact = total = 0
    for search in (0, 1, 2):             # DIRECT1, DIRECT7, KDTREE
        for _ in range(4):
            p = orc.ndt_matrix_to_pose(guess_matrix(synth, truth + np.r_[rng.uniform(-0.03, 0.03, 3), rng.uniform(-0.5, 0.5, 3)]))
            got = gpu_ctx.ndt_derivatives(2, 2, p, NdtOpts(search=search))
            ref = orc.ndt_derivatives(tg, src, p, orc.NdtOpts(resolution=1.0, search=search, threads=8))
            # same float terms, same summation order: identical up to the (float-rounded) transcendental functions
            assert np.allclose(got, ref, rtol=1e-13, atol=1e-13 * np.abs(ref).max()), np.abs(got - ref).max()
            exact += int(np.array_equal(got, ref)); total += 1
    assert exact >= total - 1, (exact, total)


def _pose_err(T, Tref):
    dt = np.abs(T[:3, 3] - Tref[:3, 3]).max()
    dR = np.abs(T[:3, :3] - Tref[:3, :3]).max()
    return dt, dR


@pytest.mark.parametrize("search", [0, 1, 2])
def test_align_matches_oracle(gpu_ctx, orc, synth, ndt_scene, search):
    from liloc_b200 import NdtOpts
    sub, src, truth = ndt_scene
    gpu_ctx.ndt_target_put(3, sub, 1.0)
    gpu_ctx.ndt_source_put(3, src)
    tg = orc.NdtTarget(sub, 1.0)
    rng = np.random.default_rng(5)
    guesses = [guess_matrix(synth, truth + np.r_[rng.uniform(-0.02, 0.02, 3), rng.uniform(-0.4, 0.4, 3)]) for _ in range(12)]
    guesses.append(guess_matrix(synth, truth + np.array([0, 0, 0, 500.0, 0, 0])))                 # no overlap at all
    J = len(guesses)
    T, score, iters, conv, sweeps = gpu_ctx.ndt_align_batch([3] * J, [3] * J, guesses, NdtOpts(search=search))
    mism = []
    for j in range(J):
        Tr, rr = orc.ndt_align(tg, src, guesses[j], orc.NdtOpts(resolution=1.0, search=search, threads=8))
        dt, dR = _pose_err(T[j], Tr)
        assert dt <= 1e-4 and dR <= 1e-5, (j, dt, dR)          # 1e-4 m; rotation-matrix entries within 1e-5 (~1e-5 rad)
        if (iters[j], conv[j], sweeps[j]) != (rr.iterations, rr.converged, rr.sweeps):
            mism.append((j, iters[j], sweeps[j], rr.iterations, rr.sweeps))
        assert abs(score[j] - rr.score) <= 1e-6 * max(1.0, abs(rr.score))
    assert not mism, mism
    assert np.abs(T[0, :3, 3] - truth[3:]).max() < 0.03


def test_two_pass_align_and_batch_independence(gpu_ctx, orc, synth, ndt_scene):
    """n_align = 2 (relocalisation init, :282-290) equals two chained single aligns; a batch equals its jobs run alone."""
    from liloc_b200 import NdtOpts
    sub, src, truth = ndt_scene
    gpu_ctx.ndt_target_put(4, sub, 1.0)
    gpu_ctx.ndt_target_put(5, sub[: len(sub) // 2], 1.0)
    gpu_ctx.ndt_source_put(4, src)
    gpu_ctx.ndt_source_put(5, src[::2])
    g = [guess_matrix(synth, truth + np.array([0.01, 0, 0.03, 0.2, -0.1, 0.05])), guess_matrix(synth, truth + np.array([0, 0.01, -0.02, -0.15, 0.2, 0.0]))]
    T2, s2, it2, c2, sw2 = gpu_ctx.ndt_align_batch([4, 5], [4, 5], g, NdtOpts(n_align=2))
    for j, (t, s) in enumerate(((4, 4), (5, 5))):
        Ta, _, ia, _, swa = gpu_ctx.ndt_align_batch([t], [s], [g[j]], NdtOpts())
        Tb, sb, ib, _, swb = gpu_ctx.ndt_align_batch([t], [s], [Ta[0]], NdtOpts())
        assert np.array_equal(Tb[0], T2[j]) and it2[j] == ia[0] + ib[0] and sw2[j] == swa[0] + swb[0]
    tg = orc.NdtTarget(sub, 1.0)
    Tr, _ = orc.ndt_align(tg, src, g[0], orc.NdtOpts(resolution=1.0, threads=8))
    Tr2, _ = orc.ndt_align(tg, src, Tr, orc.NdtOpts(resolution=1.0, threads=8))
    dt, dR = _pose_err(T2[0], Tr2)
    assert dt <= 1e-4 and dR <= 1e-5


def _bench_workload(name):
    import types
    import bench
    return bench.make_ndt_workload(types.SimpleNamespace(seed={"relo": 4000, "jfgo": 5000}[name], workload=name, items=0, submaps=64))


def _check_workload_against_oracle(gpu_ctx, orc, w, res=1.0):
    """Every job of a bench workload: iterations / sweeps / convergence exact, pose inside the north_star tolerance."""
    from liloc_b200 import NdtOpts
    gpu_ctx.ndt_clear()
    for t, pts in w["targets"].items():
        gpu_ctx.ndt_target_put(int(t), pts, res)
    for s_, pts in w["sources"].items():
        gpu_ctx.ndt_source_put(int(s_), pts)
    T, score, iters, conv, sweeps = gpu_ctx.ndt_align_batch(w["tid"], w["sid"], w["guesses"], NdtOpts(trans_eps=0.01, search=0, n_align=w["n_align"]))
    o = orc.NdtOpts(resolution=res, trans_eps=0.01, search=0, threads=8)
    tgs = {}
    bad, worst_t, worst_r, n_max_iter, n_far = [], 0.0, 0.0, 0, 0
    for j in range(len(w["guesses"])):
        t = int(w["tid"][j])
        if t not in tgs:
            tgs[t] = orc.NdtTarget(w["targets"][t], np.float32(res))
        Tr, it, sw = w["guesses"][j][:3], 0, 0
        for _ in range(w["n_align"]):
            Tr, rr = orc.ndt_align(tgs[t], w["sources"][int(w["sid"][j])], Tr, o)
            it += rr.iterations; sw += rr.sweeps
        dt, dR = _pose_err(T[j], Tr)
        worst_t, worst_r = max(worst_t, dt), max(worst_r, dR)
        n_max_iter += int(rr.iterations >= 36)
        n_far += int(np.abs(Tr[:3, 3] - w["guesses"][j][:3, 3]).max() > 0.5)
        if (iters[j], sweeps[j], conv[j]) != (it, sw, rr.converged) or dt > 1e-4 or dR > 1e-5:
            bad.append((j, iters[j], sweeps[j], it, sw, dt, dR))
        assert abs(score[j] - rr.score) <= 1e-9 * max(1.0, abs(rr.score)), (j, score[j], rr.score)
    assert not bad, bad[:8]
    return worst_t, worst_r, n_max_iter, n_far


def test_bench_relo_workload_matches_oracle(gpu_ctx, orc):
    """BASELINE configs[3] exactly as `bench.py --workload relo` runs it: 256 hypotheses (16 yaw offsets over the full circle
    x 16 positions), 2 x align each (src/mapOptmization.cpp:282-290) -- EVERY job against the oracle, including the
    hypotheses half a turn off, the ones that run into max_iterations and the sign-flipped directions."""
    w = _bench_workload("relo")
    assert len(w["guesses"]) == 256 and w["n_align"] == 2
    worst_t, worst_r, n_max_iter, n_far = _check_workload_against_oracle(gpu_ctx, orc, w)
    assert n_far > 20                          # the far hypotheses really move (or wander) a long way


def test_bench_jfgo_workload_matches_oracle(gpu_ctx, orc):
    """BASELINE configs[4]a as `bench.py --workload jfgo` runs it: one scan against K = 64 prior submaps
    (anchorOptimize.hpp:394-407) -- every job against the oracle (most submaps do not overlap the scan at all)."""
    w = _bench_workload("jfgo")
    assert len(w["guesses"]) == 64
    _check_workload_against_oracle(gpu_ctx, orc, w)


def test_ndt_errors(gpu_ctx):
    from liloc_b200 import LilocError
    with pytest.raises(LilocError):
        gpu_ctx.ndt_align_batch([999], [999], [np.eye(4, dtype=np.float32)])


def test_device_against_the_numpy_golden(gpu_ctx):
    """The device against the oracle-independent golden vectors (tests/golden/make_golden_ndt.py): leaves, the float32
    restatement's sums, the float64 symbolic sums and the float64 registration."""
    import os
    from liloc_b200 import NdtOpts
    from test_ndt_golden import pose_to_guess
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ndt_np.npz"))
    pad = lambda a: np.ascontiguousarray(np.c_[a, np.zeros(len(a), np.float32)], np.float32)
    gpu_ctx.ndt_clear()
    n = gpu_ctx.ndt_target_put(0, pad(g["target"]), float(g["resolution"]))
    means, icovs, counts, keys = gpu_ctx.ndt_target_get(0)
    assert n == len(g["leaf_means"]) and np.array_equal(counts, g["leaf_counts"])
    assert np.allclose(means, g["leaf_means"], rtol=0, atol=1e-12)
    assert np.allclose(icovs, g["leaf_icovs"].reshape(-1, 9).astype(np.float32), rtol=2e-6, atol=0)
    gpu_ctx.ndt_source_put(0, pad(g["source"]))
    gpu_ctx.ndt_source_put(1, pad(g["source_safe"]))
    for p, w32, w64 in zip(g["poses"], g["sums_f32"], g["sums_f64_safe"]):
        got = gpu_ctx.ndt_derivatives(0, 0, p, NdtOpts(search=0))
        assert np.allclose(got, w32, rtol=1e-12, atol=1e-12 * np.abs(w32).max())
        got = gpu_ctx.ndt_derivatives(0, 1, p, NdtOpts(search=0))
        assert np.abs(got - w64).max() <= 1e-4 * np.abs(w64).max()
    guesses = np.stack([pose_to_guess(p) for p in g["guesses"]])
    T, score, iters, conv, sweeps = gpu_ctx.ndt_align_batch([0] * 3, [0] * 3, guesses, NdtOpts(trans_eps=float(g["trans_eps"]), search=0))
    assert np.array_equal(iters, g["final_iterations"]) and conv.all()
    for j in range(3):
        want = pose_to_guess(g["finals"][j])
        assert np.abs(T[j, :3, 3] - want[:, 3]).max() <= 1e-5 and np.abs(T[j, :3, :3] - want[:, :3]).max() <= 1e-5
