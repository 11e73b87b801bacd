"""The NDT oracle against golden vectors made WITHOUT it (tests/golden/make_golden_ndt.py -> ndt_np.npz): voxel Gaussians
and score / gradient / Hessian in numpy float64 with SYMBOLIC (sympy) derivatives of T(p) x; an independent numpy float32
restatement of ndt_omp's per-point terms; a float64 registration with a from-the-paper More-Thuente line search.
ndt_omp is un-vendored and unpinned in the reference (README.md:39,47): this is the pin the NDT path has."""
import os
import sys

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    g = np.load(os.path.join(GOLD, "ndt_np.npz"))
    pad = lambda a: np.ascontiguousarray(np.c_[a, np.zeros(len(a), np.float32)], np.float32)
    return g, pad(g["target"]), pad(g["source"]), pad(g["source_safe"])


def pose_to_guess(p):
    sys.path.insert(0, GOLD)
    import make_golden_ndt as mg
    return mg.pose_matrix(p)[:3].astype(np.float32)


def test_voxel_gaussians_equal_the_numpy_golden(orc, gold):
    g, tgt, _, _ = gold
    m, ic, c = orc.NdtTarget(tgt, np.float32(g["resolution"])).leaves()
    assert len(m) == len(g["leaf_means"]) > 400 and np.array_equal(c, g["leaf_counts"])
    assert np.allclose(m, g["leaf_means"], rtol=0, atol=1e-12)
    assert np.allclose(ic, g["leaf_icovs"].reshape(-1, 9).astype(np.float32), rtol=2e-6, atol=0)        # Jacobi here, LAPACK eigh there


def test_derivative_sums_against_symbolic_float64(orc, gold):
    """(A) float64 mathematics with sympy-derived Jacobian / second derivatives; the oracle's per-point terms are float32."""
    g, tgt, _, src_safe = gold
    tg = orc.NdtTarget(tgt, np.float32(g["resolution"]))
    for p, want in zip(g["poses"], g["sums_f64_safe"]):
        for search in (0,):
            got = orc.ndt_derivatives(tg, src_safe, p, orc.NdtOpts(resolution=1.0, search=search, threads=4))
            assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max(), (np.abs(got - want).max(), np.abs(want).max())
            assert np.allclose(got[1:7], want[1:7], rtol=0, atol=2e-4 * np.abs(want[1:7]).max())      # the gradient on its own scale


def test_derivative_sums_against_numpy_float32_restatement(orc, gold):
    """(B) the same float32 operation sequence, written independently and vectorised in numpy: sums in ascending point order
    to double rounding (bit-identical unless numpy's exp and glibc's round one argument differently), the canonical
    tile-ordered sums to ~1e-13."""
    g, tgt, src, _ = gold
    tg = orc.NdtTarget(tgt, np.float32(g["resolution"]))
    exact = 0
    for p, want in zip(g["poses"], g["sums_f32"]):
        seq = orc.ndt_derivatives_sequential(tg, src, p, orc.NdtOpts(resolution=1.0, search=0))
        assert np.allclose(seq, want, rtol=1e-9, atol=1e-9 * np.abs(want).max())
        exact += int(np.array_equal(seq, want))
        canon = orc.ndt_derivatives(tg, src, p, orc.NdtOpts(resolution=1.0, search=0, threads=3))
        assert np.allclose(canon, want, rtol=1e-12, atol=1e-12 * np.abs(want).max())
    assert exact >= 2


def test_registration_against_float64_newton_and_more_thuente(orc, gold):
    """(C) numpy.linalg.solve Newton step + More-Thuente from the paper: same number of iterations, same optimum."""
    g, tgt, src, _ = gold
    tg = orc.NdtTarget(tgt, np.float32(g["resolution"]))
    o = orc.NdtOpts(resolution=1.0, trans_eps=float(g["trans_eps"]), search=0, threads=4)
    for guess, final, its in zip(g["guesses"], g["finals"], g["final_iterations"]):
        T, r = orc.ndt_align(tg, src, pose_to_guess(guess), o)
        assert r.iterations == its and r.converged == 1
        assert np.abs(orc.ndt_matrix_to_pose(T) - final).max() <= 1e-5
        assert np.abs(final - g["truth"])[:3].max() < 0.02 and np.abs(final - g["truth"])[3:].max() < 0.005
