"""One full LMOptimization step (src/mapOptmization.cpp:1061-1181) of the oracle against golden vectors made WITHOUT it
(tests/golden/make_golden_lm.py): numpy float32 Jacobian rows + the reference's own OpenCV calls (cv2.gemm, cv2.solve
DECOMP_QR, cv2.eigen, cv2.invert) through the Python binding.  Also pins the repo's canonical choice for A^T A (float
products accumulated in double, rounded once) against what OpenCV's gemm returns."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _case(g, name):
    return {k.split(".", 1)[1]: g[k] for k in g.files if k.startswith(name + ".")}


@pytest.mark.parametrize("name", ["room_iter0", "corridor_iter0", "corridor_iter1", "few", "converged"])
def test_lm_step_against_cv2(orc, name):
    g = np.load(os.path.join(GOLD, "lm_step_cv2.npz"))
    c = _case(g, name)
    st = orc.LmState()
    st.is_degenerate = int(c["deg_in"])
    for i, v in enumerate(np.asarray(c["P_in"], np.float32).reshape(36)):
        st.matP[i] = float(v)
    ori = np.c_[c["ori"], np.zeros(len(c["ori"]), np.float32)]
    rc, pose, AtA, AtB, st = orc.lm_step(ori, c["coef"], int(c["iter"]), c["pose"], st)
    if not bool(c["ran"]):
        assert rc == -1 and np.array_equal(pose, c["pose"])          # < 50 correspondences: no update (:1079-1082)
        return
    assert rc == (1 if bool(c["converged"]) else 0)
    # A^T A, A^T b: float32 sums of a few hundred products on OpenCV's side, double accumulation here
    assert np.allclose(AtA, c["AtA"], rtol=2e-5, atol=2e-5 * np.abs(c["AtA"]).max())
    assert np.allclose(AtB, c["AtB"], rtol=0, atol=2e-5 * max(np.abs(c["AtB"]).max(), 1.0))
    assert st.is_degenerate == int(c["deg_out"])
    X_ref = np.asarray(c["X"], np.float64)
    X = (pose.astype(np.float64) - np.asarray(c["pose"], np.float64))
    assert np.allclose(X, X_ref, rtol=0, atol=1e-4 * max(np.abs(X_ref).max(), 1e-3) + 2e-6), (X, X_ref)
    if int(c["iter"]) == 0:
        P, P_ref = np.array(list(st.matP), np.float32).reshape(6, 6), np.asarray(c["P_out"], np.float32)
        assert np.allclose(P, P_ref, atol=2e-3), np.abs(P - P_ref).max()
        if bool(c["deg_out"]):
            assert abs(X[3]) < 1e-4                                 # the unobservable direction (x) receives no update
