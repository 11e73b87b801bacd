#!/usr/bin/env python
"""Golden vectors for one full LMOptimization step (src/mapOptmization.cpp:1061-1181), made WITHOUT the oracle:
the Jacobian rows of :1093-1125 restated in numpy float32 (same expressions, same operand order) and the linear algebra
done by the very OpenCV calls the reference makes -- cv2.transpose, cv2.gemm (matAt * matA, matAt * matB), cv2.solve(...,
DECOMP_QR), cv2.eigen, cv2.invert (matV.inv()), matP * matX -- through OpenCV's Python binding (the C++ library itself
cannot be linked here: no headers).  Cases: a well-conditioned scene, a degenerate corridor (eigenvalues below the
threshold of 100 -> isDegenerate, matP projects the update), an iteration > 0 that re-uses matP, fewer than 50
correspondences (no update) and a converged step.

Writes tests/golden/lm_step_cv2.npz.  Run:  python tests/golden/make_golden_lm.py
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
F = np.float32


def jacobian_rows(pose6, ori, coef):
    """matA / matB of :1093-1125 in float32.  pose6 = transformTobeMapped [roll, pitch, yaw, x, y, z]."""
    srx, crx = F(np.sin(F(pose6[2]))), F(np.cos(F(pose6[2])))        # lidar -> camera relabelling (:1072-1077)
    sry, cry = F(np.sin(F(pose6[1]))), F(np.cos(F(pose6[1])))
    srz, crz = F(np.sin(F(pose6[0]))), F(np.cos(F(pose6[0])))
    px, py, pz = (ori[:, k].astype(F) for k in range(3))
    cx, cy, cz, ci = (coef[:, k].astype(F) for k in range(4))
    arx = (-srx * cry * px - (srx * sry * srz + crx * crz) * py + (crx * srz - srx * sry * crz) * pz) * cx \
        + (crx * cry * px - (srx * crz - crx * sry * srz) * py + (crx * sry * crz + srx * srz) * pz) * cy
    ary = (-crx * sry * px + crx * cry * srz * py + crx * cry * crz * pz) * cx \
        + (-srx * sry * px + srx * sry * srz * py + srx * cry * crz * pz) * cy \
        + (-cry * px - sry * srz * py - sry * crz * pz) * cz
    arz = ((crx * sry * crz + srx * srz) * py + (srx * crz - crx * sry * srz) * pz) * cx \
        + ((-crx * srz + srx * sry * crz) * py + (-srx * sry * srz - crx * crz) * pz) * cy \
        + (cry * crz * py - cry * srz * pz) * cz
    A = np.stack([arz, ary, arx, cx, cy, cz], 1).astype(F)
    B = (-ci).astype(F).reshape(-1, 1)
    return A, B


def lm_step_cv2(pose6, ori, coef, iter_count, is_degenerate, matP):
    """Returns (ran, pose_out, X, AtA, AtB, is_degenerate, matP, converged)."""
    n = len(ori)
    pose = np.array(pose6, F)
    if n < 50:                                                     # :1079-1082
        return False, pose, np.zeros(6, F), None, None, is_degenerate, matP, False
    A, B = jacobian_rows(pose, ori, coef)
    At = cv2.transpose(A)
    AtA = cv2.gemm(At, A, 1.0, None, 0.0)                          # matAt * matA
    AtB = cv2.gemm(At, B, 1.0, None, 0.0)
    ok, X = cv2.solve(AtA, AtB, flags=cv2.DECOMP_QR)               # :1130
    assert ok
    if iter_count == 0:                                            # :1132-1154
        ok, E, V = cv2.eigen(AtA)
        V2 = V.copy()
        is_degenerate = False
        E = E.reshape(-1)
        for i in range(5, -1, -1):
            if E[i] < 100:
                V2[i, :] = 0
                is_degenerate = True
            else:
                break
        ok, Vinv = cv2.invert(V)
        matP = cv2.gemm(Vinv, V2, 1.0, None, 0.0)
    if is_degenerate:                                              # :1156-1160
        X = cv2.gemm(matP, X.copy(), 1.0, None, 0.0)
    X = X.reshape(6).astype(F)
    pose = (pose + X).astype(F)                                    # :1162-1167
    dR = F(np.sqrt(np.sum(np.rad2deg(X[:3].astype(F)).astype(np.float64) ** 2)))
    dT = F(np.sqrt(np.sum((X[3:] * F(100)).astype(np.float64) ** 2)))
    return True, pose, X, AtA, AtB.reshape(6), is_degenerate, matP, bool(dR < 0.05 and dT < 0.05)


def make_correspondences(rng, n, kind, pose6, offset):
    """Body-frame points on planes of a room (kind 'room') or of a corridor whose walls and floor all contain the x
    direction ('corridor': translation along x is unobservable).  coeff = (s n, s pd2) as surfOptimization emits it, with the
    residual pd2 produced by displacing the planes by `offset` (map frame) relative to where the pose puts the points."""
    cr, sr, cp, spp, cy, sy = (np.cos(pose6[0]), np.sin(pose6[0]), np.cos(pose6[1]), np.sin(pose6[1]), np.cos(pose6[2]), np.sin(pose6[2]))
    R = np.array([[cy * cp, cy * spp * sr - sy * cr, sy * sr + cy * spp * cr], [sy * cp, cy * cr + sy * spp * sr, sy * spp * cr - cy * sr], [-spp, cp * sr, cp * cr]])
    normals = np.array([[0, 0, 1.0], [0, 1.0, 0], [1.0, 0, 0]]) if kind == "room" else np.array([[0, 0, 1.0], [0, 1.0, 0], [0, 0.6, 0.8]])
    ori, coef = [], []
    for i in range(n):
        nrm = normals[i % 3] + rng.normal(0, 0.01, 3)           # measured normals: nearly, not exactly, degenerate
        nrm = nrm / np.linalg.norm(nrm)
        p_map = rng.uniform(-15, 15, 3)
        p_body = R.T @ (p_map - pose6[3:])
        pd2 = nrm @ offset + rng.normal(0, 0.01)
        s = 1 - 0.9 * abs(pd2) / np.sqrt(np.linalg.norm(p_body))
        ori.append(p_body); coef.append(np.r_[s * nrm, s * pd2])
    return np.array(ori, F), np.array(coef, F)


def main():
    rng = np.random.default_rng(1130)
    cases = []
    pose = np.array([0.01, -0.02, 0.7, 3.0, -2.0, 0.5])
    o, c = make_correspondences(rng, 400, "room", pose, np.array([0.08, -0.05, 0.03]))
    cases.append(("room_iter0", pose, o, c, 0, False, np.eye(6, dtype=F)))
    o2, c2 = make_correspondences(rng, 300, "corridor", pose, np.array([0.0, 0.06, -0.04]))
    cases.append(("corridor_iter0", pose, o2, c2, 0, False, np.eye(6, dtype=F)))
    cases.append(("few", pose, o[:49], c[:49], 0, False, np.eye(6, dtype=F)))
    o3, c3 = make_correspondences(rng, 200, "room", pose, np.array([1e-5, -1e-5, 2e-5]))
    c3[:, 3] *= F(0.01)
    cases.append(("converged", pose, o3, c3, 3, False, np.eye(6, dtype=F)))
    out = {}
    state = {}
    for name, p, oo, cc, it, deg, P in cases:
        ran, pose_out, X, AtA, AtB, deg_out, P_out, conv = lm_step_cv2(p, oo, cc, it, deg, P)
        state[name] = (deg_out, P_out)
        out[name] = dict(pose=np.array(p, F), ori=oo, coef=cc, iter=it, deg_in=deg, P_in=P, ran=ran, pose_out=pose_out, X=X,
                         AtA=AtA if AtA is not None else np.zeros((6, 6), F), AtB=AtB if AtB is not None else np.zeros(6, F),
                         deg_out=deg_out, P_out=P_out, converged=conv)
    # iteration 1 of the corridor scene re-uses the projector of its iteration 0 (the members persist, :90-91)
    deg, P = state["corridor_iter0"]
    p1 = out["corridor_iter0"]["pose_out"]
    o4, c4 = make_correspondences(rng, 300, "corridor", p1.astype(np.float64), np.array([0.0, 0.02, -0.01]))
    ran, pose_out, X, AtA, AtB, deg_out, P_out, conv = lm_step_cv2(p1, o4, c4, 1, deg, P)
    out["corridor_iter1"] = dict(pose=p1, ori=o4, coef=c4, iter=1, deg_in=deg, P_in=P, ran=ran, pose_out=pose_out, X=X, AtA=AtA, AtB=AtB,
                                 deg_out=deg_out, P_out=P_out, converged=conv)
    assert out["corridor_iter0"]["deg_out"] and not out["room_iter0"]["deg_out"] and out["converged"]["converged"] and not out["few"]["ran"]
    flat = {f"{k}.{f}": v for k, d in out.items() for f, v in d.items()}
    np.savez_compressed(os.path.join(HERE, "lm_step_cv2.npz"), names=np.array(list(out)), cv2_version=np.array(cv2.__version__), **flat)
    for k, d in out.items():
        print(k, "ran", d["ran"], "deg", d["deg_out"], "conv", d["converged"], "X", np.round(d["X"], 5).tolist())


if __name__ == "__main__":
    main()
