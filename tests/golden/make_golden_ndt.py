#!/usr/bin/env python
"""Golden vectors for the relo-mode NDT path, made WITHOUT the oracle (oracle/ndt_oracle.cpp) and without the device.

ndt_omp is un-vendored and unpinned in the reference (README.md:39,47) and nothing in the reference tests it, so parity
for this path is unpinned at the source.  What this script pins instead, from the published algorithm (Magnusson 2009,
"The Three-Dimensional Normal-Distributions Transform", ch. 6; PCL / ndt_omp lineage as summarised in SURVEY.md A.6):

  (A) float64 mathematics: voxel Gaussians (VoxelGridCovariance: >= 6 points, covariance formula, eigenvalue clamp at
      0.01 * max, inverse) in numpy float64; the score, its gradient and Hessian at three poses with the Jacobian and the
      second derivatives of T(p) x obtained SYMBOLICALLY (sympy) from T(p) = Trans * Rx * Ry * Rz -- no hand-written
      derivative tables.  The oracle / the device evaluate the per-point terms in float32, so they must agree with these
      sums to ~1e-5 relative.  Source points whose transformed position lies within 1e-3 of a voxel face are left out, so
      the voxel assignment cannot depend on float32 / float64 rounding.
  (B) float32 arithmetic: the per-(point, voxel) terms of ndt_omp's updateDerivatives restated in numpy float32, vectorised,
      with the same operand order (left-to-right sums of products, exp in double rounded to float, the angle tables of
      computeAngleDerivatives in double rounded to float), summed in double in ascending point order.  The oracle's
      sequential sums must reproduce these to double rounding; its canonical (tile-ordered) sums and the device's to ~1e-12.
  (C) one full registration in float64 (Newton step by numpy.linalg.solve, More-Thuente line search written from the 1994
      paper as PCL restates it) from three guesses: the oracle's and the device's final poses must land on the same
      optimum within the line search's own resolution.

Writes tests/golden/ndt_np.npz.  Run:  python tests/golden/make_golden_ndt.py
"""
import os

import numpy as np
import sympy as sp

HERE = os.path.dirname(os.path.abspath(__file__))
RES = 1.0
OUTLIER = 0.55
STEP_MAX, EPS, MAX_IT = 0.1, 0.01, 35


# ---------------------------------------------------------------------------------------------------------------------
def make_scene(seed=7):
    """A corridor corner: floor, two walls and a box, sampled with noise (target, map frame); the source is a noisy sample
    of the same surfaces seen from a pose 0.3 m / 0.03 rad away (body frame)."""
    rng = np.random.default_rng(seed)

    def plane(n, origin, u, v, su, sv):
        a, b = rng.uniform(0, su, n), rng.uniform(0, sv, n)
        return np.asarray(origin)[None] + a[:, None] * np.asarray(u)[None] + b[:, None] * np.asarray(v)[None]

    def surfaces(n):
        return np.concatenate([plane(n, (-8, -6, 0), (1, 0, 0), (0, 1, 0), 16, 12),          # floor
                               plane(n // 2, (-8, -6, 0), (1, 0, 0), (0, 0, 1), 16, 3),        # wall y = -6
                               plane(n // 2, (8, -6, 0), (0, 1, 0), (0, 0, 1), 12, 3),         # wall x = 8
                               plane(n // 4, (1, 1, 0), (1, 0, 0), (0, 0, 1), 2, 1.5),         # box faces
                               plane(n // 4, (1, 1, 0), (0, 1, 0), (0, 0, 1), 2, 1.5)])
    tgt = surfaces(6000) + rng.normal(0, 0.02, (6000 + 3000 + 3000 + 1500 + 1500, 3))
    src_map = surfaces(700) + rng.normal(0, 0.02, (700 + 350 + 350 + 175 + 175, 3))
    truth = np.array([0.3, -0.2, 0.05, 0.01, -0.015, 0.03])                                       # [tx ty tz rx ry rz]
    T = pose_matrix(truth)
    src = (src_map - T[:3, 3]) @ T[:3, :3]                                                       # body frame: R^T (x - t)
    return tgt.astype(np.float32), src.astype(np.float32), truth


def rot_xyz(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]); Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rx @ Ry @ Rz


def pose_matrix(p):
    T = np.eye(4)
    T[:3, :3] = rot_xyz(*p[3:]); T[:3, 3] = p[:3]
    return T


# ---------------------------------------------------------------------------------------------------------------------
def voxel_gaussians(tgt):
    """pclomp::VoxelGridCovariance at leaf RES, float64.  Returns dict voxel (i, j, k) -> (mean, icov, n)."""
    inv = np.float32(1.0) / np.float32(RES)
    ijk = np.floor(tgt * inv).astype(np.int64)
    out = {}
    order = np.lexsort((np.arange(len(tgt)), ijk[:, 0], ijk[:, 1], ijk[:, 2]))
    ijk_s = ijk[order]
    brk = np.flatnonzero(np.any(np.diff(ijk_s, axis=0) != 0, axis=1)) + 1
    for a, b in zip(np.r_[0, brk], np.r_[brk, len(order)]):
        if b - a < 6:
            continue
        P = tgt[order[a:b]].astype(np.float64)
        n = float(b - a)
        s = np.zeros(3); ss = np.zeros((3, 3))
        for x in P:                                                   # the accumulation order of the filter: source order
            s += x; ss += np.outer(x, x)
        mean = s / n
        cov = ((ss - 2.0 * np.outer(s, mean)) / n + np.outer(mean, mean)) * ((n - 1.0) / n)
        w, V = np.linalg.eigh(0.5 * (cov + cov.T))
        if w[0] < 0 or w[1] < 0 or w[2] <= 0:
            continue
        if w[0] < 0.01 * w[2]:
            w = w.copy(); w[0] = 0.01 * w[2]
            if w[1] < 0.01 * w[2]:
                w[1] = 0.01 * w[2]
            cov = V @ np.diag(w) @ V.T
        icov = np.linalg.inv(cov)
        if not np.isfinite(icov).all():
            continue
        out[tuple(ijk_s[a])] = (mean, icov, int(n))
    return out


def gauss_consts():
    c1, c2 = 10.0 * (1.0 - OUTLIER), OUTLIER / RES ** 3
    d3 = -np.log(c2)
    d1 = -np.log(c1 + c2) - d3
    d2 = -2.0 * np.log((-np.log(c1 * np.exp(-0.5) + c2) - d3) / d1)
    return d1, d2


# ---------------------------------------------------------------------------------------------------------------------
# (A) symbolic derivatives of T(p) x
def symbolic_derivatives():
    tx, ty, tz, rx, ry, rz, x, y, z = sp.symbols("tx ty tz rx ry rz x y z", real=True)
    Rx = sp.Matrix([[1, 0, 0], [0, sp.cos(rx), -sp.sin(rx)], [0, sp.sin(rx), sp.cos(rx)]])
    Ry = sp.Matrix([[sp.cos(ry), 0, sp.sin(ry)], [0, 1, 0], [-sp.sin(ry), 0, sp.cos(ry)]])
    Rz = sp.Matrix([[sp.cos(rz), -sp.sin(rz), 0], [sp.sin(rz), sp.cos(rz), 0], [0, 0, 1]])
    f = Rx * Ry * Rz * sp.Matrix([x, y, z]) + sp.Matrix([tx, ty, tz])
    p = [tx, ty, tz, rx, ry, rz]
    J = f.jacobian(p)                                                   # 3 x 6
    H = [[sp.diff(J[:, i], p[j]) for j in range(6)] for i in range(6)]  # 6 x 6 of 3-vectors
    args = (tx, ty, tz, rx, ry, rz, x, y, z)
    fJ = sp.lambdify(args, J, "numpy")
    fH = [[sp.lambdify(args, H[i][j], "numpy") for j in range(6)] for i in range(6)]
    return fJ, fH


def score_derivatives_f64(leaves, src, p, fJ, fH, margin=1e-3):
    """Score (as ndt_omp sums it: -d1 * exp(..)), gradient and Hessian of one sweep at pose p, DIRECT1, float64.
    Returns (sums28, used_mask)."""
    d1, d2 = gauss_consts()
    T = pose_matrix(p)
    X = src.astype(np.float64)
    Xt = X @ T[:3, :3].T + T[:3, 3]
    frac = Xt / RES - np.floor(Xt / RES)
    safe = np.all((frac > margin) & (frac < 1 - margin), axis=1)
    score, g, Hm = 0.0, np.zeros(6), np.zeros((6, 6))
    used = np.zeros(len(src), bool)
    for i in np.flatnonzero(safe):
        key = tuple(np.floor(Xt[i] / RES).astype(np.int64))
        if key not in leaves:
            used[i] = True                                             # safe and without a Gaussian: contributes nothing anywhere
            continue
        mean, icov, _ = leaves[key]
        # the leaf as the implementations store it: inverse covariance in float32
        C = icov.astype(np.float32).astype(np.float64)
        q = Xt[i] - mean
        e = np.exp(-d2 * (q @ C @ q) / 2.0)
        used[i] = True
        score += -d1 * e
        ge = d2 * e
        if ge > 1 or ge < 0 or np.isnan(ge):
            continue
        ge *= d1
        J = np.asarray(fJ(*p, *X[i]), np.float64).reshape(3, 6)
        qCJ = q @ C @ J
        g += ge * qCJ
        for a in range(6):
            for b in range(6):
                hab = np.asarray(fH[a][b](*p, *X[i]), np.float64).reshape(3)
                Hm[a, b] += ge * (-d2 * qCJ[a] * qCJ[b] + q @ C @ hab + J[:, b] @ C @ J[:, a])
    sums = np.r_[score, g, [Hm[a, b] for a in range(6) for b in range(a, 6)]]
    return sums, used


# ---------------------------------------------------------------------------------------------------------------------
# (B) float32 restatement of updateDerivatives (vectorised over the hits)
F = np.float32


def angle_tables(p):
    """computeAngleDerivatives: j_ang (8 x 3) and h_ang (15 x 3) in double, rounded to float32; angles below 1e-4 are
    treated as zero (cos 1, sin 0)."""
    def cs(a):
        return (1.0, 0.0) if abs(a) < 10e-5 else (np.cos(a), np.sin(a))
    (cx, sx), (cy, sy), (cz, sz) = cs(p[3]), cs(p[4]), cs(p[5])
    J = np.array([[-sx * sz + cx * sy * cz, -sx * cz - cx * sy * sz, -cx * cy],
                  [cx * sz + sx * sy * cz, cx * cz - sx * sy * sz, -sx * cy],
                  [-sy * cz, sy * sz, cy],
                  [sx * cy * cz, -sx * cy * sz, sx * sy],
                  [-cx * cy * cz, cx * cy * sz, -cx * sy],
                  [-cy * sz, -cy * cz, 0],
                  [cx * cz - sx * sy * sz, -cx * sz - sx * sy * cz, 0],
                  [sx * cz + cx * sy * sz, cx * sy * cz - sx * sz, 0]])
    H = np.array([[-cx * sz - sx * sy * cz, -cx * cz + sx * sy * sz, sx * cy],
                  [-sx * sz + cx * sy * cz, -cx * sy * sz - sx * cz, -cx * cy],
                  [cx * cy * cz, -cx * cy * sz, cx * sy],
                  [sx * cy * cz, -sx * cy * sz, sx * sy],
                  [-sx * cz - cx * sy * sz, sx * sz - cx * sy * cz, 0],
                  [cx * cz - sx * sy * sz, -sx * sy * cz - cx * sz, 0],
                  [-cy * cz, cy * sz, sy],
                  [-sx * sy * cz, sx * sy * sz, sx * cy],
                  [cx * sy * cz, -cx * sy * sz, -cx * cy],
                  [sy * sz, sy * cz, 0],
                  [-sx * cy * sz, -sx * cy * cz, 0],
                  [cx * cy * sz, cx * cy * cz, 0],
                  [-cy * cz, cy * sz, 0],
                  [-cx * sz - sx * sy * cz, -cx * cz + sx * sy * sz, 0],
                  [-sx * sz + cx * sy * cz, -cx * sy * sz - sx * cz, 0]])
    return J.astype(F), H.astype(F)


def pose_matrix_f32(p):
    """T(p) in float32 as computeStepLengthMT builds it: angles rounded to float, cos / sin in double rounded to float,
    (Rx * Ry) * Rz with float products, translation rounded to float."""
    a, b, c = (float(F(v)) for v in p[3:])
    ca, sa, cb, sb, cc, sc = (F(np.cos(a)), F(np.sin(a)), F(np.cos(b)), F(np.sin(b)), F(np.cos(c)), F(np.sin(c)))
    o, z = F(1), F(0)
    Rx = np.array([[o, z, z], [z, ca, -sa], [z, sa, ca]], F); Ry = np.array([[cb, z, sb], [z, o, z], [-sb, z, cb]], F)
    Rz = np.array([[cc, -sc, z], [sc, cc, z], [z, z, o]], F)

    def mm(A, B):
        R = np.zeros((3, 3), F)
        for i in range(3):
            for j in range(3):
                R[i, j] = F(F(F(A[i, 0] * B[0, j]) + F(A[i, 1] * B[1, j])) + F(A[i, 2] * B[2, j]))
        return R
    R = mm(mm(Rx, Ry), Rz)
    return R, np.array([F(p[0]), F(p[1]), F(p[2])], F)


def dot3(a0, b0, a1, b1, a2, b2):
    return (a0 * b0 + a1 * b1) + a2 * b2                             # float32 arrays: every product and sum rounds to float32


def sweep_f32(leaf_table, src, p):
    """updateDerivatives over all (point, voxel) hits, DIRECT1: float32 terms, double sums in ascending point order."""
    d1, d2d = gauss_consts()
    d2 = F(d2d)
    R, t = pose_matrix_f32(p)
    x = src.astype(F)
    xt = np.stack([dot3(R[i, 0], x[:, 0], R[i, 1], x[:, 1], R[i, 2], x[:, 2]) + t[i] for i in range(3)], 1).astype(F)
    inv = F(1.0) / F(RES)
    ijk = np.floor(xt * inv).astype(np.int64)
    hit = np.array([tuple(k) in leaf_table for k in ijk])
    idx = np.flatnonzero(hit)
    mean = np.array([leaf_table[tuple(ijk[i])][0] for i in idx])                      # float64
    C = np.array([leaf_table[tuple(ijk[i])][1].astype(F).reshape(9) for i in idx], F)  # float32, row-major
    x, xt = x[idx], xt[idx]
    q = (xt.astype(np.float64) - mean).astype(F)
    q0, q1, q2 = q[:, 0], q[:, 1], q[:, 2]
    qC = [dot3(q0, C[:, 0 + c], q1, C[:, 3 + c], q2, C[:, 6 + c]) for c in range(3)]
    qCq = dot3(q0, qC[0], q1, qC[1], q2, qC[2])
    e = np.exp((F(-1) * d2 * qCq * F(0.5)).astype(np.float64)).astype(F)
    score_inc = (-d1 * e.astype(np.float64)).astype(F)
    e = d2 * e
    ok = ~((e > 1) | (e < 0) | np.isnan(e))
    e = (e.astype(np.float64) * d1).astype(F)
    Jt, Ht = angle_tables(p)
    xj = [dot3(Jt[r, 0], x[:, 0], Jt[r, 1], x[:, 1], Jt[r, 2], x[:, 2]) for r in range(8)]
    xh = [dot3(Ht[r, 0], x[:, 0], Ht[r, 1], x[:, 1], Ht[r, 2], x[:, 2]) for r in range(15)]
    one, zero = np.ones(len(idx), F), np.zeros(len(idx), F)
    # point gradient (3 x 6): identity | column 3 = (0, xj0, xj1) | column 4 = (xj2, xj3, xj4) | column 5 = (xj5, xj6, xj7)
    g = [[one, zero, zero, zero, xj[2], xj[5]], [zero, one, zero, xj[0], xj[3], xj[6]], [zero, zero, one, xj[1], xj[4], xj[7]]]
    Cg = [[dot3(C[:, 3 * r], g[0][c], C[:, 3 * r + 1], g[1][c], C[:, 3 * r + 2], g[2][c]) for c in range(6)] for r in range(3)]
    qCg = [dot3(q0, Cg[0][c], q1, Cg[1][c], q2, Cg[2][c]) for c in range(6)]
    terms = np.zeros((len(idx), 28), F)
    terms[:, 0] = score_inc
    for c in range(6):
        terms[:, 1 + c] = e * qCg[c]
    ha, hb, hc = (zero, xh[0], xh[1]), (zero, xh[2], xh[3]), (zero, xh[4], xh[5])
    hd, he, hf = (xh[6], xh[7], xh[8]), (xh[9], xh[10], xh[11]), (xh[12], xh[13], xh[14])
    blk = [[ha, hb, hc], [hb, hd, he], [hc, he, hf]]
    k = 7
    for i in range(6):
        for j in range(i, 6):
            ht = zero
            if i >= 3 and j >= 3:
                v = blk[i - 3][j - 3]
                ht = dot3(qC[0], v[0], qC[1], v[1], qC[2], v[2])
            gCg = dot3(g[0][j], Cg[0][i], g[1][j], Cg[1][i], g[2][j], Cg[2][i])
            terms[:, k] = e * ((F(-1) * d2 * qCg[i] * qCg[j] + ht) + gCg)
            k += 1
    terms[~ok] = 0
    sums = np.zeros(28)
    for row in terms.astype(np.float64):                                                 # ascending point order
        sums += row
    return sums, len(idx)


# ---------------------------------------------------------------------------------------------------------------------
# (C) a full registration in float64
def mt_trial(a_l, f_l, g_l, a_u, f_u, g_u, a_t, f_t, g_t):
    if f_t > f_l:
        z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l; w = np.sqrt(z * z - g_t * g_l)
        a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w)
        a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t))
        return a_c if abs(a_c - a_l) < abs(a_q - a_l) else 0.5 * (a_q + a_c)
    if g_t * g_l < 0:
        z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l; w = np.sqrt(z * z - g_t * g_l)
        a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w)
        a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l
        return a_c if abs(a_c - a_t) >= abs(a_s - a_t) else a_s
    if abs(g_t) <= abs(g_l):
        z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l; w = np.sqrt(z * z - g_t * g_l)
        a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w)
        a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l
        a_n = a_c if abs(a_c - a_t) < abs(a_s - a_t) else a_s
        return min(a_t + 0.66 * (a_u - a_t), a_n) if a_t > a_l else max(a_t + 0.66 * (a_u - a_t), a_n)
    z = 3 * (f_t - f_u) / (a_t - a_u) - g_t - g_u; w = np.sqrt(z * z - g_t * g_u)
    return a_u + (a_t - a_u) * (w - g_u - z) / (g_t - g_u + 2 * w)


def align_f64(leaf_table, src, p0):
    def sweep(p):
        s, _ = sweep_f32(leaf_table, src, p)          # the objective itself: the float32-term sums (what all implementations minimise)
        Hm = np.zeros((6, 6)); k = 7
        for i in range(6):
            for j in range(i, 6):
                Hm[i, j] = Hm[j, i] = s[k]; k += 1
        return s[0], s[1:7], Hm
    p = np.array(p0, np.float64)
    score, g, Hm = sweep(p)
    it, mu, nu = 0, 1e-4, 0.9
    step_min = EPS / 2
    while True:
        dp = np.linalg.solve(Hm, -g)
        nrm = np.linalg.norm(dp)
        if nrm == 0 or np.isnan(nrm):
            break
        d = dp / nrm
        phi0, dphi0 = -score, -(g @ d)
        if dphi0 >= 0:
            if dphi0 == 0:
                a_t = 0.0
                conv = it > MAX_IT or (it and 0.0 < EPS); it += 1
                if conv:
                    break
                continue
            dphi0, d = -dphi0, -d
        a_l = a_u = 0.0
        f_l = f_u = 0.0; g_l = g_u = dphi0 - mu * dphi0
        a_t = max(min(nrm, STEP_MAX), step_min)
        open_iv, iv_conv, k = True, False, 0
        score, g, Hm = sweep(p + d * a_t)
        phi_t, dphi_t = -score, -(g @ d)
        psi_t, dpsi_t = phi_t - phi0 - mu * dphi0 * a_t, dphi_t - mu * dphi0
        while not iv_conv and k < 10 and not (psi_t <= 0 and dphi_t <= -nu * dphi0):
            a_t = mt_trial(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, dpsi_t) if open_iv else mt_trial(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, dphi_t)
            a_t = max(min(a_t, STEP_MAX), step_min)
            score, g, Hm = sweep(p + d * a_t)
            phi_t, dphi_t = -score, -(g @ d)
            psi_t, dpsi_t = phi_t - phi0 - mu * dphi0 * a_t, dphi_t - mu * dphi0
            if open_iv and psi_t <= 0 and dpsi_t >= 0:
                open_iv = False
                f_l += phi0 - mu * dphi0 * a_l; g_l += mu * dphi0
                f_u += phi0 - mu * dphi0 * a_u; g_u += mu * dphi0
            ft, gt = (psi_t, dpsi_t) if open_iv else (phi_t, dphi_t)
            if ft > f_l:
                a_u, f_u, g_u = a_t, ft, gt
            elif gt * (a_l - a_t) > 0:
                a_l, f_l, g_l = a_t, ft, gt
            elif gt * (a_l - a_t) < 0:
                a_u, f_u, g_u = a_l, f_l, g_l
                a_l, f_l, g_l = a_t, ft, gt
            else:
                iv_conv = True
            k += 1
        p = p + d * a_t
        conv = it > MAX_IT or (it and abs(a_t) < EPS)
        it += 1
        if conv:
            break
    return p, it


# ---------------------------------------------------------------------------------------------------------------------
def main():
    tgt, src, truth = make_scene()
    leaves = voxel_gaussians(tgt)
    keys = np.array(sorted(leaves, key=lambda k: (k[2], k[1], k[0])), np.int64)
    means = np.array([leaves[tuple(k)][0] for k in keys]); icovs = np.array([leaves[tuple(k)][1] for k in keys]); counts = np.array([leaves[tuple(k)][2] for k in keys])
    fJ, fH = symbolic_derivatives()
    rng = np.random.default_rng(11)
    poses = [truth.copy(), truth + np.r_[rng.uniform(-0.3, 0.3, 3), rng.uniform(-0.03, 0.03, 3)], truth + np.r_[rng.uniform(-0.5, 0.5, 3), rng.uniform(-0.05, 0.05, 3)]]
    # the source subset that is safely inside its voxel at every pose (A) is evaluated at
    masks = []
    for p in poses:
        T = pose_matrix(p)
        Xt = src.astype(np.float64) @ T[:3, :3].T + T[:3, 3]
        fr = Xt / RES - np.floor(Xt / RES)
        masks.append(np.all((fr > 1e-3) & (fr < 1 - 1e-3), axis=1))
    safe = np.all(masks, axis=0)
    src_a = src[safe]
    sums64 = [score_derivatives_f64(leaves, src_a, p, fJ, fH)[0] for p in poses]
    sums32, hits = zip(*[sweep_f32(leaves, src, p) for p in poses])
    sums32_a = [sweep_f32(leaves, src_a, p)[0] for p in poses]
    for a, b in zip(sums64, sums32_a):                      # the two restatements made here agree with each other
        assert np.allclose(a, b, rtol=2e-4, atol=2e-4 * np.abs(a).max()), np.abs(a - b).max()
    guesses = [truth + np.array([0.15, -0.1, 0.02, 0.01, -0.01, 0.02]), truth + np.array([-0.2, 0.15, -0.03, -0.015, 0.01, -0.03]),
               truth + np.array([0.05, 0.25, 0.0, 0.0, 0.02, 0.04])]
    finals, its = zip(*[align_f64(leaves, src, g) for g in guesses])
    np.savez_compressed(os.path.join(HERE, "ndt_np.npz"), target=tgt, source=src, source_safe=src_a, truth=truth, poses=np.array(poses),
                        leaf_keys=keys, leaf_means=means, leaf_icovs=icovs, leaf_counts=counts,
                        sums_f64_safe=np.array(sums64), sums_f32=np.array(sums32), sums_f32_safe=np.array(sums32_a), hits=np.array(hits),
                        guesses=np.array(guesses), finals=np.array(finals), final_iterations=np.array(its),
                        resolution=RES, outlier_ratio=OUTLIER, step_size=STEP_MAX, trans_eps=EPS)
    print("leaves", len(keys), "source", len(src), "safe", len(src_a), "hits", hits, "iterations", its)
    print("final - truth:", [np.round(f - truth, 4).tolist() for f in finals])


if __name__ == "__main__":
    main()
