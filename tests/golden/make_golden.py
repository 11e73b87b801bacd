"""Generates tests/golden/*.npz -- independent known answers for the third-party arithmetic the
reference delegates to.  The reference itself has no tests or golden vectors and
cannot be built here, so these come from libraries that ARE in this container:

  * cv2 (OpenCV, the library the reference calls at src/mapOptmization.cpp:1127-1153):
        cv2.solve(A, b, flags=cv2.DECOMP_QR), cv2.eigen(A) on 6x6 float32 systems
  * scipy.spatial.cKDTree: exact 5-NN (stand-in for pcl::KdTreeFLANN, :992)
  * numpy.linalg.lstsq: the 5x3 plane fit (stand-in for Eigen colPivHouseholderQr, :1009)

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    import cv2
    from scipy.spatial import cKDTree
    from liloc_b200 import synth
    import oracle_lib as orc

    rng = np.random.default_rng(20240919)

    # ---- 6x6 normal equations: realistic J^T J from a synthetic scene + random SPD + degenerate ones
    world = synth.make_world(seed=3)
    traj = synth.trajectory_line(8)
    kfs = [orc.voxelgrid(synth.scan(world, traj[k], synth.VLP16, 50 + k), 0.4) for k in range(6)]
    m = orc.localmap_build(kfs, traj[:6], 0.5)
    sc = orc.voxelgrid(synth.scan(world, traj[6], synth.VLP16, 77), 0.4)
    As, bs = [], []
    for s in range(6):
        guess = synth.perturb(traj[6], 100 + s)
        _, _, ori, coef, _ = orc.surf_pass(m, sc, guess, stride=2)
        rows = np.array([np.append(*orc.jacobian_row(guess, o, c)) for o, c in zip(ori, coef)], dtype=np.float64)
        J, r = rows[:, :6], rows[:, 6]
        As.append((J.T @ J).astype(np.float32)); bs.append((J.T @ r).astype(np.float32))
    for s in range(6):
        B = rng.normal(size=(40, 6)) * rng.uniform(0.1, 30, 6)
        As.append((B.T @ B).astype(np.float32)); bs.append(rng.normal(size=6).astype(np.float32) * 10)
    for s in range(4):                      # rank-deficient directions: eigenvalues below the 100 threshold
        B = rng.normal(size=(60, 6)) * np.array([40, 40, 40, 30, 0.2 * (s + 1), 0.05])
        As.append((B.T @ B).astype(np.float32)); bs.append(rng.normal(size=6).astype(np.float32))
    A = np.stack(As); b = np.stack(bs)
    xs, ws, Vs = [], [], []
    for Ai, bi in zip(A, b):
        ok, x = cv2.solve(Ai, bi.reshape(6, 1), flags=cv2.DECOMP_QR)
        assert ok
        ok, w, V = cv2.eigen(Ai)
        xs.append(x.reshape(6)); ws.append(w.reshape(6)); Vs.append(V)
    np.savez_compressed(os.path.join(HERE, "lm6x6_cv2.npz"), A=A, b=b, x=np.stack(xs), w=np.stack(ws), V=np.stack(Vs),
                        cv2_version=np.array(cv2.__version__))

    # ---- exact 5-NN on tie-free data (continuous noise)
    map_pts = rng.uniform(-20, 20, size=(4000, 3)).astype(np.float32)
    map_pts[:, 2] *= 0.1
    q = rng.uniform(-20, 20, size=(300, 3)).astype(np.float32)
    q[:, 2] *= 0.1
    d, i = cKDTree(map_pts.astype(np.float64)).query(q.astype(np.float64), k=5)
    np.savez_compressed(os.path.join(HERE, "knn5_ckdtree.npz"), map=map_pts, q=q, idx=i.astype(np.int32), dist=d.astype(np.float64))

    # ---- 5x3 plane fits
    P = []
    for s in range(64):
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        c = rng.uniform(-50, 50, 3)
        u = np.cross(n, [1, 0, 0.3]); u /= np.linalg.norm(u); v = np.cross(n, u)
        pts = c + rng.uniform(-0.6, 0.6, (5, 1)) * u + rng.uniform(-0.6, 0.6, (5, 1)) * v + rng.normal(0, 0.02, (5, 1)) * n
        P.append(pts.astype(np.float32))
    P = np.stack(P)
    X = np.stack([np.linalg.lstsq(p.astype(np.float64), -np.ones(5), rcond=None)[0] for p in P])
    np.savez_compressed(os.path.join(HERE, "plane5x3_lstsq.npz"), P=P, X=X)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()


def reference_params():
    """tests/golden/reference_params.json: the values of every ParamServer key (include/utility.h:195-304)
    as written in the reference's own config/*.yaml, parsed with pyyaml in this container."""
    import json
    import yaml
    ref = "/root/reference/config"
    out = {}
    for name in ("lio_sam_default.yaml", "nclt.yaml", "m2dgr.yaml", "lio_sam_mid360.yaml"):
        with open(os.path.join(ref, name)) as f:
            y = yaml.safe_load(f)
        flat = {}
        for sec in ("System", "Sensors", "Mapping"):
            for k, v in y[sec].items():
                if k in ("savePCDDirectory", "saveSessionDirectory"):
                    continue
                flat[f"{sec}/{k}"] = v
        out[name] = flat
    with open(os.path.join(HERE, "reference_params.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__" and os.path.isdir("/root/reference/config"):
    reference_params()
