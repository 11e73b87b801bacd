"""Generates tests/golden/corner3x3_cv2.npz -- known answers for the 3x3 symmetric eigen-decomposition that
upstream LIO-SAM's cornerOptimization delegates to cv::eigen (the reference itself has no
corner path).  cv2 is the Python build of the same OpenCV routine.

Run from the repo root:  python tests/golden/make_golden_corner.py
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    import cv2
    rng = np.random.default_rng(20261019)
    A = []
    for s in range(48):
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        c = rng.uniform(-60, 60, 3)
        if s < 32:      # five points along an edge with LiDAR-like noise (line-like: l0 >> l1)
            pts = c + rng.uniform(-0.7, 0.7, (5, 1)) * d + rng.normal(0, 0.02, (5, 3))
        else:           # blob / planar patches (must fail the l0 > 3 l1 test most of the time)
            pts = c + rng.normal(0, 0.3, (5, 3)) * np.array([1.0, 1.0, 0.05 if s % 2 else 1.0])
        pts = pts.astype(np.float32)
        m = pts.mean(0, dtype=np.float32)
        q = pts - m
        A.append(((q.T @ q) / np.float32(5)).astype(np.float32))
    A = np.stack(A)
    ws, Vs = [], []
    for Ai in A:
        ok, w, V = cv2.eigen(Ai)
        assert ok
        ws.append(w.reshape(3)); Vs.append(V)
    np.savez_compressed(os.path.join(HERE, "corner3x3_cv2.npz"), A=A, w=np.stack(ws), V=np.stack(Vs), cv2_version=np.array(cv2.__version__))
    print("written", os.path.join(HERE, "corner3x3_cv2.npz"))


if __name__ == "__main__":
    main()
