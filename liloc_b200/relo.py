"""Relo-mode work items over the C ABI: what LiLoc's relocalisation init (src/mapOptmization.cpp:261-305) and
AnchorOptimization::addScanMatchingFactor (include/egoOptimization/anchorOptimize.hpp:382-421) ask of
`registration` (pcl::Registration, :118), batched and sharded.

A work item is (target submap, source scan, initial guess 4x4).  Items are independent, so they shard across
ranks with no data-path collective; the only exchange is the final all-gather of 8 floats per item
(liloc_b200.parallel.gather_results).  Nothing here computes: it packs guesses and calls liloc_ndt_*.
"""
from __future__ import annotations

import numpy as np

from . import parallel
from .abi import Context, NdtOpts, shard_indices


def euler_to_matrix(roll: float, pitch: float, yaw: float, x: float, y: float, z: float) -> np.ndarray:
    """init_guess of the relocalisation (src/mapOptmization.cpp:276-280): eulerToRotation (include/math_tools.h:229-237,
    R = Rz(yaw) Ry(pitch) Rx(roll)) with the translation in the last column; float32 4x4, row-major."""
    cr, sr, cp, sp, cy, sy = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = [[cy * cp, cy * sp * sr - sy * cr, sy * sr + cy * sp * cr],
                 [sy * cp, cy * cr + sy * sp * sr, sy * sp * cr - cy * sr],
                 [-sp, cp * sr, cp * cr]]
    T[:3, 3] = (x, y, z)
    return T


def matrix_to_pose6(T: np.ndarray) -> np.ndarray:
    """RotMtoEuler (include/math_tools.h:92-111) + translation -> [roll, pitch, yaw, x, y, z] (:288-303)."""
    R = np.asarray(T, np.float64)
    sy = np.sqrt(R[0, 0] * R[0, 0] + R[1, 0] * R[1, 0])
    if sy >= 1e-6:
        roll, pitch, yaw = np.arctan2(R[2, 1], R[2, 2]), np.arctan2(-R[2, 0], sy), np.arctan2(R[1, 0], R[0, 0])
    else:
        roll, pitch, yaw = np.arctan2(-R[1, 2], R[1, 1]), np.arctan2(-R[2, 0], sy), 0.0
    return np.array([roll, pitch, yaw, R[0, 3], R[1, 3], R[2, 3]], np.float32)


def matrices_to_pose6(T: np.ndarray) -> np.ndarray:
    """matrix_to_pose6 over a (n, 4, 4) stack."""
    R = np.asarray(T, np.float64)
    sy = np.sqrt(R[:, 0, 0] ** 2 + R[:, 1, 0] ** 2)
    ok = sy >= 1e-6
    roll = np.where(ok, np.arctan2(R[:, 2, 1], R[:, 2, 2]), np.arctan2(-R[:, 1, 2], R[:, 1, 1]))
    pitch = np.arctan2(-R[:, 2, 0], sy)
    yaw = np.where(ok, np.arctan2(R[:, 1, 0], R[:, 0, 0]), 0.0)
    return np.stack([roll, pitch, yaw, R[:, 0, 3], R[:, 1, 3], R[:, 2, 3]], 1).astype(np.float32)


def hypothesis_grid(center6, n_yaw: int = 16, n_xy: int = 16, xy_radius: float = 2.0, seed: int = 0) -> np.ndarray:
    """Multi-hypothesis relocalisation guesses (BASELINE configs[3]): /initialpose only fixes x, y and yaw
    (src/mapOptmization.cpp:764-769), so hypotheses spread n_yaw yaw offsets over the full circle times n_xy
    positions inside a disc of xy_radius around the given pose.  Returns (n_yaw * n_xy, 4, 4) float32."""
    c = np.asarray(center6, np.float64)
    rng = np.random.default_rng(seed)
    r = xy_radius * np.sqrt(rng.uniform(0, 1, n_xy)); a = rng.uniform(0, 2 * np.pi, n_xy)
    r[0] = 0.0
    out = []
    for iy in range(n_yaw):
        yaw = c[2] + 2 * np.pi * iy / n_yaw
        for k in range(n_xy):
            out.append(euler_to_matrix(c[0], c[1], yaw, c[3] + r[k] * np.cos(a[k]), c[4] + r[k] * np.sin(a[k]), c[5]))
    return np.stack(out)


def align_sharded(ctx: Context, target_ids, source_ids, guesses, opts: NdtOpts | None = None, device=None, round_robin: bool = False,
                  want_local: bool = True):
    """Run this rank's share of the items and all-gather the packed results -- all inside the C ABI
    (liloc_ndt_align_batch_sharded): the kernel packs {pose6, score / n, iterations} on the device, one ncclAllGather on
    the context's stream, one device-to-host copy.  The context must have been given its communicator
    (parallel.comm_init) when WORLD_SIZE > 1.

    Every rank passes the SAME full item list; targets / sources named by a rank's items must already be resident in
    its context (liloc_ndt_target_put / liloc_ndt_source_put).  Contiguous blocks by default; round_robin for lists
    whose cost correlates with position.  Returns (rows (n_items, 8) = pose6, score / n_source, iterations; local dict)."""
    rank, world = ctx.comm_world()
    n = len(guesses)
    mine = shard_indices(n, rank, world, round_robin)
    local = {"items": mine, "T": np.zeros((0, 4, 4), np.float32), "sweeps": np.zeros(0, np.int64), "iterations": np.zeros(0, np.int64)}
    if want_local:
        rows, outs = ctx.ndt_align_batch_sharded(target_ids, source_ids, guesses, opts, round_robin=round_robin, want_local=True)
        outs = outs[: len(mine)]
        local.update(T=outs["T"].reshape(-1, 4, 4).copy(), sweeps=outs["sweeps"].copy(), iterations=outs["iterations"].copy(),
                     converged=outs["converged"].copy(), score=outs["score"].copy())
    else:
        rows = ctx.ndt_align_batch_sharded(target_ids, source_ids, guesses, opts, round_robin=round_robin)
    return rows, local
