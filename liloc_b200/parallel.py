"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in the CPU tests).

A single frame does not shard: up to 20 dependent Gauss-Newton iterations over a
working set that fits in L2.  What shards is naturally independent work -- whole lio sequences
(replicas), pose hypotheses, prior submaps, (scan, submap, guess) triples -- so ranks own disjoint
blocks of work items, run them with no data-path collective, and exchange ONLY the packed results
(8 floats per item: 6-DoF pose, score/iters, flags) in one all-gather at the end.
"""
from __future__ import annotations

import os

import numpy as np

RESULT_WIDTH = 8   # floats per work item: pose6, a, b


def env_world() -> tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1 process if absent)."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def init(backend: str | None = None):
    """Initialise torch.distributed if WORLD_SIZE > 1.  Returns (rank, world, local_rank)."""
    rank, world, local = env_world()
    if world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(local)
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def comm_init(ctx) -> None:
    """Give a liloc context its NCCL communicator (liloc_comm_init): rank 0 creates the unique id through the C ABI, the
    launcher's process group (torch.distributed) only carries those 128 bytes to the other ranks."""
    rank, world, _ = env_world()
    if world == 1:
        ctx.comm_init(0, 1, None)
        return
    import torch.distributed as dist
    from . import abi
    ids = [abi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_init(rank, world, ids[0])


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of n_items owned by `rank`; blocks differ by at most one item."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_round_robin(n_items: int, rank: int, world: int) -> np.ndarray:
    """Items rank, rank + world, ... (used when item cost correlates with position, e.g. submap ids)."""
    return np.arange(rank, n_items, world, dtype=np.int64)


def pack_results(poses6: np.ndarray, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """(n, 8) float32 rows {roll, pitch, yaw, x, y, z, a, b}."""
    n = len(poses6)
    out = np.zeros((n, RESULT_WIDTH), np.float32)
    out[:, :6] = np.asarray(poses6, np.float32).reshape(n, 6)
    out[:, 6] = np.asarray(a, np.float32)
    out[:, 7] = np.asarray(b, np.float32)
    return out


def gather_results(local_rows: np.ndarray, n_items: int, device=None) -> np.ndarray:
    """All-gather of the per-rank result blocks (contiguous shard_range layout) -> (n_items, 8) on every rank.
    The only collective of the path: n_items * 32 bytes."""
    rank, world, _ = env_world()
    local_rows = np.ascontiguousarray(local_rows, np.float32).reshape(-1, RESULT_WIDTH)
    if world == 1:
        assert len(local_rows) == n_items
        return local_rows
    import torch
    import torch.distributed as dist
    per = (n_items + world - 1) // world                 # pad every block to the same size for all_gather_into_tensor
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                             if dist.get_backend() == "nccl" else torch.device("cpu"))
    send = torch.zeros((per, RESULT_WIDTH), dtype=torch.float32, device=dev)
    send[: len(local_rows)] = torch.from_numpy(local_rows).to(dev)
    recv = torch.empty((world * per, RESULT_WIDTH), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(recv, send)
    recv = recv.cpu().numpy().reshape(world, per, RESULT_WIDTH)
    out = np.empty((n_items, RESULT_WIDTH), np.float32)
    for r in range(world):
        lo, hi = shard_range(n_items, r, world)
        out[lo:hi] = recv[r, : hi - lo]
    return out


def gather_padded(local_rows: np.ndarray, per: int, device=None) -> np.ndarray:
    """All-gather of per-rank blocks padded to `per` rows -> (world, per, 8) on every rank."""
    rank, world, _ = env_world()
    local_rows = np.ascontiguousarray(local_rows, np.float32).reshape(-1, RESULT_WIDTH)
    if world == 1:
        out = np.zeros((1, per, RESULT_WIDTH), np.float32)
        out[0, : len(local_rows)] = local_rows
        return out
    import torch
    import torch.distributed as dist
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                             if dist.get_backend() == "nccl" else torch.device("cpu"))
    send = torch.zeros((per, RESULT_WIDTH), dtype=torch.float32, device=dev)
    send[: len(local_rows)] = torch.from_numpy(local_rows).to(dev)
    recv = torch.empty((world * per, RESULT_WIDTH), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(recv, send)
    return recv.cpu().numpy().reshape(world, per, RESULT_WIDTH)


def max_over_ranks(value: float, device=None) -> float:
    rank, world, _ = env_world()
    if world == 1:
        return float(value)
    import torch
    import torch.distributed as dist
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                             if dist.get_backend() == "nccl" else torch.device("cpu"))
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    rank, world, _ = env_world()
    if world == 1:
        return float(value)
    import torch
    import torch.distributed as dist
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                             if dist.get_backend() == "nccl" else torch.device("cpu"))
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    _, world, _ = env_world()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
