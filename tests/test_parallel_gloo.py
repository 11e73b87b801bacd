"""CPU test of the N > 1 host path: two processes, gloo backend, 127.0.0.1 rendezvous.  Checks the
work partition and the final all-gather of packed 6-DoF results (the only collective of the path)."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_cover_everything():
    from liloc_b200 import parallel
    for n in (0, 1, 7, 64, 256, 8000):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = parallel.shard_range(n, r, world)
                assert 0 <= lo <= hi <= n and hi - lo in (n // world, n // world + 1)
                seen += list(range(lo, hi))
            assert seen == list(range(n))
            rr = np.concatenate([parallel.shard_round_robin(n, r, world) for r in range(world)])
            assert sorted(rr.tolist()) == list(range(n))
            # the C ABI's own partition (liloc_shard_indices) is the same one
            from liloc_b200 import abi
            for r in range(world):
                assert abi.shard_indices(n, r, world, False).tolist() == list(range(*parallel.shard_range(n, r, world)))
                assert abi.shard_indices(n, r, world, True).tolist() == parallel.shard_round_robin(n, r, world).tolist()


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    sys.path.insert(0, ROOT_DIR)
    from liloc_b200 import parallel
    rank, world, _ = parallel.init("gloo")
    n_items = 37
    lo, hi = parallel.shard_range(n_items, rank, world)
    ids = np.arange(lo, hi)
    poses = np.stack([ids * 1.0, ids * 2.0, ids * 3.0, ids + 0.25, ids + 0.5, ids + 0.75], 1)
    rows = parallel.pack_results(poses, ids % 5, rank * np.ones(len(ids)))
    parallel.barrier()
    allr = parallel.gather_results(rows, n_items)
    assert allr.shape == (n_items, 8)
    want = np.arange(n_items)
    assert np.array_equal(allr[:, 0], want) and np.array_equal(allr[:, 5], want + 0.75) and np.array_equal(allr[:, 6], want % 5)
    owners = allr[:, 7]
    assert set(owners.tolist()) == {0.0, 1.0}
    assert parallel.max_over_ranks(float(rank + 1)) == 2.0 and parallel.sum_over_ranks(1.5) == 3.0
    print("rank", rank, "ok")
""")


def test_two_rank_gather_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.replace('ROOT_DIR', repr(ROOT)))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for rank, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {rank} ok" in o


def test_relo_guess_helpers():
    from liloc_b200 import relo
    rng = np.random.default_rng(0)
    for _ in range(50):
        p = np.r_[rng.uniform(-1.2, 1.2, 2), rng.uniform(-3.1, 3.1), rng.uniform(-50, 50, 3)]
        T = relo.euler_to_matrix(*p)
        assert T.dtype == np.float32 and np.allclose(T[:3, :3] @ T[:3, :3].T, np.eye(3), atol=1e-6)
        assert np.allclose(relo.matrix_to_pose6(T), p, atol=2e-6 * 50)
    g = relo.hypothesis_grid([0.01, -0.02, 0.3, 5.0, -2.0, 1.8], 16, 16, 2.0, seed=3)
    assert g.shape == (256, 4, 4)
    assert np.allclose(g[0], relo.euler_to_matrix(0.01, -0.02, 0.3, 5.0, -2.0, 1.8))          # first hypothesis = the given pose
    d = np.hypot(g[:, 0, 3] - 5.0, g[:, 1, 3] + 2.0)
    assert d.max() <= 2.0 + 1e-5 and np.allclose(g[:, 2, 3], 1.8)
    yaws = np.array([relo.matrix_to_pose6(t)[2] for t in g[::16]])
    assert len(np.unique(np.round(yaws, 4))) == 16


RELO_WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    sys.path.insert(0, ROOT_DIR)
    from liloc_b200 import abi, parallel, relo
    rank, world, _ = parallel.init("gloo")
    # The host side of liloc_ndt_align_batch_sharded without a device: the share of each rank (liloc_shard_indices), a REAL
    # two-process all-gather of the padded per-rank blocks (gloo here, ncclAllGather on the GPUs) and the de-interleave
    # (liloc_unshard_rows) -- the result of an item is a function of the item alone, so the gathered table is checkable.
    n = 37
    tids, sids = np.arange(n) % 5, np.arange(n) % 3
    guesses = np.stack([relo.euler_to_matrix(0, 0, 0.01 * i, float(i), 2.0 * i, 1.0) for i in range(n)])
    per = (n + world - 1) // world
    for rr in (False, True):
        mine = abi.shard_indices(n, rank, world, rr)
        assert len(mine) in (n // world, n // world + 1)
        assert (mine[1] - mine[0] == world) if rr else (mine[1] - mine[0] == 1)
        T = guesses[mine].copy()
        T[:, 0, 3] += tids[mine] + 0.5 * sids[mine]
        rows = parallel.pack_results(relo.matrices_to_pose6(T), T[:, 1, 3], tids[mine] % 7)
        gathered = parallel.gather_padded(rows, per)
        full = abi.unshard_rows(gathered, n, rr)
        assert full.shape == (n, 8)
        want_x = np.arange(n) + tids + 0.5 * sids
        assert np.allclose(full[:, 3], want_x) and np.allclose(full[:, 4], 2.0 * np.arange(n)) and np.array_equal(full[:, 7], tids % 7)
    print("rank", rank, "ok")
""")


def test_relo_items_shard_and_gather_gloo(tmp_path):
    script = tmp_path / "relo_worker.py"
    script.write_text(RELO_WORKER.replace('ROOT_DIR', repr(ROOT)))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for rank, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {rank} ok" in o
