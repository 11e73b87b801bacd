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
