"""GPU tests of the sharded NDT entry point (liloc_ndt_align_batch_sharded) and of the C-ABI gather
(liloc_comm_init / liloc_gather_results): the multi-hypothesis relocalisation (src/mapOptmization.cpp:261-305 for many
initial poses) and the multi-submap scan matching (anchorOptimize.hpp:382-421) partitioned over ranks, one ncclAllGather of
8 floats per item.  The 2-rank test needs two GPUs (one process and one context per GPU, NCCL id handed over in a file);
on a single-GPU box it is skipped and the world-1 path is what runs."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from test_ndt_oracle import guess_matrix, ndt_scene  # noqa: E402,F401


def _gpu_count():
    import torch
    return torch.cuda.device_count()


def test_sharded_world1_equals_batch(gpu_ctx, synth, ndt_scene):
    """Without a communicator the sharded call runs the whole list here; its packed rows are the batch call's results
    (RotMtoEuler of the final transform, score / n, iterations)."""
    from liloc_b200 import NdtOpts, relo
    sub, src, truth = ndt_scene
    gpu_ctx.ndt_clear()
    gpu_ctx.ndt_target_put(0, sub, 1.0)
    gpu_ctx.ndt_source_put(0, src)
    guesses = relo.hypothesis_grid(truth, 4, 5, 1.0, seed=1)
    n = len(guesses)
    T, score, iters, conv, sweeps = gpu_ctx.ndt_align_batch([0] * n, [0] * n, guesses, NdtOpts(n_align=2))
    for rr in (False, True):
        rows, outs = gpu_ctx.ndt_align_batch_sharded([0] * n, [0] * n, guesses, NdtOpts(n_align=2), round_robin=rr, want_local=True)
        assert rows.shape == (n, 8)
        assert np.array_equal(outs["T"].reshape(n, 4, 4), T) and np.array_equal(outs["iterations"], iters) and np.array_equal(outs["sweeps"], sweeps)
        want = relo.matrices_to_pose6(T)
        assert np.abs(rows[:, :6] - want).max() <= 2e-6                     # double atan2 on both sides, float rounding
        assert np.array_equal(rows[:, 3:6], T[:, :3, 3])
        assert np.allclose(rows[:, 6], score / len(src), rtol=1e-6) and np.array_equal(rows[:, 7], iters.astype(np.float32))
    g = gpu_ctx.gather_results(rows[:5], 8)
    assert g.shape == (1, 8, 8) and np.array_equal(g[0, :5], rows[:5]) and not g[0, 5:].any()


WORKER = textwrap.dedent("""
    import os, sys, time
    import numpy as np
    sys.path.insert(0, ROOT_DIR); sys.path.insert(0, os.path.join(ROOT_DIR, "tests"))
    rank, world, idfile = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    from liloc_b200 import abi, relo, synth, Context, NdtOpts
    import oracle_lib as orc
    # the id travels in a file: the C ABI needs no launcher library at all
    if rank == 0:
        uid = abi.comm_unique_id()
        with open(idfile + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(idfile + ".tmp", idfile)
    else:
        for _ in range(600):
            if os.path.exists(idfile):
                break
            time.sleep(0.1)
        uid = open(idfile, "rb").read()
    world_ = synth.make_world(seed=3)
    traj = synth.trajectory_line(12, step=1.0, x0=-5)
    sub = np.concatenate([orc.transform_cloud(synth.scan(world_, traj[k], synth.VLP16, 30 + k), traj[k].astype(np.float32)) for k in range(10)])
    sub2 = sub[::2].copy()
    src = synth.scan(world_, traj[10], synth.VLP16, 77)
    guesses = relo.hypothesis_grid(traj[10], 4, 5, 1.0, seed=1)
    n = len(guesses)
    tids = np.arange(n) % 2
    with Context(device=rank) as ctx:
        ctx.comm_init(rank, world, uid)
        assert ctx.comm_world() == (rank, world)
        for rr in (False, True):
            mine = abi.shard_indices(n, rank, world, rr)
            ctx.ndt_clear()
            for t in sorted(set(tids[mine].tolist())):        # only the grids this rank's share names
                ctx.ndt_target_put(int(t), sub if t == 0 else sub2, 1.0)
            ctx.ndt_source_put(0, src)
            rows, outs = ctx.ndt_align_batch_sharded(tids, [0] * n, guesses, NdtOpts(n_align=2), round_robin=rr, want_local=True)
            # every row against the oracle, on every rank (the gathered table is complete everywhere)
            o = orc.NdtOpts(resolution=1.0, trans_eps=0.01, search=0, threads=4)
            tg = {0: orc.NdtTarget(sub, np.float32(1.0)), 1: orc.NdtTarget(sub2, np.float32(1.0))}
            for j in range(n):
                T = guesses[j][:3]
                it = 0
                for _ in range(2):
                    T, r = orc.ndt_align(tg[int(tids[j])], src, T, o)
                    it += r.iterations
                want = relo.matrix_to_pose6(np.vstack([T, [0, 0, 0, 1]]))
                assert np.abs(rows[j, 3:6] - want[3:]).max() <= 1e-4 and np.abs(rows[j, :3] - want[:3]).max() <= 1e-5, (j, rows[j], want)
                assert rows[j, 7] == it, (j, rows[j, 7], it)
            assert len(outs) >= len(mine)
        g = ctx.gather_results(np.full((3, 8), rank + 1.0, np.float32), 4)
        assert g.shape == (world, 4, 8) and all(np.all(g[r, :3] == r + 1.0) and not g[r, 3:].any() for r in range(world))
        ctx.comm_destroy()
    print("rank", rank, "ok")
""")


def test_two_rank_sharded_align_nccl(tmp_path):
    if _gpu_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.replace("ROOT_DIR", repr(ROOT)))
    idfile = str(tmp_path / "nccl_id.bin")
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2", idfile], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o[-3000:]
        assert f"rank {r} ok" in o


HOST_WORKER = textwrap.dedent("""
    import os, sys, time
    import numpy as np
    sys.path.insert(0, ROOT_DIR); sys.path.insert(0, os.path.join(ROOT_DIR, "tests"))
    rank, world, idfile = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    from liloc_b200 import abi, host_api, synth
    from test_gpu_relo import _hypothesis_scene
    uid = None
    if world > 1:
        if rank == 0:
            uid = abi.comm_unique_id()
            with open(idfile + ".tmp", "wb") as f:
                f.write(uid)
            os.replace(idfile + ".tmp", idfile)
        else:
            for _ in range(600):
                if os.path.exists(idfile):
                    break
                time.sleep(0.1)
            uid = open(idfile, "rb").read()
    prior_traj, prior_clouds, truth, scan = _hypothesis_scene(synth)
    keyposes7 = np.c_[prior_traj, 100.0 + np.arange(len(prior_traj))]
    init = np.array([0, 0, truth[2] + np.pi / 2 + 0.1, truth[3] + 1.0, truth[4] - 0.8, 0], np.float32)
    with host_api.MapOptimization("lio_sam_default.yaml", "relo", device=rank) as m:
        m.comm_init(rank, world, uid)
        m.load_prior_session(keyposes7, prior_clouds)
        m.set_relo_hypotheses(8, 6, 2.0)
        m.initialpose(init[3], init[4], init[2])
        rep = m.handle(host_api.make_infos([scan], truth.astype(np.float32)[None], np.array([600.0]))[0])
        assert rep.processed == 1 and rep.relo_initialized == 1
        print("RESULT", rank, rep.relo_hypothesis, " ".join(repr(float(v)) for v in rep.relo_pose), repr(float(rep.relo_score)))
""")


def test_host_class_multi_hypothesis_init_on_two_gpus(tmp_path):
    """The C++ host class on two GPUs: each process one `mapOptimization` with its context's communicator; the 48 hypotheses of
    the relocalisation init are split between the ranks inside liloc_ndt_align_batch_sharded and both ranks end up with the
    same winner -- the one a single GPU finds."""
    if _gpu_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    script = tmp_path / "host_worker.py"
    script.write_text(HOST_WORKER.replace("ROOT_DIR", repr(ROOT)))

    def run(world):
        idfile = str(tmp_path / f"id_{world}.bin")
        procs = [subprocess.Popen([sys.executable, str(script), str(r), str(world), idfile], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                 for r in range(world)]
        outs = [p.communicate(timeout=600)[0] for p in procs]
        res = []
        for p, o in zip(procs, outs):
            assert p.returncode == 0, o[-3000:]
            res.append([l for l in o.splitlines() if l.startswith("RESULT")][-1].split()[2:])
        return res
    one = run(1)[0]
    two = run(2)
    assert two[0] == two[1] == one, (one, two)
