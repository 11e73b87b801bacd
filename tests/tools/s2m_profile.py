import os, sys, numpy as np
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import liloc_b200
from liloc_b200 import synth
import oracle_lib as orc
name = sys.argv[1] if len(sys.argv) > 1 else "hdl32"
sensor, n_kf, scan_leaf, map_leaf = {"vlp16": (synth.VLP16, 50, 0.2, 0.5), "hdl32": (synth.HDL32, 17, 0.4, 0.4), "mid360": (synth.MID360, 130, 0.2, 0.2)}[name]
step = 0.5 if name == "mid360" else 1.0
world = synth.make_world(size_xy=(120.0, 80.0), n_pillars=40, seed=7)
traj = synth.trajectory_line(n_kf + 1, step=step, x0=-step * n_kf / 2)
kfs = [orc.voxelgrid(synth.scan(world, traj[k], sensor, 100 + k), scan_leaf) for k in range(n_kf)]
raw = synth.scan(world, traj[n_kf], sensor, 999)
guess = synth.perturb(traj[n_kf], 3)
with liloc_b200.Context(device=0, max_scan_points=len(raw) + 1024, max_raw_points=1 << 21) as ctx:
    for k, c in enumerate(kfs): ctx.keyframe_put(k, c)
    ids, poses = list(range(n_kf)), traj[:n_kf].astype(np.float32)
    rows = []
    for r in range(60):
        pose, res, rc = ctx.frame(ids, poses, map_leaf, raw, scan_leaf, guess)
        if r >= 10: rows.append(ctx.last_clocks(res.iters).astype(np.float64))
    c = np.median(np.stack(rows), axis=0) / 1.965e3   # us
    print(name, "iters", res.iters, "q_all", res.q_all)
    print("per iteration us: tiles, store+barrier, reduce, solve, refresh | kNN(last tile), fit+row(thread0), wait")
    for i in range(res.iters): print(np.round(c[i], 2))
