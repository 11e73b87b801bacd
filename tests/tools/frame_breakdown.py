"""Device-side latency breakdown of one lio frame (run on the GPU box):  python tests/tools/frame_breakdown.py [vlp16|hdl32|mid360]

(Lives under tests/ because it also times the CPU oracle on the same frame; nothing outside tests/, smoke() and
bench.py's CPU legs may touch oracle/.)

Prints the median duration of every phase of the three launches of a frame (local-map pipeline, scan pipeline,
registration) from the %globaltimer marks block 0 writes at the phase boundaries (liloc_last_marks), next to the
CUDA-event totals.  Inputs: a static synthetic scene of the chosen sensor (same generator as the tests)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    import liloc_b200
    from liloc_b200 import synth
    import oracle_lib as orc
    name = sys.argv[1] if len(sys.argv) > 1 else "hdl32"
    sensor, n_kf, scan_leaf, map_leaf = {"vlp16": (synth.VLP16, 50, 0.2, 0.5), "hdl32": (synth.HDL32, 17, 0.4, 0.4),
                                         "mid360": (synth.MID360, 130, 0.2, 0.2)}[name]
    step = 0.5 if name == "mid360" else 1.0
    world = synth.make_world(size_xy=(120.0, 80.0), n_pillars=40, seed=7)
    traj = synth.trajectory_line(n_kf + 1, step=step, x0=-step * n_kf / 2)
    kfs = [orc.voxelgrid(synth.scan(world, traj[k], sensor, 100 + k), scan_leaf) for k in range(n_kf)]
    raw = synth.scan(world, traj[n_kf], sensor, 999)
    guess = synth.perturb(traj[n_kf], 3)
    reps = int(os.environ.get("REPS", 300))
    with liloc_b200.Context(device=0, max_scan_points=len(raw) + 1024, max_raw_points=1 << 21) as ctx:
        for k, c in enumerate(kfs):
            ctx.keyframe_put(k, c)
        ids, poses = list(range(n_kf)), traj[:n_kf].astype(np.float32)
        ctx.set_timing(True)
        rows, wall = [], []
        for r in range(reps + 20):
            t0 = time.perf_counter()
            pose, res, rc = ctx.frame(ids, poses, map_leaf, raw, scan_leaf, guess)
            wall.append((time.perf_counter() - t0) * 1e6)
            mk = ctx.last_marks().astype(np.int64)
            tm = ctx.last_timing()
            if r < 20:
                continue
            m, s, g = mk[:16], mk[16:32], mk[32:]
            it = res.iters
            row = {
                "map.assemble+bbox": m[1] - m[0], "map.keys": m[2] - m[1], "map.sort": m[3] - m[2], "map.run_count": m[4] - m[3],
                "map.centroids": m[5] - m[4], "map.cell_scan": m[6] - m[5], "map.cell_scatter": m[7] - m[6], "map.total": m[7] - m[0],
                "scan.bbox": s[1] - s[0], "scan.keys": s[2] - s[1], "scan.sort": s[3] - s[2], "scan.run_count": s[4] - s[3],
                "scan.centroids": s[5] - s[4], "scan.total": s[5] - s[0],
                "scan_start_after_map_start": s[0] - m[0],
                "s2m.start_after_map_end": g[0] - m[7],
                "s2m.tiles(avg/iter)": np.mean([g[1 + 3 * i] - (g[0] if i == 0 else g[3 * i]) for i in range(it)]),
                "s2m.barrier(avg/iter)": np.mean([g[2 + 3 * i] - g[1 + 3 * i] for i in range(it)]),
                "s2m.reduce+solve(avg/iter)": np.mean([g[3 + 3 * i] - g[2 + 3 * i] for i in range(it)]),
                "s2m.total": g[3 * it] - g[0], "s2m.iters": it * 1000,
                "frame.device(map start -> s2m end)": g[3 * it] - m[0],
                "timing.scan_ms": tm["scan_ms"] * 1e6, "timing.s2m_ms": tm["s2m_ms"] * 1e6, "timing.total_ms": tm["total_ms"] * 1e6,
            }
            rows.append(row)
        ex = {"n_scan": len(raw), "q_all": res.q_all, "map_points": res.map_points, "n_raw": int(sum(len(c) for c in kfs)), "iters": res.iters}
    out = {k: float(np.median([r[k] for r in rows])) / 1e3 for k in rows[0]}
    out["wall_us(p50 host clock per frame call)"] = float(np.median(wall[20:]))
    if os.environ.get("CPU", "1") != "0":
        # the reference's CPU path on the same frame (oracle restatement, OpenMP on all host cores; faithful redundancy:
        # re-concatenation, single-threaded VoxelGrid, search tree rebuilt) -- the north_star's p50 latency comparison
        threads = max(orc.num_threads(), os.cpu_count() or 1)
        pts = np.concatenate(kfs); offs = np.zeros(n_kf + 1, np.int64); offs[1:] = np.cumsum([len(c) for c in kfs])
        cpu_ms = []
        for _ in range(5):
            t0 = time.perf_counter()
            rp, rres, rrc, M, Q, ms = orc.frame(pts, offs, poses, map_leaf, raw, scan_leaf, guess, orc.S2mOpts(knn=1, threads=threads))
            cpu_ms.append((time.perf_counter() - t0) * 1e3)
        assert (rres.iters, rres.n_sel) == (res.iters, res.n_sel) and np.abs(rp - pose).max() < 1e-4
        out["cpu_reference_ms(p50, %d threads)" % threads] = float(np.median(cpu_ms)) * 1e3      # printed in us like the rest
        out["cpu_reference.assemble+mapDS"] = float(ms[0]) * 1e3; out["cpu_reference.scanDS"] = float(ms[1]) * 1e3; out["cpu_reference.scan2map"] = float(ms[2]) * 1e3
        out["p50 latency ratio CPU reference / GPU frame call"] = float(np.median(cpu_ms)) * 1e3 / out["wall_us(p50 host clock per frame call)"]
    print(f"# frame latency breakdown, {name}, {ex}, median of {reps} frames, microseconds")
    for k, v in out.items():
        print(f"{k:45s} {v:9.2f}")
    print(json.dumps({"sensor": name, "sizes": ex, "us": out}))


if __name__ == "__main__":
    main()
