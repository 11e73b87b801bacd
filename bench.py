#!/usr/bin/env python
"""bench.py -- scan-to-map registrations/s of the LiLoc hot path on B200 (BASELINE.json metric).

Workload (config.workload): BASELINE.json configs[1] -- synthetic HDL-32E NCLT-shaped scans, nclt.yaml
parameters, lio mode, a 1000-frame sequence pushed through mapOptimization::laserCloudInfoHandler
(liloc_b200/host -> C ABI -> sm_100a kernels).  One "step" = one pass over the whole sequence
(local-map assembly + VoxelGrid, scan VoxelGrid, scan2MapOptimization, keyframe bookkeeping, for every frame).

  value   registrations/s with the raw scans already resident in HBM (device pointers in cloud_info)
  e2e     the same through the same call with the scans in pinned HOST memory: per-frame H2D of the scan and
          the keyframe descriptors and D2H of the result are inside the timed region
  N > 1   one process per GPU (torchrun); a lio sequence does not shard, so every rank runs
          its own replica sequence (different seed) -- weak scaling; the only collective is the final
          all-gather of packed poses.  value = frames of all ranks / max-over-ranks time.

`--impl reference` times the reference's CPU path (the oracle restatement, oracle/ref_pipeline.py, OpenMP on
all host cores) on a bounded sample of the same workload; rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "scan2map_registrations_per_s"
UNIT = "registrations/s"
YAML = "nclt.yaml"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], threading.Event(), None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self.stop_flag.is_set():
                    break
        except Exception:
            pass

    def finish(self) -> dict:
        self.stop_flag.set()
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(n_frames: int, seed: int, n_total: int | None = None):
    from liloc_b200 import synth, host_api
    dt = host_api.yaml_param(YAML, "timeInterval") + 0.01     # every frame passes the rate gate (:257)
    return synth.make_sequence(synth.HDL32, n_frames, seed=seed, dt=dt, n_total=n_total)


# ---------------------------------------------------------------------------------------------------------
def run_reference(args):
    """CPU reference arm: oracle restatement of the reference's lio callback, OpenMP on all host cores."""
    from liloc_b200 import parallel
    rank, world, _ = parallel.env_world()
    if rank != 0:
        return
    import oracle_lib as orc
    import ref_pipeline as rp
    threads = orc.num_threads()
    per_step = args.ref_frames
    total = (args.warmup + args.steps) * per_step
    seq = make_workload(max(total, 2), args.seed, n_total=args.frames)   # head of the same 1000-frame sequence
    lio = rp.RefLio(rp.Params(YAML), threads=threads)
    f = 0
    for _ in range(args.warmup):
        for _ in range(per_step):
            lio.handle(seq.stamps[f], seq.odom[f], seq.scans[f]); f += 1
    lat, t0 = [], time.perf_counter()
    for _ in range(args.steps):
        for _ in range(per_step):
            rec = lio.handle(seq.stamps[f], seq.odom[f], seq.scans[f]); f += 1
            lat.append(rec["ms"])
    dt = time.perf_counter() - t0
    n = args.steps * per_step
    value = n / dt
    sample = f"frames {args.warmup * per_step}..{f - 1} of the sequence ({per_step} frames per step), map built by the preceding frames"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "p50_frame_ms": float(np.percentile(lat, 50)), "p95_frame_ms": float(np.percentile(lat, 95)),
        "config": {"workload": "configs[1]: HDL-32E NCLT-shaped lio sequence, nclt.yaml", "frames_per_step": per_step, "yaml": YAML,
                   "max_iters": 20, "stride": 2, "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host": {"nproc": os.cpu_count()},
    }))


# ---------------------------------------------------------------------------------------------------------
def cpu_baseline_sample(seq, n_sample: int, budget_s: float):
    """Oracle timed on the box's host cores over the head of the same sequence (rank 0, N = 1)."""
    import oracle_lib as orc
    import ref_pipeline as rp
    threads = orc.num_threads()
    lio = rp.RefLio(rp.Params(YAML), threads=threads)
    warm = min(10, max(len(seq.scans) // 4, 1))
    for f in range(warm):
        lio.handle(seq.stamps[f], seq.odom[f], seq.scans[f])
    t0, n = time.perf_counter(), 0
    for f in range(warm, min(len(seq.scans), warm + n_sample)):
        lio.handle(seq.stamps[f], seq.odom[f], seq.scans[f]); n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": n / dt if dt > 0 else None, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"frames {warm}..{warm + n - 1} of the benchmark sequence after {warm} warm-up frames ({n} frames, {dt:.1f} s)"}


def run_ours(args):
    import torch
    from liloc_b200 import abi, host_api, parallel
    rank, world, local = parallel.init()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- liloc_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    seq = make_workload(args.frames, args.seed + 1000 * rank)
    F = len(seq.scans)
    sizes = [len(s) for s in seq.scans]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    flat = np.concatenate(seq.scans).astype(np.float32)
    host_pinned = torch.from_numpy(flat).pin_memory()                       # e2e: scans in pinned host memory
    dev_scans = host_pinned.to(dev)                                         # value: scans resident in HBM
    torch.cuda.synchronize()
    hp, dp = host_pinned.data_ptr(), dev_scans.data_ptr()
    host_views = [host_pinned.numpy()[offs[i]:offs[i + 1]] for i in range(F)]
    infos_host = host_api.make_infos(host_views, seq.odom, seq.stamps)
    infos_dev = host_api.make_infos(None, seq.odom, seq.stamps, device_ptrs=[(dp + 16 * int(offs[i]), sizes[i]) for i in range(F)])
    assert all(infos_host[i].cloud_deskewed == hp + 16 * int(offs[i]) for i in (0, F - 1))

    m = host_api.MapOptimization(YAML, "lio", device=local)
    ctxh = m.ctx_handle()
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def one_pass(infos):
        m.reset()
        l2_flush.fill_(1)                                                   # flush L2 between timed passes
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = m.run(infos)
        torch.cuda.synchronize()
        return time.perf_counter() - t0, reps

    abi.lib.liloc_set_timing(ctxh, 1)
    for _ in range(args.warmup):
        one_pass(infos_dev)
        one_pass(infos_host)

    def timed(infos):
        parallel.barrier(); torch.cuda.synchronize()
        abi.lib.liloc_stats_reset(ctxh)
        sampler = ClockSampler(local); sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ext = torch.cuda.ExternalStream(abi.lib.liloc_stream(ctxh), device=dev)
        times, lat, last = [], [], None
        ev0.record(ext)
        for _ in range(args.steps):
            dt, reps = one_pass(infos)
            times.append(dt); last = reps
            lat += [r.reg_ms for r in reps if r.processed and r.iters > 0]
        ev1.record(ext)
        torch.cuda.synchronize(); parallel.barrier()
        clocks = sampler.finish()
        stats = abi.stats_of(ctxh)
        return sum(times), lat, last, stats, clocks, ev0.elapsed_time(ev1) * 1e-3

    t_dev, lat_dev, reps, st_dev, clocks, ev_dev = timed(infos_dev)
    t_e2e, lat_e2e, _, st_e2e, _, ev_e2e = timed(infos_host)

    n_proc = sum(1 for r in reps if r.processed)
    t_dev_max = parallel.max_over_ranks(t_dev, dev)
    t_e2e_max = parallel.max_over_ranks(t_e2e, dev)
    frames_all = parallel.sum_over_ranks(n_proc * args.steps, dev)
    launches_all = parallel.sum_over_ranks(st_dev["launches"], dev)
    # the path's only collective: packed final poses of every rank (8 floats per frame)
    rows = parallel.pack_results(np.array([list(r.pose) for r in reps], np.float32), [r.iters for r in reps], [r.n_sel for r in reps])
    lo, hi = parallel.shard_range(F * world, rank, world)
    gathered = parallel.gather_results(rows, F * world, dev)
    assert gathered.shape[0] == F * world and hi - lo == F

    if rank != 0:
        return
    value = frames_all / t_dev_max
    e2e = frames_all / t_e2e_max
    peak, peak_src = load_peaks()
    # roofline of the dominant kernel by share of the frame: decided from the measured phase times
    phases = {k: st_dev[k] for k in ("assemble_ms", "map_voxel_ms", "scan_ms", "grid_ms", "s2m_ms")}
    s2m_gbs = st_dev["alg_bytes_s2m"] / (st_dev["s2m_ms"] * 1e-3) / 1e9 if st_dev["s2m_ms"] > 0 else None
    map_gbs = st_dev["alg_bytes_map"] / ((st_dev["assemble_ms"] + st_dev["map_voxel_ms"]) * 1e-3) / 1e9 if st_dev["map_voxel_ms"] > 0 else None
    scan_gbs = st_dev["alg_bytes_scan"] / (st_dev["scan_ms"] * 1e-3) / 1e9 if st_dev["scan_ms"] > 0 else None
    dominant = max(phases, key=phases.get)
    roof = {"bound": "hbm", "kernel": "k_scan2map (persistent cooperative, one launch per frame)", "achieved": s2m_gbs, "peak": peak, "unit": "GB/s",
            "frac": (s2m_gbs / peak) if s2m_gbs else None, "traffic": None, "peak_source": peak_src,
            "avg_launch_us": st_dev["s2m_ms"] * 1e3 / max(st_dev["s2m_launches"], 1),
            "avg_iters_per_launch": st_dev["s2m_iters"] / max(st_dev["s2m_launches"], 1),
            "note": "latency-bound: <= 20 dependent Gauss-Newton iterations over an L2-resident working set (SURVEY 8d)",
            "phase_share": {k: v / max(st_dev["total_ms"], 1e-9) for k, v in phases.items()}, "dominant_phase": dominant,
            "map_assemble_voxel": {"achieved": map_gbs, "frac": (map_gbs / peak) if map_gbs else None},
            "scan_voxel": {"achieved": scan_gbs, "frac": (scan_gbs / peak) if scan_gbs else None}}
    cpu = cpu_baseline_sample(seq, args.cpu_frames, args.cpu_budget) if (world == 1 and not args.no_cpu) else None
    its = [r.iters for r in reps if r.processed and r.iters > 0]
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_dev_max * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "p50_frame_ms": float(np.percentile(lat_dev, 50)), "p95_frame_ms": float(np.percentile(lat_dev, 95)),
        "e2e_p50_frame_ms": float(np.percentile(lat_e2e, 50)), "e2e_p95_frame_ms": float(np.percentile(lat_e2e, 95)),
        "config": {"workload": "configs[1]: HDL-32E NCLT-shaped lio sequence, nclt.yaml, one replica sequence per GPU",
                   "frames_per_step": F, "yaml": YAML, "max_iters": 20, "stride": 2, "seed": args.seed,
                   "l2": "flushed (256 MiB fill) between timed passes", "timer": "host clock bracketed by device synchronisation; CUDA-event cross-check in device_s",
                   "sizes": {"n_scan": int(np.mean(sizes)), "q_all": int(np.mean([r.q_all for r in reps if r.processed])),
                             "map_points": int(np.mean([r.map_points for r in reps if r.map_points > 0])),
                             "keyframes_selected": float(np.mean([r.n_selected_keyframes for r in reps if r.n_selected_keyframes > 0])),
                             "iters": float(np.mean(its)), "keyframes_total": m.num_keyframes()}},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": st_e2e["h2d_bytes"] / args.steps, "d2h_bytes_per_step": st_e2e["d2h_bytes"] / args.steps},
        "gpu_launches": int(launches_all),
        "device_s": {"value_pass": ev_dev, "e2e_pass": ev_e2e},
        "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
        "host": {"nproc": os.cpu_count()},
    }
    print(json.dumps(out))
    m.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--frames", type=int, default=1000, help="frames per sequence (BASELINE configs[1]: 1000)")
    ap.add_argument("--seed", type=int, default=2000)
    ap.add_argument("--ref-frames", type=int, default=20, help="reference arm: frames per step")
    ap.add_argument("--cpu-frames", type=int, default=300, help="cpu_baseline: at most this many frames")
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="cpu_baseline: seconds of CPU work")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and not os.environ.get("LILOC_BENCH_QUICK"):
        args.warmup = 3          # timing rule: at least 3 warm-up steps (LILOC_BENCH_QUICK=1 lifts it for profiler runs)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
