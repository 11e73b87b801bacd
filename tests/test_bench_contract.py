"""CPU tests of bench.py's contract: the reference arm (the one arm that runs without a GPU) prints exactly one JSON line on
stdout with the keys the driver reads, for the lio workload and for an NDT workload; the product arm refuses to run
without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
        "data", "config", "e2e", "cpu_baseline"}


def _run(*args, timeout=600):
    env = dict(os.environ, LILOC_BENCH_QUICK="1")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def _one_json_line(stdout):
    lines = [l for l in stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                      # nothing but the result on stdout
    return json.loads(lines[0])


def test_reference_arm_lio_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-frames", "3", "--frames", "40")
    assert r.returncode == 0, r.stderr[-2000:]
    d = _one_json_line(r.stdout)
    assert KEYS <= set(d), KEYS - set(d)
    assert d["impl"] == "reference" and d["metric"] == "scan2map_registrations_per_s" and d["unit"] == "registrations/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert "workload" in d["config"] and "model" not in d["config"]
    # both arms describe the workload with the very same `config` object (bench.lio_config)
    import argparse
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.lio_config(argparse.Namespace(frames=40, seed=2000, ref_frames=3))


def test_reference_arm_does_not_map_the_product_libraries():
    """The CPU arm must not load libliloc_gpu.so / libliloc_host.so (it imports liloc_b200.parallel and .synth only; the
    package binding is lazy and the yaml is read with PyYAML)."""
    code = ("import sys; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1', '--ref-frames', '2', '--frames', '40'];"
            f"sys.path.insert(0, {ROOT!r}); import bench; bench.main();"
            "m = open('/proc/self/maps').read(); import os; os.write(2, ('LIBS ' + ' '.join(sorted({l.split('/')[-1] for l in m.splitlines() if 'libliloc' in l})) + '\\n').encode())")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=dict(os.environ, LILOC_BENCH_QUICK="1"), cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    libs = [l for l in r.stderr.splitlines() if l.startswith("LIBS ")][-1].split()[1:]
    assert "libliloc_gpu.so" not in libs and "libliloc_host.so" not in libs, libs
    assert "libliloc_oracle.so" in libs


def test_reference_arm_ndt_line():
    r = _run("--impl", "reference", "--workload", "relo", "--items", "4", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    d = _one_json_line(r.stdout)
    assert KEYS <= set(d), KEYS - set(d)
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = _run("--steps", "1", "--warmup", "1", "--frames", "5", "--no-cpu")
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout) or "CUDA" in r.stderr


def test_reference_arm_under_torchrun_only_rank0_works():
    """Launched the way the driver launches N > 1: rank 0 alone runs and prints the line, the other rank exits 0 silently."""
    env = dict(os.environ, LILOC_BENCH_QUICK="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
           "--ref-frames", "3", "--frames", "40"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = _one_json_line(r.stdout)
    assert d["impl"] == "reference" and d["value"] > 0
