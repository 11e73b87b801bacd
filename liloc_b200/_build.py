"""Build recipes for the native pieces (in-tree .so files; they travel to the GPU box with the snapshot).

    libliloc_gpu.so    liloc_b200/csrc/liloc_gpu.cu    nvcc, sm_100a only      -- the product
    libliloc_host.so   liloc_b200/host/*.cpp           g++                     -- mapOptimization host class over the C ABI
    libliloc_synth.so  liloc_b200/synth/synth.cpp      g++ -fopenmp            -- synthetic scan generator (inputs only)
    oracle/libliloc_oracle.so                          g++ -fopenmp -ffp-contract=off -- parity oracle (tests/bench only)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "liloc_b200")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",              # arithmetic contract: no FMA contraction
    "-shared", "-Xcompiler", "-fPIC",
]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd))


def _nvcc() -> str:
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _glob(d: str, exts: tuple[str, ...]) -> list[str]:
    out = []
    for base, _, files in os.walk(d):
        for f in files:
            if f.endswith(exts):
                out.append(os.path.join(base, f))
    return sorted(out)


def build_gpu(force: bool = False) -> str:
    target = os.path.join(PKG, "libliloc_gpu.so")
    srcs = _glob(os.path.join(PKG, "csrc"), (".cu", ".cuh", ".h")) + [os.path.join(ROOT, "include", "liloc_gpu.h")]
    if force or not _newer(target, srcs):
        _run([_nvcc(), *NVCC_FLAGS, "-o", target, os.path.join(PKG, "csrc", "liloc_gpu.cu"), "-ldl"])
    return target


def build_host(force: bool = False) -> str:
    target = os.path.join(PKG, "libliloc_host.so")
    d = os.path.join(PKG, "host")
    srcs = _glob(d, (".cpp", ".hpp", ".h")) + [os.path.join(ROOT, "include", "liloc_gpu.h")]
    cpps = [s for s in srcs if s.endswith(".cpp")]
    if not cpps:
        return ""
    if force or not _newer(target, srcs):
        _run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), "-o", target, *cpps,
              "-L", PKG, "-lliloc_gpu", "-Wl,-rpath,$ORIGIN"])
    return target


def build_synth(force: bool = False) -> str:
    target = os.path.join(PKG, "libliloc_synth.so")
    src = os.path.join(PKG, "synth", "synth.cpp")
    if force or not _newer(target, [src]):
        _run(["g++", "-O3", "-std=c++17", "-fopenmp", "-shared", "-fPIC", "-o", target, src])
    return target


def build_oracle(force: bool = False) -> str:
    d = os.path.join(ROOT, "oracle")
    target = os.path.join(d, "libliloc_oracle.so")
    srcs = _glob(d, (".cpp", ".h"))
    if force or not _newer(target, srcs):
        _run(["make", "-C", d, "-s"] + (["-B"] if force else []))
    return target


def build_all(force: bool = False) -> dict[str, str]:
    out = {"gpu": build_gpu(force), "synth": build_synth(force), "oracle": build_oracle(force)}
    h = build_host(force)
    if h:
        out["host"] = h
    return out


if __name__ == "__main__":
    for k, v in build_all(force="--force" in sys.argv).items():
        print(f"{k}: {v}")
