import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # native pieces are built in-tree; (re)build what is stale so a fresh checkout can run the CPU suite
    # (the recipe is loaded by path: importing the package dlopens libliloc_gpu.so, which may be stale or missing)
    import importlib.util
    spec = importlib.util.spec_from_file_location("liloc_b200_build", os.path.join(ROOT, "liloc_b200", "_build.py"))
    _build = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(_build)
    _build.build_all()


@pytest.fixture(scope="session")
def orc():
    import oracle_lib
    return oracle_lib


@pytest.fixture(scope="session")
def synth():
    from liloc_b200 import synth as s
    return s


@pytest.fixture(scope="session")
def gpu_ctx():
    """One liloc context on cuda:0 for the whole GPU session.  No fallback: fails if the device or the
    CUDA library is unusable."""
    from liloc_b200 import Context
    ctx = Context(device=0, max_scan_points=32 * 1800, max_raw_points=1 << 20)
    yield ctx
    ctx.close()


class Scene:
    """A small seeded lio scene: K keyframes (already down-sampled, body frame), their poses, one scan
    with ground truth and a perturbed initial guess."""

    def __init__(self, synth, orc, sensor, n_kf, scan_leaf, map_leaf, seed, world=None, step=1.0):
        self.sensor, self.scan_leaf, self.map_leaf = sensor, scan_leaf, map_leaf
        self.world = synth.make_world(seed=seed) if world is None else world
        traj = synth.trajectory_line(n_kf + 1, step=step, x0=-step * n_kf / 2)
        self.kf_poses = traj[:n_kf].astype(np.float32)
        self.kf_clouds = [orc.voxelgrid(synth.scan(self.world, traj[k], sensor, seed * 1000 + k), scan_leaf)
                          for k in range(n_kf)]
        self.truth = traj[n_kf].astype(np.float32)
        self.scan_raw = synth.scan(self.world, traj[n_kf], sensor, seed * 1000 + 999)
        self.guess = synth.perturb(traj[n_kf], seed)


@pytest.fixture(scope="session")
def scene_small(synth, orc):
    return Scene(synth, orc, synth.VLP16, n_kf=6, scan_leaf=0.4, map_leaf=0.5, seed=11)


@pytest.fixture(scope="session")
def scene_hdl(synth, orc):
    return Scene(synth, orc, synth.HDL32, n_kf=15, scan_leaf=0.4, map_leaf=0.4, seed=12)
