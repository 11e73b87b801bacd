"""oracle/ref_pipeline.py -- the reference's per-frame lio callback on the CPU (TEST INFRASTRUCTURE).

Restates laserCloudInfoHandler for lio mode (src/mapOptmization.cpp:246-353): updateInitialGuess
(odometry-increment branch :865-883), extractNearby / extractCloud (:902-961), downsampleCurrentScan
(:970-975), scan2MapOptimization (:1183-1207) and the keyframe decision (:1245-1264, :1391-1461),
with the heavy lifting in oracle/liboracle (orc_frame keeps the reference's per-frame redundancy:
full re-concatenation, single-threaded VoxelGrid, search structure rebuilt every frame, OpenMP only
over the residual loop and the keyframe transform).  Used by bench.py's cpu_baseline and
`--impl reference` legs and by tests; never by the product.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_TESTS = os.path.join(os.path.dirname(_HERE), "tests")
if _TESTS not in sys.path:
    sys.path.insert(0, _TESTS)
import oracle_lib as orc  # noqa: E402


def yaml_param(yaml_name: str, key: str) -> float:
    """A ParamServer key (include/utility.h:195-304) of config/<yaml_name>, read with PyYAML: the CPU arms must not map the
    product's shared libraries (liloc_b200.host_api.yaml_param would load libliloc_host.so -> libliloc_gpu.so)."""
    import yaml
    path = yaml_name if os.path.isabs(yaml_name) else os.path.join(os.path.dirname(_HERE), "config", yaml_name)
    doc = yaml.safe_load(open(path))
    # the reference reads every real-valued key used here into a `float` member (include/utility.h:256-301)
    as_member = (lambda v: float(v)) if key in ("numberOfCores", "mappingProcessInterval") else (lambda v: float(np.float32(v)))
    for section in doc.values():
        if isinstance(section, dict) and key in section:
            return as_member(section[key])
    if key in doc:
        return as_member(doc[key])
    raise KeyError(f"{yaml_name}:{key}")


class Params:
    """The ParamServer keys the lio callback reads, from config/*.yaml (the parser is host glue, not part of the measured path)."""

    def __init__(self, yaml: str):
        g = lambda k: yaml_param(yaml, k)
        self.scan_leaf = np.float32(g("mappingSurfLeafSize"))
        self.map_leaf = np.float32(g("surroundingKeyframeMapLeafSize"))
        self.radius = np.float32(g("surroundingKeyframeSearchRadius"))
        self.density = np.float32(g("surroundingKeyframeDensity"))
        self.add_dist = np.float32(g("surroundingkeyframeAddingDistThreshold"))
        self.add_angle = np.float32(g("surroundingkeyframeAddingAngleThreshold"))
        self.time_interval = float(g("timeInterval"))
        self.cores = int(g("numberOfCores"))


class RefLio:
    """Sequential CPU lio pipeline.  threads: OpenMP threads for the residual loop / keyframe transform."""

    def __init__(self, params: Params, threads: int, max_iters=20, stride=2):
        self.p = params
        self.opts = orc.S2mOpts(max_iters=max_iters, stride=stride, min_sel=50, min_points=30, knn=1, threads=threads)
        self.state = orc.LmState()
        self.key_poses = []      # float32 [r,p,y,x,y,z]
        self.key_times = []
        self.key_clouds = []     # DS scans, body frame
        self.pose = np.zeros(6, np.float32)
        self.last_odom = None
        self.t_last = -1.0

    def _select(self, stamp):
        P = np.asarray(self.key_poses, np.float32)[:, 3:6]
        last = P[-1]
        d2 = ((last[0] - P[:, 0]) ** 2 + (last[1] - P[:, 1]) ** 2) + (last[2] - P[:, 2]) ** 2
        hit = np.flatnonzero(d2 < self.p.radius * self.p.radius)
        hit = hit[np.lexsort((hit, d2[hit]))]
        ds = orc.voxelgrid(np.c_[P[hit], hit.astype(np.float32)].astype(np.float32), self.p.density)
        sel, pos = [], []
        for c in ds:
            dd = ((c[0] - P[:, 0]) ** 2 + (c[1] - P[:, 1]) ** 2) + (c[2] - P[:, 2]) ** 2
            sel.append(int(np.argmin(dd))); pos.append(c[:3])
        for i in range(len(P) - 1, -1, -1):
            if stamp - self.key_times[i] < 5.0:
                sel.append(i); pos.append(P[i])
            else:
                break
        out = []
        for i, c in zip(sel, pos):
            if np.sqrt(np.float32((c[0] - last[0]) ** 2 + (c[1] - last[1]) ** 2 + (c[2] - last[2]) ** 2)) > self.p.radius:
                continue
            out.append(i)
        return out

    def handle(self, stamp, odom6, scan_raw):
        """Returns dict(processed, pose, iters, n_sel, M, Q, ms) ; ms = reg_time of the frame (CPU wall clock)."""
        if stamp - self.t_last < self.p.time_interval:
            return {"processed": 0}
        t0 = time.perf_counter()
        rec = {"processed": 1}
        if not self.key_poses:
            self.pose[:3] = odom6[:3]
            ds = orc.voxelgrid(scan_raw, self.p.scan_leaf)
            rec.update(iters=0, n_sel=0, M=0, Q=len(ds))
        else:
            if self.last_odom is None:
                self.last_odom = np.array(odom6, np.float32)
            else:
                self.pose = orc.update_initial_guess(self.pose, self.last_odom, odom6)      # :865-883, float32 like the reference
                self.last_odom = np.array(odom6, np.float32)
            sel = self._select(stamp)
            pts = np.concatenate([self.key_clouds[i] for i in sel])
            offs = np.zeros(len(sel) + 1, np.int64)
            offs[1:] = np.cumsum([len(self.key_clouds[i]) for i in sel])
            poses = np.asarray([self.key_poses[i] for i in sel], np.float32)
            guess = self.pose.copy()
            pose, res, rc, M, Q, ms = orc.frame(pts, offs, poses, self.p.map_leaf, scan_raw, self.p.scan_leaf, guess, self.opts, self.state)
            self.pose = pose
            rec.update(iters=res.iters, n_sel=res.n_sel, M=M, Q=Q, guess=guess, sel=sel, phases_ms=ms)
            ds = None
        rec["ms"] = (time.perf_counter() - t0) * 1e3
        # saveFrame / saveLIOKeyFramesAndFactor
        save = True
        if self.key_poses:
            save = orc.save_frame(self.key_poses[-1], self.pose, self.p.add_angle, self.p.add_dist)          # :1245-1264
        if save:
            if ds is None:
                ds = orc.voxelgrid(scan_raw, self.p.scan_leaf)
            self.key_poses.append(self.pose.copy()); self.key_times.append(stamp); self.key_clouds.append(ds)
            rec["keyframe_id"] = len(self.key_poses) - 1
        rec["pose"] = self.pose.copy()
        self.t_last = stamp
        return rec
