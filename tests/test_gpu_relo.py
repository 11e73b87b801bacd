"""GPU tests of the relo-mode host path: prior session -> on-device submaps (generateSubMaps), /initialpose ->
NDT relocalisation (src/mapOptmization.cpp:261-305), per-frame scan matching against the nearest prior submap
(anchorOptimize.hpp:382-407) -- compared with the oracle on the same inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _guess_matrix(synth, pose6):
    """The init_guess bits the host class hands to align (float32 eulerToRotation, :276-280): the More-Thuente path is only
    reproducible from the very same starting matrix."""
    from liloc_b200 import host_api
    return host_api.euler_to_matrix(pose6)[:3].copy()


def _rot_err(synth, a6, b6):
    return np.abs(synth.rot_zyx(*[float(v) for v in a6[:3]]) - synth.rot_zyx(*[float(v) for v in b6[:3]])).max()


def test_submap_from_device_keyframes_equals_host_concatenation(gpu_ctx, orc, synth):
    world = synth.make_world(seed=9)
    traj = synth.trajectory_line(10, x0=-5)
    clouds = [synth.scan(world, traj[k], synth.VLP16, 900 + k) for k in range(10)]
    gpu_ctx.keyframe_clear(); gpu_ctx.ndt_clear()
    for k, c in enumerate(clouds):
        gpu_ctx.keyframe_put(k, c)
    poses = traj.astype(np.float32)
    nl, npts = gpu_ctx.ndt_target_put_keyframes(7, list(range(10)), poses, 1.0)
    sub = np.concatenate([orc.transform_cloud(clouds[k], poses[k], threads=8) for k in range(10)])
    assert npts == len(sub)
    means, icovs, counts, keys = gpu_ctx.ndt_target_get(7)
    nl2 = gpu_ctx.ndt_target_put(8, sub, 1.0)
    m2, i2, c2, k2 = gpu_ctx.ndt_target_get(8)
    assert nl == nl2 and np.array_equal(means, m2) and np.array_equal(icovs, i2) and np.array_equal(counts, c2) and np.array_equal(keys, k2)
    rm, ri, rc = orc.NdtTarget(sub, np.float32(1.0)).leaves()
    assert np.array_equal(means, rm) and np.array_equal(icovs, ri) and np.array_equal(counts, rc)


def test_relo_mode_through_host_class(orc, synth):
    from liloc_b200 import host_api
    # prior session: 30 keyframes along a line, un-down-sampled key clouds (what save_session stores), truth poses
    # (the map frame of a LiLoc session starts at the sensor: the floor is 1.8 m below z = 0, and /initialpose sets z = 0)
    world = synth.make_world(size_xy=(80.0, 40.0), n_pillars=16, seed=13)
    world[:, [2, 5]] -= 1.8
    n_prior = 30
    prior_traj = synth.trajectory_line(n_prior, step=1.0, x0=-15, height=0.0)
    prior_clouds = [synth.scan(world, prior_traj[k], synth.VLP16, 1300 + k) for k in range(n_prior)]
    keyposes7 = np.c_[prior_traj, 100.0 + np.arange(n_prior)]
    # current session: starts near prior keyframe 14, moves on
    # (10 Hz-like spacing: the reference loses the first odometry increment after the init, :865-871)
    cur = np.repeat(prior_traj[14:15], 6, axis=0)
    cur[:, 3] += 0.15 * np.arange(6); cur[:, 4] += 0.3 + 0.02 * np.arange(6); cur[:, 2] += 0.02
    scans = [synth.scan(world, cur[k], synth.VLP16, 7700 + k) for k in range(len(cur))]
    with host_api.MapOptimization("lio_sam_default.yaml", "relo", device=0) as m:
        res = np.float32(m.param("ndtResolution")); eps = np.float32(m.param("ndtEpsilon"))
        n_sub = m.load_prior_session(keyposes7, prior_clouds)
        assert n_sub == 3
        info0 = m.submap_info(1)
        assert (info0["first"], info0["count"]) == (10, 10) and info0["points"] == sum(len(c) for c in prior_clouds[10:20])
        assert np.allclose(info0["centroid"], prior_traj[10:20, 3:].astype(np.float32).mean(0), atol=1e-4)
        odom = cur.astype(np.float32)
        stamps = 500.0 + 0.25 * np.arange(len(cur))
        infos = host_api.make_infos(scans, odom, stamps)
        # before /initialpose the frame is dropped (:262-265)
        rep = m.handle(infos[0])
        assert rep.processed == 0 and rep.relo_initialized == 0
        # /initialpose: x, y, yaw only, half a metre and 0.05 rad off (:755-769)
        init = np.array([0, 0, cur[0][2] + 0.05, cur[0][3] + 0.4, cur[0][4] - 0.3, 0], np.float32)
        m.initialpose(init[3], init[4], init[2])
        stamps2 = stamps + 10.0
        infos = host_api.make_infos(scans, odom, stamps2)
        rep = m.handle(infos[0])
        assert rep.processed == 1 and rep.status == 0 and rep.relo_initialized == 1 and rep.relo_submap == 1
        # oracle: nearest submap by centroid, 2 x align from the eulerToRotation guess
        sub = np.concatenate([orc.transform_cloud(prior_clouds[k], prior_traj[k].astype(np.float32), threads=8) for k in range(10, 20)])
        tg = orc.NdtTarget(sub, res)
        o = orc.NdtOpts(resolution=float(res), trans_eps=float(eps), search=0, threads=8)
        T, r1 = orc.ndt_align(tg, scans[0], _guess_matrix(synth, init), o)
        T, r2 = orc.ndt_align(tg, scans[0], T, o)
        got = np.array(rep.relo_pose, np.float32)
        assert rep.relo_iterations == r1.iterations + r2.iterations
        assert np.abs(got[3:] - T[:, 3]).max() <= 1e-4
        assert np.abs(synth.rot_zyx(*[float(v) for v in got[:3]]) - T[:, :3]).max() <= 1e-5
        assert np.abs(got[3:] - cur[0][3:]).max() < 0.05 and abs(got[2] - cur[0][2]) < 5e-3      # relocalised onto the truth
        assert rep.keyframe_id == 0                                                              # every relo frame is a keyframe
        # following frames: scan2map against the current session's keyframes + scan matching against the prior submap
        for f in range(1, len(cur)):
            rep = m.handle(infos[f])
            assert rep.processed == 1 and rep.status == 0 and rep.keyframe_id == f and rep.iters > 0
            pose = np.array(rep.pose, np.float32)
            sid = rep.relo_submap
            assert sid == int(np.argmin(((np.array([m.submap_info(s)["centroid"] for s in range(n_sub)]) - pose[3:]) ** 2).sum(1)))
            lo = m.submap_info(sid)["first"]
            sub = np.concatenate([orc.transform_cloud(prior_clouds[k], prior_traj[k].astype(np.float32), threads=8) for k in range(lo, lo + 10)])
            Tm, rr = orc.ndt_align(orc.NdtTarget(sub, res), scans[f], _guess_matrix(synth, pose), o)
            got = np.array(rep.relo_pose, np.float32)
            # one canonical summation order on both sides (ndt.cuh ndt_sweep_tile <-> oracle derivative_sweep): the
            # More-Thuente iteration takes the same path, so the north_star tolerance holds for every frame
            assert rep.relo_converged == rr.converged and rep.relo_iterations == rr.iterations
            assert np.abs(got[3:] - Tm[:, 3]).max() <= 1e-4
            assert np.abs(synth.rot_zyx(*[float(v) for v in got[:3]]) - Tm[:, :3]).max() <= 1e-5
            assert np.abs(got[3:5] - cur[f][3:5]).max() < 0.2       # NDT at 1 m resolution with eps 0.01: decimetre-level
            # prior vertices within 5 m (at most 3, nearest first; over the WHOLE prior trajectory -- every entry of the
            # reference's subMapVertexCloudVec_ is the same never-cleared cloud, dataLoader.hpp:311-325) and their ICP fitness
            # against the matched scan
            d2 = ((prior_traj[:, 3:].astype(np.float32) - pose[3:]) ** 2).sum(1)
            near = [int(k) for k in np.argsort(d2, kind="stable") if d2[k] < 25.0][:3]
            assert rep.relo_n_vertices == len(near) and list(rep.relo_vertex)[: len(near)] == near
            for v, vid in enumerate(near):
                want = orc.icp_fitness(prior_clouds[vid], prior_traj[vid].astype(np.float32), scans[f], got, threads=8)
                assert rep.relo_fitness[v] == np.float32(min(want, 2.0)), (f, v, rep.relo_fitness[v], want)


def test_icp_fitness_bit_exact(gpu_ctx, orc, synth):
    """liloc_icp_fitness (getICPFitnessScore, anchorOptimize.hpp:465-481): unbounded exact 1-NN on the search grid, squared
    distances summed sequentially in float -> the same bits as the oracle."""
    rng = np.random.default_rng(41)
    world = synth.make_world(seed=41)
    traj = synth.trajectory_line(4, x0=-2)
    a = synth.scan(world, traj[0], synth.VLP16, 4100)
    b = synth.scan(world, traj[2], synth.VLP16, 4102)
    cases = [(a, traj[0], b, traj[2]),                                     # two overlapping scans, 2 m apart
             (a[::7], traj[0], b, traj[2] + np.array([0, 0, 0.3, 5.0, -3.0, 0.5])),      # misaligned: larger distances, more rings
             (a[:500], traj[0] + np.array([0, 0, 0, 200.0, 0, 0]), b[::3], traj[2]),   # no overlap at all: far outside the grid
             ((rng.normal(size=(300, 4)) * 30).astype(np.float32), np.zeros(6), (rng.normal(size=(7, 4)) * 30).astype(np.float32), np.zeros(6))]
    for c1, p1, c2, p2 in cases:
        got = gpu_ctx.icp_fitness(c1, p1, c2, p2)
        ref = orc.icp_fitness(c1, np.asarray(p1, np.float32), c2, np.asarray(p2, np.float32), kdtree=True, threads=8)
        assert got == ref, (got, ref)
    assert gpu_ctx.icp_fitness(a, traj[0], a, traj[0]) == 0.0
    # resident form: cloud 1 from the keyframe store, cloud 2 from an NDT source
    gpu_ctx.keyframe_clear(); gpu_ctx.ndt_clear()
    gpu_ctx.keyframe_put(5, a)
    gpu_ctx.ndt_source_put(0, b)
    got = gpu_ctx.icp_fitness_resident(5, traj[0], 0, traj[2])
    assert got == orc.icp_fitness(a, traj[0].astype(np.float32), b, traj[2].astype(np.float32), threads=8)


def test_session_round_trip_lio_save_then_relocalise(tmp_path, orc, synth):
    """A lio session saved in the reference's on-disk format (g2o + PCDs/<k>.pcd, src/mapOptmization.cpp:420-635) is loaded
    as the prior session of a relo run (dataLoader.hpp:207-348) and the first scan of a new session relocalises on it."""
    from liloc_b200 import host_api
    seq = synth.make_sequence(synth.VLP16, 40, seed=6, world_size=(60.0, 40.0), n_pillars=14, loop_radius=(20.0, 11.0), dt=0.21)
    sdir = str(tmp_path / "session")
    with host_api.MapOptimization("lio_sam_default.yaml", "lio", device=0) as m:
        reps = m.run(host_api.make_infos(seq.scans, seq.odom, seq.stamps))
        n_kf = m.num_keyframes()
        kp = m.keyposes()
        m.save_session(sdir)
        kf_ids = [r.keyframe_id for r in reps if r.keyframe_id >= 0]
        frames_of_kf = [f for f, r in enumerate(reps) if r.keyframe_id >= 0]
    assert n_kf > 10 and sorted(os.listdir(os.path.join(sdir, "PCDs")), key=lambda s: int(s[:-4])) == [f"{k}.pcd" for k in range(n_kf)]
    # the files hold exactly what the session held
    back = host_api.load_g2o(os.path.join(sdir, "singlesession_posegraph.g2o"))
    dang = (back[:, :3] - kp[:, :3] + np.pi) % (2 * np.pi) - np.pi                  # yaw may come back as -pi instead of +pi
    assert len(back) == n_kf and np.abs(back[:, 3:6] - kp[:, 3:6]).max() < 5e-6 and np.abs(dang).max() < 5e-6
    k = kf_ids[3]
    assert np.array_equal(host_api.load_pcd(os.path.join(sdir, "PCDs", f"{k}.pcd")), orc.voxelgrid(seq.scans[frames_of_kf[3]], 0.2))
    assert len(host_api.load_pcd(os.path.join(sdir, "globalMap.pcd"))) > 1000
    # transformations.pcd: the key poses as PCL writes PointXYZIRPYT (48-byte records behind the header)
    raw = open(os.path.join(sdir, "transformations.pcd"), "rb").read()
    body = raw[raw.index(b"DATA binary\n") + len(b"DATA binary\n"):]
    assert len(body) == 48 * n_kf
    rec = np.frombuffer(body, np.dtype([("xyz", "<f4", 3), ("one", "<f4"), ("irpy", "<f4", 4), ("time", "<f8"), ("pad", "V8")]))
    assert np.array_equal(rec["xyz"], kp[:, 3:6].astype(np.float32)) and np.array_equal(rec["irpy"][:, 1:], kp[:, :3].astype(np.float32))
    assert np.array_equal(rec["irpy"][:, 0], np.arange(n_kf, dtype=np.float32)) and np.all(rec["one"] == 1.0)
    assert len(host_api.load_pcd(os.path.join(sdir, "trajectory.pcd"))) == n_kf
    # the same session at req.resolution = 0.5: every key cloud and the all-keyframe global map through VoxelGrid(0.5)
    sdir5 = str(tmp_path / "session_05")
    with host_api.MapOptimization("lio_sam_default.yaml", "lio", device=0) as m:
        m.run(host_api.make_infos(seq.scans, seq.odom, seq.stamps))
        m.save_session(sdir5, resolution=0.5)
    full = [host_api.load_pcd(os.path.join(sdir, "PCDs", f"{i}.pcd")) for i in range(n_kf)]
    for i in (0, 3, n_kf - 1):
        assert np.array_equal(host_api.load_pcd(os.path.join(sdir5, "PCDs", f"{i}.pcd")), orc.voxelgrid(full[i], 0.5))
    world_cloud = np.concatenate([orc.transform_cloud(full[i], kp[i, :6].astype(np.float32), threads=8) for i in range(n_kf)])
    assert np.array_equal(host_api.load_pcd(os.path.join(sdir5, "globalMap.pcd")), orc.voxelgrid(world_cloud, 0.5))
    # new session in relo mode against the saved one: scan of frame 12, /initialpose 0.4 m / 0.04 rad off its lio pose
    f = 12
    truth = np.array(reps[f].pose, np.float32)
    with host_api.MapOptimization("lio_sam_default.yaml", "relo", device=0) as r:
        assert r.load_prior_session_dir(sdir) == (n_kf + 9) // 10
        r.initialpose(truth[3] + 0.3, truth[4] - 0.25, truth[2] + 0.04)
        info = host_api.make_infos([seq.scans[f]], seq.odom[f:f + 1], np.array([900.0]))
        rep = r.handle(info[0])
        assert rep.processed == 1 and rep.relo_initialized == 1 and rep.relo_converged == 1
        got = np.array(rep.relo_pose, np.float32)
        assert np.abs(got[3:5] - truth[3:5]).max() < 0.1 and abs(got[2] - truth[2]) < 0.01


import os  # noqa: E402


class _PriorModel:
    """Reference model of the cloud side of AnchorOptimization::addScanMatchingFactor / addNewPrior and of
    Session::searchNearestSubMapAndVertex (include/egoOptimization/anchorOptimize.hpp:272-421, include/dataManager/
    dataLoader.hpp:350-390), written from those lines: which prior vertices a frame finds, when the frame joins the prior
    session instead of being matched, and when a new submap is cut."""

    def __init__(self, orc, prior_traj32, add_angle, add_dist, submap_size=10):
        self.orc, self.add_angle, self.add_dist, self.submap_size = orc, add_angle, add_dist, submap_size
        self.poses = [p.copy() for p in prior_traj32]          # KeyPoses6D_ as [roll, pitch, yaw, x, y, z]
        self.prior_size = len(self.poses)
        n_sub = (self.prior_size + submap_size - 1) // submap_size
        self.vertexes = [None] * n_sub                         # None: the loader's shared vertex cloud (all loaded key poses)
        self.first_add, self.last_pose, self.end, self.count, self.added = True, None, self.prior_size, 0, []

    def search_vertexes(self, submap, pose):
        ids = range(self.prior_size) if self.vertexes[submap] is None else self.vertexes[submap]
        q = pose[3:].astype(np.float32)
        hits = []
        for i in ids:
            d = self.poses[i][3:].astype(np.float32) - q
            d2 = np.float32(np.float32(d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])
            if d2 < np.float32(25.0):
                hits.append((d2, i))
        hits.sort()
        cur_new, out = len(self.poses), []
        for _, i in hits:
            if i > self.prior_size and abs(cur_new - i) <= 10:
                continue
            if len(out) == 3:
                break
            out.append(i)
        return out

    def add_new_prior(self, pose):
        """Returns (added id or -1, ids of a newly cut submap or None)."""
        if not self.first_add and not self.orc.save_frame(self.last_pose, pose, self.add_angle, self.add_dist):
            return -1, None
        was_first = self.first_add
        i = len(self.poses)
        self.poses.append(pose.copy()); self.added.append(i); self.count += 1
        self.first_add, self.last_pose = False, pose.copy()
        if not was_first and self.count % self.submap_size == 0:
            ids = list(range(self.end, len(self.poses)))
            self.vertexes.append(ids)
            self.end, self.count = len(self.poses), 0
            return i, ids
        return i, None


def test_prior_map_grows_where_the_session_leaves_it(tmp_path, orc, synth):
    """addScanMatchingFactor with fewer than two prior vertices in reach calls addNewPrior (anchorOptimize.hpp:387-390): the
    session drives out of the prior map, its frames (pose + FULL scan) join the prior session, after ten additions past the
    first a new, voxel-filtered submap is cut on the device, and the frames after that are matched against it.  Checked
    frame by frame against a model of those reference lines and against the oracle (submap grid bit for bit, NDT poses)."""
    from liloc_b200 import abi, host_api
    world = synth.make_world(size_xy=(90.0, 70.0), n_pillars=24, seed=21)
    world[:, [2, 5]] -= 1.8
    n_prior = 20
    prior_traj = synth.trajectory_line(n_prior, step=1.0, x0=-10, height=0.0)
    prior_clouds = [synth.scan(world, prior_traj[k], synth.VLP16, 2100 + k) for k in range(n_prior)]
    keyposes7 = np.c_[prior_traj, 50.0 + np.arange(n_prior)]
    # the session starts at prior keyframe 9, drives off sideways (+y) 1.25 m per frame out of the prior map, turns round and
    # comes back next to its own track: by then the vertices it added on the way out are more than ten additions old (the
    # reference skips younger ones, dataLoader.hpp:376-380) and the new submap is what it gets matched against
    n_out, n_back = 17, 10
    cur = np.repeat(prior_traj[9:10], 3 + n_out + n_back, axis=0)
    cur[3:3 + n_out, 4] += 1.25 * (1 + np.arange(n_out))
    cur[3 + n_out:, 4] += 1.25 * (n_out - 1 - np.arange(n_back))
    cur[3 + n_out:, 3] += 0.4
    cur[:, 2] += 0.01
    scans = [synth.scan(world, cur[k], synth.VLP16, 8800 + k) for k in range(len(cur))]
    with host_api.MapOptimization("lio_sam_default.yaml", "relo", device=0) as m:
        res, eps = np.float32(m.param("ndtResolution")), np.float32(m.param("ndtEpsilon"))
        model = _PriorModel(orc, prior_traj.astype(np.float32), np.float32(m.param("surroundingkeyframeAddingAngleThreshold")),
                            np.float32(m.param("surroundingkeyframeAddingDistThreshold")))
        assert m.load_prior_session(keyposes7, prior_clouds) == 2
        ctx = abi.Context.borrow(m.ctx_handle(), 0)
        m.initialpose(cur[0][3] + 0.2, cur[0][4] - 0.1, cur[0][2] + 0.02)
        infos = host_api.make_infos(scans, cur.astype(np.float32), 700.0 + 0.25 * np.arange(len(cur)))
        o = orc.NdtOpts(resolution=float(res), trans_eps=float(eps), search=0, threads=8)
        added, new_submaps, matched_new = [], [], 0
        key_clouds = {i: prior_clouds[i] for i in range(n_prior)}
        for f in range(len(cur)):
            rep = m.handle(infos[f])
            assert rep.processed == 1 and rep.status == 0, (f, rep.status)
            if f == 0:
                assert rep.relo_initialized == 1
                continue
            pose = np.array(rep.pose, np.float32)
            n_sub = len(model.vertexes)
            cents = np.array([m.submap_info(s)["centroid"] for s in range(n_sub)])
            sid = int(np.argmin(((cents - pose[3:]) ** 2).sum(1)))
            want = model.search_vertexes(sid, pose)
            assert rep.relo_n_vertices == len(want) and list(rep.relo_vertex)[: len(want)] == want, (f, want, list(rep.relo_vertex))
            if len(want) <= 1:
                aid, cut = model.add_new_prior(pose)
                assert rep.prior_added == aid, (f, rep.prior_added, aid)
                if aid >= 0:
                    added.append(aid); key_clouds[aid] = scans[f]
                    assert ctx.keyframe_size(host_api.PRIOR_KEYFRAME_BASE + aid) == len(scans[f])
                if cut is not None:
                    new_sid = len(model.vertexes) - 1
                    assert rep.prior_new_submap == new_sid
                    new_submaps.append(new_sid)
                    # the new submap: transform + concatenate + VoxelGrid(0.2) (:341-367), then the NDT grid -- bit for bit
                    sub = orc.voxelgrid(np.concatenate([orc.transform_cloud(key_clouds[i], model.poses[i], threads=8) for i in cut]), 0.2)
                    info = m.submap_info(new_sid)
                    assert (info["first"], info["count"], info["points"]) == (cut[0], len(cut), len(sub))
                    acc = np.zeros(3, np.float32)                                        # centroid: float sums / submap_size (:355-358)
                    for i in cut:
                        acc = (acc + model.poses[i][3:].astype(np.float32)).astype(np.float32)
                    assert np.array_equal(info["centroid"], (acc / np.float32(10)).astype(np.float32))
                    means, icovs, counts, keys = ctx.ndt_target_get(new_sid)
                    rm, ri, rc = orc.NdtTarget(sub, res).leaves()
                    assert np.array_equal(means, rm) and np.array_equal(icovs, ri) and np.array_equal(counts, rc)
                else:
                    assert rep.prior_new_submap == -1
            else:
                model.first_add = True
                assert rep.prior_added == -1 and rep.relo_submap == sid
                ids = range(10 * sid, min(10 * sid + 10, n_prior)) if model.vertexes[sid] is None else model.vertexes[sid]
                sub = np.concatenate([orc.transform_cloud(key_clouds[i], model.poses[i], threads=8) for i in ids])
                if model.vertexes[sid] is not None:
                    sub = orc.voxelgrid(sub, 0.2); matched_new += 1
                Tm, rr = orc.ndt_align(orc.NdtTarget(sub, res), scans[f], _guess_matrix(synth, pose), o)
                got = np.array(rep.relo_pose, np.float32)
                assert rep.relo_iterations == rr.iterations and rep.relo_converged == rr.converged
                assert np.abs(got[3:] - Tm[:, 3]).max() <= 1e-4 and np.abs(synth.rot_zyx(*[float(v) for v in got[:3]]) - Tm[:, :3]).max() <= 1e-5
                for v, vid in enumerate(want):                       # vertices this session added carry no fitness score (:431-434)
                    if vid in model.added:
                        assert rep.relo_fitness[v] == -1.0
                    else:
                        wf = orc.icp_fitness(key_clouds[vid], model.poses[vid], scans[f], got, threads=8)
                        assert rep.relo_fitness[v] == np.float32(min(wf, 2.0))
        assert len(added) >= 11 and len(new_submaps) >= 1 and matched_new >= 2, (added, new_submaps, matched_new)
        # save_session (:543-635) writes the PRIOR session with what this one joined to it
        sdir = str(tmp_path / "grown")
        m.save_session(sdir)
        n_all = n_prior + len(added)
        back = host_api.load_g2o(os.path.join(sdir, "singlesession_posegraph.g2o"))
        assert len(back) == n_all == len(os.listdir(os.path.join(sdir, "PCDs")))
        want_p = np.array([model.poses[i] for i in range(n_all)], np.float32)
        dang = (back[:, :3] - want_p[:, :3] + np.pi) % (2 * np.pi) - np.pi
        assert np.abs(back[:, 3:6] - want_p[:, 3:]).max() < 5e-6 and np.abs(dang).max() < 5e-6
        for i in (0, n_prior - 1, added[0], added[-1]):
            assert np.array_equal(host_api.load_pcd(os.path.join(sdir, "PCDs", f"{i}.pcd")), key_clouds[i])
        leaf = np.float32(m.param("mappingSurfLeafSize"))
        allc = np.concatenate([orc.transform_cloud(key_clouds[i], model.poses[i], threads=8) for i in range(n_all)])
        assert np.array_equal(host_api.load_pcd(os.path.join(sdir, "globalMap.pcd")), orc.voxelgrid(allc, leaf))


def _hypothesis_scene(synth):
    world = synth.make_world(size_xy=(80.0, 40.0), n_pillars=16, seed=13)
    world[:, [2, 5]] -= 1.8
    n_prior = 20
    prior_traj = synth.trajectory_line(n_prior, step=1.0, x0=-10, height=0.0)
    prior_clouds = [synth.scan(world, prior_traj[k], synth.VLP16, 1300 + k) for k in range(n_prior)]
    truth = prior_traj[7].copy(); truth[3] += 0.3; truth[4] += 0.4; truth[2] += 0.05
    scan = synth.scan(world, truth, synth.VLP16, 7800)
    return prior_traj, prior_clouds, truth, scan


def test_multi_hypothesis_relocalisation_init(orc, synth):
    """The /initialpose is 1.3 m and a quarter turn off: the reference's single init (2 x align from that pose) does not find
    the pose; 8 yaw x 6 position hypotheses, each through the same 2 x align (src/mapOptmization.cpp:267-290), sharded inside
    the C ABI, do.  The winner and its pose are the oracle's for the same hypothesis list."""
    from liloc_b200 import host_api
    prior_traj, prior_clouds, truth, scan = _hypothesis_scene(synth)
    keyposes7 = np.c_[prior_traj, 100.0 + np.arange(len(prior_traj))]
    n_yaw, n_xy, radius = 8, 6, 2.0
    init = np.array([0, 0, truth[2] + np.pi / 2 + 0.1, truth[3] + 1.0, truth[4] - 0.8, 0], np.float32)
    odom = truth.astype(np.float32)[None]
    with host_api.MapOptimization("lio_sam_default.yaml", "relo", device=0) as m:
        res, eps = np.float32(m.param("ndtResolution")), np.float32(m.param("ndtEpsilon"))
        assert m.load_prior_session(keyposes7, prior_clouds) == 2
        # the reference's init alone
        m.initialpose(init[3], init[4], init[2])
        rep1 = m.handle(host_api.make_infos([scan], odom, np.array([500.0]))[0])
        assert rep1.relo_initialized == 1
        single = np.array(rep1.relo_pose, np.float32)
        assert np.abs(single[3:5] - truth[3:5]).max() > 0.5 or abs(single[2] - truth[2]) > 0.2       # did not find it
        # with hypotheses
        m.reset()
        assert m.load_prior_session(keyposes7, prior_clouds) == 2
        m.set_relo_hypotheses(n_yaw, n_xy, radius)
        m.initialpose(init[3], init[4], init[2])
        rep = m.handle(host_api.make_infos([scan], odom, np.array([600.0]))[0])
        assert rep.processed == 1 and rep.relo_initialized == 1
        got = np.array(rep.relo_pose, np.float32)
        assert np.abs(got[3:5] - truth[3:5]).max() < 0.1 and abs(got[2] - truth[2]) < 0.02, (got, truth)
        cents = np.array([m.submap_info(s)["centroid"] for s in range(2)])
    # oracle over the same hypothesis list
    o = orc.NdtOpts(resolution=float(res), trans_eps=float(eps), search=0, threads=8)
    subs = [orc.NdtTarget(np.concatenate([orc.transform_cloud(prior_clouds[k], prior_traj[k].astype(np.float32), threads=8) for k in range(10 * s, 10 * s + 10)]), res) for s in range(2)]
    best, best_score, best_T = -1, -1e30, None
    for h in range(n_yaw * n_xy):
        g6 = host_api.hypothesis_pose(init, n_yaw, n_xy, radius, h)
        sid = int(np.argmin(((cents - g6[3:]) ** 2).sum(1)))
        T = host_api.euler_to_matrix(g6)[:3].copy()
        for _ in range(2):
            T, rr = orc.ndt_align(subs[sid], scan, T, o)
        sc = np.float32(rr.score / len(scan))
        if sc > best_score:
            best, best_score, best_T = h, sc, T
    assert rep.relo_hypothesis == best and abs(rep.relo_score - best_score) <= 1e-6 * abs(best_score)
    assert np.abs(got[3:] - best_T[:, 3]).max() <= 1e-4 and np.abs(synth.rot_zyx(*[float(v) for v in got[:3]]) - best_T[:, :3]).max() <= 1e-5
