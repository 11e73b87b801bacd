"""GPU test of the drop-in host class: a lio sequence pushed through mapOptimization::laserCloudInfoHandler
(liloc_b200/host, C ABI underneath) and replayed frame by frame through the CPU oracle with the same
keyframe selections, key poses and initial guesses."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_lio_sequence_through_host_class(orc, synth):
    from liloc_b200 import host_api
    seq = synth.make_sequence(synth.VLP16, 45, seed=4, world_size=(60.0, 40.0), n_pillars=14, loop_radius=(20.0, 11.0), dt=0.21)
    with host_api.MapOptimization("lio_sam_default.yaml", "lio", device=0) as m:
        scan_leaf, map_leaf = np.float32(m.param("mappingSurfLeafSize")), np.float32(m.param("surroundingKeyframeMapLeafSize"))
        infos = host_api.make_infos(seq.scans, seq.odom, seq.stamps)
        kf_clouds, n_checked, n_kf = {}, 0, 0
        for f in range(len(seq.scans)):
            rep = m.handle(infos[f])
            assert rep.processed == 1 and rep.status == 0
            sel = m.last_selection()
            keyposes = m.keyposes()
            ds = orc.voxelgrid(seq.scans[f], scan_leaf)
            if f > 0:
                assert rep.n_selected_keyframes == len(sel) > 0
                poses = keyposes[sel][:, :6].astype(np.float32)
                mp = orc.localmap_build([kf_clouds[i] for i in sel], poses, map_leaf, threads=8)
                guess = np.array(rep.guess, np.float32)
                rpose, rres, rrc, _ = orc.scan2map(mp, ds, guess, orc.S2mOpts(knn=1, threads=8))
                assert rrc == 0
                assert (rep.q_all, rep.map_points) == (len(ds), len(mp))
                assert (rep.iters, rep.n_sel, rep.converged) == (rres.iters, rres.n_sel, rres.converged)
                pose = np.array(rep.pose, np.float32)
                assert np.abs(pose[3:] - rpose[3:]).max() <= 1e-4 and np.abs(pose[:3] - rpose[:3]).max() <= 1e-5
                # the lio map frame starts at the first sensor position (:835-842), so compare relative to frame 0
                d = np.abs(pose[3:] - (seq.truth[f][3:] - seq.truth[0][3:]))
                assert d[:2].max() < 0.1 and d[2] < 0.6      # z drifts with the noisy first-frame IMU attitude
                n_checked += 1
            if rep.keyframe_id >= 0:
                assert rep.keyframe_id == n_kf
                kf_clouds[rep.keyframe_id] = ds
                n_kf += 1
        assert n_checked == len(seq.scans) - 1 and 5 < n_kf <= len(seq.scans)
        assert m.num_keyframes() == n_kf
        # publishGlobalMap (:698-753): every key pose is within the 1000 m radius; thinning at 10 m keeps a subset
        gm = m.global_map()
        kp = m.keyposes()
        pos = np.c_[kp[:, 3:6], np.arange(len(kp))].astype(np.float32)
        last = pos[-1]
        order = np.lexsort((np.arange(len(pos)), ((last[:3] - pos[:, :3]) ** 2).sum(1)))
        ds = orc.voxelgrid(pos[order], np.float32(m.param("globalMapVisualizationPoseDensity")))
        ids = [int(np.argmin(((c[:3] - pos[:, :3]) ** 2).sum(1))) for c in ds]
        want = orc.localmap_build([kf_clouds[i] for i in ids], kp[ids][:, :6].astype(np.float32), np.float32(m.param("globalMapVisualizationLeafSize")), threads=8)
        assert len(ids) >= 2 and gm.shape == want.shape and np.array_equal(gm, want)


def test_time_interval_gate_drops_frames(synth):
    from liloc_b200 import host_api
    seq = synth.make_sequence(synth.VLP16, 6, seed=5, world_size=(60.0, 40.0), n_pillars=6, loop_radius=(20.0, 11.0), dt=0.11)
    with host_api.MapOptimization("lio_sam_default.yaml", "lio", device=0) as m:     # timeInterval 0.2
        reps = m.run(host_api.make_infos(seq.scans, seq.odom, seq.stamps))
        assert [r.processed for r in reps] == [1, 0, 1, 0, 1, 0]
