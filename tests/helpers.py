"""Shared comparison code for the parity tests: CUDA path (through the C ABI) vs the numpy oracle."""
import numpy as np

from oracle import mmw_oracle as mo

# float64 state on both sides; differences come from summation order and FMA contraction only.  The spec's
# bar is 1e-4 relative on state; free-running sequences are held to 1e-6 (tiny differences are amplified by
# the filter over hundreds of frames), single stage calls to 1e-11 (test_gpu_parity.py).
STATE_RTOL = 1e-6
STATE_ATOL = 1e-9
KEYPOINT_ATOL = 1e-3        # 1 mm on joints (north_star)


def oracle_config(cfg) -> mo.OracleConfig:
    """OracleConfig with the same values as an mmw_config."""
    return mo.OracleConfig(
        s_height=cfg.s_height, s_tilt_deg=cfg.s_tilt_deg, z_max=cfg.z_max, frames_batch=cfg.frames_batch,
        db_z_weight=cfg.db_z_weight, db_range_weight=cfg.db_range_weight, db_eps=cfg.db_eps,
        db_min_samples=cfg.db_min_samples, tr_max_tracks=cfg.tr_max_tracks,
        tr_lifetime_dynamic=cfg.tr_lifetime_dynamic, tr_lifetime_static=cfg.tr_lifetime_static,
        tr_vel_thres=cfg.tr_vel_thres, tr_gate=cfg.tr_gate, kf_q_var=cfg.kf_q_var, kf_p_init=cfg.kf_p_init,
        kf_group_disp_init=cfg.kf_group_disp_init, kf_enable_est=bool(cfg.kf_enable_est), kf_a_n=cfg.kf_a_n,
        kf_est_pointnum=cfg.kf_est_pointnum, kf_spread_lim=tuple(cfg.kf_spread_lim), kf_a_spr=cfg.kf_a_spr,
        intensity_mu=cfg.intensity_mu, intensity_std=cfg.intensity_std)


def compare_frame(bt, recs, offsets, ctx, labels=None, check_keypoints=False, pose_rows=None):
    """bt: BatchedTracker after a step; recs[s]: oracle/golden record of scene s for the same frame."""
    tr, nt = bt.tracks()
    assoc = bt.point_assoc()
    ntr, nid, lastM = bt.summary()
    rc = bt.ring_counts()
    if labels is not None:
        lab, nfused = labels
    for s, r in enumerate(recs):
        c = "%s scene %d" % (ctx, s)
        assert lastM[s] == r["M"], c
        a = assoc[offsets[s]:offsets[s] + r["M"]] if r["M"] else np.zeros(0, np.int32)
        if r["M"]:
            np.testing.assert_array_equal(a, r["assoc"], err_msg=c + " assoc")
            assert np.all(assoc[offsets[s] + r["M"]:offsets[s + 1]] == -2), c
        assert nt[s] == len(r["tracks"]) == ntr[s], c + " n_tracks %d vs %d" % (nt[s], len(r["tracks"]))
        assert nid[s] == r["next_track_id"], c
        orc = np.asarray(r["ring_counts"])[-3:]
        # the reference's ring may still hold its initial empty frame; compare the non-empty suffix layout
        mine = rc[s][rc[s] >= 0]
        assert mine.sum() == orc.sum(), c + " ring total"
        np.testing.assert_array_equal(mine[mine > 0], orc[orc > 0], err_msg=c + " ring")
        if labels is not None:
            if r["labels"] is None:
                assert nfused[s] == -1, c + " dbscan should not have run"
            else:
                assert nfused[s] == len(r["labels"]), c + " fused count"
                np.testing.assert_array_equal(lab[s, :nfused[s]], r["labels"], err_msg=c + " labels")
        for k, t in enumerate(r["tracks"]):
            g = tr[s, k]
            ck = c + " track %d" % k
            assert g["id"] == t["id"], ck
            assert g["point_num"] == t["point_num"] and bool(g["is_static"]) == t["static"], ck
            np.testing.assert_array_equal(g["ring_counts"][:g["ring_frames"]],
                                          np.minimum(t["ring_counts"], 64), err_msg=ck)
            assert g["n_est"] == t["N_est"], ck
            np.testing.assert_allclose(g["lifetime"], t["lifetime"], rtol=1e-12, atol=1e-15, err_msg=ck)
            for name, key in (("x", "x"), ("P", "P"), ("spread_est", "spread_est"),
                              ("group_disp_est", "group_disp_est"), ("centroid", "centroid"),
                              ("min_vals", "min_vals"), ("max_vals", "max_vals")):
                np.testing.assert_allclose(g[name], np.asarray(t[key]).reshape(g[name].shape), rtol=STATE_RTOL,
                                           atol=STATE_ATOL, err_msg=ck + " " + name)
            if check_keypoints:
                np.testing.assert_allclose(g["keypoints"], t["keypoints"], rtol=0, atol=KEYPOINT_ATOL,
                                           err_msg=ck + " keypoints")
    if pose_rows is not None:
        si, ti, feats = pose_rows
        k = 0
        for s, r in enumerate(recs):
            if r.get("features") is None:
                continue
            f = np.asarray(r["features"])
            rows = np.nonzero(si == s)[0]
            assert len(rows) == len(f), "%s scene %d pose rows" % (ctx, s)
            np.testing.assert_array_equal(ti[rows], np.arange(len(f)))
            np.testing.assert_allclose(feats[rows], f.astype(np.float32), rtol=0, atol=2e-6,
                                       err_msg="%s scene %d features" % (ctx, s))
            k += len(f)
