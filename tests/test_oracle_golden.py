"""The numpy oracle against (a) the known-answer values recomputed from the reference's
own functions and (b) the golden traces recorded from the reference's own Tracking.py /
Utils.py (oracle/gen_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases
from oracle import mmw_oracle as mo, trace_io

KA = json.load(open(os.path.join(GOLDEN, "known_answers.json")))
FLOAT_TOL = 1e-9     # oracle vs reference, both float64: summation-order noise only


def test_constants_match_reference():
    c, cfg = KA["constants"], mo.OracleConfig()
    assert (c["S_HEIGHT"], c["S_TILT"]) == (cfg.s_height, cfg.s_tilt_deg)
    assert c["FB_FRAMES_BATCH"] == cfg.frames_batch
    assert (c["DB_Z_WEIGHT"], c["DB_RANGE_WEIGHT"], c["DB_EPS"], c["DB_MIN_SAMPLES_MIN"]) == \
        (cfg.db_z_weight, cfg.db_range_weight, cfg.db_eps, cfg.db_min_samples)
    assert (c["TR_MAX_TRACKS"], c["TR_LIFETIME_DYNAMIC"], c["TR_LIFETIME_STATIC"], c["TR_VEL_THRES"], c["TR_GATE"]) == \
        (cfg.tr_max_tracks, cfg.tr_lifetime_dynamic, cfg.tr_lifetime_static, cfg.tr_vel_thres, cfg.tr_gate)
    assert (c["KF_Q_STD"], c["KF_P_INIT"], c["KF_GROUP_DISP_EST_INIT"], c["KF_ENABLE_EST"], c["KF_A_N"],
            c["KF_EST_POINTNUM"], c["KF_A_SPR"]) == \
        (cfg.kf_q_var, cfg.kf_p_init, cfg.kf_group_disp_init, cfg.kf_enable_est, cfg.kf_a_n, cfg.kf_est_pointnum,
         cfg.kf_a_spr)
    assert list(c["KF_SPREAD_LIM"]) == list(cfg.kf_spread_lim)
    assert (c["INTENSITY_MU"], c["INTENSITY_STD"]) == (cfg.intensity_mu, cfg.intensity_std)
    np.testing.assert_array_equal(KA["default_posture"], mo.MODEL_DEFAULT_POSTURE)


def test_known_answers():
    d = mo.pair_distance_matrix(np.array([[0.1, 2.0, 1.0], [0.4, 2.5, 0.2]]))
    assert d[0, 1] == KA["altered_dist"] == d[1, 0]
    world, keep = mo.normalize_points(np.array([[.5, 3, -.4, .62, 120], [1, -1, 0, .1, 5], [0, 0, 0, .3, 7]]))
    assert keep.tolist() == [True, False, False]
    np.testing.assert_allclose(world, np.array(KA["normalize"]), rtol=0, atol=1e-15)
    np.testing.assert_array_equal(mo.kf_Q(1.0), np.array(KA["Q_dt1"]))
    np.testing.assert_array_equal(mo.kf_F(0.25), np.array(KA["F_dt0.25"]))
    P = (mo.kf_F(0.1) @ (np.eye(9) * 0.1)) @ mo.kf_F(0.1).T + mo.kf_Q(0.1)
    np.testing.assert_allclose(np.diag(P), KA["predict_diagP_dt0.1"], rtol=1e-15)


def test_projection_points_match_reference():
    """Row f-2: oracle restatement of Utils.calc_projection_points against the reference's own outputs."""
    pp = KA["projection_points"]
    for q, want in zip(pp["inputs"], pp["outputs"]):
        assert list(mo.projection_point(*q)) == want
    wc = KA["window_constants"]
    from mmwave_msc_b200 import constants as c
    assert (c.M_X, c.M_Y, c.M_Z) == (wc["M_X"], wc["M_Y"], wc["M_Z"])
    assert (c.V_SCREEN_FADE_SIZE_MAX, c.V_SCREEN_FADE_SIZE_MIN, c.V_SCREEN_FADE_WEIGHT) == \
        (wc["V_SCREEN_FADE_SIZE_MAX"], wc["V_SCREEN_FADE_SIZE_MIN"], wc["V_SCREEN_FADE_WEIGHT"])
    # calc_fade_square (Visualizer.py:14-29) on the default posture: clamp active on both sides
    kp = np.asarray(KA["default_posture"])
    (cx, cz), size = mo.fade_square(np.array([0.5, 0.2, 1.0]), kp)
    assert size == 0.3 - (0.2 + kp[12]) * 0.08 or size in (0.2, 0.3)
    assert mo.fade_square(np.array([0.5, 9.0, 1.0]), kp)[1] == 0.2
    assert mo.fade_square(np.array([0.5, -9.0, 1.0]), kp)[1] == 0.3


@pytest.mark.parametrize("name", ["c1_s4", "churn_s308"])
def test_track0_export_matches_reference(name):
    """Row f-3: oracle.track0_export against the reference's own relative_coordinates + format_batched_frames run
    under the loop body of preprocessing.py:185-216 (tests/golden/track0_export.npz, oracle/gen_golden.py)."""
    ex = np.load(os.path.join(GOLDEN, "track0_export.npz"))
    g = trace_io.unpack(np.load(os.path.join(GOLDEN, name + ".npz")))
    nf = int(ex[name + "_frames"])
    so = mo.SceneOracle()
    k = 0
    for f in range(nf):
        rec = so.step(g["frames"][f], g["dts"][f])
        got = mo.track0_export(so) if rec["ran"] else None
        assert (got is not None) == bool(ex[name + "_valid"][f]), "frame %d" % f
        if got is not None:
            np.testing.assert_allclose(got[0], ex[name + "_rows"][k], rtol=0, atol=FLOAT_TOL)
            np.testing.assert_allclose(got[1], ex[name + "_centroid"][k], rtol=0, atol=FLOAT_TOL)
            # everything but the two centroid-relative columns is copied, not computed: exact
            np.testing.assert_array_equal(got[0][:, 2:], ex[name + "_rows"][k][:, 2:])
            k += 1
    assert k == len(ex[name + "_rows"])
    assert not ex[name + "_valid"].all() or name == "c1_s4"      # the churn case covers invalid frames


def _run_oracle(g, max_tracks):
    so = mo.SceneOracle(mo.OracleConfig(tr_max_tracks=max_tracks))
    out = []
    for fr, dt, gone in zip(trace_io.reference_frames(g), g["dts"], g["missing"]):
        if gone:
            so.ring.pop()                                # preprocessing.py:262-264
        out.append(so.step(fr, dt))
    return out


@pytest.mark.parametrize("case", golden_cases())
def test_oracle_reproduces_reference_trace(case):
    d = np.load(os.path.join(GOLDEN, case + ".npz"))
    g = trace_io.unpack(d)
    ora = _run_oracle(g, int(d["max_tracks"]))
    for f, (a, b) in enumerate(zip(g["recs"], ora)):
        ctx = "%s frame %d" % (case, f)
        # decisions: bit-exact
        assert a["M"] == b["M"], ctx
        np.testing.assert_array_equal(a["assoc"], b["assoc"], err_msg=ctx)
        assert (a["labels"] is None) == (b["labels"] is None), ctx
        if a["labels"] is not None:
            np.testing.assert_array_equal(a["labels"], b["labels"], err_msg=ctx)
        np.testing.assert_array_equal(a["ring_counts"], b["ring_counts"][-len(a["ring_counts"]):] if len(
            a["ring_counts"]) else b["ring_counts"], err_msg=ctx)
        assert a["next_track_id"] == b["next_track_id"], ctx
        assert [t["id"] for t in a["tracks"]] == [t["id"] for t in b["tracks"]], ctx
        for ta, tb in zip(a["tracks"], b["tracks"]):
            assert ta["lifetime"] == tb["lifetime"] and ta["N_est"] == tb["N_est"], ctx
            assert ta["point_num"] == tb["point_num"] and ta["static"] == tb["static"], ctx
            np.testing.assert_array_equal(ta["ring_counts"], tb["ring_counts"], err_msg=ctx)
            for k in ("x", "P", "spread_est", "group_disp_est", "centroid", "min_vals", "max_vals"):
                np.testing.assert_allclose(ta[k], tb[k], rtol=FLOAT_TOL, atol=FLOAT_TOL, err_msg=ctx + " " + k)
        if a["features"] is not None:
            np.testing.assert_allclose(a["features"], b["features"], rtol=0, atol=1e-12, err_msg=ctx)


def test_dbscan_semantics_small():
    """Border point reachable from two clusters takes the lower label; noise = -1; self counts."""
    cfg = mo.OracleConfig(db_min_samples=4, db_eps=0.05)      # y = 0 -> weight 1, radius sqrt(.05) = 0.2236
    pts = np.zeros((10, 8))
    # index 0: border point between two 4-point groups (neighbourhood = itself + one core of each group = 3 < 4)
    pts[:, 0] = [0.35, 0.0, 0.05, 0.1, 0.15, 0.55, 0.6, 0.65, 0.7, 5.0]
    lab = mo.dbscan_labels(pts, cfg)
    assert lab.tolist() == [0, 0, 0, 0, 0, 1, 1, 1, 1, -1]
    # same cloud, groups swapped in index order: the border now goes to the group that is numbered first
    pts[:, 0] = [0.35, 0.55, 0.6, 0.65, 0.7, 0.0, 0.05, 0.1, 0.15, 5.0]
    assert mo.dbscan_labels(pts, cfg).tolist() == [0, 0, 0, 0, 0, 1, 1, 1, 1, -1]
    # empty and singleton inputs
    assert mo.dbscan_labels(np.zeros((0, 8)), cfg).shape == (0,)
    assert mo.dbscan_labels(np.zeros((1, 8)), cfg).tolist() == [-1]


def test_dbscan_oracle_against_sklearn_on_random_clouds():
    """The oracle's labelling against the DBSCAN the reference itself calls (scikit-learn, installed in this image;
    Utils.py:272-278) with exact neighbourhoods (algorithm="brute") and the reference's metric formula
    (Utils.py:242-247), on clouds small and dense enough for several clusters, border ties and noise."""
    from sklearn.cluster import DBSCAN
    cfg = mo.OracleConfig(db_min_samples=5, db_eps=0.02)

    def metric(p1, p2):
        w = 1 - ((p1[1] + p2[1]) / 2) * cfg.db_range_weight
        return w * ((p1[0] - p2[0]) ** 2 + (p1[1] - p2[1]) ** 2 + cfg.db_z_weight * ((p1[2] - p2[2]) ** 2))

    rng = np.random.default_rng(11)
    seen_border_choice = 0
    for trial in range(12):
        k = int(rng.integers(2, 6))
        centres = rng.uniform([-1, 1, 0.5], [1, 3, 1.5], size=(k, 3))
        pts = np.zeros((140, 8))
        own = rng.integers(0, k, size=140)
        pts[:, :3] = centres[own] + rng.normal(0, 0.09, size=(140, 3))
        pts[:20, :3] = rng.uniform([-1.5, 0.5, 0], [1.5, 3.5, 2], size=(20, 3))     # clutter
        pts[:, :3] = np.round(pts[:, :3] * 512) / 512                               # sensor lattice: exact ties occur
        want = DBSCAN(eps=cfg.db_eps, min_samples=cfg.db_min_samples, metric=metric,
                      algorithm="brute").fit_predict(pts[:, :3])
        got = mo.dbscan_labels(pts, cfg)
        np.testing.assert_array_equal(got, want, err_msg="trial %d" % trial)
        seen_border_choice += int(want.max() >= 1)
    assert seen_border_choice >= 6                   # most trials have at least two clusters


def test_features_layout():
    cfg = mo.OracleConfig()
    cloud = np.zeros((70, 8))
    cloud[:, 0] = np.linspace(1, -1, 70)
    cloud[:, 7] = 100
    tr = mo.Track(cloud, 0, cfg)
    f = mo.pose_features(tr, cfg)
    assert f.shape == (3, 8, 8, 5)
    assert np.all(f[1:] == 0)                                  # absent ring frames stay zero (Utils.py:493)
    x = f[0].reshape(64, 5)[:, 0]
    assert np.all(np.diff(x) >= 0) and len(x) == 64            # cut to the first 64, sorted by x
    f0 = mo.pose_features(tr, mo.OracleConfig(frames_batch=0))
    assert f0.shape == (8, 8, 5)
