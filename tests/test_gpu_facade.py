"""The drop-in classes (mmwave_msc_b200.Tracking / Utils, same names as the reference's modules) driven the way
offline_main.py drives them, compared with the oracle frame by frame."""
import numpy as np
import pytest

from helpers import KEYPOINT_ATOL
from mmwave_msc_b200 import pose_weights as pw, synth
from oracle import mmw_oracle as mo

pytestmark = pytest.mark.gpu


def _detobj(raw):
    return {"x": raw[:, 0].tolist(), "y": raw[:, 1].tolist(), "z": raw[:, 2].tolist(),
            "doppler": raw[:, 3].tolist(), "peakVal": raw[:, 4].tolist()}


def test_facade_loop_matches_oracle():
    from mmwave_msc_b200 import Tracking, Utils
    sc = synth.gen_scene(5, 40)
    W = pw.make_pose_weights(pw.VARIANT_3D)
    model = Tracking.PoseModel(W, pw.VARIANT_3D)
    tb, batch = Tracking.TrackBuffer(), Tracking.BatchedData()
    so = mo.SceneOracle(pose_weights=W)
    for f, (raw, dt) in enumerate(zip(sc.frames, sc.dts())):
        tb.dt = dt
        eff = Utils.normalize_data(_detobj(raw))
        rec = so.step(raw, dt)
        np.testing.assert_array_equal(np.asarray(eff), rec["world"])
        if eff.shape[0] != 0:
            tb.track(eff, batch)
            tb.estimate_posture(model)
        assert len(tb.effective_tracks) == len(rec["tracks"]), "frame %d" % f
        assert tb.next_track_id == rec["next_track_id"]
        assert [len(x) for x in batch.buffer][-len(so.ring.frames):] == [len(x) for x in so.ring.frames] or \
            sum(len(x) for x in batch.buffer) == sum(len(x) for x in so.ring.frames)
        np.testing.assert_array_equal(batch.effective_data.reshape(-1, 8) if batch.effective_data.size else
                                      np.empty((0, 8)), so.ring.fused())
        for trk, t, ot in zip(tb.effective_tracks, rec["tracks"], so.tracks):
            assert trk.id == t["id"] and trk.state.x.shape == (9, 1) and trk.state.P.shape == (9, 9)
            np.testing.assert_allclose(trk.state.x[:, 0], t["x"], rtol=1e-6, atol=1e-9)
            np.testing.assert_allclose(trk.state.P, t["P"], rtol=1e-6, atol=1e-9)
            np.testing.assert_allclose(trk.cluster.centroid, t["centroid"], rtol=1e-9, atol=1e-12)
            assert trk.cluster.point_num == t["point_num"] and trk.cluster.status == t["static"]
            assert trk.lifetime == pytest.approx(t["lifetime"], abs=1e-12)
            np.testing.assert_allclose(trk.keypoints, t["keypoints"], rtol=0, atol=KEYPOINT_ATOL)
            # the clouds the GUI / dataset builder read (Visualizer.py:235, preprocessing.py:193-205)
            np.testing.assert_array_equal(trk.batch.effective_data, ot.ring.fused())
            assert [len(x) for x in trk.batch.buffer] == [len(x) for x in ot.ring.frames]
            np.testing.assert_array_equal(trk.cluster.pointcloud, ot.cloud)


def test_apply_dbscan_and_opaque_model():
    from mmwave_msc_b200 import Tracking, Utils
    rng = np.random.default_rng(0)
    pc = np.zeros((220, 8))
    pc[:100, :3] = rng.normal([0, 2, 1], [0.12, 0.12, 0.4], size=(100, 3))
    pc[100:200, :3] = rng.normal([1.5, 3, 1], [0.12, 0.12, 0.4], size=(100, 3))
    pc[200:, :3] = rng.uniform([-2.5, 0.3, 0], [2.5, 4.5, 2.5], size=(20, 3))
    clusters = Utils.apply_DBscan(pc)
    exp = mo.clusters_from_labels(pc, mo.dbscan_labels(pc))
    assert len(clusters) == len(exp) == 2
    for c, e in zip(clusters, exp):
        np.testing.assert_array_equal(np.array(c), e)
    assert Utils.altered_EuclideanDist([0.1, 2.0, 1.0], [0.4, 2.5, 0.2]) == 0.5557700000000001

    # a model object that only has .predict (Tracking.py:732): features from the device, inference by the caller
    class Opaque:
        def __init__(self):
            self.calls = []

        def predict(self, x):
            self.calls.append(np.array(x))
            return np.tile(np.arange(57, dtype=np.float64), (len(x), 1)) + np.arange(len(x))[:, None]

    sc = synth.gen_scene(1, 6)
    tb, batch, m = Tracking.TrackBuffer(), Tracking.BatchedData(), Opaque()
    so = mo.SceneOracle()
    for raw, dt in zip(sc.frames, sc.dts()):
        tb.dt = dt
        eff = Utils.normalize_data(_detobj(raw))
        rec = so.step(raw, dt)
        tb.track(eff, batch)
        tb.estimate_posture(m)
    assert len(tb.effective_tracks) >= 1
    np.testing.assert_allclose(m.calls[-1], rec["features"].astype(np.float32), rtol=0, atol=2e-6)
    for k, trk in enumerate(tb.effective_tracks):
        np.testing.assert_array_equal(trk.keypoints, np.arange(57, dtype=np.float32) + k)


def test_replay_of_reference_csv_log(tmp_path):
    """offline_main.py's control flow over a CSV log in the reference's format, vs the oracle fed by the same reader."""
    from mmwave_msc_b200 import Tracking, Utils, replay
    sc = synth.gen_scene(4, 50)
    synth.write_reference_csv(sc, str(tmp_path), frames_per_file=40)
    W = pw.make_pose_weights(pw.VARIANT_3D)
    so = mo.SceneOracle(pose_weights=W)
    om = Utils.OfflineManager(str(tmp_path))
    state = {"t": None, "first": True}
    seen = []

    def on_frame(frame_no, tb, batch, eff):
        seen.append((frame_no, len(tb.effective_tracks), [t.id for t in tb.effective_tracks]))

    tb, frames = replay.offline_replay(str(tmp_path), Tracking.PoseModel(W, pw.VARIANT_3D), on_frame)
    exp = []
    t_prev = None
    while not om.is_finished():
        ok, no, det = om.get_data()
        if not ok:
            continue
        dt = 0.1 if t_prev is None else det["posix"][0] / 1000 - t_prev
        t_prev = det["posix"][0] / 1000
        raw = np.stack([det[k] for k in ("x", "y", "z", "doppler", "peakVal")], axis=1)
        r = so.step(raw, dt)
        exp.append((no, len(r["tracks"]), [t["id"] for t in r["tracks"]]))
    assert frames == 50 and seen == exp
    for trk, ot in zip(tb.effective_tracks, so.tracks):
        np.testing.assert_allclose(trk.state.x[:, 0], ot.x, rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(trk.keypoints, ot.keypoints, rtol=0, atol=KEYPOINT_ATOL)
