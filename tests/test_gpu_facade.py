"""The drop-in classes (mmwave_msc_b200.Tracking / Utils, same names as the reference's modules) driven the way
offline_main.py drives them, compared with the oracle frame by frame."""
import os

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


def test_facade_takes_float64_doppler_of_a_real_radar_profile(monkeypatch):
    """detObj as the reference's reader produces it: doppler = dopplerIdx * dopplerResolutionMps in float64
    (ReadDataIWR1443.py:163-165), not fp32-representable.  The facade recovers the index, the device multiplies in
    float64: decisions and states equal the reference's own trace (tests/golden/c1_s31_doppler_idx.npz)."""
    import os
    from conftest import GOLDEN
    from oracle import trace_io
    from mmwave_msc_b200 import Tracking, Utils, constants
    g = trace_io.unpack(np.load(os.path.join(GOLDEN, "c1_s31_doppler_idx.npz")))
    monkeypatch.setattr(constants, "DOPPLER_RESOLUTION", None)
    monkeypatch.setattr(Utils, "_stage_ctx", None)
    tb, batch = Tracking.TrackBuffer(), Tracking.BatchedData()
    for f, (fr, dt, rec) in enumerate(zip(trace_io.reference_frames(g), g["dts"], g["recs"])):
        det = {k: fr[:, i].tolist() for i, k in enumerate(("x", "y", "z", "doppler", "peakVal"))}
        tb.dt = float(dt)
        eff = Utils.normalize_data(det)
        assert eff.shape[0] == rec["M"]
        assert np.isin(np.asarray(eff)[:, 6], fr[:, 3]).all()        # Doppler column: the float64 input values, exactly
        if eff.shape[0]:
            tb.track(eff, batch)
        assert [t.id for t in tb.effective_tracks] == [t["id"] for t in rec["tracks"]], "frame %d" % f
        for trk, t in zip(tb.effective_tracks, rec["tracks"]):
            np.testing.assert_allclose(trk.state.x[:, 0], t["x"], rtol=1e-6, atol=1e-9, err_msg="frame %d" % f)
    assert constants.DOPPLER_RESOLUTION == g["doppler_res"] and len(tb.effective_tracks) >= 1
    assert not np.array_equal(fr[:, 3].astype(np.float32).astype(np.float64), fr[:, 3])   # fp32 could not hold them


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


# ---- dataset-builder mode (preprocessing.py:148-275) -----------------------------------
def _oracle_preprocess(frames, dts_posix, ok_mask, pairs):
    """preprocess_dataset's frame loop (preprocessing.py:167-259) on the oracle: returns (valid frame numbers,
    export blocks, centroids, invalid frame numbers)."""
    so = mo.SceneOracle()
    out_f, out_r, out_c, invalid = [], [], [], []
    first, t_prev = True, 0.0
    for i, raw in enumerate(frames):
        fn, valid = i + 1, False
        if pairs is None or fn in pairs:
            if ok_mask[i]:
                dt = 0.1 if first else dts_posix[i] / 1000 - t_prev
                first, t_prev = False, dts_posix[i] / 1000
                rec = so.step(raw, dt)
                ex = mo.track0_export(so) if rec["ran"] else None
                if ex is not None:
                    valid = True
                    out_f.append(fn); out_r.append(ex[0]); out_c.append(ex[1])
            else:
                so.ring.pop()
        if not valid:
            invalid.append(fn)
    return out_f, out_r, out_c, invalid


def test_export_track0_matches_oracle_on_sequences():
    from mmwave_msc_b200.batched import BatchedTracker
    S, F = 6, 50
    batches = synth.gen_batch([3, 4, 5, 308, 311, 321], F)
    bt = BatchedTracker(S)
    oracles = [mo.SceneOracle() for _ in range(S)]
    n_valid = n_invalid = 0
    for b in batches:
        bt.step(b.points, b.offsets, b.dt, pose=False)
        rows, valid, cen = bt.export_track0()
        for s, o in enumerate(oracles):
            rec = o.step(b.points[b.offsets[s]:b.offsets[s + 1]], b.dt[s])
            ex = mo.track0_export(o) if rec["ran"] else None
            assert bool(valid[s]) == (ex is not None)
            if ex is None:
                n_invalid += 1
                assert not rows[s].any()
                continue
            n_valid += 1
            np.testing.assert_allclose(rows[s], ex[0], rtol=0, atol=1e-9)       # centroid-relative columns
            np.testing.assert_array_equal(rows[s][:, 2:], ex[0][:, 2:])          # copied columns: exact
            np.testing.assert_array_equal((rows[s] != 0).any(1), (ex[0] != 0).any(1))
            np.testing.assert_allclose(cen[s], ex[1], rtol=1e-9, atol=1e-12)
    assert n_valid > 100


def test_preprocess_dataset_batched_matches_oracle_loop(tmp_path):
    """Three recorded experiments (reference CSV logs, one with dropped frames and a frame-pair filter) through
    preprocess_dataset_batched vs the same loop on the oracle."""
    from mmwave_msc_b200 import preprocessing as pp
    specs = [(4, 47, None, None), (308, 55, {7, 8, 23}, None), (5, 38, None, set(range(1, 39, 2)) | {2, 4, 6})]
    dirs, want = [], []
    for sid, nf, drop, pairs in specs:
        sc = synth.gen_scene(sid, nf)
        ok = np.ones(nf, bool)
        if drop:
            for fnum in drop:
                ok[fnum - 1] = False
        d = str(tmp_path / ("exp%d" % sid))
        # dropped frames simply have no rows in the log -> OfflineManager reports dataOk = False for them
        import copy
        sc2 = copy.copy(sc)
        sc2.frames = [fr if ok[i] else fr[:0] for i, fr in enumerate(sc.frames)]
        synth.write_reference_csv(sc2, d, frames_per_file=20)
        dirs.append(d)
        want.append((sc, ok, pairs))
    got = pp.preprocess_dataset_batched(dirs, frame_pairs=[w[2] for w in want])
    for g, (sc, ok, pairs) in zip(got, want):
        # what the reference's reader actually delivers (including its read-buffer quirk, Q27) is replayed for
        # the oracle loop too, so both sides see the same frames
        from mmwave_msc_b200.Utils import OfflineManager
        rd = OfflineManager(os.path.join(str(tmp_path), g.name))
        frames, posix, okm = [], [], []
        while not rd.is_finished():
            dok, fn, det = rd.get_data()
            okm.append(dok)
            if dok:
                frames.append(np.stack([np.asarray(det[k], np.float32) for k in
                                        ("x", "y", "z", "doppler", "peakVal")], axis=1))
                posix.append(det["posix"][0])
            else:
                frames.append(np.zeros((0, 5), np.float32)); posix.append(0)
        wf, wr, wc, winv = _oracle_preprocess(frames, posix, okm, pairs)
        assert g.frames == wf and g.invalid_frames == winv, g.name
        for a, b in zip(g.rows, wr):
            np.testing.assert_allclose(a, b, rtol=0, atol=1e-9)
        for a, b in zip(g.centroids, wc):
            np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12)
        assert len(wf) > 10 and len(winv) > 0
    paths = pp.write_export_csv(got[0], str(tmp_path / "out"))
    import csv as _csv
    rows = list(_csv.reader(open(paths[0])))
    assert len(rows) == 192 * min(len(got[0].frames), 200) and len(rows[0]) == 6
    assert int(rows[0][0]) == got[0].frames[0] and float(rows[0][1]) == got[0].rows[0][0, 0]
