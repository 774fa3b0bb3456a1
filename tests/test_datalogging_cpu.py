"""Experiment logs on disk (mmwave_msc_b200/DataLogging.py): the recorder's point schema is what the reference's
reader replays (DataLogging.py:60-89 <-> Utils.py:86-142), the result log round-trips packed result records."""
import os

import numpy as np

from mmwave_msc_b200 import DataLogging, Utils, _lib, synth


def test_point_logger_is_replayed_by_the_offline_manager(tmp_path):
    sc = synth.gen_scene(3, 25)
    with DataLogging.PointLogger(str(tmp_path), frames_per_file=10, buffer_rows=40) as log:
        for f, fr in enumerate(sc.frames):
            det = {k: fr[:, i] for i, k in enumerate(("x", "y", "z", "doppler", "peakVal"))}
            log.log(f + 1, det, int(sc.posix_ms[f]))
    assert sorted(os.listdir(tmp_path)) == ["1.csv", "2.csv", "3.csv"]
    om = Utils.OfflineManager(str(tmp_path))
    seen = 0
    while not om.is_finished():
        ok, no, det = om.get_data()
        if ok and len(det["x"]) == len(sc.frames[no - 1]):        # (the reader's Q27 frame is delivered cut short)
            np.testing.assert_array_equal(np.array(det["x"], np.float32), sc.frames[no - 1][:, 0])
            np.testing.assert_array_equal(np.array(det["doppler"], np.float32), sc.frames[no - 1][:, 3])
            assert det["posix"][0] == int(sc.posix_ms[no - 1])
            seen += 1
    assert seen >= 23


def test_result_logger_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    S, T = 3, 8
    frames = []
    with DataLogging.ResultLogger(str(tmp_path), max_tracks=T, frames_per_file=4, buffer_rows=5) as log:
        for f in range(9):
            rec = np.zeros((S, T, _lib.RESULT_FLOATS), np.float32)
            rec[:, :, 0] = -1
            for s in range(S):
                nt = int(rng.integers(0, 4))
                rec[s, :, 1] = nt
                for k in range(nt):
                    rec[s, k, 0] = 10 * s + k
                    rec[s, k, 2:71] = rng.normal(size=69).astype(np.float32)
            frames.append(rec)
            log.log_packed(f + 1, rec.reshape(-1), 1000 + 83 * f)
    back = dict(DataLogging.read_results(str(tmp_path)))
    assert len(os.listdir(tmp_path)) == 3
    for f, rec in enumerate(frames):
        live = rec[rec[:, :, 0] >= 0]
        if len(live) == 0:
            assert f + 1 not in back
            continue
        rows = back[f + 1]
        assert rows.shape == (len(live), len(DataLogging.RESULT_COLUMNS))
        np.testing.assert_array_equal(rows[:, 2], live[:, 0])                       # track ids
        np.testing.assert_array_equal(rows[:, 1], np.nonzero(rec[:, :, 0] >= 0)[0])  # scenes
        np.testing.assert_array_equal(rows[:, 4:73].astype(np.float32), live[:, 2:71])
        assert (rows[:, -1] == 1000 + 83 * f).all()
