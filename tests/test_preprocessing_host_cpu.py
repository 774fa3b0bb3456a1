"""Host logic of the dataset-builder driver (mmwave_msc_b200/preprocessing.py) on CPU: the device context is replaced
by a stand-in with the same methods that steps one oracle per scene, so the frame pairing, first-frame dt,
pop-on-missing-frame and invalid-frame bookkeeping of preprocess_dataset (preprocessing.py:167-259) are checked
without a GPU.  (The same driver against the real device context: tests/test_gpu_facade.py.)"""
import copy
import os

import types

import numpy as np
import pytest

from mmwave_msc_b200 import preprocessing as pp, synth
from oracle import mmw_oracle as mo


class _OracleTracker:
    """BatchedTracker look-alike: step / ring_pop / export_track0 on SceneOracle objects."""
    instances = []

    def __init__(self, n_scenes, max_points=256, max_tracks=8, device=0, config=None):
        self.S = n_scenes
        self.scenes = [mo.SceneOracle() for _ in range(n_scenes)]
        self.ran = [False] * n_scenes
        self.log = []
        self.cfg = types.SimpleNamespace(doppler_res=1.0)
        _OracleTracker.instances.append(self)

    def set_doppler_resolution(self, res):
        self.cfg.doppler_res = float(res)

    def ring_pop(self, scene):
        self.log.append(("pop", scene))
        self.scenes[scene].ring.pop()

    def step(self, points, offsets, dt, pose=True, record_labels=False):
        assert not pose
        for s, o in enumerate(self.scenes):
            raw = np.asarray(points[offsets[s]:offsets[s + 1]], np.float64).copy()
            raw[:, 3] *= self.cfg.doppler_res            # the rows carry the Doppler index (the device multiplies in float64)
            self.ran[s] = False
            if len(raw):
                self.ran[s] = bool(o.step(raw, dt[s])["ran"])
                self.log.append(("step", s, len(raw), float(dt[s])))

    def export_track0(self):
        rows = np.zeros((self.S, 192, 5))
        valid = np.zeros(self.S, bool)
        cen = np.zeros((self.S, 2))
        for s, o in enumerate(self.scenes):
            ex = mo.track0_export(o) if self.ran[s] else None
            if ex is not None:
                rows[s], cen[s], valid[s] = ex[0], ex[1], True
        return rows, valid, cen


def test_driver_control_flow(tmp_path, monkeypatch):
    monkeypatch.setattr(pp, "BatchedTracker", _OracleTracker)
    monkeypatch.setattr(pp, "config_from_constants", lambda const: None)
    specs = [(4, 30, {5, 6}, None), (308, 26, None, {1, 2, 3, 4, 10, 11, 12, 20, 21, 22, 23, 24})]
    dirs = []
    for sid, nf, drop, pairs in specs:
        sc = synth.gen_scene(sid, nf)
        sc2 = copy.copy(sc)
        sc2.frames = [fr if not (drop and i + 1 in drop) else fr[:0] for i, fr in enumerate(sc.frames)]
        d = str(tmp_path / ("exp%d" % sid))
        synth.write_reference_csv(sc2, d, frames_per_file=10)
        dirs.append(d)
    out = pp.preprocess_dataset_batched(dirs, frame_pairs=[s[3] for s in specs])
    trk = _OracleTracker.instances[-1]
    # experiment 0: every frame paired; frames 5 and 6 are missing from the log -> ring pops, listed invalid; so is the
    # reader's trailing call past the last frame (get_data() returns dataOk = False once before is_finished() turns
    # true, Utils.py:158-166 -- with frame_pairs=None that frame number counts as paired, like any other)
    assert [e for e in trk.log if e[0] == "pop"] == [("pop", 0)] * 3
    assert {5, 6, 31} <= set(out[0].invalid_frames)
    assert sorted(out[0].frames + out[0].invalid_frames) == list(range(1, 32))
    steps0 = [e for e in trk.log if e[0] == "step" and e[1] == 0]
    assert steps0[0][3] == 0.1                                       # first processed frame: dt = 0.1 (:178-182)
    sc = synth.gen_scene(4, 30)
    assert steps0[1][3] == pytest.approx((sc.posix_ms[1] - sc.posix_ms[0]) / 1000)
    # the frame after the gap: dt spans the gap (posix difference to the last PROCESSED frame)
    after_gap = [e for e in steps0 if e[3] > 0.2]
    assert len(after_gap) == 1 and after_gap[0][3] == pytest.approx((sc.posix_ms[6] - sc.posix_ms[3]) / 1000)
    # experiment 1: only the paired frames are stepped; all others are invalid without touching the tracker
    steps1 = [e for e in trk.log if e[0] == "step" and e[1] == 1]
    assert len(steps1) == len(specs[1][3])
    assert set(out[1].frames) <= specs[1][3]
    assert set(range(1, 28)) - specs[1][3] <= set(out[1].invalid_frames)
    # exports are the oracle's blocks
    assert len(out[0].frames) > 15 and all(r.shape == (192, 5) for r in out[0].rows)
    # the shorter experiment finishes first; the other keeps running
    assert max(out[0].frames + out[0].invalid_frames) == 31 and max(out[1].frames + out[1].invalid_frames) == 27
    paths = pp.write_export_csv(out[0], str(tmp_path / "csv"))
    assert os.path.isfile(paths[0])
