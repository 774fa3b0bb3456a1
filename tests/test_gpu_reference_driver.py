"""SURVEY section 4 tier 5, north_star "offline_main.py drives it unchanged": the reference's OWN script
(`/root/reference/src/offline_main.py:21-68`, executed with runpy, not restated) runs against the GPU drop-in modules.

The script imports `constants`, `Utils`, `Tracking` as top-level modules; `mmwave_msc_b200/dropin/` first on sys.path
binds those names to the GPU package.  Its GUI / keep-awake / Keras imports are stubbed
(PyQt5.QtWidgets.QApplication, wakepy.keep.presenting, Visualizer.VisualManager, keras.models.load_model); the
Visualizer stub records what `visual.update(trackbuffer, detObj)` sees after every frame, which is compared with the
oracle stepping the same CSV log frame by frame: track ids, Kalman states, keypoints.

The script comes from `/root/reference/src` in the build container and from the unmodified copy staged by
`oracle/make_ref.py` under `oracle/_ref/src` (git-ignored, travels with the snapshot) on the GPU box."""
import contextlib
import os
import runpy
import sys
import types

import numpy as np
import pytest

from helpers import KEYPOINT_ATOL
from mmwave_msc_b200 import pose_weights as pw, synth
from oracle import make_ref, mmw_oracle as mo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_modules(model, seen):
    qt, qtw = types.ModuleType("PyQt5"), types.ModuleType("PyQt5.QtWidgets")

    class QApplication:
        def __init__(self, argv):
            self.argv = argv

    qtw.QApplication = QApplication
    qt.QtWidgets = qtw
    wakepy = types.ModuleType("wakepy")
    wakepy.keep = types.SimpleNamespace(presenting=lambda: contextlib.nullcontext())
    vis = types.ModuleType("Visualizer")

    class VisualManager:
        def update(self, trackbuffer, detObj):
            seen.append({"module": type(trackbuffer).__module__,
                         "ids": [t.id for t in trackbuffer.effective_tracks],
                         "x": [t.state.x[:, 0].copy() for t in trackbuffer.effective_tracks],
                         "keypoints": [np.array(t.keypoints) for t in trackbuffer.effective_tracks],
                         "n_points": len(detObj["x"])})

    vis.VisualManager = VisualManager
    keras, kmodels = types.ModuleType("keras"), types.ModuleType("keras.models")
    kmodels.load_model = lambda path: model
    keras.models = kmodels
    return {"PyQt5": qt, "PyQt5.QtWidgets": qtw, "wakepy": wakepy, "Visualizer": vis, "keras": keras,
            "keras.models": kmodels}


def test_reference_offline_main_drives_the_dropin_modules(tmp_path, monkeypatch):
    ref = make_ref.staged_dir()
    if not ref:
        pytest.skip("reference sources neither at /root/reference/src nor staged under oracle/_ref/src "
                    "(run __graft_entry__.build() where the reference is present)")
    from mmwave_msc_b200 import Tracking as gpu_tracking, Utils as gpu_utils
    n_frames = 60
    sc = synth.gen_scene(4, n_frames)
    exp = tmp_path / "dataset" / "log" / "mmWave" / "A21"        # EXPERIMENT_PATH of the script, relative to the cwd
    exp.mkdir(parents=True)
    synth.write_reference_csv(sc, str(exp), frames_per_file=40)
    monkeypatch.chdir(tmp_path)
    W = pw.make_pose_weights(pw.VARIANT_3D)
    model = gpu_tracking.PoseModel(W, pw.VARIANT_3D)             # what keras.load_model would hand back: get_weights()
    seen = []
    names = ["constants", "Utils", "Tracking"]
    saved = {k: sys.modules.get(k) for k in names + list(_stub_modules(model, []))}
    saved_path = list(sys.path)
    try:
        for k in names:                                          # e.g. the reference's own modules from the oracle harness
            sys.modules.pop(k, None)
        sys.modules.update(_stub_modules(model, seen))
        sys.path.insert(0, os.path.join(ROOT, "mmwave_msc_b200", "dropin"))
        runpy.run_path(os.path.join(ref, "offline_main.py"), run_name="__main__")
        assert sys.modules["Tracking"] is gpu_tracking and sys.modules["Utils"] is gpu_utils
    finally:
        sys.path[:] = saved_path
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v

    # the oracle over the same log, through the same reader and the script's dt rule (offline_main.py:45-51)
    so = mo.SceneOracle(pose_weights=W)
    om = gpu_utils.OfflineManager(str(exp))
    expected = []
    t_prev = None
    while not om.is_finished():
        ok, _, det = om.get_data()
        if not ok:
            continue
        dt = 0.1 if t_prev is None else det["posix"][0] / 1000 - t_prev
        t_prev = det["posix"][0] / 1000
        raw = np.stack([det[k] for k in ("x", "y", "z", "doppler", "peakVal")], axis=1)
        so.step(raw, dt)
        expected.append({"ids": [t.id for t in so.tracks], "x": [t.x.copy() for t in so.tracks],
                         "keypoints": [np.array(t.keypoints) for t in so.tracks], "n_points": len(raw)})
    assert len(seen) == len(expected) == n_frames
    assert all(s["module"] == "mmwave_msc_b200.Tracking" for s in seen)
    assert max(len(e["ids"]) for e in expected) >= 1
    for f, (s, e) in enumerate(zip(seen, expected)):
        assert s["n_points"] == e["n_points"] and s["ids"] == e["ids"], "frame %d" % f
        for k in range(len(e["ids"])):
            np.testing.assert_allclose(s["x"][k], e["x"][k], rtol=1e-6, atol=1e-9, err_msg="frame %d track %d" % (f, k))
            np.testing.assert_allclose(s["keypoints"][k], e["keypoints"][k], rtol=0, atol=KEYPOINT_ATOL,
                                       err_msg="frame %d track %d keypoints" % (f, k))
