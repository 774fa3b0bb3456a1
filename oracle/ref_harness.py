"""TEST INFRASTRUCTURE ONLY -- drives the reference's own, unmodified modules.

Runs ``Utils.normalize_data`` / ``Tracking.TrackBuffer.track`` /
``TrackBuffer.estimate_posture`` imported from ``/root/reference/src`` (read
only; present in the build container, absent on the GPU box) with

* the filterpy restatement of ``oracle/filterpy_shim`` on ``sys.path``
  (filterpy==1.4.5 is pinned by the reference but not installed),
* ``Utils.DBSCAN`` wrapped to ``algorithm="brute"`` so neighbourhoods are the
  exact epsilon-neighbourhoods of the reference's own metric function
  (SURVEY.md section 8(c), Appendix A Q2); ``algorithm="auto"`` (the shipped BallTree
  behaviour) can be selected to measure the deviation,
* ``numpy.argsort`` inside ``Utils.format_single_frame`` forced to
  ``kind="stable"`` (canonical tie order, Appendix A Q21) -- optional.

and records every stage's outputs per frame.  It is used by
``oracle/gen_golden.py`` to produce ``tests/golden/*.npz`` and by the tests that
pin ``oracle/mmw_oracle.py`` when ``/root/reference`` is available.  Nothing in
``mmwave_msc_b200/`` imports it.
"""
from __future__ import annotations

import functools
import os
import sys
from typing import Dict, List, Optional

import numpy as np

REF_SRC = os.environ.get("MMW_REFERENCE_SRC", "/root/reference/src")
if not os.path.isfile(os.path.join(REF_SRC, "Tracking.py")):      # GPU box: the copy staged by oracle/make_ref.py
    _staged = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "src")
    if os.path.isfile(os.path.join(_staged, "Tracking.py")):
        REF_SRC = _staged
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "filterpy_shim")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "Tracking.py"))


_mods = None


def load_reference():
    """Import the reference's constants/Utils/Tracking (once)."""
    global _mods
    if _mods is not None:
        return _mods
    if not reference_available():
        raise RuntimeError("reference sources not found at %s" % REF_SRC)
    for p in (_SHIM, REF_SRC):
        if p not in sys.path:
            sys.path.insert(0, p)
    import constants as ref_const   # noqa
    import Utils as ref_utils       # noqa
    import Tracking as ref_tracking  # noqa
    _mods = (ref_const, ref_utils, ref_tracking)
    return _mods


class _StableArgsortNumpy:
    """Proxy for the ``np`` name inside Utils: argsort defaults to kind='stable'."""

    def __init__(self, real):
        self._real = real

    def __getattr__(self, name):
        return getattr(self._real, name)

    def argsort(self, a, *args, **kw):
        kw.setdefault("kind", "stable")
        return self._real.argsort(a, *args, **kw)


class patched_reference:
    """Context manager applying the harness patches to the loaded reference."""

    def __init__(self, dbscan_algorithm: str = "brute", stable_sort: bool = True,
                 max_tracks: Optional[int] = None):
        self.alg = dbscan_algorithm
        self.stable = stable_sort
        self.max_tracks = max_tracks

    def __enter__(self):
        import sklearn.cluster
        const, utils, tracking = load_reference()
        self._saved = (utils.DBSCAN, utils.np, const.TR_MAX_TRACKS)
        self.labels_log: List[np.ndarray] = []
        log = self.labels_log

        class _DBSCAN(sklearn.cluster.DBSCAN):
            def fit_predict(self_inner, X, y=None, sample_weight=None):
                lab = super().fit_predict(X, y, sample_weight)
                log.append(np.asarray(lab).copy())
                return lab

        utils.DBSCAN = functools.partial(_DBSCAN, algorithm=self.alg)
        if self.stable:
            utils.np = _StableArgsortNumpy(np)
        if self.max_tracks is not None:
            const.TR_MAX_TRACKS = self.max_tracks
        return self

    def __exit__(self, *exc):
        const, utils, tracking = load_reference()
        utils.DBSCAN, utils.np, const.TR_MAX_TRACKS = self._saved
        return False


def detobj_from_raw(raw: np.ndarray) -> Dict[str, list]:
    """fp32 (N,5) sensor points -> the dict normalize_data takes (exact up-cast)."""
    r = np.asarray(raw, dtype=np.float64)
    return {"x": r[:, 0].tolist(), "y": r[:, 1].tolist(), "z": r[:, 2].tolist(),
            "doppler": r[:, 3].tolist(), "peakVal": r[:, 4].tolist()}


class _ModelAdapter:
    """Object with .predict(ndarray) -> (n, 57), as Tracking.py:732 expects."""

    def __init__(self, fn):
        self.fn = fn
        self.last_input = None

    def predict(self, x):
        self.last_input = np.array(x, copy=True)
        return self.fn(x)


def run_reference_scene(frames: List[np.ndarray], dts: np.ndarray, pose_fn=None,
                        dbscan_algorithm: str = "brute", stable_sort: bool = True,
                        max_tracks: Optional[int] = None, export0: bool = False,
                        missing: Optional[List[bool]] = None) -> List[dict]:
    """Run the offline_main.py loop body (offline_main.py:45-60) on one scene.

    Returns one record per frame with the decisions and states after the frame.
    Track ids are the value of ``next_track_id`` at spawn (Tracking.py:587-588).
    ``missing[f]``: the frame has no sensor data -- the dataset builder's loop then calls ``batch.pop_frame()``
    instead of the loop body (preprocessing.py:177-264); ``frames[f]`` is ignored.
    """
    const, utils, tracking = load_reference()
    out: List[dict] = []
    with patched_reference(dbscan_algorithm, stable_sort, max_tracks) as pr:
        tb = tracking.TrackBuffer()
        batch = tracking.BatchedData()
        model = _ModelAdapter(pose_fn) if pose_fn is not None else None
        ids: Dict[int, int] = {}
        assoc_log: List[np.ndarray] = []
        orig_calc = tb._calc_dist_fun

        def calc_spy(full_set):
            r = orig_calc(full_set)
            assoc_log.append(np.array([-1 if a is None else int(a) for a in r], dtype=np.int32))
            return r

        tb._calc_dist_fun = calc_spy
        for f, raw in enumerate(frames):
            tb.dt = float(dts[f])
            gone = missing is not None and bool(missing[f])
            if gone:
                batch.pop_frame()                                # preprocessing.py:262-264
                raw = np.zeros((0, 5))
            eff = utils.normalize_data(detobj_from_raw(raw))
            rec = {"M": int(eff.shape[0]), "world": eff.copy()}
            n_lab0, n_as0 = len(pr.labels_log), len(assoc_log)
            ran = eff.shape[0] != 0
            if ran:
                next_before = tb.next_track_id
                tb.track(eff, batch)
                # ids of tracks spawned in this frame: they are appended in cluster order
                n_new = tb.next_track_id - next_before
                for k, tr in enumerate(tb.effective_tracks[len(tb.effective_tracks) - n_new:]):
                    ids[id(tr)] = next_before + k
                if model is not None:
                    model.last_input = None
                    tb.estimate_posture(model)
            if export0:
                # the body of preprocess_dataset's frame loop (preprocessing.py:185-216), calling the reference's own
                # relative_coordinates / format_batched_frames (preprocessing.py itself runs at import and needs
                # wakepy + the authors' dataset, so its loop body is restated here around the reference's functions)
                rec["export0"] = None
                if ran and len(tb.effective_tracks) > 0:
                    t0 = tb.effective_tracks[0]
                    if t0.lifetime == 0 and len(t0.batch.effective_data) > 0:
                        fr = utils.relative_coordinates(list(t0.batch.buffer), t0.cluster.centroid)
                        rec["export0"] = (np.array(utils.format_batched_frames(fr), dtype=np.float64),
                                          np.array(t0.cluster.centroid[:2], dtype=np.float64))
            rec["ran"] = ran
            rec["assoc"] = assoc_log[n_as0].copy() if len(assoc_log) > n_as0 else np.zeros(0, np.int32)
            rec["labels"] = pr.labels_log[n_lab0].copy() if len(pr.labels_log) > n_lab0 else None
            rec["ring_counts"] = np.array([len(fr) for fr in batch.buffer], dtype=np.int32)
            rec["next_track_id"] = int(tb.next_track_id)
            tracks = []
            for tr in tb.effective_tracks:
                tracks.append({
                    "id": ids[id(tr)],
                    "x": tr.state.x[:, 0].copy(), "P": tr.state.P.copy(),
                    "lifetime": float(tr.lifetime),
                    "spread_est": np.array(tr.spread_est, dtype=np.float64).copy(),
                    "group_disp_est": tr.group_disp_est.copy(),
                    "N_est": float(tr.N_est), "point_num": int(tr.cluster.point_num),
                    "centroid": tr.cluster.centroid.copy(),
                    "min_vals": tr.cluster.min_vals.copy(), "max_vals": tr.cluster.max_vals.copy(),
                    "static": bool(tr.cluster.status),
                    "ring_counts": np.array([len(fr) for fr in tr.batch.buffer], dtype=np.int32),
                    "keypoints": np.array(tr.keypoints, dtype=np.float64).copy(),
                })
            rec["tracks"] = tracks
            rec["features"] = None if (model is None or model.last_input is None) else model.last_input
            out.append(rec)
    return out
