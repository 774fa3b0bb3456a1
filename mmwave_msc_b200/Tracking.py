"""Drop-in for the reference's ``Tracking.py`` (TrackBuffer / BatchedData / track objects) on top of the GPU
path.  ``offline_main.py`` / ``main.py`` of the reference use exactly this surface:

    trackbuffer = TrackBuffer(); batch = BatchedData()
    trackbuffer.dt = ...; trackbuffer.t = ...
    trackbuffer.track(normalize_data(detObj), batch)        # Tracking.py:664-703
    trackbuffer.estimate_posture(model)                     # Tracking.py:705-734
    for track in trackbuffer.effective_tracks: track.state.x, track.keypoints, track.batch.effective_data, ...

All arithmetic (gating, association statistics, Kalman filter, DBSCAN, feature maps, CNN) runs in libmmw on the
device, one single-scene context per TrackBuffer.  The objects below only hold host-side *views*: the point
clouds that callers read through ``.buffer`` / ``.effective_data`` are slices of the ``normalize_data`` output
selected by the device's decisions, and the numeric attributes are copies of the device records.
"""
from __future__ import annotations

import time
from typing import List, Optional

import numpy as np

from . import _lib
from . import constants as const
from . import pose_weights as pw
from .batched import BatchedTracker, config_from_constants
from .Utils import RingBuffer, WorldPoints

ACTIVE = 1
INACTIVE = 0
STATIC = True
DYNAMIC = False


class BatchedData(RingBuffer):
    """Ring of the last ``FB_FRAMES_BATCH + 1`` frames and their concatenation (Tracking.py:21-71).  When it is
    the global ring handed to ``TrackBuffer.track`` its pops/clears are replayed on the device ring."""

    def __init__(self, init_data=None):
        init = np.empty((0, 8)) if init_data is None else init_data
        super().__init__(const.FB_FRAMES_BATCH + 1, init_val=init)
        self.effective_data = np.concatenate(list(self.buffer), axis=0)
        self._pending = []            # operations done by the caller since the last track(): "pop" / "clear"
        self._external_add = False

    def _refresh(self):
        self.effective_data = np.concatenate(list(self.buffer), axis=0) if len(self.buffer) else np.array([])

    def add_frame(self, new_data, _from_tracker: bool = False):
        while len(self.buffer) >= self.size:
            self.pop_frame(_from_tracker=True)
        super().append(new_data)
        self._refresh()
        if not _from_tracker:
            self._external_add = True

    def clear(self, _from_tracker: bool = False):
        self.buffer.clear()
        self.effective_data = np.array([])
        if not _from_tracker:
            self._pending.append("clear")

    def change_buffer_size(self, new_size):
        self.size = new_size

    def pop_frame(self, _from_tracker: bool = False):
        if len(self.buffer) > 0:
            self.buffer.popleft()
            if not _from_tracker:
                self._pending.append("pop")


class KalmanState:
    """View of a track's filter state (Tracking.py:74-97): x is (9, 1), P is (9, 9)."""

    def __init__(self, x, P):
        self.x = np.asarray(x, dtype=np.float64).reshape(9, 1).copy()
        self.P = np.asarray(P, dtype=np.float64).reshape(9, 9).copy()
        self.H = const.MOTION_MODEL.KF_H


class PointCluster:
    """View of the last associated cloud and its statistics (Tracking.py:100-136)."""

    def __init__(self, pointcloud, rec=None):
        self.pointcloud = pointcloud
        self.point_num = int(rec["point_num"]) if rec is not None else int(pointcloud.shape[0])
        if rec is not None:
            self.centroid = np.array(rec["centroid"])
            self.min_vals = np.array(rec["min_vals"])
            self.max_vals = np.array(rec["max_vals"])
            self.status = STATIC if rec["is_static"] else DYNAMIC


class ClusterTrack:
    """One element of ``TrackBuffer.effective_tracks`` (Tracking.py:139-407), refreshed from the device after
    every ``track()`` / ``estimate_posture()``."""

    def __init__(self, rec, cloud):
        self.id = int(rec["id"])
        self.status = ACTIVE
        self.batch = BatchedData(cloud)
        self.color = np.random.rand(3)
        self._load(rec, cloud)

    def _load(self, rec, cloud=None):
        if cloud is not None or not hasattr(self, "cluster"):
            self.cluster = PointCluster(cloud if cloud is not None else np.empty((0, 8)), rec)
        self.state = KalmanState(rec["x"], rec["P"])
        self.predict_x = self.state.x          # visualisation only in the reference (never read by its callers)
        self.N_est = float(rec["n_est"])
        self.spread_est = np.array(rec["spread_est"])
        self.group_disp_est = np.array(rec["group_disp_est"])
        self.lifetime = float(rec["lifetime"])
        self.keypoints = np.array(rec["keypoints"])

    def get_Rm(self):
        return np.diag((self.spread_est / 2) ** 2)


class PoseModel:
    """Carrier of the keypoint regressor's weights for ``estimate_posture`` (what ``keras.load_model`` returns
    in the reference): ``weights`` in Keras ``get_weights()`` order (see pose_weights.py)."""

    def __init__(self, weights, variant: Optional[int] = None):
        self.weights = [np.asarray(w, dtype=np.float32) for w in weights]
        self.variant = variant

    def get_weights(self):
        return self.weights

    @classmethod
    def seeded(cls, seed: int = 7):
        v = pw.VARIANT_3D if const.FB_FRAMES_BATCH == 2 else pw.VARIANT_2D
        return cls(pw.make_pose_weights(v, seed), v)


class TrackBuffer:
    def __init__(self):
        self.effective_tracks: List[ClusterTrack] = []
        self.next_track_id = 0
        self.dt = 0
        self.t = time.time()
        self._bt: Optional[BatchedTracker] = None
        self._model_key = None
        self._doppler_seen = False          # a frame with a non-zero Doppler value went to the device

    # ------------------------------------------------------------------------------------------------
    def _ctx(self, n_points: int) -> BatchedTracker:
        if self._bt is None:
            cap = 256
            while cap < n_points:
                cap *= 2
            self._bt = BatchedTracker(1, max_points=cap, max_tracks=max(8, 2 * int(const.TR_MAX_TRACKS) + 2),
                                      config=config_from_constants(const))
        elif n_points > self._bt.ncap:
            raise _lib.MmwError("frame has %d points but the context was created for %d per frame"
                                % (n_points, self._bt.ncap))
        return self._bt

    def has_active_tracks(self) -> bool:
        return len(self.effective_tracks) > 0

    def track(self, pointcloud, batch: BatchedData):
        """Predict, gate/associate, maintain, update, cluster the remainder, spawn (Tracking.py:664-703)."""
        raw = getattr(pointcloud, "raw", None)
        if not isinstance(pointcloud, WorldPoints) or raw is None or raw.shape[0] != pointcloud.shape[0]:
            raise TypeError("track() needs the array returned by this package's Utils.normalize_data "
                            "(the device tracker works from the sensor rows it carries)")
        world = np.asarray(pointcloud)
        bt = self._ctx(raw.shape[0])
        res = float(getattr(pointcloud, "doppler_res", 1.0))
        if bt.cfg.doppler_res != res:       # unit of the rows' Doppler column (Utils._doppler_units)
            if self._doppler_seen:
                raise ValueError("the Doppler resolution of the input changed mid-sequence (%r -> %r): set "
                                 "constants.DOPPLER_RESOLUTION" % (bt.cfg.doppler_res, res))
            bt.set_doppler_resolution(res)
        self._doppler_seen = self._doppler_seen or bool(np.any(raw[:, 3] != 0))
        if batch._external_add:
            raise NotImplementedError("frames added to the global ring by the caller are not mirrored on the device")
        for op in batch._pending:                       # preprocessing.py:263-264 pops the ring between frames
            bt.ring_pop(0) if op == "pop" else bt.ring_clear(0)
        batch._pending = []
        offsets = np.array([0, raw.shape[0]], dtype=np.int32)
        bt.step(raw, offsets, np.array([float(self.dt)]), pose=False, record_labels=True)
        recs, nt = bt.tracks()
        recs, nt = recs[0], int(nt[0])
        assoc = bt.point_assoc()
        labels, nfused = bt.labels()
        old = self.effective_tracks
        # association views (Tracking.py:648-653)
        for j, trk in enumerate(old):
            sel = assoc == j
            if sel.any():
                cloud = world[sel]
                trk.batch.add_frame(cloud, _from_tracker=True)
                trk._cloud = cloud
            else:
                trk._cloud = None
        by_id = {t.id: t for t in old}
        batch.add_frame(world[assoc == -1], _from_tracker=True)      # Tracking.py:691
        fused = batch.effective_data
        new_tracks: List[ClusterTrack] = []
        spawned = 0
        for k in range(nt):
            rec = recs[k]
            trk = by_id.get(int(rec["id"]))
            if trk is not None:
                trk._load(rec, trk._cloud)
            else:                                                     # spawned this frame (Tracking.py:576-589)
                cloud = fused[labels[0, :nfused[0]] == spawned]
                trk = ClusterTrack(rec, cloud)
                spawned += 1
            new_tracks.append(trk)
        if spawned:
            batch.clear(_from_tracker=True)                           # Tracking.py:699-700
        self.effective_tracks[:] = new_tracks
        self.next_track_id = int(bt.summary()[1][0])

    def estimate_posture(self, model):
        """Feature maps + CNN for every track (Tracking.py:705-734).  ``model`` may be a PoseModel / anything with
        ``get_weights()`` in Keras order (inference on the device), or an object with only ``predict`` (the
        feature maps come from the device, the caller's model does the inference)."""
        if self._bt is None or not self.effective_tracks:
            return
        bt = self._bt
        if hasattr(model, "get_weights"):
            key = id(model)
            if self._model_key != key:
                bt.load_pose_weights(model.get_weights(), getattr(model, "variant", None))
                self._model_key = key
            bt.estimate_posture()
        else:
            bt.pose_features_only()
            si, ti, feats = bt.pose_rows()
            if len(ti):
                kp = np.asarray(model.predict(feats.astype(np.float64)), dtype=np.float32)
                for row, k in enumerate(ti):
                    bt.set_keypoints(0, int(k), kp[row])
        recs, nt = bt.tracks()
        for k, trk in enumerate(self.effective_tracks[:int(nt[0])]):
            trk.keypoints = np.array(recs[0][k]["keypoints"])
