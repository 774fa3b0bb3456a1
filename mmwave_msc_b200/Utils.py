"""Drop-in for the hot-path part of the reference's ``Utils.py``: same names, arguments and return types,
computed on the GPU through libmmw (include/mmw.h).

    normalize_data(detObj)        Utils.py:342-434   -> mmw_preprocess
    apply_DBscan(pc, eps, ms)     Utils.py:250-291   -> mmw_dbscan (exact epsilon-neighbourhoods, sklearn labelling)
    altered_EuclideanDist(p, q)   Utils.py:222-247   (scalar formula, kept for callers that print/plot it)
    RingBuffer                    Utils.py:10-50
    OfflineManager                Utils.py:53-177    reader of the reference's CSV experiment logs

There is no CPU fallback: without a CUDA device the first call raises ``_lib.MmwError``.
"""
from __future__ import annotations

import csv
import os
from collections import deque
from typing import Optional

import numpy as np

from . import constants as const
from .batched import BatchedTracker, config_from_constants


class WorldPoints(np.ndarray):
    """(M, 8) float64 array returned by normalize_data that remembers the fp32 sensor rows it came from, which
    is what the device tracker consumes (rings are stored as raw rows, SURVEY/DESIGN 'data layout'), and the unit
    of their Doppler column."""

    def __new__(cls, world: np.ndarray, raw: np.ndarray, doppler_res: float = 1.0):
        obj = np.asarray(world, dtype=np.float64).view(cls)
        obj.raw = raw
        obj.doppler_res = doppler_res
        return obj

    def __array_finalize__(self, obj):
        self.raw = None          # any derived view/copy loses the tag
        self.doppler_res = 1.0


_stage_ctx: Optional[BatchedTracker] = None


def _stage(min_points: int = 0) -> BatchedTracker:
    """Shared single-scene context for the stateless entry points."""
    global _stage_ctx
    need = max(256, int(min_points))
    if _stage_ctx is None or _stage_ctx.ncap < need:
        cap = 256
        while cap < need:
            cap *= 2
        _stage_ctx = BatchedTracker(1, max_points=cap, max_tracks=8, config=config_from_constants(const))
    return _stage_ctx


def _doppler_units(d: np.ndarray):
    """Doppler column of the device rows and its unit.  The sensor reports dopplerIdx and the reference forms
    doppler = dopplerIdx * dopplerResolutionMps in float64 (ReadDataIWR1443.py:163-165; DataLogging.py writes that
    value to the CSV logs) -- not representable in fp32.  The device therefore takes the INDEX (exact in fp32) and
    multiplies by the resolution in float64.  The resolution is ``constants.DOPPLER_RESOLUTION`` when set; else, if
    the values are fp32-representable they are passed as they are (unit 1.0); else it is inferred from the data once
    (every value must be an exact integer multiple) and kept in ``constants.DOPPLER_RESOLUTION``."""
    res = getattr(const, "DOPPLER_RESOLUTION", None)
    if res is None:
        d32 = d.astype(np.float32)
        if np.array_equal(d32.astype(np.float64), d):
            return d32, 1.0
        nz = np.abs(d[d != 0])
        for k in range(1, 65):
            r = float(nz.min()) / k
            if np.array_equal(np.rint(d / r) * r, d):
                res = const.DOPPLER_RESOLUTION = r
                break
        else:
            raise ValueError("doppler values are neither fp32-representable nor integer multiples of one resolution: "
                             "set constants.DOPPLER_RESOLUTION to the sensor's dopplerResolutionMps")
    idx = np.rint(d / res)
    if not np.array_equal(idx * res, d) or (len(idx) and np.abs(idx).max() >= 2 ** 24):
        raise ValueError("doppler values are not integer multiples of constants.DOPPLER_RESOLUTION = %r" % res)
    return idx.astype(np.float32), float(res)


def normalize_data(detObj) -> np.ndarray:
    """Sensor dict {x, y, z, doppler, peakVal} -> (M, 8) [x y z vx vy vz doppler peakVal] in the room frame,
    restricted to the scene bounds (same contract as the reference)."""
    raw = np.stack([np.asarray(detObj[k], dtype=np.float64) for k in ("x", "y", "z", "doppler", "peakVal")], axis=1)
    raw32 = raw.astype(np.float32)
    if not np.array_equal(raw32[:, [0, 1, 2, 4]].astype(np.float64), raw[:, [0, 1, 2, 4]]):
        raise ValueError("x, y, z, peakVal must be representable in float32 (they are int16 / 2^Q lattice values)")
    raw32[:, 3], res = _doppler_units(raw[:, 3])
    if raw32.shape[0] == 0:
        return WorldPoints(np.empty((0, 8)), raw32, res)
    st = _stage()
    if st.cfg.doppler_res != res:
        st.set_doppler_resolution(res)
    world, keep = st.preprocess(raw32)
    return WorldPoints(world[keep], np.ascontiguousarray(raw32[keep]), res)


def altered_EuclideanDist(p1, p2) -> float:
    w = 1 - ((p1[1] + p2[1]) / 2) * const.DB_RANGE_WEIGHT
    return w * ((p1[0] - p2[0]) ** 2 + (p1[1] - p2[1]) ** 2 + const.DB_Z_WEIGHT * ((p1[2] - p2[2]) ** 2))


def apply_DBscan(pointcloud, eps=None, min_samples=None):
    """List of clusters (ascending label), each a list of the input rows in input order; noise dropped."""
    pc = np.asarray(pointcloud, dtype=np.float64)
    if pc.ndim != 2 or pc.shape[0] == 0:
        return []
    labels = _stage(-(-pc.shape[0] // 3)).dbscan([pc[:, :3]], eps=const.DB_EPS if eps is None else eps,
                                                min_samples=const.DB_MIN_SAMPLES_MIN if min_samples is None
                                                else min_samples)[0]
    return [[pointcloud[i] for i in np.nonzero(labels == c)[0]] for c in range(int(labels.max()) + 1)]


class RingBuffer:
    """Fixed-size FIFO (deque) as in the reference; ``BatchedData`` derives from it."""

    def __init__(self, size, init_val=None):
        self.size = size
        self.buffer = deque(maxlen=size)
        self.append(0 if init_val is None else init_val)

    def append(self, item):
        self.buffer.append(item)

    def get_max(self):
        return np.max(self.buffer)

    def get_mean(self):
        return np.mean(self.buffer)


class OfflineManager:
    """Replays a reference experiment log: directory with 1.csv, 2.csv, ... rows
    ``frame,x,y,z,doppler,peakVal,posix_ms`` (DataLogging.py:60-89).  ``get_data()`` hands out one frame at a
    time as (exists, frame_count, dict) and ``is_finished()`` turns true after the last file."""

    KEYS = ("x", "y", "z", "doppler", "peakVal", "posix")

    def __init__(self, experiment_path):
        self.experiment_path = experiment_path
        self.frame_count = 0
        self._file = 1
        self._row = 0
        self.pointclouds = {}
        self.last_frame = None
        self.read_next_frames()

    def read_next_frames(self):
        self.pointclouds, self.last_frame = {}, None
        while len(self.pointclouds) < const.FB_READ_BUFFER_SIZE:
            path = os.path.join(self.experiment_path, "%d.csv" % self._file)
            if not os.path.isfile(path):
                break
            filled = False
            with open(path, newline="") as fh:
                for idx, row in enumerate(csv.reader(fh)):
                    if idx < self._row:
                        continue
                    num = int(row[0])
                    vals = [float(v) for v in row[1:6]] + [int(row[6])]
                    slot = self.pointclouds.setdefault(num, {k: [] for k in self.KEYS})
                    for k, v in zip(self.KEYS, vals):
                        slot[k].append(v)
                    self.last_frame = num
                    if len(self.pointclouds) >= const.FB_READ_BUFFER_SIZE:
                        self._row = idx + 1
                        filled = True
                        break
            if not filled:
                self._row = 0
                self._file += 1

    def get_data(self):
        self.frame_count += 1
        if self.last_frame is not None and self.frame_count > self.last_frame:
            self.read_next_frames()
        if self.frame_count in self.pointclouds:
            return True, self.frame_count, self.pointclouds[self.frame_count]
        return False, self.frame_count, None

    def is_finished(self) -> bool:
        return self.last_frame is None
