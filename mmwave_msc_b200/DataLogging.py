"""Experiment logs on disk, in the layout of the reference's recorder (``DataLogging.py:47-89``): header-less CSV
files ``1.csv, 2.csv, ...`` inside one directory per experiment, ``FB_EXPERIMENT_FILE_SIZE`` frames per file, rows
buffered and appended in blocks.

  PointLogger    the recorder's own schema ``Frame,X,Y,Z,Doppler,Intensity,Timestamp`` (DataLogging.py:60-68) -- what
                 ``Utils.OfflineManager`` replays;
  ResultLogger   the inverse of the replay: what the perception path produced for a frame, one row per live track,
                 ``Frame,Scene,TrackId,NTracks,x[9],keypoints[57],FadeX,FadeZ,FadeSize,Timestamp`` -- the fields of a
                 packed result record (``mmw_pack_results`` / ``mmw_read_results_async``, include/mmw.h);
  read_results   reads such a log back.

Host-side I/O only (control plane of the path): no arithmetic, no device calls.
"""
from __future__ import annotations

import csv
import os
from typing import Dict, Iterator, List, Tuple

import numpy as np

from . import _lib
from . import constants as const


class _RotatingCsv:
    def __init__(self, data_path: str, frames_per_file=None, buffer_rows=None):
        self.path = data_path
        os.makedirs(data_path, exist_ok=True)
        self.per_file = int(frames_per_file or getattr(const, "FB_EXPERIMENT_FILE_SIZE", 200))
        self.buffer_rows = int(buffer_rows or getattr(const, "FB_WRITE_BUFFER_SIZE", 40))
        self.rows: List[str] = []
        self.file_index, self.frames_in_file = 1, 0

    def _flush(self):
        if self.rows:
            with open(os.path.join(self.path, "%d.csv" % self.file_index), "a", newline="") as fh:
                fh.write("".join(self.rows))
            self.rows = []

    def add_frame(self, lines: List[str]):
        """One frame's rows (DataLogging.py:70-87: flush on a full buffer or a full file, then rotate)."""
        self.rows.extend(lines)
        self.frames_in_file += 1
        if len(self.rows) >= self.buffer_rows or self.frames_in_file >= self.per_file:
            self._flush()
            if self.frames_in_file >= self.per_file:
                self.frames_in_file = 0
                self.file_index += 1

    def close(self):
        self._flush()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class PointLogger(_RotatingCsv):
    """Sensor frames in the recorder's schema (DataLogging.py:60-68)."""

    def log(self, frame_number: int, detObj: Dict, timestamp_ms: int):
        n = len(detObj["x"])
        self.add_frame(["%d,%r,%r,%r,%r,%r,%d\n" % (frame_number, float(detObj["x"][i]), float(detObj["y"][i]),
                                                     float(detObj["z"][i]), float(detObj["doppler"][i]),
                                                     float(detObj["peakVal"][i]), int(timestamp_ms)) for i in range(n)])


RESULT_COLUMNS = (["Frame", "Scene", "TrackId", "NTracks"] + ["x%d" % i for i in range(9)] +
                  ["kp%d" % i for i in range(57)] + ["FadeX", "FadeZ", "FadeSize", "Timestamp"])


class ResultLogger(_RotatingCsv):
    """Tracks and keypoints per frame.  ``log_packed`` takes the buffer ``BatchedTracker.read_results_async`` fills
    ([S * max_tracks, 72] float32: id, n_tracks, x[9], keypoints[57], fade square, pad); ``log_trackbuffer`` takes a
    drop-in ``Tracking.TrackBuffer`` (one scene)."""

    def __init__(self, data_path: str, max_tracks: int = 8, **kw):
        super().__init__(data_path, **kw)
        self.max_tracks = int(max_tracks)            # rows per scene of a packed buffer ([scene][track slot])

    def log_packed(self, frame_number: int, packed: np.ndarray, timestamp_ms: int):
        rec = np.asarray(packed, np.float32).reshape(-1, _lib.RESULT_FLOATS)
        lines = []
        for r in np.nonzero(rec[:, 0] >= 0)[0]:
            v = rec[r]
            lines.append("%d,%d,%d,%d,%s,%d\n" % (frame_number, r // self.max_tracks, int(v[0]), int(v[1]),
                                                  ",".join(repr(float(x)) for x in v[2:71]), int(timestamp_ms)))
        self.add_frame(lines)

    def log_trackbuffer(self, frame_number: int, trackbuffer, timestamp_ms: int, scene: int = 0):
        lines = []
        nt = len(trackbuffer.effective_tracks)
        for t in trackbuffer.effective_tracks:
            x = np.asarray(t.state.x, np.float64).reshape(9)
            kp = np.asarray(t.keypoints, np.float64).reshape(57)
            vals = [float(np.float32(v)) for v in x] + [float(np.float32(v)) for v in kp] + [0.0, 0.0, 0.0]
            lines.append("%d,%d,%d,%d,%s,%d\n" % (frame_number, scene, int(t.id), nt,
                                                  ",".join(repr(v) for v in vals), int(timestamp_ms)))
        self.add_frame(lines)


def read_results(data_path: str) -> Iterator[Tuple[int, np.ndarray]]:
    """Yields (frame number, rows [n, len(RESULT_COLUMNS)] float64) of a ResultLogger directory, in file order.
    Frames without a live track leave no rows (like frames without points in the recorder's logs)."""
    k = 1
    cur, rows = None, []
    while os.path.isfile(os.path.join(data_path, "%d.csv" % k)):
        with open(os.path.join(data_path, "%d.csv" % k), newline="") as fh:
            for row in csv.reader(fh):
                f = int(row[0])
                if cur is not None and f != cur:
                    yield cur, np.array(rows, np.float64)
                    rows = []
                cur = f
                rows.append([float(v) for v in row])
        k += 1
    if cur is not None:
        yield cur, np.array(rows, np.float64)
