"""Offline replay driver with the same control flow as the reference's ``offline_main.py:40-62`` (minus GUI):
reads an experiment log, keeps dt from the log timestamps, runs normalize -> track -> estimate_posture per frame.
It exists so the drop-in classes can be exercised end to end where the reference's own script (which needs
PyQt5 / keras / wakepy) is not available; with those installed, the reference's offline_main.py drives the
same classes unchanged."""
from __future__ import annotations

import os
from typing import Callable, Optional

from . import Tracking, Utils


def offline_replay(experiment_path: str, model, on_frame: Optional[Callable] = None, first_dt: float = 0.1):
    if not os.path.exists(experiment_path):
        raise ValueError(f"No experiment file found in the path: {experiment_path}")
    sensor_data = Utils.OfflineManager(experiment_path)
    trackbuffer = Tracking.TrackBuffer()
    batch = Tracking.BatchedData()
    first = True
    frames = 0
    while not sensor_data.is_finished():
        ok, frame_no, detObj = sensor_data.get_data()
        if not ok:
            continue
        if first:
            trackbuffer.dt = first_dt
            first = False
        else:
            trackbuffer.dt = detObj["posix"][0] / 1000 - trackbuffer.t
        trackbuffer.t = detObj["posix"][0] / 1000
        effective_data = Utils.normalize_data(detObj)
        if effective_data.shape[0] != 0:
            trackbuffer.track(effective_data, batch)
            trackbuffer.estimate_posture(model)
        frames += 1
        if on_frame is not None:
            on_frame(frame_no, trackbuffer, batch, effective_data)
    return trackbuffer, frames
