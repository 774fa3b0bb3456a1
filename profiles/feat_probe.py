#!/usr/bin/env python
"""Times pose_feature_kernel alone (C2 state after 15 frames, L2 flushed before every launch, CUDA events on the
library stream).  MMW_FEAT_DBG selects what the kernel leaves out (pose.cuh)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmwave_msc_b200 import pose_weights as pw, synth  # noqa: E402
from mmwave_msc_b200.batched import BatchedTracker  # noqa: E402

S = 1024
bt = BatchedTracker(S, max_points=256, max_tracks=8, device=0)
bt.load_pose_weights(pw.make_pose_weights(pw.VARIANT_3D))
for b in synth.gen_batch(range(S), 15):
    bt.step(b.points, b.offsets, b.dt, pose=True)
bt.sync()
stream = torch.cuda.ExternalStream(bt.stream, device=0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ms = []
for i in range(24):
    with torch.cuda.stream(stream):
        flush.fill_(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        bt.pose_features_only()
        e1.record(stream)
    e1.synchronize()
    ms.append(e0.elapsed_time(e1) * 1e3)
warm = []
for i in range(24):
    with torch.cuda.stream(stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        bt.pose_features_only()
        e1.record(stream)
    e1.synchronize()
    warm.append(e0.elapsed_time(e1) * 1e3)
print("MMW_FEAT_DBG=%s  flushed L2: p50 %.1f us  min %.1f | warm L2: p50 %.1f us min %.1f" % (
    os.environ.get("MMW_FEAT_DBG", "0"), np.median(ms[4:]), min(ms[4:]), np.median(warm[4:]), min(warm[4:])))
