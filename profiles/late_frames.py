#!/usr/bin/env python
"""How the C2 workload and the kernel times drift over a long sequence (tracks accumulate slowly): every 30 frames the
per-step counters and the per-kernel CUDA-event times of 5 steps."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmwave_msc_b200 import pose_weights as pw, synth  # noqa: E402
from mmwave_msc_b200.batched import BatchedTracker  # noqa: E402

S, NF = 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 340
bt = BatchedTracker(S, max_points=256, max_tracks=8, device=0)
bt.load_pose_weights(pw.make_pose_weights(pw.VARIANT_3D))
batches = synth.gen_batch(range(S), NF)
dev = [(torch.from_numpy(b.points).cuda(), torch.from_numpy(b.offsets).cuda(), torch.from_numpy(b.dt).cuda()) for b in batches]
names = ["scene_frames", "N", "M", "U", "Bf", "tracks", "ring_rows", "pose_rows"]


def step(f, pipeline=False):
    p, o, d = dev[f]
    bt.step_device(p.data_ptr(), o.data_ptr(), d.data_ptr(), p.shape[0], pose=True, pipeline=pipeline)


stream = torch.cuda.ExternalStream(bt.stream, device=0)
res = torch.empty(S * bt.tcap * 72, dtype=torch.float32, device="cuda")


f = 0
while f + 30 <= NF:
    for _ in range(5):
        step(f); f += 1
    bt.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(20):                      # throughput mode: 4 untimed, 16 timed steps
        if i == 4:
            e0.record(stream)
        step(f, pipeline=True); f += 1
    bt.pack_results(res.data_ptr())
    e1.record(stream)
    bt.sync()
    tp_us = e0.elapsed_time(e1) / 16 * 1e3
    bt.counters(reset=True)
    k = bt.profile_kernels(lambda i: step(f + i), 5)
    f += 5
    bt.sync()
    c = bt.counters(reset=True)
    _, nt = bt.tracks()
    print("frame %3d  tracks/scene %.2f (max %d)  pose_rows %.0f  Bf %.0f  | us: %s  sum %.0f | throughput mode %.0f us/step" % (
        f, nt.mean(), nt.max(), c[7] / 5, c[4] / 5, {n: round(v * 1e3, 1) for n, v in k.items()}, sum(k.values()) * 1e3, tp_us))
