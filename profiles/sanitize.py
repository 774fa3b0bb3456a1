#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): 16 scenes x 8 frames with pose in serial
mode (fused steps: the feature kernel beside dbscan_big_kernel, spawns, dense 2 on clusters), then 10 more frames through
mmw_run_frames_compact in throughput mode (int16 rows, grouped uploads, compact download into pinned memory).
Usage: compute-sanitizer --tool racecheck python profiles/sanitize.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mmwave_msc_b200 import _lib, synth, pose_weights as pw
from mmwave_msc_b200.batched import BatchedTracker, default_config

S, F, F2 = 16, 8, 10
RES = 0.0626
b = synth.gen_batch(range(S), F + F2)
i16 = []
for fb in b:
    p = fb.points.astype(np.float64)
    q = np.stack([np.rint(p[:, 0] * 512), np.rint(p[:, 1] * 512), np.rint(p[:, 2] * 512), np.rint(p[:, 3] / RES), p[:, 4]], axis=1)
    i16.append(np.ascontiguousarray(q.astype(np.int16)))
    fb.points = np.ascontiguousarray(np.stack([q[:, 0] / 512, q[:, 1] / 512, q[:, 2] / 512, q[:, 3], q[:, 4]], 1).astype(np.float32))
bt = BatchedTracker(S, config=default_config(doppler_res=RES, xyz_q_format=9))
bt.load_pose_weights(pw.make_pose_weights(pw.VARIANT_3D))
for f in range(F):
    bt.step(b[f].points, b[f].offsets, b[f].dt, pose=True, record_labels=True)
tr, nt = bt.tracks()
print("serial: tracks", int(nt.sum()), "launches", bt.launch_count())
n = S * bt.tcap * _lib.RESULT_FLOATS
rows = torch.from_numpy(np.concatenate(i16[F:])).pin_memory().numpy()
fro = np.cumsum([0] + [len(x) for x in i16[F:]]).astype(np.int64)
res = torch.zeros((F2, n), dtype=torch.float32).pin_memory().numpy()
cnt = torch.zeros(F2, dtype=torch.int32).pin_memory().numpy()
bt.run_frames(rows, fro, np.stack([x.offsets for x in b[F:]]), np.stack([x.dt for x in b[F:]]), res, n_records=cnt)
tr, nt = bt.tracks()
print("run_frames_compact: tracks", int(nt.sum()), "records per frame", cnt.tolist(), "launches", bt.launch_count())
