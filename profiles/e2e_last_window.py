#!/usr/bin/env python
"""Why is the last end-to-end window of bench.py slow?  Same run_frames windows, executed in allocation order, then
all again (second use of every buffer), then in reverse order."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mmwave_msc_b200 import _lib, pose_weights as pw, synth  # noqa: E402
from mmwave_msc_b200.batched import BatchedTracker, default_config  # noqa: E402

S, K, REPS = 1024, 20, 6
TRAIL = int(os.environ.get("TRAIL", "0"))         # frames generated after the last window
batches = synth.gen_batch(range(S), 12 + 5 + REPS * K + TRAIL)
rows16 = []
for b in batches:
    b.points, r16 = bench.lattice_rows(b.points)
    rows16.append(r16)
bt = BatchedTracker(S, max_points=256, max_tracks=8, device=0, config=default_config(doppler_res=bench.DOPPLER_RES, xyz_q_format=9))
bt.load_pose_weights(pw.make_pose_weights(pw.VARIANT_3D))
per_frame = S * bt.tcap * _lib.RESULT_FLOATS


def block(lo, hi):
    rows = torch.from_numpy(np.concatenate(rows16[lo:hi])).pin_memory()
    fro = np.cumsum([0] + [len(r) for r in rows16[lo:hi]]).astype(np.int64)
    offs = torch.from_numpy(np.stack([b.offsets for b in batches[lo:hi]])).pin_memory()
    dts = torch.from_numpy(np.stack([b.dt for b in batches[lo:hi]])).pin_memory()
    res = torch.empty((hi - lo, per_frame), dtype=torch.float32).pin_memory()
    return rows.numpy(), fro, offs.numpy(), dts.numpy(), res.numpy()


warm = block(0, 17)
blks = [block(17 + r * K, 17 + (r + 1) * K) for r in range(REPS)]
bt.run_frames(*warm)


def run(order, tag):
    out = []
    for i in order:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bt.run_frames(*blks[i])
        torch.cuda.synchronize()
        out.append(round(S * K / (time.perf_counter() - t0) / 1e6, 2))
        if os.environ.get("TT"):
            t = torch.tensor([1.0], dtype=torch.float64, device="cuda")
            t.item()
        c = bt.counters(reset=True)
        print("   window", i, "per step: U %.0f Bf %.0f tracks %.0f" % (c[3] / K, c[4] / K, c[5] / K))
    print(tag, order, out)


run(list(range(REPS)), "allocation order     ")
_, nt = bt.tracks()
print("tracks/scene after the windows: %.2f" % nt.mean(), "counters", bt.counters(reset=True))
