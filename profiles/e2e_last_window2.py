#!/usr/bin/env python
"""Where inside the sequence the end-to-end loop slows down: run_frames over sub-blocks of 4 frames."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mmwave_msc_b200 import _lib, pose_weights as pw, synth  # noqa: E402
from mmwave_msc_b200.batched import BatchedTracker, default_config  # noqa: E402

S, K, REPS, SUB = 1024, 20, 6, 4
batches = synth.gen_batch(range(S), 12 + 5 + REPS * K)
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


mode = os.environ.get("MODE", "big")          # big: 6 blocks of 20 frames, executed in sub-calls of 4 frames
warm = block(0, 17)
blks = [block(17 + r * K, 17 + (r + 1) * K) for r in range(REPS)]
if os.environ.get("EXTRA"):
    extra = block(17, 17 + K)                 # one more pinned block that is never used
bt.run_frames(*warm)
out = []
for rows, fro, offs, dts, res in blks:
    for j in range(0, K, SUB):
        sub = (rows[fro[j]:fro[j + SUB]], (fro[j:j + SUB + 1] - fro[j]).astype(np.int64), offs[j:j + SUB], dts[j:j + SUB], res[j:j + SUB])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bt.run_frames(*sub)
        torch.cuda.synchronize()
        out.append(round((time.perf_counter() - t0) / SUB * 1e6))
print("us per frame, sub-blocks of %d frames:" % SUB, out)
