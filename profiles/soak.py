#!/usr/bin/env python
"""Long full-size consistency run: C2 (1024 scenes) x NF frames.  (a) frame by frame in serial mode (fp32 rows, full
records read back), (b) mmw_run_frames_compact in throughput mode (int16 rows, grouped uploads, compact records) on a
second context.  The row count crosses the tile-width edges of dense 1 on the way (176 -> 192 -> 256 wide tiles), tracks
time out and slots are re-used.  Every frame's records and the final track state must agree bit for bit; no scene may
raise an overflow flag."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mmwave_msc_b200 import _lib, pose_weights as pw, synth  # noqa: E402
from mmwave_msc_b200.batched import BatchedTracker, default_config  # noqa: E402

S, NF = 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 360
t0 = time.time()
batches = synth.gen_batch(range(S), NF)
f32, i16 = zip(*[bench.lattice_rows(b.points) for b in batches])
print("generated %d frames in %.0f s" % (NF, time.time() - t0), flush=True)
cfg = default_config(doppler_res=bench.DOPPLER_RES, xyz_q_format=9)
W = pw.make_pose_weights(pw.VARIANT_3D)
n = S * 8 * _lib.RESULT_FLOATS

a = BatchedTracker(S, config=cfg)
a.load_pose_weights(W)
ra = torch.zeros((NF, n), dtype=torch.float32).pin_memory().numpy()
rows_a = []
for f in range(NF):
    a.step(f32[f], batches[f].offsets, batches[f].dt, pose=True)
    a.wait_results(a.read_results_async(ra[f]))
    if f % 60 == 59:
        rows_a.append(int(a.counters(reset=True)[7] / 60))
ta, na = a.tracks()
print("serial: pose rows per frame (means of 60 frames):", rows_a, "| tracks at the end %d | status flags %d" % (na.sum(), int(a.status().any())))

b = BatchedTracker(S, config=cfg)
b.load_pose_weights(W)
R = torch.from_numpy(np.concatenate(i16)).pin_memory().numpy()
fro = np.cumsum([0] + [len(r) for r in i16]).astype(np.int64)
O = torch.from_numpy(np.stack([x.offsets for x in batches])).pin_memory().numpy()
D = torch.from_numpy(np.stack([x.dt for x in batches])).pin_memory().numpy()
rb = torch.zeros((NF, n), dtype=torch.float32).pin_memory().numpy()
cnt = torch.zeros(NF, dtype=torch.int32).pin_memory().numpy()
t0 = time.perf_counter()
b.run_frames(R, fro, O, D, rb, n_records=cnt)
dt = time.perf_counter() - t0
tb, nb = b.tracks()
bad = 0
for f in range(NF):
    if b.expand_compact_results(rb[f], int(cnt[f])).tobytes() != ra[f].tobytes():
        bad += 1
print("run_frames_compact (throughput mode): %.2f M scene-frames/s end to end over %d frames; records per frame %d -> %d" % (
    S * NF / dt / 1e6, NF, cnt[0], cnt[-1]))
print("frames whose records differ from the serial run: %d of %d; final track state identical: %s; status flags %d" % (
    bad, NF, ta.tobytes() == tb.tobytes() and np.array_equal(na, nb), int(b.status().any())))
assert bad == 0 and ta.tobytes() == tb.tobytes()
