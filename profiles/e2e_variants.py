#!/usr/bin/env python
"""End-to-end variants of the C2 workload on one GPU (pinned host buffers in, packed records out):
python loop vs mmw_run_frames, fp32 rows vs int16 rows, serial vs throughput mode.  Wall clock over K frames."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mmwave_msc_b200 import _lib, pose_weights as pw, synth
from mmwave_msc_b200.batched import BatchedTracker, default_config
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

S, PRIME, K = 1024, 12, 40
batches = synth.gen_batch(range(S), PRIME + 2 * K)
f32, i16 = zip(*[bench.lattice_rows(b.points) for b in batches])
W = pw.make_pose_weights(pw.VARIANT_3D)
n = S * 8 * _lib.RESULT_FLOATS

def pin(a):
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()

for name, rows, loop, pipe in (("python loop, fp32 rows, serial", f32, "py", False), ("python loop, fp32 rows, pipeline", f32, "py", True),
                               ("python loop, int16 rows, pipeline", i16, "py", True), ("run_frames, fp32 rows, pipeline", f32, "c", True),
                               ("run_frames, int16 rows, pipeline", i16, "c", True), ("run_frames, int16 rows, serial", i16, "c", False),
                               ("run_frames, int16 rows, pipeline, frames 4 KiB aligned", i16, "ca", True),
                               ("run_frames, fp32 rows, pipeline, frames 4 KiB aligned", f32, "ca", True)):
    bt = BatchedTracker(S, config=default_config(doppler_res=bench.DOPPLER_RES, xyz_q_format=9))
    bt.load_pose_weights(W)
    for f in range(PRIME):
        bt.step(rows[f], batches[f].offsets, batches[f].dt, pose=True)
    bt.sync()
    for rep in range(2):                     # rep 0 = warm-up
        lo = PRIME + rep * K
        if loop == "py":
            pr = [(pin(rows[f]), pin(batches[f].offsets), pin(batches[f].dt)) for f in range(lo, lo + K)]
            res = [pin(np.zeros(n, np.float32)) for _ in range(2)]
            torch.cuda.synchronize(); t0 = time.perf_counter()
            prev = None
            for i, (p, o, d) in enumerate(pr):
                bt.step(p, o, d, pose=True, pipeline=pipe)
                slot = bt.read_results_async(res[i & 1])
                if prev is not None:
                    bt.wait_results(prev)
                prev = slot
            bt.wait_results(prev)
            dtm = time.perf_counter() - t0
        elif loop == "ca":
            # every frame's rows start on a 4 KiB boundary of the pinned buffer (2048 rows of 10 / 20 bytes = a multiple)
            starts, pos = [], 0
            for r in rows[lo:lo + K]:
                starts.append(pos); pos = (pos + len(r) + 2047) // 2048 * 2048
            big = np.zeros((pos, 5), rows[0].dtype)
            for st, r in zip(starts, rows[lo:lo + K]):
                big[st:st + len(r)] = r
            R = pin(big); fro = np.array(starts + [pos], np.int64)
            O = pin(np.stack([b.offsets for b in batches[lo:lo + K]])); D = pin(np.stack([b.dt for b in batches[lo:lo + K]]))
            res = pin(np.zeros((K, n), np.float32))
            torch.cuda.synchronize(); t0 = time.perf_counter()
            bt.run_frames(R, fro, O, D, res, pose=True, pipeline=pipe)
            dtm = time.perf_counter() - t0
        else:
            R = pin(np.concatenate(rows[lo:lo + K])); fro = np.cumsum([0] + [len(r) for r in rows[lo:lo + K]]).astype(np.int64)
            O = pin(np.stack([b.offsets for b in batches[lo:lo + K]])); D = pin(np.stack([b.dt for b in batches[lo:lo + K]]))
            res = pin(np.zeros((K, n), np.float32))
            torch.cuda.synchronize(); t0 = time.perf_counter()
            bt.run_frames(R, fro, O, D, res, pose=True, pipeline=pipe)
            dtm = time.perf_counter() - t0
    print("%-40s %.3f ms/frame  %.2f M scene-frames/s" % (name, 1e3 * dtm / K, S * K / dtm / 1e6), flush=True)
    bt.close()
