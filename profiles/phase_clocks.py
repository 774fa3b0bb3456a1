#!/usr/bin/env python
"""Cycles per phase of the fused tracker step (thread 0 of every CTA, mmw_phase_clocks), C2 workload.
First with the tracker alone (state stays in L2 between steps), then with the pose network between steps
(its ~300 MB of scratch traffic evicts the tracker state, like in the real pipeline)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mmwave_msc_b200 import synth, pose_weights as pw
from mmwave_msc_b200.batched import BatchedTracker

S, F = 1024, 40
b = synth.gen_batch(range(S), F)
names = ["load (bulk copies)", "predict", "gate matrices", "transform+filter", "gate+counts", "lists+pushes",
         "statistics", "assoc results", "maintain", "update", "screen", "writeback"]
for pose in (False, True):
    bt = BatchedTracker(S)
    if pose:
        bt.load_pose_weights(pw.make_pose_weights(pw.VARIANT_3D))
    for f in range(20):
        bt.step(b[f].points, b[f].offsets, b[f].dt, pose=pose)
    bt.sync(); bt.phase_clocks(True)
    for f in range(20, 40):
        bt.step(b[f].points, b[f].offsets, b[f].dt, pose=pose)
    pc = bt.phase_clocks(False).astype(float)
    tot = pc[1:13].sum()
    print("== pose between steps: %s" % pose)
    for i, n in enumerate(names):
        print("%-20s %8.0f cyc/scene-frame %5.1f%%" % (n, pc[i + 1] / (S * 20), 100 * pc[i + 1] / tot))
    print("total cycles per scene-frame %.0f" % (tot / (S * 20)))

# distribution of per-scene cycles in one frame (the kernel ends when the slowest scene does)
import ctypes
from mmwave_msc_b200 import _lib
bt = BatchedTracker(S)
for f in range(25):
    bt.step(b[f].points, b[f].offsets, b[f].dt, pose=False)
bt.sync(); bt.phase_clocks(True)
bt.step(b[25].points, b[25].offsets, b[25].dt, pose=False)
raw = np.zeros(3 * S, np.uint64)
_lib.check(bt.lib.mmw_scene_cycles(bt._h, _lib.ptr(raw)))
bt.phase_clocks(False)
cyc = raw[:S].astype(float)
t0 = raw[S:2 * S].astype(np.int64); t1 = raw[2 * S:].astype(np.int64)
base = t0.min()
print("CTA start offsets [us]: p50 %.1f p90 %.1f max %.1f | CTA end offsets [us]: p50 %.1f p90 %.1f max %.1f | CTA duration [us]: p50 %.1f max %.1f" % (
    np.percentile(t0 - base, 50) / 1e3, np.percentile(t0 - base, 90) / 1e3, (t0 - base).max() / 1e3,
    np.percentile(t1 - base, 50) / 1e3, np.percentile(t1 - base, 90) / 1e3, (t1 - base).max() / 1e3,
    np.percentile(t1 - t0, 50) / 1e3, (t1 - t0).max() / 1e3))
n, nid, m = bt.summary()
rc = bt.ring_counts(); fused = np.where(rc > 0, rc, 0).sum(1)
print("per-scene cycles: mean %.0f  p50 %.0f  p90 %.0f  p99 %.0f  max %.0f" % (cyc.mean(), np.percentile(cyc, 50),
      np.percentile(cyc, 90), np.percentile(cyc, 99), cyc.max()))
worst = np.argsort(-cyc)[:8]
for w in worst:
    print("  scene %4d cycles %7.0f tracks %d ring(after) %s M %d" % (w, cyc[w], n[w], rc[w].tolist(), m[w]))

# phase breakdown of the slowest scene alone (same frames replayed in a 1-scene context)
w = int(worst[0])
b1 = synth.gen_batch([w], 26)
bt1 = BatchedTracker(1)
for f in range(25):
    bt1.step(b1[f].points, b1[f].offsets, b1[f].dt, pose=False)
bt1.sync(); bt1.phase_clocks(True)
bt1.step(b1[25].points, b1[25].offsets, b1[25].dt, pose=False, record_labels=True)
pc = bt1.phase_clocks(False).astype(float)
lab, nf = bt1.labels()
print("slowest scene %d: fused points %d, clusters %d" % (w, nf[0], lab[0, :max(nf[0], 0)].max() + 1 if nf[0] > 0 else 0))
for i, nme in enumerate(names):
    print("   %-20s %8.0f" % (nme, pc[i + 1]))
