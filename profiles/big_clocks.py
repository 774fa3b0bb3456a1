#!/usr/bin/env python
"""Where dbscan_big_kernel spends its time (C2 workload): cycles of thread 0 per part, per deferred scene."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mmwave_msc_b200 import synth, _lib
from mmwave_msc_b200.batched import BatchedTracker

S, F = 1024, 40
b = synth.gen_batch(range(S), F)
bt = BatchedTracker(S)
for f in range(20):
    bt.step(b[f].points, b[f].offsets, b[f].dt, pose=False)
bt.sync(); bt.phase_clocks(True)
out = np.zeros(16, np.uint64)
_lib.check(bt.lib.mmw_dbscan_big_clocks(bt._h, _lib.ptr(out)))
for f in range(20, 40):
    bt.step(b[f].points, b[f].offsets, b[f].dt, pose=False)
_lib.check(bt.lib.mmw_dbscan_big_clocks(bt._h, _lib.ptr(out)))
bt.phase_clocks(False)
o = out.astype(float)
n = max(o[2], 1)
print("deferred scenes in 20 frames: %d (%.1f per frame), mean fused points %.0f" % (o[2], o[2] / 20, o[7] / n))
for name, k in (("ring load", 0), ("counts", 3), ("unions", 4), ("rank+relabel", 5), ("border sweep", 6),
                ("labels+spawn+write-back", 1)):
    print("%-26s %8.0f cycles per deferred scene" % (name, o[k] / n))
print("  inside rank+relabel: first hop %.0f, single-blob check %.0f, propagation loop %.0f cycles; %d of %d scenes took "
      "the single-blob shortcut, %.1f propagation passes per scene" % (o[8] / n, o[9] / n, o[10] / n, o[12], o[2], o[13] / n))
print("total %.0f cycles = %.1f us at 1.965 GHz per deferred scene" % ((o[0] + o[1] + o[3] + o[4] + o[5] + o[6]) / n,
      (o[0] + o[1] + o[3] + o[4] + o[5] + o[6]) / n / 1965))
