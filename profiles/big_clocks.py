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
for name, k in (("ring load", 0), ("adjacency + counts", 3), ("components (mask frontier)", 5), ("border points", 6),
                ("labels+spawn+write-back", 1)):
    print("%-26s %8.0f cycles per deferred scene" % (name, o[k] / n))
print("total %.0f cycles = %.1f us at 1.965 GHz per deferred scene" % ((o[0] + o[1] + o[3] + o[4] + o[5] + o[6]) / n,
      (o[0] + o[1] + o[3] + o[4] + o[5] + o[6]) / n / 1965))
