#!/usr/bin/env python
"""Phase timeline of a few CTAs of the dense-1 pair kernel (MMW_GEMM_DBG=4 -> device printf), C2 workload."""
import os, sys
os.environ["MMW_GEMM_DBG"] = os.environ.get("MMW_GEMM_DBG", "4")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmwave_msc_b200 import synth, pose_weights as pw
from mmwave_msc_b200.batched import BatchedTracker

S, F = 1024, 16
b = synth.gen_batch(range(S), F)
bt = BatchedTracker(S)
bt.load_pose_weights(pw.make_pose_weights(pw.VARIANT_3D))
for f in range(F - 2):
    bt.step(b[f].points, b[f].offsets, b[f].dt, pose=False)
bt.sync()
print("--- two steps with pose ---", flush=True)
for f in range(F - 2, F):
    bt.step(b[f].points, b[f].offsets, b[f].dt, pose=True)
    bt.sync()
    print("--- step done ---", flush=True)
