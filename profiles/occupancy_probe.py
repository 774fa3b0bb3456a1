#!/usr/bin/env python
"""How does the tracker step scale with resident CTAs per SM?  Same light workload (80-120 points, 1-2 people,
1024 scenes) with two shared-memory footprints (N_cap 256/T_cap 8 vs N_cap 128/T_cap 4); run it against the
default build and a register-capped build (MMW_LIB=...) to vary the register limit as well."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmwave_msc_b200 import synth
from mmwave_msc_b200.batched import BatchedTracker

S, F = 1024, 40
spec = synth.SceneSpec(pts_min=80, pts_max=120, people_min=1, people_max=2)
b = synth.gen_batch(range(S), F, spec)
for ncap, tcap in ((256, 8), (128, 4)):
    bt = BatchedTracker(S, max_points=ncap, max_tracks=tcap)
    for f in range(20):
        bt.step(b[f].points, b[f].offsets, b[f].dt, pose=False)
    k = bt.profile_kernels(lambda i: bt.step(b[20 + i].points, b[20 + i].offsets, b[20 + i].dt, pose=False), 20)
    print("lib=%s ncap=%d tcap=%d step=%.1f us overflow=%d" % (os.path.basename(os.environ.get("MMW_LIB", "default")),
                                                              ncap, tcap, 1e3 * k["step"], int((bt.status() != 0).sum())))
