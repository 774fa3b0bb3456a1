#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): 16 scenes x 8 frames with pose,
plus one spawn-heavy first frame.  Usage: compute-sanitizer --tool racecheck python profiles/sanitize.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mmwave_msc_b200 import synth, pose_weights as pw
from mmwave_msc_b200.batched import BatchedTracker

S, F = 16, 8
b = synth.gen_batch(range(S), F)
bt = BatchedTracker(S)
bt.load_pose_weights(pw.make_pose_weights(pw.VARIANT_3D))
for f in range(F):
    bt.step(b[f].points, b[f].offsets, b[f].dt, pose=True, record_labels=True)
tr, nt = bt.tracks()
print("tracks", int(nt.sum()), "launches", bt.launch_count())
