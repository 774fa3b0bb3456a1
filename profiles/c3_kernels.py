#!/usr/bin/env python
"""Kernel split of the dense config C3 (1024 scenes = 256 generated x 4, 1000 points/frame, 10 targets,
TR_MAX_TRACKS = 10): CUDA events around every launch (mmw_profile), plus how many scenes per frame reach
dbscan_big_kernel and how large their fused clouds are."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from mmwave_msc_b200 import pose_weights as pw, synth
from mmwave_msc_b200.batched import BatchedTracker, default_config

gen = synth.gen_batch(range(200_000, 200_256), 12 + 3 + 8, synth.SceneSpec.dense())
tiled = bench._tile_batches(gen, 4)
bt = BatchedTracker(1024, max_points=1024, max_tracks=16, config=default_config(tr_max_tracks=10))
bt.load_pose_weights(pw.make_pose_weights(pw.VARIANT_3D))
for pts, off, dt in tiled[:15]:
    bt.step(pts, off, dt, pose=True)
bt.sync()
k = bt.profile_kernels(lambda i: bt.step(*tiled[15 + i], pose=True), 8)
print({n: round(v * 1e3, 1) for n, v in k.items()}, "us per launch; sum %.0f us" % (1e3 * sum(k.values())))
lab, nf = bt.labels()
rc = bt.ring_counts()
_, nt = bt.tracks()
fused = np.where(rc > 0, rc, 0).sum(1)
print("tracks/scene %.2f; scenes whose DBSCAN ran last frame: %d of 1024; fused cloud sizes p50 %d p90 %d max %d" % (
    nt.mean(), int((nf >= 0).sum()), np.median(nf[nf >= 0]) if (nf >= 0).any() else 0,
    np.percentile(nf[nf >= 0], 90) if (nf >= 0).any() else 0, nf.max()))
dbg = np.zeros(16, np.uint64)
