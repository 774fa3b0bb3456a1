#!/usr/bin/env python
"""BASELINE config C4: MARS-style pose regression alone -- 65,536 per-track 8x8x5 feature maps -> 19 keypoints
(2-D net, define_CNN).  Times the tensor-core pose kernels with CUDA events on the library stream (mmw_profile)
and checks a sample of rows against the float64 oracle.  Usage: python profiles/pose_c4.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmwave_msc_b200 import _lib, pose_weights as pw
from mmwave_msc_b200.batched import BatchedTracker, default_config
from oracle import mmw_oracle as mo

N = 65536
rng = np.random.default_rng(4)
feats = np.zeros((N, 64, 5), np.float32)
k = rng.integers(20, 65, size=N)
for i in range(N):
    r = np.zeros((64, 5), np.float32)
    r[:k[i], :3] = rng.normal([0, 0, 1.0], [0.15, 0.15, 0.4], size=(k[i], 3))
    r[:k[i], 3] = rng.normal(0, 0.3, k[i])
    r[:k[i], 4] = (np.floor(rng.gamma(0.5, 54.0, k[i])) + 1 - 27.0187) / 70.351
    feats[i] = r[np.argsort(r[:, 0], kind="stable")]
feats = feats.reshape(N, 8, 8, 5)
W = pw.make_pose_weights(pw.VARIANT_2D)
bt = BatchedTracker(8192, max_tracks=8, config=default_config(frames_batch=0))
bt.load_pose_weights(W, pw.VARIANT_2D)
out = bt.pose(feats)                       # warm-up + correctness
ref = mo.pose_forward(W, feats[:512], np.float64)
err = float(np.abs(out[:512] - ref).max())
_lib.check(bt.lib.mmw_profile(bt._h, 1))
REP = 10
for _ in range(REP):
    bt.pose(feats)
ms = np.zeros(8); calls = np.zeros(8, np.uint64)
_lib.check(bt.lib.mmw_get_kernel_ms(bt._h, _lib.ptr(ms), _lib.ptr(calls)))
per = {n: float(ms[i] / calls[i]) for i, n in enumerate(_lib.KERNEL_NAMES) if calls[i]}
total = sum(per.values())
flops = 2837504.0 * N
print(json.dumps({"config": "C4", "rows": N, "kernel_ms": per, "pose_ms": total, "maps_per_s": N / (total / 1e3),
                  "algorithmic_tflops": flops / (total / 1e3) / 1e12, "max_joint_err_m_vs_fp64_oracle": err}))
