#!/usr/bin/env python
"""Where the end-to-end step time goes: the C2 workload of bench.py stepped four ways (inputs resident / uploaded from
pinned memory) x (results left on the device / packed and downloaded), with the host time spent inside each API call.
Diagnostic only -- not a bench value."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mmwave_msc_b200 import _lib, pose_weights as pw, synth
from mmwave_msc_b200.batched import BatchedTracker

S, PRIME, W, K = 1024, 12, 5, 40
batches = synth.gen_batch(list(range(S)), PRIME + 4 * (W + K))
bt = BatchedTracker(S, max_points=256, max_tracks=8, device=0)
bt.load_pose_weights(pw.make_pose_weights(pw.VARIANT_3D))
bt.set_dense_path(True)
dev = [(torch.from_numpy(b.points).cuda(), torch.from_numpy(b.offsets).cuda(), torch.from_numpy(b.dt).cuda()) for b in batches]
pin = [(torch.from_numpy(b.points).pin_memory().numpy(), torch.from_numpy(b.offsets).pin_memory().numpy(),
        torch.from_numpy(b.dt).pin_memory().numpy()) for b in batches]
res = [torch.empty(S * bt.tcap * _lib.RESULT_FLOATS, dtype=torch.float32).pin_memory().numpy() for _ in range(2)]
f = 0
for _ in range(PRIME):
    p, o, d = dev[f]; bt.step_device(p.data_ptr(), o.data_ptr(), d.data_ptr(), p.shape[0]); f += 1
bt.sync()

def run(host_in, read, lo, hi, acc=None):
    prev = None
    for i in range(lo, hi):
        t0 = time.perf_counter()
        if host_in:
            bt.step(*pin[i], pose=True)
        else:
            p, o, d = dev[i]; bt.step_device(p.data_ptr(), o.data_ptr(), d.data_ptr(), p.shape[0])
        t1 = time.perf_counter()
        if read:
            slot = bt.read_results_async(res[i & 1])
            t2 = time.perf_counter()
            if prev is not None:
                bt.wait_results(prev)
            prev = slot
            t3 = time.perf_counter()
        else:
            t2 = t3 = t1
        if acc is not None:
            acc.append((t1 - t0, t2 - t1, t3 - t2))
    if prev is not None:
        bt.wait_results(prev)

out = {}
for name, host_in, read in (("device_in/no_read", False, False), ("device_in/read", False, True),
                            ("host_in/no_read", True, False), ("host_in/read", True, True)):
    run(host_in, read, f, f + W); f += W
    bt.sync(); torch.cuda.synchronize()
    acc = []
    t0 = time.perf_counter()
    run(host_in, read, f, f + K, acc); f += K
    bt.sync()
    dt_ = time.perf_counter() - t0
    a = np.array(acc) * 1e3
    out[name] = {"ms_per_step": dt_ / K * 1e3, "host_ms_in_step_call": float(np.median(a[:, 0])),
                 "host_ms_in_read_call": float(np.median(a[:, 1])), "host_ms_in_wait": float(np.median(a[:, 2]))}
    print(name, json.dumps(out[name]))
