#!/usr/bin/env python
"""The other BASELINE configurations on one B200 (bench.py measures C2): device-resident inputs, CUDA events on the
library stream around K steps of new frames, L2 flushed between timed steps.
  C1  one scene (what a live sensor sees): per-frame latency
  C3  dense 8.5 m config: 1000 points/frame, 10 targets, TR_MAX_TRACKS = 10
  C5  8192 scenes on ONE GPU (the 8-GPU run shards them 1024 per GPU = C2)
Usage: python profiles/configs.py  ->  one JSON line per configuration"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmwave_msc_b200 import pose_weights as pw, synth
from mmwave_msc_b200.batched import BatchedTracker, default_config

PRIME, WARM, K = 12, 3, 20


def run(name, S, spec, cfg, max_points, max_tracks):
    batches = synth.gen_batch(range(S), PRIME + WARM + K, spec)
    bt = BatchedTracker(S, max_points=max_points, max_tracks=max_tracks, config=cfg)
    bt.load_pose_weights(pw.make_pose_weights(pw.VARIANT_3D))
    stream = torch.cuda.ExternalStream(bt.stream)
    dev = [(torch.from_numpy(b.points).cuda(), torch.from_numpy(b.offsets).cuda(), torch.from_numpy(b.dt).cuda())
           for b in batches]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ms = []
    for f, (p, o, d) in enumerate(dev):
        timed = f >= PRIME + WARM
        with torch.cuda.stream(stream):
            if timed:
                flush.fill_(f & 0xff)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
            bt.step_device(p.data_ptr(), o.data_ptr(), d.data_ptr(), p.shape[0], pose=True)
            if timed:
                e1.record(stream)
                e1.synchronize()
                ms.append(e0.elapsed_time(e1))
    bt.sync()
    _, nt = bt.tracks()
    ms = np.array(ms)
    print(json.dumps({"config": name, "scenes": S, "points_per_frame": int(np.mean([len(b.points) for b in batches]) / S),
                      "tracks_per_scene": float(nt.mean()), "ms_per_step_mean": float(ms.mean()),
                      "ms_per_step_p50": float(np.median(ms)), "scene_frames_per_s": float(S / (ms.mean() / 1e3))}),
          flush=True)
    bt.close()


if __name__ == "__main__":
    run("C1 single scene", 1, synth.SceneSpec(), default_config(), 256, 8)
    run("C3 dense, 64 scenes", 64, synth.SceneSpec.dense(), default_config(tr_max_tracks=10), 1024, 16)
    run("C3 dense, 1024 scenes", 1024, synth.SceneSpec.dense(), default_config(tr_max_tracks=10), 1024, 16)
    run("C5 8192 scenes on one GPU", 8192, synth.SceneSpec(), default_config(), 256, 8)
