#!/usr/bin/env python
"""Host <-> device copy bandwidth from pinned memory at the sizes the end-to-end loop uses, and the box's NUMA layout."""
import os
import subprocess
import time

import torch

for cmd in (["nvidia-smi", "topo", "-m"], ["numactl", "-H"], ["lscpu"]):
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout
        keep = [l for l in out.splitlines() if cmd[0] != "lscpu" or any(k in l for k in ("NUMA", "Socket", "Model name", "CPU(s):"))]
        print("$", " ".join(cmd)); print("\n".join(keep[:24]))
    except Exception as e:
        print(cmd, "unavailable:", e)
print("affinity:", sorted(os.sched_getaffinity(0))[:8], "... n =", len(os.sched_getaffinity(0)))
torch.cuda.set_device(0)
for mb in (2, 8, 64, 512):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, src, dst in (("H2D", h, d), ("D2H", d, h)):
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        reps = max(4, 256 // mb)
        t0 = time.perf_counter()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        print("%s %4d MB pinned: %.1f GB/s  (%.0f us per copy)" % (name, mb, n / dt / 1e9, dt * 1e6))
# both directions at once (two streams), 2 MB each, like the replay loop
h1, h2 = torch.empty(2 << 20, dtype=torch.uint8).pin_memory(), torch.empty(2 << 20, dtype=torch.uint8).pin_memory()
d1, d2 = torch.empty(2 << 20, dtype=torch.uint8, device="cuda"), torch.empty(2 << 20, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    with torch.cuda.stream(s1):
        d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 200
print("H2D 2 MB + D2H 2 MB concurrently: %.0f us per pair (%.1f GB/s each way)" % (dt * 1e6, (2 << 20) / dt / 1e9))
