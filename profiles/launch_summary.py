#!/usr/bin/env python
"""Markdown summary of an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of bench.py:
python profiles/launch_summary.py launches.csv "title" > profiles/rNN_launches.md"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h, data = rows[hi], rows[hi + 1:]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
seq = [(r[ki], float(r[vi].replace(",", "")) / (1e3 if r[ui] == "ns" else 1.0)) for r in data if len(r) > vi]
# timed steps = the (at most 7) launches that follow each torch fill kernel (the L2 flush of a device-timed step)
fills = [i for i, (k, _) in enumerate(seq) if "FillFunctor" in k]
timed, allk = collections.defaultdict(list), collections.defaultdict(list)
for a in fills:
    j = a + 1
    while j < len(seq) and "FillFunctor" not in seq[j][0] and j - a <= 7:
        timed[seq[j][0]].append(seq[j][1])
        j += 1
for k, t in seq:
    allk[k].append(t)
tot = sum(sum(v) / len(v) for v in timed.values())
print("# %s\n" % (sys.argv[2] if len(sys.argv) > 2 else "ncu launch list"))
print("(cold-cache, serialised per-launch times: compare SHARES with bench.py's kernel_ms, not absolutes; raw list: %s."
      % sys.argv[1].split("/")[-1])
print("12 priming frames (every scene's first cloud goes through dbscan_big_kernel), 3 warm-up, 5 device-timed steps -- the")
print("launches that follow each L2-flush fill kernel, column 'timed steps' -- then the end-to-end, latency and per-kernel passes.)\n")
print("| kernel | launches | mean us (all) | mean us (timed steps) | share of a timed step |")
print("|---|---|---|---|---|")
mean = lambda v: sum(v) / len(v)
for k, v in sorted(allk.items(), key=lambda kv: -(mean(timed[kv[0]]) if kv[0] in timed else 0)):
    t = timed.get(k)
    print("| `%s` | %d | %.1f | %s | %s |" % (k[:90], len(v), mean(v), ("%.1f" % mean(t)) if t else "-",
                                           ("%.1f%%" % (100 * mean(t) / tot)) if t else "-"))
