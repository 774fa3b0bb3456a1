#!/usr/bin/env python
"""Opcode mix and stall-sample breakdown of a kernel from an ncu report's SASS source page:
   python profiles/opcode_mix.py rep.ncu-rep [kernel-substring]"""
import collections
import csv
import subprocess
import sys


def main(path, pick=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    i = 0
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name":
            name = rows[i][1]
            h = rows[i + 1]
            j = i + 2
            while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
                j += 1
            if pick is None or pick in name:
                report(name, h, rows[i + 2:j])
            i = j
        else:
            i += 1


def report(name, h, body):
    si, ii = h.index("Source"), h.index("Instructions Executed")
    stall_cols = [(k, c) for k, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    ops, stalls = collections.Counter(), collections.Counter()
    total = 0
    for r in body:
        if len(r) <= ii:
            continue
        toks = r[si].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        op = op.split(".")[0]
        n = int(float(r[ii] or 0))
        ops[op] += n
        total += n
        for k, c in stall_cols:
            stalls[c] += int(float(r[k] or 0))
    print("## %s" % name.split("(")[0])
    print("warp instructions executed: %d" % total)
    print()
    print("| opcode | warp insts | share |")
    print("|---|---|---|")
    for op, n in ops.most_common(24):
        print("| %s | %d | %.1f %% |" % (op, n, 100.0 * n / max(total, 1)))
    ssum = sum(stalls.values())
    print()
    print("| stall reason (all samples) | samples | share |")
    print("|---|---|---|")
    for c, n in stalls.most_common(10):
        print("| %s | %d | %.1f %% |" % (c, n, 100.0 * n / max(ssum, 1)))
    print()


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
