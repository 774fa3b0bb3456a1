#!/usr/bin/env python
"""Summarise an ncu report (raw page) into a small markdown table:  python profiles/summarize_ncu.py rep.ncu-rep"""
import csv
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("smsp__inst_executed.sum", "warp_insts"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    ki = h.index("Kernel Name")
    cols = [(h.index(m), n) for m, n in METRICS if m in h]
    print("| kernel | " + " | ".join("%s [%s]" % (n, units[i]) if units[i] else n for i, n in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for r in rows[2:]:
        print("| %s | " % r[ki].split("(")[0][:40] + " | ".join(r[i] for i, _ in cols) + " |")


def traffic(path):
    """{kernel: dram bytes per launch} (mean over the captured launches) as JSON on stdout."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    ki, ri, wi = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    acc = {}
    for r in rows[2:]:
        name = r[ki]
        key = ("step_kernel" if "step_kernel" in name else "gemm_tc_kernel_dense1" if "gemm_tc_kernel<192" in name or
               "gemm_tc_kernel<256" in name or "gemm_tc_pair_kernel" in name else "gemm_tc_kernel_dense2" if "gemm_tc_kernel<64" in name else
               "conv_slab_conv1" if "conv_slab_kernel<16" in name else "conv_slab_conv2" if "conv_slab_kernel<32" in name
               else name.split("(")[0])
        b = float(r[ri]) * scale.get(units[ri], 1.0) + float(r[wi]) * scale.get(units[wi], 1.0)
        acc.setdefault(key, []).append(b)
    res = {k: sum(v) / len(v) for k, v in acc.items()}
    if "smsp__inst_executed.sum" in h:                 # warp instructions of the tracker step (bench.py issue roofline)
        ii = h.index("smsp__inst_executed.sum")
        wi = [float(r[ii]) for r in rows[2:] if "step_kernel" in r[ki]]
        if wi:
            res["step_kernel_warp_insts"] = sum(wi) / len(wi)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "--traffic":
        traffic(sys.argv[2])
    else:
        main(sys.argv[1])
