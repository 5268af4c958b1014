#!/usr/bin/env python
"""Summarise ncu output into small text files under profiles/ (the .ncu-rep files themselves stay in gpurun_out/).

  python scripts/ncu_summary.py launches gpurun_out/<tag>_launches.csv  profiles/<tag>_launches.txt
  python scripts/ncu_summary.py full     gpurun_out/<tag>_<kernel>.ncu-rep profiles/<tag>_<kernel>.txt
"""
import collections
import csv
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_tmem.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed")


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
    total = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# source: {src}  (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)\n")
        f.write(f"# unit: {rows[1][ui]}; share = kernel total / all launches in the capture\n")
        f.write(f"{'kernel':90s} {'n':>4s} {'last':>14s} {'total':>14s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k[:90]:90s} {len(v):4d} {v[-1]:14.0f} {sum(v):14.0f} {100 * sum(v) / total:6.2f}%\n")


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# source: {src}  (ncu --set full --clock-control none --import-source on)\n")
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            f.write(f"## {name}\n")
            for i, h in enumerate(hdr):
                if h in KEEP or "stalled" in h and "per_issue_active" in h:
                    f.write(f"{h:80s} {vals[i]:>20s} {units[i]}\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
