#!/usr/bin/env python3
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the few numbers DESIGN.md cites."""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second",
]
STALL = "smsp__pcsamp_warps_issue_stalled_"


def summarise(h, units, v):
    print("kernel:", v[h.index("Kernel Name")])
    for i, n in enumerate(h):
        if n in WANT:
            print(f"  {n:75s} {v[i]:>18s} {units[i]}")
    stalls = {n[len(STALL):]: int(float(v[i])) for i, n in enumerate(h) if n.startswith(STALL) and not n.endswith("_not_issued")}
    tot = sum(stalls.values()) or 1
    print("  warp-state samples (share):", ", ".join(f"{k} {100 * c / tot:.1f}%" for k, c in sorted(stalls.items(), key=lambda x: -x[1]) if c > 0.01 * tot))


for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    print(f"== {rep}")
    for v in rows[2:]:
        summarise(rows[0], rows[1], v)
