#!/usr/bin/env python
"""Summarise ncu output brought back in gpurun_out/ into small text files for profiles/.

    python tools/ncu_summary.py launches <launches.csv>            # per-kernel time shares
    python tools/ncu_summary.py full <report.ncu-rep> [kernel-re]  # key metrics of one capture
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_issued.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        k = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1.0)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':58s} {'launches':>8s} {'total ms':>10s} {'avg ms':>9s} {'share':>7s}")
    for k, a in agg.items():
        print(f"{k:58s} {a[0]:8d} {a[1]:10.3f} {a[1] / a[0]:9.3f} {a[1] / tot * 100:6.1f}%")
    print(f"{'all':58s} {sum(a[0] for a in agg.values()):8d} {tot:10.3f}")


def full(path, kre=None):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    H, U = rows[0], rows[1]
    ki = H.index("Kernel Name")
    for r in rows[2:]:
        if kre and not re.search(kre, r[ki]):
            continue
        print("kernel:", r[ki][:100])
        for i, h in enumerate(H):
            if h in KEYS or ("issue_stalled" in h and h.endswith("per_warp_active.pct")):
                print(f"  {h:88s} {r[i]:>16s} {U[i]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
