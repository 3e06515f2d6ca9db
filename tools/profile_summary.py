#!/usr/bin/env python
"""Summarise ncu captures into profiles/ (markdown): launch-list shares and the
headline metrics of selected kernels.

    python tools/profile_summary.py <launches.csv> <prof.ncu-rep> [<prof2.ncu-rep> ...] > profiles/<name>.md
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    cols = rows[h]
    ki, vi = cols.index("Kernel Name"), cols.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) > vi:
            agg.setdefault(r[ki].split("(")[0], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print(f"## Launch list `{path}` (ncu --metrics gpu__time_duration.sum --clock-control none)\n")
    print("per-launch times are cold-cache and serialised: compare SHARES\n")
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        if sum(v) / tot > 0.0005:
            print(f"| `{k}` | {len(v)} | {sum(v) / 1e6:.3f} | {100 * sum(v) / tot:.1f}% |")
    print()


def details(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"## `{rep}` (ncu --set full --clock-control none --import-source on)\n")
    for r in rows[2:]:
        print(f"### `{r[hdr.index('Kernel Name')]}`\n")
        print("| metric | value | unit |\n|---|---|---|")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"| {m} | {r[i]} | {units[i]} |")
        print()


if __name__ == "__main__":
    launches(sys.argv[1])
    for rep in sys.argv[2:]:
        details(rep)
