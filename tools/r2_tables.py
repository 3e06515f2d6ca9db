#!/usr/bin/env python
"""Markdown tables of profiles/r2_summary.md from the raw ncu pages and the launch list:

    python tools/r2_tables.py kernels  sw:raw_sw.csv lw:raw_lw.csv ha:raw_ha.csv     one row per kernel of a chunk
    python tools/r2_tables.py shares   launches.csv                                   launch-list shares of a step
    python tools/r2_tables.py traffic  profiles/r2_traffic.json                       DRAM bytes per column
"""
import collections
import csv
import json
import re
import sys

COLS = [("gpu__time_duration.sum", "ms", 3), ("launch__registers_per_thread", "regs", 0),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %", 0),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 0),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %", 0),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 0)]
STALLS = ["short_scoreboard", "long_scoreboard", "wait", "mio_throttle", "no_instruction"]


def short(name):
    name = name.replace("void ", "")
    m = re.match(r"(\w+)(<[^>]*>)?", name)
    return m.group(1) + (m.group(2) or "")


def kernels(specs):
    print("| workload | kernel | " + " | ".join(c[1] for c in COLS) + " | stalls per issue: short / long / wait / mio / no-inst |")
    print("|---" * (len(COLS) + 3) + "|")
    for spec in specs:
        w, path = spec.split(":", 1)
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        seen = 0
        for r in rows[2:]:
            name = short(r[hdr.index("Kernel Name")])
            if name.startswith("k_fp64_probe"):
                continue
            if name == "k_prologue":
                seen += 1
                if seen > 1:
                    break
            vals = []
            for key, _, nd in COLS:
                v = float(r[hdr.index(key)].replace(",", ""))
                if key == "gpu__time_duration.sum":
                    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[hdr.index(key)], 1.0)
                vals.append(f"{v:.{nd}f}")
            st = " / ".join(f"{float(r[hdr.index('smsp__average_warps_issue_stalled_' + s + '_per_issue_active.ratio')]):.2f}"
                            for s in STALLS)
            print(f"| {w} | `{name}` | " + " | ".join(vals) + f" | {st} |")


def shares(path):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    cols = rows[h]
    ki, vi, ui = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) > vi:
            scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
            name = short(r[ki])
            if not name.startswith("k_"):
                name = "PyTorch fills / copies between the calls"
            agg.setdefault(name, []).append(float(r[vi].replace(",", "")) * scale)
    agg.pop("k_fp64_probe", None)
    tot = sum(sum(v) for v in agg.values())
    print("| kernel | launches | total ms | share of kernel time |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"| `{k}` | {len(v)} | {sum(v):.2f} | {100 * sum(v) / tot:.1f} % |")


def traffic(path):
    d = json.load(open(path))
    print("| workload | kernel: bytes | total |\n|---|---|---|")
    for w in ("sw", "lw", "ha"):
        per = {k: v for k, v in d[w].items() if not k.startswith("_") and v > 1000}
        print(f"| {w} | " + ", ".join(f"`{k}` {v / 1e6:.3f} MB" for k, v in sorted(per.items())) + f" | {sum(per.values()) / 1e6:.2f} MB |")


if __name__ == "__main__":
    {"kernels": lambda: kernels(sys.argv[2:]), "shares": lambda: shares(sys.argv[2]), "traffic": lambda: traffic(sys.argv[2])}[sys.argv[1]]()
