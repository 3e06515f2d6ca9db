#!/usr/bin/env python
"""Per-source-line stall-reason breakdown of one kernel in an .ncu-rep (see ncu_lines.py).

    python tools/ncu_stalls.py report.ncu-rep <kernel-name-regex> [top]
"""
import csv
import io
import subprocess
import sys

REASONS = ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_no_inst", "stall_mio", "stall_math", "stall_branch_resolving",
           "stall_lg", "stall_dispatch", "stall_not_selected", "stall_selected"]


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur, hdr, lines = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
            idx = [hdr.index(x) for x in REASONS]
            ii = hdr.index("Instructions Executed")
        elif hdr and r[0].isdigit():
            try:
                lines.append((cur, int(r[0]), float(r[ii]), [float(r[i] or 0) for i in idx], r[1].strip()))
            except (ValueError, IndexError):
                pass
    tot = [sum(x[3][k] for x in lines) for k in range(len(REASONS))]
    print("totals: " + "  ".join(f"{n[6:]}={t:.0f}" for n, t in zip(REASONS, tot)))
    for k, name in enumerate(REASONS[:5]):
        print(f"--- top lines for {name}")
        for f, ln, ins, st, txt in sorted(lines, key=lambda x: -x[3][k])[:top]:
            if st[k] <= 0:
                break
            print(f"{100 * st[k] / (tot[k] or 1):5.1f}%  {f}:{ln:<4d} {txt[:90]}")


if __name__ == "__main__":
    main()
