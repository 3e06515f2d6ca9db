#!/usr/bin/env python
"""Per-source-line instruction / stall summary of one kernel in an .ncu-rep
(needs a report captured with --import-source on and code built with -lineinfo).

    python tools/ncu_lines.py report.ncu-rep <kernel-name-regex> [top]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, hdr, lines = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
            ii, si, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), 1
        elif hdr and r[0].isdigit():
            try:
                lines.append((cur_file, int(r[0]), float(r[ii]), float(r[si]), r[ti].strip()))
            except (ValueError, IndexError):
                pass
    tot_i = sum(x[2] for x in lines) or 1
    tot_s = sum(x[3] for x in lines) or 1
    print(f"total warp instructions {tot_i:.4g}, stall samples {tot_s:.4g}")
    for f, ln, ins, smp, txt in sorted(lines, key=lambda x: -x[2])[:top]:
        print(f"{100 * ins / tot_i:5.1f}% inst {100 * smp / tot_s:5.1f}% stall  {f}:{ln:<4d} {txt[:100]}")


if __name__ == "__main__":
    main()
