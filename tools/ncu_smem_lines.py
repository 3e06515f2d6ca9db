#!/usr/bin/env python
"""Shared-memory wavefronts per source line (ideal / excessive = bank conflicts) of one kernel in an .ncu-rep captured
with --import-source on:   python tools/ncu_smem_lines.py report.ncu-rep <kernel-name-regex> [top]"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    hdr, cur, lines = None, None, {}
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] in ("File Path", "File Name"):
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
            iw, ie, ii = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Excessive"), hdr.index("Instructions Executed")
        elif hdr and r[0].isdigit() and len(r) == len(hdr):
            def f(x):
                try:
                    return float(x)
                except ValueError:
                    return 0.0
            key = (cur, int(r[0]))
            w, e, n, src = lines.get(key, (0.0, 0.0, 0.0, r[1].strip()))
            lines[key] = (w + f(r[iw]), e + f(r[ie]), n + f(r[ii]), src)
    tot = sum(v[0] for v in lines.values())
    exc = sum(v[1] for v in lines.values())
    print(f"shared wavefronts {tot:.4g}, excessive (bank conflicts) {exc:.4g} = {100 * exc / max(tot, 1):.1f} %")
    for (fn, ln), (w, e, n, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * w / tot:5.1f}% wavefronts  {100 * e / max(exc, 1):5.1f}% of conflicts  {w / max(n, 1):4.1f}/inst  {fn}:{ln}  {src[:80]}")


if __name__ == "__main__":
    main()
