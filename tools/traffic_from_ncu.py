#!/usr/bin/env python
"""profiles/r2_traffic.json from ncu captures:  python tools/traffic_from_ncu.py out.json workload:columns:rep.ncu-rep|raw.csv ...

Every capture is `ncu --set full` over `bench.py --workload W --columns C --chunk C` (one launch = C columns), so
DRAM bytes per column = (dram__bytes_read.sum + dram__bytes_write.sum) / C, per kernel (template arguments dropped;
a kernel launched more than once per chunk -- the fallback passes -- is summed).  bench.py multiplies this by the
chunk size of the timed run for `roofline.traffic`."""
import csv
import io
import json
import re
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    out_path, specs = sys.argv[1], sys.argv[2:]
    result = {}
    for spec in specs:
        workload, cols, rep = spec.split(":", 2)
        if rep.endswith(".csv"):   # the raw page exported on the GPU box (ncu -i rep --page raw --csv)
            raw = open(rep).read()
        else:
            raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        ir, iw, inm, it = (hdr.index(k) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "Kernel Name",
                                                   "gpu__time_duration.sum"))
        per = result.setdefault(workload, {})
        meta = per.setdefault("_capture", {"columns": int(cols), "kernels": {}})
        prologues = 0
        for r in rows[2:]:
            name = re.sub(r"[<(].*", "", r[inm].replace("void ", "")).strip()
            if name == "k_fp64_probe":   # bench.py's peak measurement, not part of the path
                continue
            if name == "k_prologue":     # one chunk = the launches from one prologue to the next
                prologues += 1
                if prologues > 1:
                    break
            b = float(r[ir].replace(",", "")) * UNIT[units[ir]] + float(r[iw].replace(",", "")) * UNIT[units[iw]]
            per[name] = per.get(name, 0.0) + b / int(cols)
            k = meta["kernels"].setdefault(name, {"launches": 0, "dram_bytes": 0.0, "time_" + units[it]: 0.0})
            k["launches"] += 1
            k["dram_bytes"] += b
            k["time_" + units[it]] += float(r[it].replace(",", ""))
    json.dump(result, open(out_path, "w"), indent=1, sort_keys=True)
    for w, per in result.items():
        tot = sum(v for k, v in per.items() if not k.startswith("_"))
        print(w, {k: round(v) for k, v in per.items() if not k.startswith("_")}, "total bytes/column", round(tot))


if __name__ == "__main__":
    main()
