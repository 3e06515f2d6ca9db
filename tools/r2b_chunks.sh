#!/bin/bash
# SW: device-resident and end-to-end rate against the columns per pydisort() call
for c in 4096 8192 16384; do
timeout 600 python bench.py --workload sw --steps 3 --warmup 3 --chunk $c --no-cpu --no-others > gpurun_out/q_chunk_$c.json 2>/dev/null
python - $c <<'PY'
import json, sys
d=json.loads(open(f"gpurun_out/q_chunk_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], "device %.4g e2e %.4g ratio %.3f compact %.4g" % (d["value"], d["e2e"]["value"], d["e2e"]["value"]/d["value"], d["e2e_compact_inputs"]["value"]))
PY
done
