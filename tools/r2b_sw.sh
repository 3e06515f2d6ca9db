#!/bin/bash
# SW only: parity tests nearest to the solver, then the bench line
timeout 900 python -m pytest tests -m gpu -q -x -k "ensembles_vs_reference_golden or generic_kernels or batched or live_oracle" 2>&1 | grep -E "^E  |passed|failed" | head
timeout 600 python bench.py --workload sw --steps 2 --warmup 3 --no-cpu --no-others > gpurun_out/q_bench_sw.json 2> gpurun_out/q_bench_sw.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/q_bench_sw.json").read().strip().splitlines()[-1])
print("sw %.4g %.4g" % (d["value"], d["e2e"]["value"]), "%.2f ms" % d["ms_per_step"], {k: round(v, 2) for k, v in d["roofline"]["kernel_ms_per_step_all"].items()})
PY
