#!/bin/bash
# quick GPU check of a change: the new tests, then bench lines of the three ensembles (no CPU arm)
timeout 900 python -m pytest tests -m gpu -q -x -k "interface or ensembles_vs_reference_golden or generic_kernels or smoke" 2>&1 | tail -5 > gpurun_out/q_tests.log
for w in sw lw ha; do
timeout 600 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu --no-others > gpurun_out/q_bench_$w.json 2> gpurun_out/q_bench_$w.err
done
cat gpurun_out/q_tests.log
python - <<'PY'
import json
for w in ("sw","lw","ha"):
    try:
        d=json.loads(open(f"gpurun_out/q_bench_{w}.json").read().strip().splitlines()[-1])
        print(w, d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step_all"])
    except Exception as e:
        print(w, "failed", e)
PY
