#!/bin/bash
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
timeout 2400 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -12 > gpurun_out/r2_gpu_tests.log
timeout 1500 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
cat gpurun_out/r2_smoke.log gpurun_out/r2_gpu_tests.log; tail -3 gpurun_out/r2_bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1])
print("sw", round(d['value']), round(d['e2e']['value']), d['ms_per_step'], round(d['cpu_baseline']['value'],1), d['roofline']['kernel_ms_per_step_all'])
for k,v in d['other_workloads'].items():
    print(k, round(v['value']), round(v['e2e']['value']), v['ms_per_step'], round(v['cpu_baseline']['value'],1), v['roofline']['kernel'], v['roofline']['kernel_ms_per_step_all'])
PY
