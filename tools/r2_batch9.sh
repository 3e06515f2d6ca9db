#!/bin/bash
timeout 1500 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
tail -3 gpurun_out/r2_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/r2_bench_reference_arm.json').read().strip().splitlines()[-1])
print("same config:", d['config']==r['config'], r['value'])
print("sw", round(d['value']), round(d['e2e']['value']), d['ms_per_step'], round(d['cpu_baseline']['value'],1), d['config']['chunk_columns'], d['roofline']['kernel_ms_per_step_all'])
for k,v in d['other_workloads'].items():
    print(k, round(v['value']), round(v['e2e']['value']), v['ms_per_step'], round(v['cpu_baseline']['value'],1), v['roofline']['kernel'], v['config']['chunk_columns'])
PY
