#!/bin/bash
set -x
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2b_gpu_tests.log
timeout 900 python bench.py --cpu-seconds 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_err1.log
ncu --set full --clock-control none --import-source on -k regex:"k_stage_b_add" -c 1 \
    -o gpurun_out/r2b_prof_sw python bench.py --steps 1 --warmup 0 --columns 2048 --chunk 2048 --no-cpu --no-others > gpurun_out/r2b_ncu_sw.log 2>&1
cat gpurun_out/r2b_gpu_tests.log
tail -5 gpurun_out/r2b_err1.log
python -c "
import json; d=json.loads(open('gpurun_out/r2b_bench.json').read().strip().splitlines()[-1])
print('sw', round(d['value']), round(d['e2e']['value']), d['roofline']['kernel_ms_per_step_all'], d['cpu_baseline'])
for k,v in d.get('other_workloads',{}).items(): print(k, round(v['value']), round(v['e2e']['value']), v['roofline']['kernel_ms_per_step_all'], v['cpu_baseline']['value'])
"
du -sh gpurun_out
