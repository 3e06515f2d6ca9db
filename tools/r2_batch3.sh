#!/bin/bash
set -x
python bench.py --workload lw --steps 3 --no-cpu > gpurun_out/r2c_bench_lw.json 2> gpurun_out/r2c_err1.log
ncu --set full --clock-control none --import-source on -k regex:"k_stage_b_add" -c 1 \
    -o gpurun_out/r2c_prof_sw python bench.py --steps 1 --warmup 0 --columns 2048 --chunk 2048 --no-cpu --no-others > gpurun_out/r2c_ncu_sw.log 2>&1
tail -3 gpurun_out/r2c_ncu_sw.log
python -c "
import json; d=json.loads(open('gpurun_out/r2c_bench_lw.json').read().strip().splitlines()[-1])
print('lw', round(d['value']), round(d['e2e']['value']), d['roofline']['kernel_ms_per_step_all'])"
