#!/bin/bash
timeout 2400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sw or tp1 or suite" 2>&1 | tail -3
python bench.py --workload sw --steps 3 --warmup 3 --no-cpu --no-others 2>gpurun_out/r2g_err_sw.log | tee gpurun_out/r2g_bench_sw.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('sw', round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d['roofline']['kernel_ms_per_step_all'])"
tail -2 gpurun_out/r2g_err_sw.log
ncu --set full --clock-control none --import-source on -k regex:"k_stage_a_sym" -c 1 -o gpurun_out/r2_prof_sw_a2 python bench.py --steps 1 --warmup 0 --columns 2048 --chunk 2048 --no-cpu --no-others > gpurun_out/r2_ncu_sw_a2.log 2>&1
