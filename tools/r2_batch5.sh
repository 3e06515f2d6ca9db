#!/bin/bash
set -x
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ha or 6 or hapke" 2>&1 | tail -25 > gpurun_out/r2e_tests_ha.log
cat gpurun_out/r2e_tests_ha.log
timeout 600 python bench.py --workload ha --steps 2 --warmup 1 --no-cpu --no-others > gpurun_out/r2e_bench_ha.json 2> gpurun_out/r2e_err.log
tail -3 gpurun_out/r2e_err.log
python -c "
import json; d=json.loads(open('gpurun_out/r2e_bench_ha.json').read().strip().splitlines()[-1])
print('ha', round(d['value']), round(d['e2e']['value']), d['roofline']['kernel_ms_per_step_all'])"
