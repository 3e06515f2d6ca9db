#!/bin/bash
for v in base ls8m6 ls8m8; do
  if [ $v = base ]; then unset PD_LIB_PATH; else export PD_LIB_PATH=$PWD/gpurun_in_$v.so; fi
  python bench.py --workload sw --columns 16384 --chunk 8192 --steps 2 --warmup 1 --no-cpu --no-others 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['value']), d['roofline']['kernel_ms_per_step_all'])"
done
export PD_LIB_PATH=$PWD/gpurun_in_ls8m8.so
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sw or tp or suite" 2>&1 | tail -3
