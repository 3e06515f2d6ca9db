#!/bin/bash
for v in r200 t128m2 t64m4; do
  export PD_LIB_PATH=$PWD/gpurun_in_$v.so
  python bench.py --workload ha --columns 2048 --chunk 1024 --steps 2 --warmup 1 --no-cpu --no-others 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['value']), d['roofline']['kernel_ms_per_step_all'])"
done
export PD_LIB_PATH=$PWD/gpurun_in_t128m2.so
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ha or hapke" 2>&1 | tail -5
