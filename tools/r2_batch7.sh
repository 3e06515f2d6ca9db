#!/bin/bash
for v in base m6; do
  if [ $v = base ]; then unset PD_LIB_PATH; else export PD_LIB_PATH=$PWD/gpurun_in_$v.so; fi
  python bench.py --workload lw --columns 262144 --chunk 131072 --steps 3 --warmup 1 --no-cpu --no-others 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['value']), d['roofline']['kernel_ms_per_step_all'])"
done
