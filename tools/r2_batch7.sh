#!/bin/bash
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for w in sw lw ha; do
python bench.py --workload $w --steps 3 --warmup 2 --no-cpu --no-others 2>gpurun_out/r2g_err_$w.log | tee gpurun_out/r2g_bench_$w.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d['roofline']['kernel_ms_per_step_all'])"
done
