#!/bin/bash
# Round 2, first GPU batch: tests, benches of the three ensembles, launch list, ncu capture of the new stage B kernel.
set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2a_gpu_tests.log
python bench.py --no-cpu > gpurun_out/r2a_bench_sw.json 2> gpurun_out/r2a_err1.log
python bench.py --workload ha --steps 2 --no-cpu > gpurun_out/r2a_bench_ha.json 2> gpurun_out/r2a_err2.log
python bench.py --workload lw --steps 3 --no-cpu > gpurun_out/r2a_bench_lw.json 2> gpurun_out/r2a_err3.log
ncu --set full --clock-control none --import-source on -k regex:"k_stage_b_add" -c 1 \
    -o gpurun_out/r2a_prof_sw python bench.py --steps 1 --warmup 0 --columns 2048 --chunk 2048 --no-cpu > gpurun_out/r2a_ncu_sw.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_stage_b_add|k_stage_a" -c 2 \
    -o gpurun_out/r2a_prof_ha python bench.py --workload ha --steps 1 --warmup 0 --columns 512 --chunk 512 --no-cpu > gpurun_out/r2a_ncu_ha.log 2>&1
cat gpurun_out/r2a_smoke.log gpurun_out/r2a_gpu_tests.log
for f in sw ha lw; do python -c "
import json; d=json.loads(open('gpurun_out/r2a_bench_${f}.json').read().strip().splitlines()[-1]); print('$f', round(d['value']), round(d['e2e']['value']), d['roofline']['kernel_ms_per_step_all'])"; done
du -sh gpurun_out
