#!/bin/bash
# End-of-round measurement batch for profiles/ (run under gpurun from the repo root; outputs in gpurun_out/, < 64 MiB).
set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r1_gpu_tests.log
python bench.py > gpurun_out/r1_bench_sw_1gpu.json 2> gpurun_out/err1.log
python bench.py --workload ha --steps 2 > gpurun_out/r1_bench_ha_1gpu.json 2> gpurun_out/err2.log
python bench.py --workload lw --steps 3 > gpurun_out/r1_bench_lw_1gpu.json 2> gpurun_out/err3.log
python bench.py --workload sw_flux --steps 3 > gpurun_out/r1_bench_swflux_1gpu.json 2> gpurun_out/err4.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_bench_reference_arm.json 2> gpurun_out/err5.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_final.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_stage_a_sym|k_stage_b_mma|k_eval_u|k_nt_tab" -c 4 \
    -o gpurun_out/r1_prof_final python bench.py --steps 1 --warmup 0 --columns 2048 --chunk 2048 --no-cpu > gpurun_out/ncu_f.log 2>&1
cat gpurun_out/r1_smoke.log gpurun_out/r1_gpu_tests.log
for f in sw ha lw swflux; do python -c "
import json; d=json.loads(open('gpurun_out/r1_bench_${f}_1gpu.json').read().strip().splitlines()[-1]); print('$f', round(d['value']), round(d['e2e']['value']), round(d['roofline']['whole_path']['frac_of_fp64_peak'],4), round(d['roofline']['frac'],4), d['cpu_baseline']['value'], d['roofline']['kernel_ms_per_step_all'])"; done
du -sh gpurun_out
