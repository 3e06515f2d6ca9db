#!/bin/bash
# Round 2, final batch: smoke, all GPU tests, the default bench line (with CPU arm and the other workloads), the
# reference arm, launch lists and ncu --set full of every kernel of the SW / LW / HA steps (raw pages as CSV).
set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
timeout 2400 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -12 > gpurun_out/r2_gpu_tests.log
timeout 1200 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null
for w in "sw 2048" "lw 32768" "ha 256"; do set -- $w
ncu --set full --clock-control none -k regex:"^k_(prologue|stage|layer|eval|nt|interp)" -c 12 \
    -o /tmp/prof_$1 python bench.py --workload $1 --steps 1 --warmup 0 --columns $2 --chunk $2 --no-cpu --no-others > gpurun_out/r2_ncu_$1.log 2>&1
ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/r2_raw_$1.csv
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_sw.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-others > /dev/null 2>&1
cat gpurun_out/r2_smoke.log gpurun_out/r2_gpu_tests.log
tail -3 gpurun_out/r2_bench_default.err
du -sh gpurun_out
