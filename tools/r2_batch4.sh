#!/bin/bash
# Round 2, batch 4: hardened parity tests on the GPU; ncu --set full of every kernel of the SW / LW / HA steps
# (raw pages exported as CSV on the box; the .ncu-rep files are too big to bring back).
set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2d_smoke.log 2>&1
timeout 2400 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -40 > gpurun_out/r2d_gpu_tests.log
for w in "sw 2048" "lw 32768" "ha 256"; do set -- $w
ncu --set full --clock-control none -k regex:"^k_" -c 12 \
    -o /tmp/prof_$1 python bench.py --workload $1 --steps 1 --warmup 0 --columns $2 --chunk $2 --no-cpu --no-others > gpurun_out/r2d_ncu_$1.log 2>&1
ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/r2d_raw_$1.csv
done
cat gpurun_out/r2d_smoke.log gpurun_out/r2d_gpu_tests.log
du -sh gpurun_out
