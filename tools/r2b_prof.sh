#!/bin/bash
# ncu --set full with source of the SW step's three heavy kernels (one 2,048-column chunk); the failing test's message
ncu --set full --import-source on --clock-control none -k regex:"^k_(stage_a_sym|stage_b_add|layer_ops)" -c 3 \
    -o gpurun_out/r2b_prof_sw python bench.py --workload sw --steps 1 --warmup 0 --columns 2048 --chunk 2048 --no-cpu --no-others > gpurun_out/r2b_ncu_sw.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -x -k "interface" 2>&1 | grep -E "^E  |passed|failed" | head -20
