#!/bin/bash
# compute-sanitizer over small solves through every kernel of the build (tools/sanitize_run.py)
timeout 1500 compute-sanitizer --tool memcheck python tools/sanitize_run.py > gpurun_out/r2_memcheck.log 2>&1
tail -3 gpurun_out/r2_memcheck.log
timeout 2400 compute-sanitizer --tool racecheck python tools/sanitize_run.py > gpurun_out/r2_racecheck.log 2>&1
tail -3 gpurun_out/r2_racecheck.log
timeout 1500 compute-sanitizer --tool synccheck python tools/sanitize_run.py > gpurun_out/r2_synccheck.log 2>&1
tail -3 gpurun_out/r2_synccheck.log
