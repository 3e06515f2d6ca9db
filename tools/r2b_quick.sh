#!/bin/bash
# quick GPU check of a change: the tests nearest to it, then bench lines of the three ensembles (no CPU arm);
# optional argument: variant libraries (exp/lib_*.so) to run the SW bench on as well
timeout 900 python -m pytest tests -m gpu -q -x -k "interface or described or ensembles_vs_reference_golden or generic_kernels or batched" 2>&1 | grep -E "^E  |passed|failed" | head > gpurun_out/q_tests.log
for w in sw lw ha; do
timeout 600 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu --no-others > gpurun_out/q_bench_$w.json 2> gpurun_out/q_bench_$w.err
done
for v in "$@"; do
PD_LIB_PATH=$PWD/exp/lib_$v.so timeout 600 python bench.py --workload sw --steps 2 --warmup 3 --no-cpu --no-others > gpurun_out/q_bench_sw_$v.json 2> gpurun_out/q_bench_sw_$v.err
done
cat gpurun_out/q_tests.log
python - "$@" <<'PY'
import json, sys
for w in ["sw","lw","ha"] + ["sw_" + v for v in sys.argv[1:]]:
    try:
        d=json.loads(open(f"gpurun_out/q_bench_{w}.json").read().strip().splitlines()[-1])
        c = d.get("e2e_compact_inputs") or {}
        print(w, "%.4g %.4g" % (d["value"], d["e2e"]["value"]), "compact %s h2d %s" % (c.get("value"), c.get("h2d_bytes_per_step")), "%.2f ms" % d["ms_per_step"], {k: round(v, 2) for k, v in d["roofline"]["kernel_ms_per_step_all"].items()})
    except Exception as e:
        print(w, "failed", e)
PY
