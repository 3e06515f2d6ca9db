#!/usr/bin/env python
"""Test infrastructure (it runs the oracle): parity table for profiles/: the CUDA path (through the public API / C ABI) against the oracle on seeded ensemble
slices, per output field: worst  max|X - X_ref| / max|X_ref|  over the columns (the SURVEY 8(c) metric)."""
import multiprocessing as mp
import os
import sys
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))                   # tests/ (golden_io, parity_suite)
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "1")
import numpy as np  # noqa: E402


def _oracle(job):
    name, first, ncol = job
    warnings.simplefilter("ignore")
    from oracle import disort_oracle
    from pythonic_disort_b200 import synthetic
    ens = synthetic.make(name, ncol, first)
    return synthetic.run_reference_like(disort_oracle.pydisort, ens)


def main():
    import golden_io
    import parity_suite
    import pythonic_disort_b200 as pd
    from pythonic_disort_b200 import synthetic
    warnings.simplefilter("ignore")
    cores = os.cpu_count()
    print("| ensemble | columns | field | worst error / scale | worst pointwise relative (entries > 1e-3 of scale) |")
    print("|---|---|---|---|---|")
    with mp.get_context("spawn").Pool(cores) as pool:
        for name, ncol, first in (("sw", 256, 40000), ("lw", 1024, 300000), ("ha", 32, 2000), ("tp1", 6, 0), ("tp9c", 2, 0)):
            ens = synthetic.make(name, ncol, first)
            got = parity_suite.run_batched(pd.pydisort, ens)
            per = max(1, ncol // cores)
            jobs = [(name, first + lo, min(per, ncol - lo)) for lo in range(0, ncol, per)]
            parts = pool.map(_oracle, jobs)
            ref = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
            for key in got:
                worst = worst_pw = 0.0
                for b in range(ncol):
                    scale = None
                    if key.startswith("flux"):
                        scale = golden_io.group_scale([ref[k][b] for k in ("flux_up", "flux_down_diffuse", "flux_down_direct")])
                    err, pw, _ = golden_io.parity(got[key][b], ref[key][b], scale=scale)
                    worst, worst_pw = max(worst, err), max(worst_pw, pw)
                print(f"| {name} | {ncol} | {key} | {worst:.2e} | {worst_pw:.2e} |")


if __name__ == "__main__":
    main()
