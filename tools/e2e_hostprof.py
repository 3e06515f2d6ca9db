#!/usr/bin/env python
"""cProfile of the host side of one ensemble.solve_ensemble(wait=False) call: which Python / torch / ctypes calls the
host spends (or is blocked for) the time of a step in.  python tools/e2e_hostprof.py [workload] [columns] [chunk]"""
import cProfile
import os
import pstats
import sys
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pythonic_disort_b200 import ensemble, synthetic  # noqa: E402

warnings.simplefilter("ignore")
name = sys.argv[1] if len(sys.argv) > 1 else "sw"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
ens = synthetic.make(name, B)
if os.environ.get("PD_ONLY_FLUX"):   # the flux-only variant of a workload (bench.py: sw_flux)
    ens["kwargs"]["only_flux"] = True
    ens["kwargs"].pop("NT_cor", None)
    ens["outputs"] = ("flux",)
pin = lambda x: ensemble.pinned(x) if hasattr(x, "shape") and getattr(x, "ndim", 0) > 0 else x  # noqa: E731
args = [pin(a) for a in ens["args"]]
kw = {k: ([pin(m) for m in v] if k == "BDRF_Fourier_modes" else pin(v)) for k, v in ens["kwargs"].items()}
tau = pin(ens["tau_eval"])
outputs = ("flux_up", "flux_down") + (("u",) if "u" in ens["outputs"] else ())
res = None
for _ in range(3):
    res = ensemble.solve_ensemble(*args, tau=tau, phi=ens["phi_eval"], outputs=outputs, chunk=chunk, out=res, **kw)
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
r = ensemble.solve_ensemble(*args, tau=tau, phi=ens["phi_eval"], outputs=outputs, chunk=chunk, out=res, wait=False, **kw)
pr.disable()
t1 = time.perf_counter()
r.wait()
t2 = time.perf_counter()
print(f"enqueue {1e3 * (t1 - t0):.1f} ms, wait {1e3 * (t2 - t1):.1f} ms")
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
