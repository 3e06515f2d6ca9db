#!/usr/bin/env python
"""Small solves through every production kernel, for compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py

Covers SW (N = 8: symmetric stage A, tensor-core stage B, u + tabulated NT, interpolate), LW (N = 4: three-row
register stage B, thermal source) and HA (N = 16: general stage A, tensor-core stage B, BDRF surface)."""
import os
import sys
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import pythonic_disort_b200 as pd  # noqa: E402
from pythonic_disort_b200 import synthetic  # noqa: E402

warnings.simplefilter("ignore")
for name, ncol in (("sw", 6), ("lw", 20), ("ha", 2)):
    ens = synthetic.make(name, ncol)
    out = pd.pydisort(*ens["args"], **ens["kwargs"])
    t = ens["tau_eval"]
    Fp = out[1](t)
    Fm = out[2](t)
    u0 = out[3](t)
    chk = float(np.sum(Fp)) + float(np.sum(Fm[0])) + float(np.sum(u0))
    if "u" in ens["outputs"]:
        u = out[4](t, ens["phi_eval"])
        um = pd.subroutines.interpolate(out[4])(np.array([0.3, -0.3]), t, ens["phi_eval"])
        chk += float(np.sum(u)) + float(np.sum(um))
    assert np.isfinite(chk), name
    print(name, "ok", chk)
print("sanitizer run finished")
