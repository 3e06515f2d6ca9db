#!/usr/bin/env python
"""Small solves through every production kernel, for compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py

Covers the input helpers (Planck band integrals, s_poly coefficients, Hapke Fourier modes), SW (N = 8: one-thread
symmetric eigen stage, lane-group elimination over interface radiances, u + tabulated NT, interpolate), LW (N = 4:
one-thread-per-system boundary-condition stage, thermal source), HA (N = 16: eight-lane symmetric eigen stage with its
shuffles and shared-memory hand-offs, 16-lane boundary-condition stage, BDRF surface) and the size-generic kernels
(general eigen stage, pivoted band solver) through the PD_FLAG_GENERIC_KERNELS test bit."""
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
    # levels between the interfaces: the output functions assemble G (C * exp) there instead of reading the sweep's
    # interface radiances (a grid mixing both kinds sends some warps of k_eval_flux down each path)
    tm = np.sort(np.concatenate([t, 0.5 * (t[:, :-1] + t[:, 1:])[:, ::2]], axis=1), axis=1)
    chk += float(np.sum(out[1](tm))) + float(np.sum(out[3](tm)))
    if "u" in ens["outputs"]:
        chk += float(np.sum(out[4](tm, ens["phi_eval"])))
    # the same ensemble described per layer (pd_hg_moments, pd_level_source)
    args, kw = list(ens["args"]), dict(ens["kwargs"])
    for nm, obj in ens["compact"].items():
        if nm == "Leg_coeffs_all":
            args[3] = obj
        else:
            kw[nm] = obj
    chk += float(np.sum(pd.pydisort(*args, **kw)[1](t)))
    assert np.isfinite(chk), name
    print(name, "ok", chk)
from pythonic_disort_b200 import _lib  # noqa: E402

for name, ncol in (("sw", 2), ("tp9c", 1)):
    ens = synthetic.make(name, ncol)
    out = pd.pydisort(*ens["args"], _kernel_flags=_lib.PD_FLAG_GENERIC_KERNELS, **ens["kwargs"])
    chk = float(np.sum(out[1](ens["tau_eval"]))) + float(np.sum(out[4](ens["tau_eval"], ens["phi_eval"])))
    assert np.isfinite(chk), name
    print(name, "generic kernels ok", chk)
import torch  # noqa: E402

T = torch.linspace(0.0, 320.0, 257, device="cuda", dtype=torch.float64)
em = pd.subroutines.blackbody_contrib_to_BCs(T, 600.0, 700.0)
tau = torch.cumsum(torch.rand(5, 12, device="cuda", dtype=torch.float64) + 0.1, dim=1)
tem = torch.linspace(200.0, 300.0, 13, device="cuda", dtype=torch.float64).repeat(5, 1)
sp = pd.subroutines.generate_s_poly_coeffs(tau, tem, 10.0, 3000.0)
modes = pd.subroutines.hapke_BDRF_Fourier_modes(8, 16, torch.tensor([0.3, 0.6, 0.9]))
assert bool(torch.isfinite(em).all()) and bool(torch.isfinite(sp).all()) and bool(torch.isfinite(modes[3].q0).all())
print("input helpers ok", float(em.sum()), float(sp.sum()), float(modes[0].q.sum()))
print("sanitizer run finished")
