"""Loader for the golden vectors written by tests/golden/make_golden.py and the
parity metric of SURVEY.md section 8(c)."""
import os

import numpy as np

from pythonic_disort_b200.subroutines import TabulatedBDRF

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SCALAR_ARGS = ("NQuad", "mu0", "I0", "phi0", "NLeg", "NFourier", "only_flux", "NT_cor")
INT_ARGS = ("NQuad", "NLeg", "NFourier")
BOOL_ARGS = ("only_flux", "NT_cor")
POSITIONAL = ("tau_arr", "omega_arr", "NQuad", "Leg_coeffs_all", "mu0", "I0", "phi0")


def suite_names():
    return sorted(f[:-4] for f in os.listdir(os.path.join(GOLDEN, "pydisotest")) if f.endswith(".npz"))


def load_test(name):
    d = np.load(os.path.join(GOLDEN, "pydisotest", name + ".npz"))
    records = []
    for k in range(int(d["n_records"])):
        pre = f"r{k}_arg_"
        named = {key[len(pre):]: d[key] for key in d.files if key.startswith(pre)}
        for a in list(named):
            if a in INT_ARGS:
                named[a] = int(named[a])
            elif a in BOOL_ARGS:
                named[a] = bool(named[a])
            elif a in SCALAR_ARGS:
                named[a] = float(named[a])
        named.pop("use_banded_solver_NLayers", None)
        named.pop("autograd_compatible", None)
        q, q0, sc = d[f"r{k}_bdrf_q"], d[f"r{k}_bdrf_q0"], d[f"r{k}_bdrf_scalar"]
        modes = [float(q[m, 0, 0]) if sc[m] else TabulatedBDRF(q[m], q0[m]) for m in range(len(sc))]
        calls = []
        for j in range(int(d[f"r{k}_n_calls"])):
            cp = f"r{k}_c{j}_"
            calls.append(dict(fn=str(d[cp + "fn"]), tau=d[cp + "tau"], phi=d[cp + "phi"], anti=bool(d[cp + "anti"]),
                              outs=[d[cp + f"out{i}"] for i in range(int(d[cp + "n_outs"]))]))
        args = tuple(named.pop(a) for a in POSITIONAL)
        kwargs = dict(named)
        if modes:
            kwargs["BDRF_Fourier_modes"] = modes
        records.append(dict(args=args, kwargs=kwargs, calls=calls))
    compares = []
    for j in range(int(d["n_compares"])):
        compares.append(dict(record=int(d[f"cmp{j}_record"]), file=str(d[f"cmp{j}_file"]),
                             mu_to_compare=d[f"cmp{j}_mu_to_compare"], reorder_mu=d[f"cmp{j}_reorder_mu"],
                             has_u=bool(d[f"cmp{j}_has_u"])))
    return records, compares


def stamnes(file):
    return np.load(os.path.join(GOLDEN, "stamnes", file))


def group_scale(refs):
    """Common magnitude of a family of outputs (e.g. diffuse and direct downward
    flux): a diffuse flux that is physically zero at the top of the atmosphere
    comes out of the reference as +-1e-16 noise and has no scale of its own."""
    vals = [np.max(np.abs(np.asarray(r, dtype=float)[np.isfinite(r)]), initial=0.0) for r in refs]
    return max(vals) if vals else 0.0


PW_FLOOR = 1e-6  # the pointwise criterion applies where |X_ref| > 1e-6 * max|X_ref|


def parity(mine, ref, floor=PW_FLOOR, scale=None):
    """(scale-relative error, worst pointwise relative error over entries with
    |ref| > floor * scale, number of non-finite reference entries masked);
    scale defaults to max|ref|."""
    mine = np.asarray(mine, dtype=float)
    ref = np.asarray(ref, dtype=float)
    if mine.shape != ref.shape:  # e.g. the reference returns a scalar 0 direct flux when there is no beam
        mine, ref = np.broadcast_arrays(mine, ref)
    ok = np.isfinite(ref)
    if not ok.any():
        return 0.0, 0.0, int(ref.size)
    if scale is None:
        scale = np.max(np.abs(ref[ok]))
    if scale == 0:
        return float(np.max(np.abs(mine[ok]))), 0.0, int((~ok).sum())
    err = np.abs(mine - ref)
    big = ok & (np.abs(ref) > floor * scale)
    pw = float(np.max(err[big] / np.abs(ref[big]))) if big.any() else 0.0
    return float(np.max(err[ok]) / scale), pw, int((~ok).sum())


def conditioning_tolerance(omega, base=1e-9):
    """The reference itself is only reproducible to about eps/(1-omega_max):
    re-ordering a few floating-point operations in it moves its conservative
    test problems (omega = 1 - 1e-6) by 1e-10...1e-8 ('Parity').
    Tolerance: ``base`` up to omega = 0.9999, then growing like 1/(1-omega)."""
    wmax = float(np.max(omega))
    return base * max(1.0, 1e-4 / max(1.0 - wmax, 1e-12))


def perturb_inputs(args, kwargs, seed):
    """The same problem with every floating-point input array moved by one unit in the last place, up or down at random
    (optical depths, single-scattering albedos, phase-function moments above the zeroth, source polynomials).
    Used to measure how far the REFERENCE ALGORITHM itself moves under a perturbation no caller can see
    (``reference_sensitivity``): no re-ordered implementation can be expected to sit closer to the reference's
    output than that.  The lowest level stays where it is so that recorded query depths remain inside the
    atmosphere."""
    rng = np.random.default_rng(seed)

    def p(x):
        x = np.array(x, dtype=float)
        up = rng.integers(0, 2, x.shape) > 0
        return np.nextafter(x, np.where(up, np.inf, -np.inf))

    a = list(args)
    tau = np.array(a[0], dtype=float)
    scalar_tau = tau.ndim == 0
    tau = np.atleast_1d(tau)
    tp = p(tau)
    tp[..., -1] = tau[..., -1]
    a[0] = float(tp[0]) if scalar_tau else tp
    a[1] = p(a[1]) if np.ndim(a[1]) else float(p(a[1]))
    leg = np.array(a[3], dtype=float)
    lp = p(leg)
    lp[..., 0] = leg[..., 0]
    a[3] = lp
    kw = dict(kwargs)
    if kw.get("s_poly_coeffs") is not None:
        kw["s_poly_coeffs"] = p(kw["s_poly_coeffs"])
    return tuple(a), kw


def run_calls(outputs, rec):
    """Evaluate the recorded calls on a (mu, flux_up, flux_down, u0[, u]) tuple;
    yields (call, list of produced arrays)."""
    fns = dict(zip(("flux_up", "flux_down", "u0", "u"), outputs[1:]))
    for c in rec["calls"]:
        fn = fns[c["fn"]]
        tau = c["tau"] if c["tau"].size > 1 else float(c["tau"][0])
        if c["fn"] == "u":
            res = fn(tau, c["phi"] if c["phi"].size > 1 else float(c["phi"][0]), c["anti"])
        else:
            res = fn(tau, c["anti"])
        yield c, list(res) if isinstance(res, tuple) else [res]
