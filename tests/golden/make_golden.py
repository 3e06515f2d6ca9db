"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

What it does
  1. runs every test function of the reference's own suite
     (/root/reference/pydisotest/*_test.py) with a spy wrapped around
     ``PythonicDISORT.pydisort``: the exact inputs of every call, the
     (tau, phi) points at which the tests evaluate the returned functions, and
     the reference's FP64 outputs there are written to
     ``tests/golden/pydisotest/<test>.npz``; BDRF callables are tabulated at the
     quadrature nodes (that is all the solver ever asks of them);
  2. copies the Stamnes DISORT 4.0.99 result files the tests compare against
     (data, not code) to ``tests/golden/stamnes/``;
  6. (``hapke``) builds the Fourier modes of the Hapke BDRF the way the reference's test problem 6b does (quad_vec over the
     relative azimuth, here with a tight tolerance) -> ``tests/golden/hapke_modes.npz``;
  5. (``thermal``) runs the reference's ``blackbody_contrib_to_BCs`` / ``generate_s_poly_coeffs`` on fixed temperatures,
     bands and atmospheres -> ``tests/golden/thermal_inputs.npz``;
  4. (``interpolate``) runs the reference's ``subroutines.interpolate`` on its own ``u`` / ``u0`` at the user
     polar angles of config 5 -> ``tests/golden/interpolate.npz``;
  3. runs the reference on fixed subsets of the three synthetic ensembles of
     SURVEY.md section 8(d) (generator: pythonic_disort_b200/synthetic.py) and
     stores inputs-by-seed + outputs in ``tests/golden/ensemble_<name>.npz``.
"""
import importlib.util
import os
import shutil
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "src"))
sys.path.insert(0, ROOT)

import PythonicDISORT  # noqa: E402
import PythonicDISORT.subroutines as ref_sub  # noqa: E402

_orig_pydisort = PythonicDISORT.pydisort
_orig_compare = ref_sub._compare
_orig_load = np.load

ARG_NAMES = ["tau_arr", "omega_arr", "NQuad", "Leg_coeffs_all", "mu0", "I0", "phi0", "NLeg", "NFourier",
             "b_pos", "b_neg", "only_flux", "f_arr", "NT_cor", "BDRF_Fourier_modes", "s_poly_coeffs",
             "use_banded_solver_NLayers", "autograd_compatible"]

state = {"records": [], "loads": [], "compares": []}
MAX_CALLS = 24


def _tabulate_bdrf(modes, NQuad, mu0, beam):
    N = NQuad // 2
    mu_pos = ref_sub.Gauss_Legendre_quad(N)[0]
    n = len(modes)
    q = np.zeros((n, N, N))
    q0 = np.zeros((n, N))
    is_scalar = np.zeros(n, dtype=bool)
    for m, fm in enumerate(modes):
        if np.isscalar(fm):
            is_scalar[m] = True
            q[m] = fm
            q0[m] = fm
        else:
            q[m] = fm(mu_pos, mu_pos)
            if beam:
                q0[m] = np.asarray(fm(mu_pos, np.array([mu0])))[:, 0]
    return q, q0, is_scalar


def spy_pydisort(*args, **kwargs):
    named = dict(zip(ARG_NAMES, args))
    named.update(kwargs)
    rec = {"args": {}, "calls": [], "keys": set()}
    for k, v in named.items():
        if k == "BDRF_Fourier_modes":
            continue
        if v is None:
            continue
        rec["args"][k] = np.array(v, dtype=float, copy=True)
    modes = named.get("BDRF_Fourier_modes", [])
    q, q0, sc = _tabulate_bdrf(modes, int(named["NQuad"]), float(named["mu0"]), float(named["I0"]) > 0)
    rec["bdrf_q"], rec["bdrf_q0"], rec["bdrf_scalar"] = q, q0, sc
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = _orig_pydisort(*args, **kwargs)
    names = ["flux_up", "flux_down", "u0", "u"][: len(out) - 1]

    def wrap(fn, name):
        def wrapped(tau, *a, **k):
            res = fn(tau, *a, **k)
            pos = list(a)
            phi = None
            if name == "u":
                phi = pos.pop(0) if pos else k.get("phi")
            anti = bool(pos[0]) if pos else bool(k.get("is_antiderivative_wrt_tau", False))
            extra = (len(pos) > 1 and any(pos[1:])) or k.get("return_Fourier_error") or k.get("return_tau_arr") \
                or k.get("_return_act_dscale_for_reclass") or k.get("_return_l")
            if not extra:
                t = np.atleast_1d(np.asarray(tau, dtype=float))
                ph = np.atleast_1d(np.asarray(phi, dtype=float)) if phi is not None else np.zeros(0)
                key = (name, t.tobytes(), ph.tobytes(), anti)
                small = t.size == 1 and not anti
                n_small = sum(1 for c in rec["calls"] if c["small"] and c["fn"] == name)
                if key not in rec["keys"] and len(rec["calls"]) < MAX_CALLS and not (small and n_small >= 2):
                    rec["keys"].add(key)
                    outs = res if isinstance(res, tuple) else (res,)
                    rec["calls"].append({"fn": name, "tau": t, "phi": ph, "anti": anti, "small": small,
                                         "outs": [np.array(o, dtype=float) for o in outs]})
            return res
        return wrapped

    wrapped = []
    for fn, name in zip(out[1:], names):
        w = wrap(fn, name)
        wrapped.append(w)
    rec["flux_up_fn"] = wrapped[0]
    state["records"].append(rec)
    return (out[0],) + tuple(wrapped)


def spy_compare(results, mu_to_compare, reorder_mu, flux_up, flux_down, u=None):
    owner = [i for i, r in enumerate(state["records"]) if r["flux_up_fn"] is flux_up]
    state["compares"].append({"record": owner[0], "file": state["loads"][-1] if state["loads"] else "",
                              "mu_to_compare": np.array(mu_to_compare), "reorder_mu": np.array(reorder_mu),
                              "has_u": u is not None})
    import io
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        return _orig_compare(results, mu_to_compare, reorder_mu, flux_up, flux_down, u)


def spy_load(file, *a, **k):
    if isinstance(file, str) and "Stamnes_results" in file:
        state["loads"].append(os.path.basename(file))
    return _orig_load(file, *a, **k)


def dump_test(name, outdir):
    d = {}
    recs = state["records"]
    d["n_records"] = np.array(len(recs))
    for k, rec in enumerate(recs):
        for an, av in rec["args"].items():
            d[f"r{k}_arg_{an}"] = av
        d[f"r{k}_bdrf_q"] = rec["bdrf_q"]
        d[f"r{k}_bdrf_q0"] = rec["bdrf_q0"]
        d[f"r{k}_bdrf_scalar"] = rec["bdrf_scalar"]
        d[f"r{k}_n_calls"] = np.array(len(rec["calls"]))
        for j, c in enumerate(rec["calls"]):
            d[f"r{k}_c{j}_fn"] = np.array(c["fn"])
            d[f"r{k}_c{j}_tau"] = c["tau"]
            d[f"r{k}_c{j}_phi"] = c["phi"]
            d[f"r{k}_c{j}_anti"] = np.array(c["anti"])
            d[f"r{k}_c{j}_n_outs"] = np.array(len(c["outs"]))
            for i, o in enumerate(c["outs"]):
                d[f"r{k}_c{j}_out{i}"] = o
    d["n_compares"] = np.array(len(state["compares"]))
    for j, c in enumerate(state["compares"]):
        d[f"cmp{j}_record"] = np.array(c["record"])
        d[f"cmp{j}_file"] = np.array(c["file"])
        d[f"cmp{j}_mu_to_compare"] = c["mu_to_compare"]
        d[f"cmp{j}_reorder_mu"] = c["reorder_mu"]
        d[f"cmp{j}_has_u"] = np.array(c["has_u"])
    np.savez_compressed(os.path.join(outdir, name + ".npz"), **d)


def run_reference_suite():
    outdir = os.path.join(HERE, "pydisotest")
    os.makedirs(outdir, exist_ok=True)
    PythonicDISORT.pydisort = spy_pydisort
    ref_sub._compare = spy_compare
    np.load = spy_load
    cwd = os.getcwd()
    os.chdir(os.path.join(REF, "pydisotest"))
    sys.path.insert(0, os.getcwd())
    np.random.seed(20260111)  # test_11a draws its tau points from the global RNG
    try:
        for fname in sorted(os.listdir(".")):
            if not fname.endswith("_test.py"):
                continue
            spec = importlib.util.spec_from_file_location("ref_" + fname[:-3], fname)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            for tname in sorted(n for n in dir(mod) if n.startswith("test_")):
                state["records"], state["loads"], state["compares"] = [], [], []
                import io
                import contextlib
                with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    getattr(mod, tname)()
                dump_test(tname[5:], outdir)
                print(f"{tname}: {len(state['records'])} solves, "
                      f"{sum(len(r['calls']) for r in state['records'])} evaluations, "
                      f"{len(state['compares'])} Stamnes comparisons")
    finally:
        os.chdir(cwd)
        PythonicDISORT.pydisort = _orig_pydisort
        ref_sub._compare = _orig_compare
        np.load = _orig_load


def copy_stamnes():
    dst = os.path.join(HERE, "stamnes")
    os.makedirs(dst, exist_ok=True)
    src = os.path.join(REF, "pydisotest", "Stamnes_results")
    for f in sorted(os.listdir(src)):
        if f.endswith(".npz"):
            shutil.copy(os.path.join(src, f), os.path.join(dst, f))


def run_ensembles():
    from pythonic_disort_b200 import synthetic
    for name, ncol in (("sw", 24), ("lw", 64), ("ha", 2), ("tp9c16", 1)):
        ens = synthetic.make(name, ncol)
        outs = synthetic.run_reference_like(_orig_pydisort, ens)
        np.savez_compressed(os.path.join(HERE, f"ensemble_{name}.npz"), ncol=np.array(ncol), **outs)
        print(f"ensemble {name}: {ncol} columns")


MU_USER = np.array([0.1, 0.5, 0.9, -0.1, -0.5, -0.9])  # user polar angles


def run_interpolate():
    """Row f1: the reference's ``subroutines.interpolate`` (subroutines.py:614-705) on its own ``u`` / ``u0`` for
    columns of the HA ensemble (config 5) and of the SW ensemble (NT-corrected intensities)."""
    from pythonic_disort_b200 import synthetic
    out = {"mu_user": MU_USER}
    for name, ncol in (("ha", 2), ("sw", 3)):
        ens = synthetic.make(name, ncol)
        um, u0m = [], []
        for b in range(ncol):
            args, kwargs = synthetic.column_call(ens, b)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                res = _orig_pydisort(*args, **kwargs)
            t = ens["tau_eval"][b]
            um.append(ref_sub.interpolate(res[4])(MU_USER, t, ens["phi_eval"]))
            u0m.append(ref_sub.interpolate(res[3])(MU_USER, t))
        out[f"{name}_ncol"] = np.array(ncol)
        out[f"{name}_u"] = np.array(um)
        out[f"{name}_u0"] = np.array(u0m)
        print(f"interpolate {name}: {ncol} columns", out[f"{name}_u"].shape)
    np.savez_compressed(os.path.join(HERE, "interpolate.npz"), **out)


def run_thermal():
    """Row f2: the reference's thermal-source helpers (subroutines.py:322-454) on fixed inputs: band emission for a
    range of temperatures and bands (narrow, wide, far tail) and s_poly_coeffs for a few atmospheres."""
    rng = np.random.default_rng(20260104)
    T = np.concatenate([[0.0, 2.7, 50.0], np.linspace(150.0, 340.0, 24), [1000.0, 5800.0]])
    bands = np.array([[999.0, 1000.0], [100.0, 110.0], [600.0, 700.0], [0.01, 50000.0], [2500.0, 2600.0], [10.0, 3000.0]])
    em = np.array([ref_sub.blackbody_contrib_to_BCs(T, lo, hi) for lo, hi in bands])
    ncol, L = 5, 12
    tau = np.cumsum(rng.uniform(0.05, 2.0, (ncol, L)), axis=1)
    temper = np.sort(rng.uniform(180.0, 310.0, (ncol, L + 1)), axis=1)
    sp_bands = np.array([[999.0, 1000.0], [600.0, 700.0], [10.0, 3000.0]])
    sp = np.array([[ref_sub.generate_s_poly_coeffs(tau[b], temper[b], lo, hi) for b in range(ncol)] for lo, hi in sp_bands])
    np.savez_compressed(os.path.join(HERE, "thermal_inputs.npz"), T=T, bands=bands, emission=em, tau=tau, temper=temper,
                        sp_bands=sp_bands, s_poly=sp)
    print("thermal inputs:", em.shape, sp.shape)


def run_hapke():
    """Row f4: Fourier modes of the Hapke BDRF exactly as the reference's test problem 6b builds them
    (pydisotest/6_test.py:11-24 and :193-201: quad_vec over the relative azimuth), at the quadrature nodes and at two
    beam cosines."""
    spec = importlib.util.spec_from_file_location("ref_test6", os.path.join(REF, "pydisotest", "6_test.py"))
    t6 = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(t6)
    N, NF = 8, 10
    B0, HH, W = 1, 0.06, 0.6
    mu = ref_sub.Gauss_Legendre_quad(N)[0]
    cols = np.concatenate([mu, [0.5, 0.3137]])
    import scipy.integrate
    modes = np.array([scipy.integrate.quad_vec(lambda dphi: t6.Hapke(mu, cols, dphi, B0, HH, W) * np.cos(m * dphi), 0, 2 * np.pi,
                                               epsrel=1e-12, limit=4000)[0] / ((1 + (m == 0)) * np.pi) for m in range(NF)])
    np.savez_compressed(os.path.join(HERE, "hapke_modes.npz"), N=np.array(N), mu0=cols[N:], params=np.array([B0, HH, W]),
                        modes=modes)
    print("hapke modes:", modes.shape)


def run_actinic():
    """Row a13: the reference's actinic-flux functions (``generate_diff_act_flux_funcs``, subroutines.py:258-318, on
    the ``u0`` of ``_assemble_intensity_and_fluxes.py:334-433`` with its delta-scaling reclassification term,
    :360-371) on the level grid of SW columns (delta-M: the reclassification is active), LW columns (thermal,
    no beam), test problem 9c and a single-layer delta-M problem with its tau-antiderivative."""
    from pythonic_disort_b200 import synthetic
    out = {}
    for name, ncol in (("sw", 3), ("lw", 4), ("tp9c", 1)):
        ens = synthetic.make(name, ncol)
        up, dn = [], []
        for b in range(ncol):
            args, kwargs = synthetic.column_call(ens, b)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                res = _orig_pydisort(*args, **kwargs)
            fu, fd = ref_sub.generate_diff_act_flux_funcs(res[3])
            t = ens["tau_eval"][b]
            up.append(fu(t))
            dn.append(fd(t))
        out[f"{name}_ncol"] = np.array(ncol)
        out[f"{name}_up"] = np.array(up)
        out[f"{name}_down"] = np.array(dn)
        print(f"actinic {name}: {ncol} columns", out[f"{name}_up"].shape)
    # one layer, delta-M scaled, with the antiderivative (the reference's multi-layer antiderivatives are broken)
    NQuad = 8
    leg = 0.8 ** np.arange(NQuad + 4)
    args1 = (np.array([1.7]), np.array([0.9]), NQuad, leg[None, :], 0.6, 2.0, 0.3)
    kw1 = dict(f_arr=np.array([leg[NQuad]]), NT_cor=False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = _orig_pydisort(*args1, **kw1)
    fu, fd = ref_sub.generate_diff_act_flux_funcs(res[3])
    t1 = np.linspace(0, 1.7, 9)
    out.update(one_tau=t1, one_up=fu(t1), one_down=fd(t1), one_up_anti=fu(t1, True), one_down_anti=fd(t1, True),
               one_leg=leg, one_f=np.array([leg[NQuad]]))
    np.savez_compressed(os.path.join(HERE, "actinic.npz"), **out)


if __name__ == "__main__":
    what = sys.argv[1:] or ["suite", "stamnes", "ensembles", "interpolate", "thermal", "hapke", "actinic"]
    if "suite" in what:
        run_reference_suite()
    if "stamnes" in what:
        copy_stamnes()
    if "ensembles" in what:
        run_ensembles()
    if "interpolate" in what:
        run_interpolate()
    if "thermal" in what:
        run_thermal()
    if "hapke" in what:
        run_hapke()
    if "actinic" in what:
        run_actinic()
