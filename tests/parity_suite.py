"""Parity checks shared by the CPU (host build of the kernels) and GPU test files."""
import os
import warnings

import numpy as np

import golden_io
from pythonic_disort_b200 import api, parallel, synthetic
from pythonic_disort_b200 import ensemble as pd_ensemble
from pythonic_disort_b200.subroutines import _compare


def to_np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


# How many times the reference algorithm's own response to a 1-ulp change of its inputs an implementation may differ
# from the reference by, before the 1e-9 bars of SURVEY.md 8(c) give way.  Scale criterion: 4 (the omega = 1 - 1e-6
# test problems: CUDA path 6.6e-9 / oracle 2.5e-9 / 1-ulp response 4.9e-9 on 1b).  Pointwise criterion: 64 -- an entry
# a million times below the field's maximum that is the cancellation residue of the thermal source polynomial of a
# micrometre-thin layer (8ARTS_A: s0 + s1 tau with |s1 tau| ~ 1e7 |result|) carries one rounding of the large terms per
# evaluation and per layer it is carried through; three random 1-ulp draws move it by 1-3 such quanta, two correct
# implementations differ by a few tens of them (measured: up to 37 on the host build, 28 on the GPU).  In backward-
# error terms: the result is the reference's for inputs that differ by at most 64 ulp.
SENS_FACTOR = 4.0
SENS_FACTOR_PW = 64.0
SENS_DRAWS = 3


def _oracle():
    from oracle import disort_oracle  # test infrastructure (this module lives under tests/)
    return disort_oracle.pydisort


def record_sensitivity(rec, draws=SENS_DRAWS):
    """Per recorded call and output: (scale error, pointwise error) by which the reference algorithm (the pinned
    oracle) moves when its inputs move by one unit in the last place -- the conditioning of the problem as posed,
    measured, not modelled.  Zero for well-conditioned problems (where the plain 1e-9 bars apply), ~1e-8 for the
    omega = 1 - 1e-6 test problems, ~1e-5 pointwise where an output entry is the cancellation residue of terms a
    million times larger (8ARTS_A's micrometre-thin layers)."""
    oracle = _oracle()

    def evaluate(args, kwargs):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = oracle(*args, **kwargs)
            return [[np.squeeze(np.asarray(g, dtype=float)) for g in got] for _, got in golden_io.run_calls(out, rec)]

    base = evaluate(rec["args"], rec["kwargs"])
    sens = [[(0.0, 0.0)] * len(c) for c in base]
    for d in range(draws):
        moved = evaluate(*golden_io.perturb_inputs(rec["args"], rec["kwargs"], seed=1000 + d))
        for ci, call in enumerate(rec["calls"]):
            scale = golden_io.group_scale(call["outs"])
            for k in range(len(base[ci])):
                e, pw, _ = golden_io.parity(moved[ci][k], base[ci][k], scale=scale)
                sens[ci][k] = (max(sens[ci][k][0], e), max(sens[ci][k][1], pw))
    return sens


def check_suite_record_outputs(pydisort, name, max_records=None, base_tol=1e-9):
    """Every evaluation the reference test `name` performs, against the reference's own FP64 output, by both
    criteria of SURVEY.md 8(c): max|X - X_ref| <= 1e-9 max|X_ref| per field, and pointwise
    |X - X_ref| <= 1e-9 |X_ref| wherever |X_ref| > 1e-6 max|X_ref|.  Where the problem itself is ill-conditioned
    both bars widen to SENS_FACTOR (scale) / SENS_FACTOR_PW (pointwise) x the reference algorithm's own response to a
    1-ulp change of its inputs (``record_sensitivity``); the scale criterion additionally never exceeds the analytic omega -> 1 bound of
    ``golden_io.conditioning_tolerance``."""
    records, _ = golden_io.load_test(name)
    worst = 0.0
    for rec in records[:max_records]:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = pydisort(*rec["args"], **rec["kwargs"])
        cap = golden_io.conditioning_tolerance(rec["args"][1], base_tol)
        sens = record_sensitivity(rec)
        for ci, (call, got) in enumerate(golden_io.run_calls(out, rec)):
            scale = golden_io.group_scale(call["outs"])
            for k, (g, r) in enumerate(zip(got, call["outs"])):
                err, pw, _ = golden_io.parity(np.squeeze(to_np(g)), np.squeeze(r), scale=scale)
                tol = min(cap, max(base_tol, SENS_FACTOR * sens[ci][k][0]))
                tol_pw = max(base_tol, SENS_FACTOR_PW * sens[ci][k][1])
                where = (name, call["fn"], "anti" if call["anti"] else "")
                assert err <= tol, where + ("scale criterion", err, tol)
                assert pw <= tol_pw, where + ("pointwise criterion", pw, tol_pw)
                worst = max(worst, err / tol)
    return worst


def check_stamnes(pydisort, name):
    """The reference suite's own pass criteria against DISORT 4.0.99 (pydisotest/1_test.py:78-81)."""
    records, compares = golden_io.load_test(name)
    results = []
    for cmp in compares:
        rec = records[cmp["record"]]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = pydisort(*rec["args"], **rec["kwargs"])
        results.append(_compare(golden_io.stamnes(cmp["file"]), cmp["mu_to_compare"], cmp["reorder_mu"],
                                out[1], out[2], out[4] if cmp["has_u"] else None))
    if name == "9corrections":  # pydisotest/9_test.py:368-370
        plain, corrected = results
        assert np.mean(plain[0] - corrected[0]) > 0
        assert np.mean(plain[2] - corrected[2]) > 0
        assert np.mean(plain[6] - corrected[6]) > 0
        return
    for res in results:
        for d, ratio in zip(res[0:6:2], res[1:6:2]):
            assert np.max(ratio[d > 1e-3], initial=0) < 1e-3
        if len(res) > 6:
            assert np.max(res[7][res[6] > 1e-3], initial=0) < 1e-2


def run_batched(pydisort, ens):
    """One batched call + evaluation on the ensemble's grid -> dict of numpy arrays [B, ...]."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = pydisort(*ens["args"], **ens["kwargs"])
    t = ens["tau_eval"]
    dn = out[2](t)
    res = dict(flux_up=to_np(out[1](t)), flux_down_diffuse=to_np(dn[0]), flux_down_direct=to_np(dn[1]),
               u0=to_np(out[3](t)))
    if "u" in ens["outputs"]:
        res["u"] = to_np(out[4](t, ens["phi_eval"]))
    return res


def _field_scale(ref, key, b):
    if key.startswith("flux"):
        return golden_io.group_scale([ref[k][b] for k in ("flux_up", "flux_down_diffuse", "flux_down_direct")])
    return None


def compare_fields(got, ref, ncol, tol=1e-9, label="", sens=None, base_tol=1e-9):
    """Per column and field, both criteria of SURVEY.md 8(c): max|X - X_ref| <= bar * scale, and pointwise
    |X - X_ref| <= bar_pw * |X_ref| wherever |X_ref| > 1e-6 * scale.  ``tol`` caps the scale bar (the analytic
    omega -> 1 bound); with ``sens`` (``ensemble_sensitivity``) both bars are 1e-9 unless SENS_FACTOR / SENS_FACTOR_PW
    times the reference algorithm's own response to a 1-ulp change of this column's inputs is larger.  Without ``sens``
    (CUDA path against itself) only the scale criterion is asserted.  Returns the worst errors."""
    worst, worst_pw, masked = 0.0, 0.0, 0
    for b in range(ncol):
        for key in got:
            err, pw, nmask = golden_io.parity(got[key][b], ref[key][b], scale=_field_scale(ref, key, b))
            bar = tol
            if sens is not None:
                bar = min(tol, max(base_tol, SENS_FACTOR * sens[key][b][0]))
                bar_pw = max(base_tol, SENS_FACTOR_PW * sens[key][b][1])
                assert pw <= bar_pw, (label, key, b, "pointwise criterion", pw, bar_pw)
            assert err <= bar, (label, key, b, "scale criterion", err, bar)
            worst, worst_pw, masked = max(worst, err), max(worst_pw, pw), masked + nmask
    return worst, worst_pw, masked


def _oracle_job(job):
    """Oracle on columns [first, first + ncol) of ensemble `name`; `seed` = None for the inputs as they are, else the
    1-ulp perturbation of golden_io.perturb_inputs.  Top-level so that a process pool can run it."""
    name, first, ncol, seed = job
    ens = synthetic.make(name, ncol, first)
    if seed is not None:
        a, k = golden_io.perturb_inputs(ens["args"], ens["kwargs"], seed)
        ens = dict(ens, args=a, kwargs=k)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return synthetic.run_reference_like(_oracle(), ens)


def oracle_on_ensemble(name, ncol, first=0, seeds=(None,), pool=None):
    """[oracle outputs for every seed]; columns are spread over `pool` (a multiprocessing pool) when given."""
    workers = pool._processes if pool is not None else 1
    per = max(1, -(-ncol // workers))
    jobs = [(name, first + lo, min(per, ncol - lo), seed) for seed in seeds for lo in range(0, ncol, per)]
    parts = pool.map(_oracle_job, jobs) if pool is not None else [_oracle_job(j) for j in jobs]
    nper = len(jobs) // len(seeds)
    return [{k: np.concatenate([p[k] for p in parts[i * nper:(i + 1) * nper]]) for k in parts[0]}
            for i in range(len(seeds))]


def ensemble_sensitivity(base, moved):
    """field -> per column (scale error, pointwise error) of the perturbed oracle runs against the unperturbed one."""
    sens = {}
    for key in base:
        rows = []
        for b in range(len(base[key])):
            scale = _field_scale(base, key, b)
            e = [golden_io.parity(m[key][b], base[key][b], scale=scale)[:2] for m in moved]
            rows.append((max(x[0] for x in e), max(x[1] for x in e)))
        sens[key] = rows
    return sens


def check_ensemble_vs_golden(pydisort, name, tol=1e-9):
    """First columns of a synthetic ensemble against the outputs of the unmodified reference (tests/golden/ensemble_*)."""
    gold = np.load(os.path.join(golden_io.GOLDEN, f"ensemble_{name}.npz"))
    ncol = int(gold["ncol"])
    ens = synthetic.make(name, ncol)
    got = run_batched(pydisort, ens)
    runs = oracle_on_ensemble(name, ncol, 0, seeds=(None,) + tuple(range(2000, 2000 + SENS_DRAWS)))
    sens = ensemble_sensitivity(runs[0], runs[1:])
    return compare_fields(got, {k: gold[k] for k in got}, ncol, tol, name, sens=sens)


def check_ensemble_vs_live_oracle(pydisort, name, ncol, first, pool=None):
    """A larger slice (256 SW / 1,024 LW / 64 HA columns) against the pinned oracle run live."""
    ens = synthetic.make(name, ncol, first)
    got = run_batched(pydisort, ens)
    runs = oracle_on_ensemble(name, ncol, first, seeds=(None,) + tuple(range(3000, 3000 + SENS_DRAWS)), pool=pool)
    sens = ensemble_sensitivity(runs[0], runs[1:])
    tol = golden_io.conditioning_tolerance(ens["args"][1])
    return compare_fields(got, runs[0], ncol, tol, name, sens=sens)


def _solution_of(out_fn):
    """The solved state an output closure of pydisort() holds (test access to switch its interface table off)."""
    for cell in out_fn.__closure__ or ():
        if type(cell.cell_contents).__name__ == "_Solution":
            return cell.cell_contents
    raise AssertionError("output function holds no _Solution")


def check_interface_levels_vs_assembled(pydisort, name, ncol, first=0, typical=1e-11):
    """Query points that are layer interfaces are answered from the interface radiances of the boundary-condition
    sweep (pd_state.Uif), every other point from G (C * exp) + particular.  Both formulations on the same levels
    (the second with the table switched off) and a grid mixing interfaces with interior points: the median column
    agrees to ``typical`` (measured 1e-14 at NQuad = 16, 4e-13 at NQuad = 32 with 100 layers) and every column within
    the parity bar (measured worst: 2.5e-11 on SW column 3163 and 1.1e-10 on HA column 921, columns where an eigenvalue
    lies within 1e-7 of 1/mu0 and the beam particular solution is conditioned like 1/(1/mu0^2 - k^2); on the SW column
    the two formulations are 2.9e-11 and 1.7e-11 from the oracle).  A
    self-consistency test of the CUDA path, not a parity test (the goldens and the oracle runs exercise the interface
    path through the ensembles' level grids)."""
    ens = synthetic.make(name, ncol, first)
    tol = golden_io.conditioning_tolerance(ens["args"][1], 1e-9)
    t_if = np.asarray(ens["tau_eval"], dtype=np.float64)
    tau = np.asarray(ens["args"][0], dtype=np.float64)
    if tau.ndim == 1:
        tau = np.tile(tau, (ncol, 1))
    if t_if.ndim == 1:
        t_if = np.tile(t_if, (ncol, 1))
    edges = np.concatenate([np.zeros((ncol, 1)), tau], axis=1)
    mids = 0.5 * (edges[:, :-1] + edges[:, 1:])
    mixed = np.sort(np.concatenate([edges, mids[:, ::3]], axis=1), axis=1)
    hits = np.mean(np.isin(t_if, edges))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = pydisort(*ens["args"], **ens["kwargs"])
    sol = _solution_of(out[1])
    assert sol.Uif is not None

    def evaluate(t):
        dn = out[2](t)
        res = dict(flux_up=to_np(out[1](t)), flux_down_diffuse=to_np(dn[0]), u0=to_np(out[3](t)))
        if "u" in ens["outputs"]:
            res["u"] = to_np(out[4](t, ens["phi_eval"]))
        return res

    grids = [edges, mixed] + ([t_if] if hits < 1.0 else [])
    with_table = [evaluate(t) for t in grids]
    table, sol.Uif = sol.Uif, None
    try:
        assembled = [evaluate(t) for t in grids]
    finally:
        sol.Uif = table
    worst = 0.0
    for a, c in zip(with_table, assembled):
        for key in a:
            # scale of a flux = the column's flux family, as in compare_fields (the diffuse downward flux under an
            # optically thick absorbing column is 1e-14 of it: both formulations carry it to absolute accuracy only)
            scale = [max(np.max(np.abs(c[k][b])) for k in c if k.startswith("flux") == key.startswith("flux"))
                     for b in range(ncol)]
            errs = [np.max(np.abs(a[key][b] - c[key][b])) / scale[b] for b in range(ncol)]
            worst = max(worst, max(errs))
            assert max(errs) <= tol and np.median(errs) <= typical, (name, key, int(np.argmax(errs)), max(errs), np.median(errs))
    # the table must actually have been used: with it, interface levels no longer depend on C
    keep = sol.C.clone()
    sol.C.zero_()
    try:
        again = evaluate(edges)
    finally:
        sol.C.copy_(keep)
    for key in again:
        assert np.array_equal(again[key], with_table[0][key]), key
    return worst


def with_compact(ens):
    """(args, kwargs) of the ensemble with its compact input descriptions in place of the arrays."""
    args, kw = list(ens["args"]), dict(ens["kwargs"])
    for name, obj in ens["compact"].items():
        if name in api.POSITIONAL:
            args[api.POSITIONAL.index(name)] = obj
        else:
            kw[name] = obj
    return tuple(args), kw


def check_compact_inputs(pd, name, ncol, chunk, first=0):
    """inputs.HenyeyGreenstein / LevelSource in place of Leg_coeffs_all / s_poly_coeffs: same results as the arrays
    (to the rounding of the moments), through pydisort(), solve_ensemble() and the sharding rule; only the description
    is uploaded."""
    ens = synthetic.make(name, ncol, first)
    ref = run_batched(pd.pydisort, ens)
    cargs, ckw = with_compact(ens)
    got = run_batched(pd.pydisort, dict(ens, args=cargs, kwargs=ckw))
    for key in ref:
        # moments differ by an ulp between pow implementations; the NQuad = 32 columns answer that with 1.5e-11 of scale
        np.testing.assert_allclose(got[key], ref[key], rtol=0, atol=1e-10 * np.max(np.abs(ref[key])))
    outputs = ("flux_up", "flux_down", "u0") + (("u",) if "u" in ens["outputs"] else ())
    res = pd_ensemble.solve_ensemble(*cargs, tau=ens["tau_eval"], phi=ens["phi_eval"], outputs=outputs, chunk=chunk, **ckw)
    full = pd_ensemble.solve_ensemble(*ens["args"], tau=ens["tau_eval"], phi=ens["phi_eval"], outputs=outputs, chunk=chunk,
                                   **ens["kwargs"])
    for key in res:
        np.testing.assert_array_equal(res[key], got[key])
    assert res.h2d_bytes < 0.5 * full.h2d_bytes
    a, k, (lo, hi) = parallel.shard_inputs(ncol, cargs, ckw, rank=1, world=2)
    part = run_batched(pd.pydisort, dict(ens, args=a, kwargs=k, tau_eval=ens["tau_eval"][lo:hi]))
    np.testing.assert_array_equal(part["flux_up"], got["flux_up"][lo:hi])
    # one column, reference-style call
    a1, k1 = synthetic.column_call(ens, 0)
    a1, k1 = list(a1), dict(k1)
    for nm, obj in ens["compact"].items():
        if nm in api.POSITIONAL:
            a1[api.POSITIONAL.index(nm)] = obj[0]
        else:
            k1[nm] = obj[0]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        one = pd.pydisort(*a1, **k1)
    np.testing.assert_allclose(one[1](ens["tau_eval"][0]), got["flux_up"][0], rtol=1e-12)


def check_actinic_vs_golden(pd_module, tol=1e-9):
    """Row a13: ``generate_diff_act_flux_funcs`` (subroutines.py:258-318) on this package's ``u0`` -- including the
    delta-scaling reclassification term of ``_assemble_intensity_and_fluxes.py:360-371`` -- against the reference's
    actinic fluxes on SW columns (delta-M active), LW columns (thermal), test problem 9c and a one-layer delta-M
    problem with its tau-antiderivative (tests/golden/actinic.npz, made by make_golden.py from the unmodified
    reference).  Both 8(c) criteria; the upward and downward actinic flux of a column share one scale."""
    gold = np.load(os.path.join(golden_io.GOLDEN, "actinic.npz"))
    sub = pd_module.subroutines
    worst = 0.0

    def compare(got, ref, where):
        nonlocal worst
        scale = golden_io.group_scale(ref)
        for g, r in zip(got, ref):
            err, pw, _ = golden_io.parity(np.squeeze(to_np(g)), np.squeeze(r), scale=scale)
            assert err <= tol and pw <= tol, where + (err, pw)
            worst = max(worst, err)

    for name in ("sw", "lw", "tp9c"):
        ncol = int(gold[f"{name}_ncol"])
        ens = synthetic.make(name, ncol)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = pd_module.pydisort(*ens["args"], **ens["kwargs"])
        up, down = sub.generate_diff_act_flux_funcs(out[3])
        got_up, got_dn = np.atleast_2d(to_np(up(ens["tau_eval"]))), np.atleast_2d(to_np(down(ens["tau_eval"])))
        assert got_up.shape == gold[f"{name}_up"].shape
        for b in range(ncol):
            compare((got_up[b], got_dn[b]), (gold[f"{name}_up"][b], gold[f"{name}_down"][b]), (name, b))
            args, kwargs = synthetic.column_call(ens, b)   # the reference-style, unbatched call gives the same
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                one = pd_module.pydisort(*args, **kwargs)
            u1, d1 = sub.generate_diff_act_flux_funcs(one[3])
            compare((u1(ens["tau_eval"][b]), d1(ens["tau_eval"][b])), (gold[f"{name}_up"][b], gold[f"{name}_down"][b]),
                    (name, b, "unbatched"))
    NQuad, leg, t1 = 8, gold["one_leg"], gold["one_tau"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = pd_module.pydisort(np.array([1.7]), np.array([0.9]), NQuad, leg[None, :], 0.6, 2.0, 0.3,
                                 f_arr=gold["one_f"], NT_cor=False)
    up, down = sub.generate_diff_act_flux_funcs(out[3])
    compare((up(t1), down(t1)), (gold["one_up"], gold["one_down"]), ("one layer",))
    compare((up(t1, True), down(t1, True)), (gold["one_up_anti"], gold["one_down_anti"]), ("one layer, antiderivative",))
    val, tau_back = up(t1, False, True)
    np.testing.assert_array_equal(np.squeeze(tau_back), 1.7)   # `return_tau_arr` hands back the layer grid
    return worst


def check_interpolate_vs_golden(pd_module, name, tol=1e-9):
    """Row f1: ``subroutines.interpolate`` on the batched output functions (GPU: ``pd_interp_mu``) against the
    reference's own ``interpolate`` (tests/golden/interpolate.npz, made by make_golden.py), and against the
    host (SciPy) interpolation of the same closures."""
    gold = np.load(os.path.join(golden_io.GOLDEN, "interpolate.npz"))
    ncol, mu = int(gold[f"{name}_ncol"]), gold["mu_user"]
    ens = synthetic.make(name, ncol)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = pd_module.pydisort(*ens["args"], **ens["kwargs"])
    t, phi = ens["tau_eval"], ens["phi_eval"]
    got_u = to_np(pd_module.subroutines.interpolate(out[4])(mu, t, phi))
    got_u0 = to_np(pd_module.subroutines.interpolate(out[3])(mu, t))
    assert got_u.shape == gold[f"{name}_u"].shape and got_u0.shape == gold[f"{name}_u0"].shape
    worst = 0.0
    for b in range(ncol):
        for g, r in ((got_u[b], gold[f"{name}_u"][b]), (got_u0[b], gold[f"{name}_u0"][b])):
            err, _, _ = golden_io.parity(g, r)
            assert err <= tol, (name, b, err)
            worst = max(worst, err)
    # the same closures through the host path (a foreign callable has no `at_mu`)
    plain = lambda tau, phi_, a=False, f=False, r=False: out[4](tau, phi_, a, f, r)  # noqa: E731
    plain.kind, plain.batched = "u", True
    host = pd_module.subroutines.interpolate(plain)(mu, t, phi)
    np.testing.assert_allclose(got_u, host, rtol=1e-10, atol=1e-12 * np.max(np.abs(host)))
    return worst


def check_thermal_inputs_vs_golden(pd_module, tol=1e-9):
    """Row f2: band-integrated Planck emission and s_poly_coeffs from level temperatures through the C ABI
    (pd_planck_band, pd_s_poly_coeffs) against the reference's quad_vec-based helpers
    (tests/golden/thermal_inputs.npz).  The reference integrates to a relative 1e-8 only, so the bar is the usual
    1e-9 of scale per output array, not per element."""
    import torch
    from pythonic_disort_b200 import api
    sub = pd_module.subroutines
    gold = np.load(os.path.join(golden_io.GOLDEN, "thermal_inputs.npz"))
    dev = api._backend()[1]
    worst = 0.0
    T = torch.as_tensor(gold["T"], device=dev)
    for (lo, hi), ref in zip(gold["bands"], gold["emission"]):
        got = to_np(sub.blackbody_contrib_to_BCs(T, lo, hi))
        assert got.shape == ref.shape
        err = np.max(np.abs(got - ref)) / np.max(np.abs(ref))
        assert err <= tol, ("emission", lo, hi, err)
        big = np.abs(ref) > 1e-6 * np.max(np.abs(ref))  # pointwise where the value matters
        assert np.max(np.abs(got[big] - ref[big]) / np.abs(ref[big])) <= 1e-7
        worst = max(worst, err)
    tau, tem = torch.as_tensor(gold["tau"], device=dev), torch.as_tensor(gold["temper"], device=dev)
    for (lo, hi), ref in zip(gold["sp_bands"], gold["s_poly"]):
        got = to_np(sub.generate_s_poly_coeffs(tau, tem, lo, hi))
        assert got.shape == ref.shape
        for k in range(2):
            err = np.max(np.abs(got[..., k] - ref[..., k])) / np.max(np.abs(ref[..., k]))
            assert err <= 10 * tol, ("s_poly", lo, hi, k, err)  # slopes difference two integrals accurate to 1e-8 each
            worst = max(worst, err)
        one = to_np(sub.generate_s_poly_coeffs(tau[1], tem[1], lo, hi))  # unbatched tensor call
        np.testing.assert_array_equal(one, got[1])
        host = sub.generate_s_poly_coeffs(gold["tau"], gold["temper"], lo, hi)  # NumPy input: host route, batched
        np.testing.assert_allclose(host, ref, rtol=1e-7, atol=1e-9 * np.max(np.abs(ref)))
    return worst


def check_hapke_modes_vs_golden(pd_module):
    """Row f4: Fourier modes of the Hapke BDRF on the device (pd_hapke_modes) against the quad_vec construction of the
    reference's test problem 6b (tests/golden/hapke_modes.npz, integrated to 1e-12)."""
    import torch
    sub = pd_module.subroutines
    gold = np.load(os.path.join(golden_io.GOLDEN, "hapke_modes.npz"))
    N, ref = int(gold["N"]), gold["modes"]
    B0, HH, W = gold["params"]
    NF = ref.shape[0]
    modes = sub.hapke_BDRF_Fourier_modes(N, NF, torch.as_tensor(gold["mu0"]), B0, HH, W)
    got = np.stack([np.concatenate([to_np(fm.q), to_np(fm.q0)], axis=1) for fm in modes])
    assert got.shape == ref.shape
    err = np.abs(got - ref) / np.max(np.abs(ref[0]))
    assert err.max() <= 1e-9, err.max()
    return err.max()
