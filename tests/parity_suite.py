"""Parity checks shared by the CPU (host build of the kernels) and GPU test files."""
import os
import warnings

import numpy as np

import golden_io
from pythonic_disort_b200 import synthetic
from pythonic_disort_b200.subroutines import _compare


def to_np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def check_suite_record_outputs(pydisort, name, max_records=None, base_tol=1e-9):
    """Every evaluation the reference test `name` performs, against the reference's own FP64 output."""
    records, _ = golden_io.load_test(name)
    worst = 0.0
    for rec in records[:max_records]:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = pydisort(*rec["args"], **rec["kwargs"])
        tol = golden_io.conditioning_tolerance(rec["args"][1], base_tol)
        for call, got in golden_io.run_calls(out, rec):
            scale = golden_io.group_scale(call["outs"])
            for g, r in zip(got, call["outs"]):
                err, _, _ = golden_io.parity(np.squeeze(to_np(g)), np.squeeze(r), scale=scale)
                assert err <= tol, (name, call["fn"], "anti" if call["anti"] else "", err, tol)
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


def compare_fields(got, ref, ncol, tol=1e-9, label=""):
    """Per column and field: max|X - X_ref| <= tol * scale; returns the worst ratio and the
    worst pointwise-relative error over entries larger than 1e-3 of the field's scale."""
    worst, worst_pw, masked = 0.0, 0.0, 0
    for b in range(ncol):
        fscale = golden_io.group_scale([ref[k][b] for k in ("flux_up", "flux_down_diffuse", "flux_down_direct")])
        for key in got:
            scale = fscale if key.startswith("flux") else None
            err, pw, nmask = golden_io.parity(got[key][b], ref[key][b], scale=scale)
            assert err <= tol, (label, key, b, err)
            worst, worst_pw, masked = max(worst, err), max(worst_pw, pw), masked + nmask
    return worst, worst_pw, masked


def check_ensemble_vs_golden(pydisort, name, tol=1e-9):
    gold = np.load(os.path.join(golden_io.GOLDEN, f"ensemble_{name}.npz"))
    ncol = int(gold["ncol"])
    ens = synthetic.make(name, ncol)
    got = run_batched(pydisort, ens)
    return compare_fields(got, {k: gold[k] for k in got}, ncol, tol, name)


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
