"""GPU tests (-m gpu): the CUDA path, called through the C ABI by the Python
host wrapper, against (a) the golden vectors the unmodified reference produced
for every evaluation of its own 42-test suite, (b) the Stamnes DISORT 4.0.99
fixtures with the reference suite's pass criteria, (c) reference outputs on
subsets of the synthetic ensembles, (d) the oracle run live on more columns.

Tolerance: max|X - X_ref| <= 1e-9 * scale per column and field;
for omega > 0.9999 the tolerance grows like 1e-13/(1-omega) because the
reference itself is only reproducible to that level there."""
import warnings

import numpy as np
import pytest

import golden_io
import parity_suite
import pythonic_disort_b200 as pd
from pythonic_disort_b200 import _lib, api, synthetic

pytestmark = pytest.mark.gpu
SUITE = golden_io.suite_names()


@pytest.fixture(scope="module", autouse=True)
def cuda_backend():
    import torch
    assert torch.cuda.is_available()
    lib, dev = api._backend()
    assert dev.type == "cuda" and not hasattr(lib, "pd_is_hostsim"), "GPU tests must run the CUDA library, not the host build"
    yield


@pytest.mark.parametrize("name", SUITE)
def test_reference_suite_outputs(name):
    parity_suite.check_suite_record_outputs(pd.pydisort, name)


@pytest.mark.parametrize("name", [n for n in SUITE if golden_io.load_test(n)[1]])
def test_stamnes_criteria(name):
    parity_suite.check_stamnes(pd.pydisort, name)


@pytest.mark.parametrize("name", ["sw", "lw", "ha", "tp9c16"])
def test_ensembles_vs_reference_golden(name):
    parity_suite.check_ensemble_vs_golden(pd.pydisort, name)


@pytest.mark.parametrize("name", ["ha", "sw"])
def test_interpolate_at_user_polar_angles_on_the_device(name):
    """Row f1 (subroutines.py:614-705): pd_interp_mu against the reference's own interpolate()."""
    parity_suite.check_interpolate_vs_golden(pd, name)


def test_actinic_fluxes_vs_reference_golden():
    """Row a13 (subroutines.py:258-318, _assemble_intensity_and_fluxes.py:360-371): actinic fluxes from k_eval_u0
    with the delta-scaling reclassification, against the unmodified reference's (tests/golden/actinic.npz)."""
    parity_suite.check_actinic_vs_golden(pd)


def test_thermal_source_inputs_on_the_device():
    """Row f2 (subroutines.py:322-454): pd_planck_band / pd_s_poly_coeffs against the reference's helpers; the
    coefficients stay on the device and feed pydisort() directly."""
    import torch
    parity_suite.check_thermal_inputs_vs_golden(pd)
    ens = synthetic.make("lw", 32)
    tau = torch.as_tensor(ens["args"][0], device="cuda")
    temper = torch.linspace(210.0, 290.0, tau.shape[1] + 1, device="cuda", dtype=torch.float64).repeat(32, 1)
    sp_dev = pd.subroutines.generate_s_poly_coeffs(tau, temper, 600.0, 700.0)
    sp_host = pd.subroutines.generate_s_poly_coeffs(ens["args"][0], temper.cpu().numpy(), 600.0, 700.0)
    kw = dict(ens["kwargs"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = pd.pydisort(*ens["args"], **{**kw, "s_poly_coeffs": sp_dev})[1](ens["tau_eval"])
        b = pd.pydisort(*ens["args"], **{**kw, "s_poly_coeffs": sp_host})[1](ens["tau_eval"])
    np.testing.assert_allclose(parity_suite.to_np(a), b, rtol=1e-7)


def test_hapke_fourier_modes_on_the_device_and_per_column_beam_directions():
    """Row f4 (pydisotest/6_test.py:11-24, :193-201): pd_hapke_modes against the reference-style quad_vec tables, and
    the per-column use: every column its own mu0, one batched call with per-column tables against column-by-column
    calls with that column's tables."""
    import torch
    parity_suite.check_hapke_modes_vs_golden(pd)
    B, L, NQuad = 6, 3, 16
    rng = np.random.default_rng(7)
    tau = np.cumsum(rng.uniform(0.1, 0.6, (B, L)), axis=1)
    omega = rng.uniform(0.2, 0.9, (B, L))
    Leg = (rng.uniform(0.3, 0.7, (B, L))[:, :, None]) ** np.arange(NQuad + 1)[None, None, :]
    mu0 = rng.uniform(0.25, 0.95, B)
    modes = pd.subroutines.hapke_BDRF_Fourier_modes(NQuad // 2, NQuad, torch.as_tensor(mu0))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = pd.pydisort(tau, omega, NQuad, Leg, mu0, np.full(B, 3.0), 0.0, BDRF_Fourier_modes=modes)
        Fp = out[1](tau)
        for b in range(B):
            mine = pd.subroutines.hapke_BDRF_Fourier_modes(NQuad // 2, NQuad, float(mu0[b]))
            one = pd.pydisort(tau[b], omega[b], NQuad, Leg[b], mu0[b], 3.0, 0.0, BDRF_Fourier_modes=mine)
            np.testing.assert_allclose(Fp[b], one[1](tau[b]), rtol=1e-12)
            trap = synthetic.hapke_fourier_tables(NQuad // 2, NQuad, mu0[b], n_phi=8192)  # independent rule, 1/n^2 on the diagonal
            two = pd.pydisort(tau[b], omega[b], NQuad, Leg[b], mu0[b], 3.0, 0.0, BDRF_Fourier_modes=trap)
            np.testing.assert_allclose(Fp[b], two[1](tau[b]), rtol=1e-6)


@pytest.fixture(scope="module")
def oracle_pool():
    """Host cores for the live oracle (BLAS threads pinned to one per process by conftest)."""
    import multiprocessing as mp
    import os
    with mp.get_context("spawn").Pool(min(os.cpu_count() or 1, 32)) as pool:
        yield pool


@pytest.mark.parametrize("name,ncol,first", [("sw", 256, 1000), ("lw", 1024, 5000), ("ha", 64, 700), ("tp1", 6, 0),
                                             ("tp9c", 2, 0)])
def test_ensembles_vs_live_oracle(name, ncol, first, oracle_pool):
    """SURVEY.md 8(d) sample sizes (256 SW / 1,024 LW / 64 HA columns): the CUDA path against the pinned oracle run
    live on the host cores, both 8(c) criteria per column and field (parity_suite.compare_fields)."""
    parity_suite.check_ensemble_vs_live_oracle(pd.pydisort, name, ncol, first, pool=oracle_pool)


def test_nquad32_with_thermal_source_and_sparse_phase_functions():
    """NQuad = 32 outside the HA ensemble: the eight-lane symmetric eigen kernel (pd_stage_a_sym16.cuh) hands the
    mode-0 items of a thermal problem to the general kernel, takes the shortcut for layers without scattering
    (_solve_for_gen_and_part_sols.py:119, :162-168) and pads the last CTA when the item count is not a multiple of
    its eight items; checked against the oracle."""
    from oracle import disort_oracle
    rng = np.random.default_rng(11)
    B, L, NQuad = 3, 7, 32
    tau = np.cumsum(rng.uniform(0.05, 0.8, (B, L)), axis=1)
    omega = rng.uniform(0.1, 0.95, (B, L))
    omega[:, 2] = 0.0          # a purely absorbing layer: every mode takes the shortcut
    g = rng.uniform(0.2, 0.8, (B, L))
    Leg = g[:, :, None] ** np.arange(NQuad + 1)[None, None, :]
    Leg[0, 4, 6:] = 0.0        # a phase function with few moments: the high modes of that layer take the shortcut
    planck = np.sort(rng.uniform(20, 90, (B, L + 1)), axis=1)
    lev = np.concatenate([np.zeros((B, 1)), tau], axis=1)
    slope = np.diff(planck, axis=1) / np.diff(lev, axis=1)
    s_poly = np.stack([planck[:, :-1] - slope * lev[:, :-1], slope], axis=2)
    mu0 = np.array([0.35, 0.62, 0.9])
    kw = dict(s_poly_coeffs=s_poly, b_pos=0.9 * planck[:, -1:], BDRF_Fourier_modes=[0.1])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = pd.pydisort(tau, omega, NQuad, Leg, mu0, np.full(B, 2.0), 0.0, **kw)
        Fp, (Fd, Fdir), u = out[1](lev), out[2](lev), out[4](lev, np.array([0.0, 2.0]))
        for b in range(B):
            ref = disort_oracle.pydisort(tau[b], omega[b], NQuad, Leg[b], mu0[b], 2.0, 0.0, s_poly_coeffs=s_poly[b],
                                         b_pos=float(0.9 * planck[b, -1]), BDRF_Fourier_modes=[0.1])
            scale = np.max(np.abs(ref[1](lev[b])))
            assert np.max(np.abs(Fp[b] - ref[1](lev[b]))) <= 1e-9 * scale
            assert np.max(np.abs(Fd[b] - ref[2](lev[b])[0])) <= 1e-9 * scale
            ur = ref[4](lev[b], np.array([0.0, 2.0]))
            assert np.max(np.abs(u[b] - ur)) <= 1e-9 * np.max(np.abs(ur))


def test_batched_equals_column_by_column():
    ens = synthetic.make("sw", 5, 77)
    got = parity_suite.run_batched(pd.pydisort, ens)
    for b in range(5):
        args, kwargs = synthetic.column_call(ens, b)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = pd.pydisort(*args, **kwargs)
        np.testing.assert_allclose(out[1](ens["tau_eval"][b]), got["flux_up"][b], rtol=1e-13, atol=0)
        np.testing.assert_allclose(out[4](ens["tau_eval"][b], ens["phi_eval"]), got["u"][b], rtol=1e-12, atol=1e-15)


def test_cuda_tensor_inputs_give_cuda_outputs():
    import torch
    ens = synthetic.make("lw", 16)
    args = tuple(torch.as_tensor(a, device="cuda") if isinstance(a, np.ndarray) else a for a in ens["args"])
    kw = dict(ens["kwargs"])
    kw["s_poly_coeffs"] = torch.as_tensor(kw["s_poly_coeffs"], device="cuda")
    out = pd.pydisort(*args, **kw)
    Fp = out[1](torch.as_tensor(ens["tau_eval"], device="cuda"))
    assert Fp.is_cuda and Fp.shape == (16, 61)
    ref = parity_suite.run_batched(pd.pydisort, ens)
    np.testing.assert_allclose(Fp.cpu().numpy(), ref["flux_up"], rtol=1e-14)


def _generic(*args, **kwargs):
    """pydisort() on the size-generic kernels (pd_config.flags test bit, include/pydisort_b200.h)."""
    return pd.pydisort(*args, _kernel_flags=_lib.PD_FLAG_GENERIC_KERNELS, **kwargs)


@pytest.mark.parametrize("name", ["sw", "lw", "tp9c"])
def test_generic_kernels_meet_the_same_bar(name):
    """The production shapes have specialised kernels (symmetric eigen stage, interface-radiance elimination,
    tabulated NT); the size-generic ones (Hessenberg-QR, pivoted band solver, per-output recurrences) serve every
    other NQuad and are the fallback.  Both must reproduce the reference."""
    if name == "tp9c":
        from oracle import disort_oracle
        ens = synthetic.make(name, 2)
        got = parity_suite.run_batched(_generic, ens)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = synthetic.run_reference_like(disort_oracle.pydisort, ens)
        parity_suite.compare_fields(got, ref, 2, 1e-9, name)
    else:
        parity_suite.check_ensemble_vs_golden(_generic, name)


@pytest.mark.parametrize("name,ncol,first", [("sw", 4096, 20000), ("ha", 96, 500), ("lw", 8192, 100000)])
def test_production_kernels_agree_with_the_generic_ones_on_large_slices(name, ncol, first):
    """Thousands of columns the oracle would take minutes for: the specialised kernels (symmetric stage A, pivot-free
    elimination over interface radiances, tabulated NT) against the size-generic ones (Hessenberg-QR, pivoted band
    elimination of the coefficients, per-output recurrences).  Two independent formulations of the same equations:
    they must agree within the parity bar on every column (typical agreement is 1e-13; columns where 1/mu0 falls
    next to an eigenvalue k of a layer are conditioned like 1/(1/mu0^2 - k^2) and move by ~1e-10 between ANY two
    implementations, the oracle included).  NOT a parity test: it compares the CUDA path with itself."""
    ens = synthetic.make(name, ncol, first)
    got = parity_suite.run_batched(pd.pydisort, ens)
    ref = parity_suite.run_batched(_generic, ens)
    tol = golden_io.conditioning_tolerance(ens["args"][1], 1e-9)
    worst, _, _ = parity_suite.compare_fields(got, ref, ncol, tol, name + " production vs generic kernels")
    assert worst < tol


@pytest.mark.parametrize("name,ncol,first", [("sw", 512, 3000), ("lw", 4096, 7000), ("ha", 48, 900), ("tp1", 6, 0), ("tp9c", 2, 0)])
def test_interface_levels_come_from_the_sweep_and_agree_with_the_assembled_solution(name, ncol, first):
    """Levels that are layer interfaces are read from the sweep's interface radiances (pd_state.Uif); the same
    levels assembled from G, C, and grids that mix interfaces with interior points, agree to 1e-11 of scale."""
    parity_suite.check_interface_levels_vs_assembled(pd.pydisort, name, ncol, first)


@pytest.mark.parametrize("name,ncol,chunk", [("sw", 96, 40), ("lw", 1000, 300), ("ha", 6, 4)])
def test_inputs_described_per_layer_and_expanded_on_the_device(name, ncol, chunk):
    """inputs.HenyeyGreenstein / LevelSource (pd_hg_moments, pd_level_source) in place of the arrays: same results to the
    rounding of the moments, through pydisort(), solve_ensemble() and the sharding rule."""
    parity_suite.check_compact_inputs(pd, name, ncol, chunk, first=4000)


def test_unphysical_phase_function_is_flagged_not_crashed():
    """Moments that make the reduced matrices indefinite: the symmetric path must hand the item to the general
    solver, which reports the non-positive k^2 (the reference returns NaN / complex garbage here)."""
    leg = np.array([1.0, 0.99, -0.99, 0.99, -0.99])
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        out = pd.pydisort(np.array([0.5, 1.0]), np.array([0.999, 0.999]), 4, np.tile(leg, (2, 1)), 0.5, 1.0, 0.0)
        out[1](0.3)
    assert any("numerical trouble" in str(x.message) for x in w)


def test_large_batch_properties():
    """Full-size-style checks that need no oracle: energy conservation for a conservative-ish, black-surface,
    beam-only ensemble slice (F_up(0) + F_down_total(tau_L) <= incident, >= 0) and monotone direct beam."""
    ens = synthetic.make("sw", 2048, first=30000)
    ens["kwargs"]["BDRF_Fourier_modes"] = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = pd.pydisort(*ens["args"], **ens["kwargs"])
    t = ens["tau_eval"]
    Fp = out[1](t)
    Fd, Fdir = out[2](t)
    incident = ens["args"][4] * ens["args"][5]  # mu0 * I0
    assert np.all(Fp >= -1e-12) and np.all(Fd >= -1e-9) and np.all(np.diff(Fdir, axis=1) <= 0)
    absorbed = incident - Fp[:, 0] - (Fd[:, -1] + Fdir[:, -1])
    assert np.all(absorbed >= -1e-9 * incident) and np.all(absorbed <= incident)
    np.testing.assert_allclose(Fdir[:, 0], incident, rtol=1e-14)
