"""CPU tests of the Python boundary (host build of the kernels underneath):
the reference's input checks and warnings (pydisort.py:223-291), output shapes
and squeezing (_assemble_intensity_and_fluxes.py:522,605), batch conventions,
the helper functions re-exported as `subroutines`."""
import warnings
from math import pi

import numpy as np
import pytest

import hostsim_backend
import pythonic_disort_b200 as pd
from pythonic_disort_b200 import subroutines as sub


@pytest.fixture(scope="module", autouse=True)
def host_build():
    with hostsim_backend.use():
        yield


LEG = 0.7 ** np.arange(9)
BASE = dict(tau_arr=np.array([0.5, 1.0, 2.0]), omega_arr=np.array([0.9, 0.8, 0.7]), NQuad=8,
            Leg_coeffs_all=np.tile(LEG, (3, 1)), mu0=0.6, I0=1.0, phi0=0.3)


def call(**over):
    kw = dict(BASE)
    kw.update(over)
    pos = [kw.pop(k) for k in ("tau_arr", "omega_arr", "NQuad", "Leg_coeffs_all", "mu0", "I0", "phi0")]
    return pd.pydisort(*pos, **kw)


@pytest.mark.parametrize("over,msg", [
    (dict(tau_arr=np.array([0.5, -1.0, 2.0])), "tau values cannot be non-positive"),
    (dict(tau_arr=np.array([0.5, 0.4, 2.0])), "Layer thicknesses cannot be non-positive"),
    (dict(omega_arr=np.array([0.9, 1.0, 0.7])), "Single-scattering albedo"),
    (dict(omega_arr=np.array([0.9, 0.7])), "omega_arr"),
    (dict(Leg_coeffs_all=np.tile(LEG, (2, 1))), "Leg_coeffs_all"),
    (dict(Leg_coeffs_all=np.tile(np.array([1, 1.2, 0, 0, 0, 0, 0, 0, 0.0]), (3, 1))), "between -1 and 1"),
    (dict(NQuad=7), "even"),
    (dict(NQuad=0, NLeg=4, NFourier=1), "at least two streams"),
    (dict(NLeg=0), "must be positive"),
    (dict(NLeg=12), "cannot be larger"),
    (dict(NFourier=0), "must be positive"),
    (dict(NFourier=9, NLeg=8), "less than or equal"),
    (dict(I0=-1.0), "cannot be negative"),
    (dict(mu0=1.5), "between 0 and 1"),
    (dict(phi0=7.0), "principal azimuthal angle"),
    (dict(b_pos=np.ones(3)), "bottom boundary condition"),
    (dict(b_neg=np.ones((2, 2))), "top boundary condition"),
    (dict(f_arr=np.array([0.1, 2.0, 0.1])), "fractional scattering"),
    (dict(f_arr=np.array([0.1, 0.2])), "f_arr"),
    (dict(use_banded_solver_NLayers=2), "use_banded_solver_NLayers"),
    (dict(s_poly_coeffs=np.ones((2, 2))), "s_poly_coeffs"),
])
def test_value_errors_match_reference(over, msg):
    with pytest.raises(ValueError, match=msg):
        call(**over)


def test_nt_refuses_mu0_on_a_quadrature_node():
    mu = sub.Gauss_Legendre_quad(4)[0]
    with pytest.raises(ValueError, match="too close to `mu0`"):
        call(mu0=float(mu[1]), NT_cor=True, f_arr=np.full(3, LEG[8]), NLeg=8,
             Leg_coeffs_all=np.tile(0.7 ** np.arange(12), (3, 1)))


def test_warnings_match_reference():
    leg = np.tile(LEG, (3, 1)).copy()
    leg[:, 0] = 0.99
    with pytest.warns(UserWarning, match="corrected to, 1"):
        call(Leg_coeffs_all=leg)
    np.testing.assert_array_equal(leg[:, 0], 0.99)  # unlike the reference, user input is not mutated
    with pytest.warns(UserWarning, match="very close to 1"):
        call(omega_arr=np.array([0.9, 1 - 1e-7, 0.7]))
    with pytest.raises(NotImplementedError):
        call(autograd_compatible=True)


def test_unbatched_shapes_follow_the_reference():
    mu_arr, Fp, Fm, u0, u = call()
    assert mu_arr.shape == (8,) and np.all(mu_arr[:4] > 0) and np.all(mu_arr[4:] < 0)
    assert np.ndim(Fp(0.3)) == 0 and Fp(np.array([0.1, 0.2])).shape == (2,)
    d, r = Fm(np.array([0.1, 0.2, 2.0]))
    assert d.shape == (3,) and r.shape == (3,)
    assert u0(0.3).shape == (8,) and u0(np.array([0.1, 0.2])).shape == (8, 2)
    assert u(0.3, 0.0).shape == (8,) and u(np.array([0.1, 0.2]), 0.0).shape == (8, 2)
    assert u(np.array([0.1, 0.2]), np.array([0.0, 1.0, 2.0])).shape == (8, 2, 3)
    val, tau_back = Fp(0.3, return_tau_arr=True)
    assert tau_back is BASE["tau_arr"]
    val, err = u(np.array([0.1]), np.array([0.0, 1.0]), return_Fourier_error=True)
    assert val.shape == (8, 2) and np.ndim(err) == 0 and err >= 0
    with pytest.raises(ValueError, match="outside the tau range"):
        Fp(2.5)
    with pytest.raises(ValueError, match="outside the tau range"):
        u(-0.1, 0.0)
    assert len(call(only_flux=True)) == 4


def test_interface_belongs_to_the_upper_layer_and_outputs_are_continuous():
    _, Fp, Fm, u0, u = call()
    eps = 1e-12
    for t in (0.5, 1.0):
        np.testing.assert_allclose(Fp(t), Fp(t + eps), rtol=1e-9)
        np.testing.assert_allclose(u(t, 0.5), u(t + eps, 0.5), rtol=1e-8, atol=1e-12)


def test_batch_mode_broadcasts_shared_inputs():
    B = 3
    tau = np.tile(BASE["tau_arr"], (B, 1)) * np.array([1.0, 1.5, 2.0])[:, None]
    out = pd.pydisort(tau, BASE["omega_arr"], 8, BASE["Leg_coeffs_all"], np.array([0.6, 0.5, 0.4]), 1.0, 0.3,
                      b_pos=np.array([0.0, 0.1, 0.2]), BDRF_Fourier_modes=[np.array([0.1, 0.2, 0.3])])
    Fp = out[1](np.array([0.0, 0.25]))
    assert Fp.shape == (B, 2)
    assert out[4](np.array([0.0, 0.25]), np.array([0.0, 1.0])).shape == (B, 8, 2, 2)
    assert out[4](0.1, 0.0).shape == (B, 8)
    assert out[1](tau[:, :2]).shape == (B, 2)  # per-column query depths
    for b in range(B):  # each column equals the corresponding single-column call
        one = pd.pydisort(tau[b], BASE["omega_arr"], 8, BASE["Leg_coeffs_all"], [0.6, 0.5, 0.4][b], 1.0, 0.3,
                          b_pos=[0.0, 0.1, 0.2][b], BDRF_Fourier_modes=[[0.1, 0.2, 0.3][b]])
        np.testing.assert_allclose(one[1](np.array([0.0, 0.25])), Fp[b], rtol=1e-13)


def test_actinic_flux_and_interpolation_helpers():
    _, Fp, Fm, u0, u = call(f_arr=np.full(3, LEG[8]) * 0 + 0.05)
    up, down = sub.generate_diff_act_flux_funcs(u0)
    t = np.array([0.0, 0.7, 2.0])
    assert up(t).shape == (3,) and down(t).shape == (3,)
    ui = sub.interpolate(u)
    mu_arr = np.concatenate([sub.Gauss_Legendre_quad(4)[0], -sub.Gauss_Legendre_quad(4)[0]])
    np.testing.assert_allclose(ui(mu_arr, t, 0.5), u(t, 0.5), rtol=1e-10, atol=1e-14)
    assert ui(np.array([0.3, -0.3]), t, np.array([0.0, 1.0])).shape == (2, 3, 2)
    u0i = sub.interpolate(u0)
    np.testing.assert_allclose(u0i(mu_arr, t), u0(t), rtol=1e-10, atol=1e-14)


def test_subroutine_helpers_against_closed_forms():
    x, w = sub.Gauss_Legendre_quad(6)
    np.testing.assert_allclose(w.sum(), 1.0, rtol=1e-14)
    np.testing.assert_allclose((w * x**3).sum(), 0.25, rtol=1e-13)
    xc, wc = sub.Clenshaw_Curtis_quad(33)
    np.testing.assert_allclose((wc * np.cos(xc) ** 2).sum(), pi, rtol=1e-12)
    c = sub.affine_transform_poly_coeffs(np.array([[1.0, 2.0, 3.0]]), np.array([2.0]), np.array([0.5]))
    y = 1.7
    xx = (y - 0.5) / 2.0
    np.testing.assert_allclose(c[0] @ y ** np.arange(3), 1 + 2 * xx + 3 * xx**2, rtol=1e-13)
    s = sub.linear_spline_coefficients(np.array([0.0, 1.0, 3.0]), np.array([1.0, 3.0, 2.0]))
    np.testing.assert_allclose(s, [[1.0, 2.0], [3.5, -0.5]])
    assert sub.generate_emissivity_from_BDRF(4, 0.3) == 0.7
    np.testing.assert_allclose(sub.calculate_nu(1.0, 0.0, 1.0, 0.0), 1.0)


def test_legendre_table_matches_scipy():
    import scipy.special
    from pythonic_disort_b200.api import norm_assoc_legendre_table
    x = np.array([0.05, 0.3, 0.77, -0.6])
    tab = norm_assoc_legendre_table(12, 16, x)
    for m in range(12):
        for l in range(m, 16):
            ref = scipy.special.lpmv(m, l, x) * np.sqrt(scipy.special.poch(l + m + 1, -2.0 * m))
            np.testing.assert_allclose(tab[m, l], ref, rtol=2e-13, atol=1e-300)
