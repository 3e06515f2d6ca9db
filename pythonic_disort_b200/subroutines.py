"""Host-side helper functions that users of the reference import from
``PythonicDISORT.subroutines`` (reference: src/PythonicDISORT/subroutines.py).

These only *prepare inputs* for ``pydisort`` or *post-process* its output
functions; none of them is on the solver hot path, so they stay NumPy/SciPy on
the host.  Names, argument order and return
conventions follow the reference so that existing scripts keep working.
"""
import warnings
from math import pi

import numpy as np
import scipy.constants
import scipy.fft
import scipy.integrate
import scipy.interpolate

__all__ = [
    "prepend", "transform_interval", "transform_weights", "calculate_nu", "Gauss_Legendre_quad",
    "Clenshaw_Curtis_quad", "atleast_2d_append", "generate_diff_act_flux_funcs", "Planck",
    "blackbody_contrib_to_BCs", "linear_spline_coefficients", "generate_s_poly_coeffs",
    "generate_emissivity_from_BDRF", "cache_BDRF_Fourier_modes", "affine_transform_poly_coeffs",
    "interpolate", "TabulatedBDRF", "hapke_BDRF_Fourier_modes",
]


def prepend(arr, arr_len, value):
    """``[value, *arr]`` as a new float array (subroutines.py:7-29)."""
    out = np.empty(arr_len + 1)
    out[0] = value
    out[1:] = arr
    return out


def transform_interval(arr, c, d, a, b):
    """Map points from [a, b] to [c, d] (subroutines.py:33-55)."""
    return (arr - a) * (d - c) / (b - a) + c


def transform_weights(weights, c, d, a, b):
    """Map quadrature weights from [a, b] to [c, d] (subroutines.py:59-81)."""
    return weights * (d - c) / (b - a)


def calculate_nu(mu, phi, mu_p, phi_p):
    """Cosine of the scattering angle; axes (mu, phi, mu_p, phi_p), squeezed
    (subroutines.py:85-112)."""
    mu, phi, mu_p, phi_p = (np.atleast_1d(x) for x in (mu, phi, mu_p, phi_p))
    a = mu[:, None, None, None] * mu_p[None, None, :, None]
    s = np.sqrt(1 - mu**2)[:, None, None, None] * np.sqrt(1 - mu_p**2)[None, None, :, None]
    return np.squeeze(a + s * np.cos(phi_p[None, None, None, :] - phi[None, :, None, None]))


def Gauss_Legendre_quad(N, c=0, d=1):
    """Gauss-Legendre nodes and weights on [c, d] (subroutines.py:116-138)."""
    x, w = np.polynomial.legendre.leggauss(int(N))
    return transform_interval(x, c, d, -1, 1), transform_weights(w, c, d, -1, 1)


def Clenshaw_Curtis_quad(Nphi, c=0, d=(2 * pi)):
    """Clenshaw-Curtis nodes and weights on [c, d]; ``Nphi`` odd and > 2
    (subroutines.py:142-175)."""
    if not (Nphi > 2 and Nphi % 2 == 1):
        raise ValueError("The number of quadrature nodes must be odd and greater than 2.")
    n = Nphi - 1
    half = n // 2
    right = np.cos(pi * np.arange(half) / n)
    nodes = np.concatenate([-right, [0.0], right[::-1]])
    moments = prepend(2 / (1 - 4 * np.arange(1, half + 1) ** 2), half, 2)
    w_half = scipy.fft.idct(moments, type=1)
    w_half[0] /= 2
    weights = np.concatenate([w_half, w_half[:-1][::-1]])
    return transform_interval(nodes, c, d, -1, 1), transform_weights(weights, c, d, -1, 1)


def atleast_2d_append(*arys):
    """Like ``numpy.atleast_2d`` but new axes go to the back (subroutines.py:219-254)."""
    out = []
    for a in arys:
        a = np.asanyarray(a)
        if a.ndim == 0:
            a = a.reshape(1, 1)
        elif a.ndim == 1:
            a = a[:, None]
        out.append(a)
    return out[0] if len(out) == 1 else out


def generate_diff_act_flux_funcs(u0):
    """Upward and downward-diffuse actinic flux functions from ``u0``
    (subroutines.py:258-318).  Works with the unbatched and the batched ``u0``
    returned by :func:`pythonic_disort_b200.pydisort` (the stream axis is the
    one before the tau axis in both)."""
    batched = getattr(u0, "batched", False)
    N = np.shape(_to_numpy(u0(0)))[1 if batched else 0] // 2
    W = Gauss_Legendre_quad(N)[1]

    def _contract(u0_val, lo, hi):
        a = _to_numpy(u0_val)
        ax = 1 if batched else 0  # stream axis
        sl = [slice(None)] * a.ndim
        sl[ax] = slice(lo, hi)
        return 2 * pi * np.tensordot(W, a[tuple(sl)], axes=([0], [ax]))

    def flux_act_up(tau, is_antiderivative_wrt_tau=False, return_tau_arr=False):
        if return_tau_arr:
            val, tau_arr = u0(tau, is_antiderivative_wrt_tau, True)
            return np.squeeze(_contract(val, 0, N))[()], tau_arr
        return np.squeeze(_contract(u0(tau, is_antiderivative_wrt_tau), 0, N))[()]

    def flux_act_down_diffuse(tau, is_antiderivative_wrt_tau=False, return_tau_arr=False):
        if return_tau_arr:
            val, tau_arr, recl = u0(tau, is_antiderivative_wrt_tau, True, _return_act_dscale_for_reclass=True)
            return np.squeeze(_contract(val, N, 2 * N) + _to_numpy(recl))[()], tau_arr
        val, recl = u0(tau, is_antiderivative_wrt_tau, False, _return_act_dscale_for_reclass=True)
        return np.squeeze(_contract(val, N, 2 * N) + _to_numpy(recl))[()]

    return flux_act_up, flux_act_down_diffuse


def _to_numpy(x):
    if hasattr(x, "detach"):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def Planck(T, WVNM):
    """Planck emission (W m^-2 per m^-1 wavenumber), DISORT units
    (subroutines.py:322-350)."""
    T = np.atleast_1d(np.asarray(T, dtype=float))
    out = np.zeros(len(T))
    hot = T != 0
    if np.any(hot):
        h, c, k = scipy.constants.h, scipy.constants.c, scipy.constants.k
        e = np.exp(-100 * h * c * WVNM / (k * T[hot]))
        out[hot] = (2e8 * h * c**2 * WVNM**3 * e) / (1 - e)
    return np.squeeze(out)[()]


def _is_tensor(x):
    return hasattr(x, "detach") and hasattr(x, "data_ptr")


def blackbody_contrib_to_BCs(T, WVNMLO, WVNMHI, **kwargs):
    """Band-integrated blackbody emission of a boundary (subroutines.py:354-377).

    A torch tensor of temperatures (any shape, e.g. one per column) is integrated on the GPU (``pd_planck_band``,
    fixed Gauss-Legendre panels in ``h c nu / k T``) and a tensor of the same shape comes back; anything else takes
    the reference's route (``scipy.integrate.quad_vec`` on the host)."""
    if _is_tensor(T):
        from . import api
        return api.planck_band(T, WVNMLO, WVNMHI)
    return np.squeeze(scipy.integrate.quad_vec(lambda w: Planck(T, w), WVNMLO, WVNMHI, **kwargs)[0])


def linear_spline_coefficients(x, y, check_inputs=True):
    """(intercept, slope) of each linear segment through (x, y)
    (subroutines.py:381-409)."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    if check_inputs:
        if not len(x) > 1:
            raise ValueError("At least 2 points are required.")
        if not len(x) == len(y):
            raise ValueError("The number of x and y points must be equal.")
        if not np.all(np.diff(x) > 0):
            raise ValueError("The x values must be sorted in ascending order.")
    slope = np.diff(y) / np.diff(x)
    return np.stack([y[:-1] - slope * x[:-1], slope], axis=1)


def generate_s_poly_coeffs(tau_arr, TEMPER, WVNMLO, WVNMHI, **kwargs):
    """DISORT-equivalent linear-in-tau thermal source coefficients from level
    temperatures (subroutines.py:413-454).

    Torch tensors ``tau_arr`` [B, NLayers] / ``TEMPER`` [B, NLayers + 1] (or the unbatched 1-D pair) are processed
    on the GPU (``pd_s_poly_coeffs``) and the coefficients stay there, ready to be passed to ``pydisort`` as
    ``s_poly_coeffs``; NumPy input takes the reference's host route (2-D input: one column per row)."""
    if _is_tensor(tau_arr) or _is_tensor(TEMPER):
        import torch
        from . import api
        return api.s_poly_coeffs(torch.as_tensor(tau_arr), torch.as_tensor(TEMPER), WVNMLO, WVNMHI)
    tau_arr = np.atleast_1d(tau_arr)
    TEMPER = np.asarray(TEMPER, dtype=float)
    if tau_arr.ndim == 2:  # batch on the host: one integration over all temperatures, one spline per column
        if TEMPER.shape != (tau_arr.shape[0], tau_arr.shape[1] + 1):
            raise ValueError("Missing temperature specification at some boundaries / interfaces.")
        em = scipy.integrate.quad_vec(lambda w: Planck(TEMPER.ravel(), w), WVNMLO, WVNMHI, **kwargs)[0]
        em = np.reshape(em, TEMPER.shape)
        lev = np.concatenate([np.zeros((tau_arr.shape[0], 1)), tau_arr], axis=1)
        slope = np.diff(em, axis=1) / np.diff(lev, axis=1)
        return np.stack([em[:, :-1] - slope * lev[:, :-1], slope], axis=2)
    if not len(TEMPER) == len(tau_arr) + 1:
        raise ValueError("Missing temperature specification at some boundaries / interfaces.")
    levels = prepend(tau_arr, len(tau_arr), 0)
    emission = scipy.integrate.quad_vec(lambda w: Planck(TEMPER, w), WVNMLO, WVNMHI, **kwargs)[0]
    return linear_spline_coefficients(levels, emission, check_inputs=False)


def generate_emissivity_from_BDRF(N, zeroth_BDRF_Fourier_mode):
    """Directional emissivity from the zeroth BDRF Fourier mode by Kirchhoff's
    law (subroutines.py:459-486)."""
    if np.isscalar(zeroth_BDRF_Fourier_mode):
        return 1 - zeroth_BDRF_Fourier_mode
    mu, W = Gauss_Legendre_quad(N)
    return 1 - 2 * zeroth_BDRF_Fourier_mode(mu, mu) * mu[None, :] @ W


class TabulatedBDRF:
    """A BDRF Fourier mode tabulated at the quadrature nodes.

    ``q`` is ``q^m(mu_i, mu_j)`` (N x N) and ``q0`` is ``q^m(mu_i, mu0)`` (N, or
    N x n_mu0 for several beam directions).  Calling it with N incidence cosines
    returns ``q``; calling it with any other number returns the ``mu0`` columns.
    ``pydisort`` recognises instances and uses the tables directly."""

    def __init__(self, q, q0=None):
        if _is_tensor(q):  # device tables (e.g. from hapke_BDRF_Fourier_modes): kept as they are
            self.q = q
            self.q0 = None if q0 is None else q0.reshape(q.shape[0], -1)
        else:
            self.q = np.asarray(q, dtype=float)
            self.q0 = None if q0 is None else np.asarray(q0, dtype=float).reshape(self.q.shape[0], -1)
        self.nodes = Gauss_Legendre_quad(self.q.shape[0])[0]

    def __call__(self, mu, neg_mup):
        neg_mup = np.atleast_1d(neg_mup)
        if len(neg_mup) == len(self.nodes) and np.allclose(neg_mup, self.nodes, rtol=0, atol=1e-14):
            return self.q
        if self.q0 is None:
            raise ValueError("this BDRF mode was tabulated without a beam direction")
        return self.q0


def hapke_BDRF_Fourier_modes(N, NFourier, mu0, B0=1.0, HH=0.06, W=0.6, n_panels=64):
    """The first ``NFourier`` Fourier modes of the Hapke surface BDRF of DISORT's test problems
    (pydisotest/6_test.py:11-24), tabulated at the ``N`` quadrature nodes and at the beam cosine(s) ``mu0`` -- what a
    user of the reference builds with ``quad_vec`` over the relative azimuth (pydisotest/6_test.py:193-201) and then
    passes to ``pydisort`` / ``cache_BDRF_Fourier_modes``.  Here the integration runs on the GPU (``pd_hapke_modes``:
    ``n_panels`` 16-point Gauss-Legendre panels on [0, pi], where the integrand is analytic) and ``mu0`` may hold one value per column: the returned list of
    ``TabulatedBDRF`` (device tables) is a valid ``BDRF_Fourier_modes`` argument for a batched ``pydisort`` call."""
    import torch
    from . import api
    mu0_t = torch.atleast_1d(torch.as_tensor(mu0, dtype=torch.float64))
    nodes = torch.as_tensor(Gauss_Legendre_quad(N)[0], dtype=torch.float64)
    tab = api.hapke_modes(N, NFourier, torch.cat([nodes, mu0_t.reshape(-1).cpu()]), B0, HH, W, n_panels)  # [NF, N, N + k]
    return [TabulatedBDRF(tab[m, :, :N], tab[m, :, N:]) for m in range(NFourier)]


def cache_BDRF_Fourier_modes(N, BDRF_Fourier_modes, mu0=0):
    """Evaluate BDRF Fourier modes once at the quadrature nodes (and optionally
    at ``mu0``) so later ``pydisort`` calls reuse the tables
    (subroutines.py:490-570)."""
    with_mu0 = 0 < mu0 <= 1
    if not with_mu0:
        warnings.warn("No caching with respect to `mu0`.")
    mu = Gauss_Legendre_quad(N)[0]
    cached = []
    for mode in BDRF_Fourier_modes:
        if np.isscalar(mode):
            cached.append(lambda mu_, neg_mup, v=mode: v)
            continue
        if with_mu0:
            both = np.asarray(mode(mu, np.append(mu, mu0)))
            cached.append(TabulatedBDRF(both[:, :-1], both[:, [-1]]))
        else:
            table = np.asarray(mode(mu, mu))
            cached.append(lambda mu_, neg_mup, f=mode, t=table: f(mu, neg_mup) if len(neg_mup) == 1 else t)
    return cached


def affine_transform_poly_coeffs(poly_coeffs, a_arr, b_arr):
    """Coefficients of a polynomial in ``x`` re-expressed in ``y = a x + b``,
    one transformation per row (subroutines.py:574-610)."""
    poly_coeffs = np.atleast_2d(np.asarray(poly_coeffs, dtype=float))
    a_arr = np.asarray(a_arr, dtype=float)
    b_arr = np.asarray(b_arr, dtype=float)
    if np.any(a_arr == 0):
        raise ValueError("The scale factors must be non-zero.")
    n = poly_coeffs.shape[1]
    out = np.zeros_like(poly_coeffs)
    binom = np.ones(1)
    for j in range(n):                       # x^j = ((y - b)/a)^j
        for i in range(j + 1):
            out[:, i] += binom[i] * (1 / a_arr) ** j * (-b_arr) ** (j - i) * poly_coeffs[:, j]
        binom = np.concatenate([[1.0], binom[1:] + binom[:-1], [1.0]])
    return out


def interpolate(u):
    """Barycentric interpolation in ``mu`` of ``u`` or ``u0`` (each hemisphere
    separately), giving ``u(mu, tau, phi)`` / ``u0(mu, tau)``
    (subroutines.py:614-705).

    For the output functions of ``pythonic_disort_b200.pydisort`` the
    interpolation runs on the GPU (``pd_interp_mu``: the stream axis is
    contracted with the barycentric weight matrix right after the evaluation
    kernel, so only the ``[nmu, ntau, nphi]`` result leaves the device).  Any
    other callable with the reference's signature is interpolated on the host."""
    kind = getattr(u, "kind", None)
    if kind is None:
        kind = {5: "u", 4: "u0"}.get(u.__code__.co_argcount)
    if kind not in ("u", "u0"):
        raise ValueError("This subroutine can only interpolate u or u0.")
    batched = getattr(u, "batched", False)
    ax = 1 if batched else 0
    at_mu = getattr(u, "at_mu", None)

    def _check_mu(mu):
        if not np.all(np.abs(mu) <= 1):
            raise ValueError("mu values must be between -1 and 1.")
        return np.atleast_1d(mu)

    if at_mu is None:  # host path for foreign callables
        probe = _to_numpy(u(0, 0) if kind == "u" else u(0))
        N = probe.shape[ax] // 2
        mu_pos = Gauss_Legendre_quad(N)[0]
        up = scipy.interpolate.BarycentricInterpolator(mu_pos)
        dn = scipy.interpolate.BarycentricInterpolator(-mu_pos)

    def _interp(mu, values):
        mu = _check_mu(mu)
        vals = np.moveaxis(_to_numpy(values), ax, 0)
        out = np.empty((len(mu),) + vals.shape[1:])
        pos = mu > 0
        if np.any(pos):
            up.set_yi(vals[:N])
            out[pos] = up(mu[pos])
        if np.any(~pos):
            dn.set_yi(vals[N:])
            out[~pos] = dn(mu[~pos])
        return np.squeeze(np.moveaxis(out, 0, ax))[()]

    if kind == "u":
        def u_interpol(mu, tau, phi, is_antiderivative_wrt_tau=False, return_Fourier_error=False,
                       return_tau_arr=False):
            if at_mu is not None:
                val = at_mu(_check_mu(mu), tau, phi, is_antiderivative_wrt_tau)
                if return_Fourier_error or return_tau_arr:
                    res = u(tau, phi, is_antiderivative_wrt_tau, return_Fourier_error, return_tau_arr)
                    return (val,) + tuple(res[1:])
                return val
            if return_Fourier_error or return_tau_arr:
                res = u(tau, phi, is_antiderivative_wrt_tau, return_Fourier_error, return_tau_arr)
                return (_interp(mu, res[0]),) + tuple(res[1:])
            return _interp(mu, u(tau, phi, is_antiderivative_wrt_tau))
    else:
        def u_interpol(mu, tau, is_antiderivative_wrt_tau=False, return_tau_arr=False):
            if at_mu is not None:
                val = at_mu(_check_mu(mu), tau, is_antiderivative_wrt_tau)
                if return_tau_arr:
                    return (val,) + tuple(u(tau, is_antiderivative_wrt_tau, True)[1:])
                return val
            if return_tau_arr:
                res = u(tau, is_antiderivative_wrt_tau, return_tau_arr)
                return (_interp(mu, res[0]),) + tuple(res[1:])
            return _interp(mu, u(tau, is_antiderivative_wrt_tau))
    return u_interpol


def _compare(results, mu_to_compare, reorder_mu, flux_up, flux_down, u=None, verbose=False):
    """Pointwise comparison against stored Stamnes DISORT results, same return
    tuple as the reference helper (subroutines.py:866-975)."""
    taus = results["tau_test_arr"]

    def diff_ratio(ref, val):
        d = np.abs(ref - _to_numpy(val))
        return d, np.divide(d, ref, out=np.zeros_like(d), where=(ref != 0))

    fd = flux_down(taus)
    out = diff_ratio(results["flup"], flux_up(taus)) + diff_ratio(results["rfldn"], fd[0]) \
        + diff_ratio(results["rfldir"], fd[1])
    if u is not None:
        uu = results["uu"]
        mine = _to_numpy(u(taus, results["phi_arr"]))[reorder_mu].reshape(np.shape(uu))
        d = np.abs(uu - mine)[mu_to_compare]
        out = out + (d, np.divide(d, uu[mu_to_compare], out=np.zeros_like(d), where=(uu[mu_to_compare] != 0)))
    if verbose:
        for name, arr in zip(("flux_up", "flux_down diffuse", "flux_down direct", "intensity"), out[::2]):
            print(name, "max abs diff", np.max(arr, initial=0))
    return out
