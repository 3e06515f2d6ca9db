"""Seeded synthetic column ensembles and
the two single-column test problems used as benchmark/parity workloads.

Every ensemble is a pure function of (name, column index): column ``i`` of a
B-column ensemble is identical to column ``i`` of any larger one, so golden
vectors computed on the first few columns stay valid for the full-size run.
Host-side NumPy only; this is workload generation, not solver code.
"""
from math import pi

import numpy as np

from . import subroutines as sub
from .inputs import HenyeyGreenstein, LevelSource

SEEDS = {"sw": 20260101, "lw": 20260102, "ha": 20260103}


def _uniform_block(seed, first, ncol, per_col):
    """Rows [first, first+ncol) of the infinite (columns x per_col) U(0,1) table."""
    rng = np.random.Generator(np.random.PCG64(seed))
    rng.bit_generator.advance(first * per_col)  # one 64-bit draw per double
    return rng.random((ncol, per_col))


def hapke(mu, neg_mup, dphi, B0=1.0, HH=0.06, W=0.6):
    """Hapke surface BDRF used by DISORT's test problems (parameters as in
    pydisotest/6_test.py:11-24)."""
    mu = np.asarray(mu, dtype=float)[:, None]
    mup = np.asarray(neg_mup, dtype=float)[None, :]
    cos_a = np.clip(mu * mup - np.sqrt(1 - mu**2) * np.sqrt(1 - mup**2) * np.cos(dphi), -1, 1)
    alpha = np.arccos(cos_a)
    phase = 1 + cos_a / 2
    surge = B0 * HH / (HH + np.tan(alpha / 2))
    gam = np.sqrt(1 - W)
    h = lambda x: (1 + 2 * x) / (1 + 2 * x * gam)
    return W / 4 / (mu + mup) * ((1 + surge) * phase + h(mup) * h(mu) - 1)


def hapke_fourier_tables(N, n_modes, mu0, n_phi=4096):
    """q^m(mu_i, mu_j) and q^m(mu_i, mu0), m < n_modes, by the periodic
    trapezoid rule in delta-phi (the tables are *inputs*; both the reference
    and the GPU path receive exactly these numbers)."""
    mu = sub.Gauss_Legendre_quad(N)[0]
    cols = np.append(mu, mu0)
    dphi = 2 * pi * np.arange(n_phi) / n_phi
    vals = np.stack([hapke(mu, cols, d) for d in dphi], axis=0)            # (n_phi, N, N+1)
    modes = []
    for m in range(n_modes):
        qm = np.tensordot(np.cos(m * dphi), vals, axes=(0, 0)) * (2 * pi / n_phi) / ((1 + (m == 0)) * pi)
        modes.append(sub.TabulatedBDRF(qm[:, :-1], qm[:, -1]))
    return modes


def make(name, ncol, first=0):
    """Build ``ncol`` columns (starting at global column ``first``) of ensemble ``name``.

    Returns a dict with batched positional arguments (``args``), batched keyword
    arguments (``kwargs``) for :func:`pythonic_disort_b200.pydisort`, the level
    grid ``tau_eval`` [B, L+1] and azimuths ``phi_eval`` on which the benchmark
    evaluates, and ``outputs`` (which fields the workload asks for).  ``compact`` (ensembles only): the inputs that
    have a few-numbers-per-layer description (``inputs.HenyeyGreenstein`` / ``LevelSource``) in that form, to be passed
    in place of the arrays of the same name."""
    if name == "sw":
        L, NQuad, NLeg_all = 60, 16, 32
        U = _uniform_block(SEEDS[name], first, ncol, 3 * L + 3)
        dt = 0.2 + 0.8 * U[:, :L]
        T = 0.5 + 9.5 * U[:, 3 * L]
        dt = dt * (T / dt.sum(axis=1))[:, None]
        tau = np.cumsum(dt, axis=1)
        omega = 0.3 + 0.699 * U[:, L:2 * L]
        g = 0.6 + 0.25 * U[:, 2 * L:3 * L]
        Leg = g[:, :, None] ** np.arange(NLeg_all)[None, None, :]
        mu0 = 0.2 + 0.8 * U[:, 3 * L + 1]
        nodes = sub.Gauss_Legendre_quad(NQuad // 2)[0]
        close = np.min(np.abs(mu0[:, None] - nodes[None, :]), axis=1) < 1e-3
        mu0 = np.where(close, mu0 + 2.5e-3, mu0)
        albedo = 0.05 + 0.35 * U[:, 3 * L + 2]
        return dict(name=name, B=ncol, L=L, NQuad=NQuad,
                    args=(tau, omega, NQuad, Leg, mu0, pi / mu0, 0.0),
                    kwargs=dict(f_arr=Leg[:, :, NQuad].copy(), NT_cor=True, BDRF_Fourier_modes=[albedo]),
                    tau_eval=np.concatenate([np.zeros((ncol, 1)), tau], axis=1),
                    phi_eval=np.array([0.0, pi / 2, pi]), outputs=("flux", "u"),
                    compact=dict(Leg_coeffs_all=HenyeyGreenstein(g, NLeg_all)))
    if name == "lw":
        L, NQuad = 60, 8
        U = _uniform_block(SEEDS[name], first, ncol, 4 * L + 1)
        tau = np.cumsum(0.01 + 0.29 * U[:, :L], axis=1)
        omega = 0.6 * U[:, L:2 * L]
        g = 0.5 * U[:, 2 * L:3 * L]
        Leg = g[:, :, None] ** np.arange(NQuad + 1)[None, None, :]
        planck = np.sort(50 + 70 * U[:, 3 * L:4 * L + 1], axis=1)          # B_0..B_L, monotone
        lev = np.concatenate([np.zeros((ncol, 1)), tau], axis=1)
        slope = np.diff(planck, axis=1) / np.diff(lev, axis=1)
        s_poly = np.stack([planck[:, :-1] - slope * lev[:, :-1], slope], axis=2)
        return dict(name=name, B=ncol, L=L, NQuad=NQuad,
                    args=(tau, omega, NQuad, Leg, np.zeros(ncol), np.zeros(ncol), 0.0),
                    kwargs=dict(only_flux=True, s_poly_coeffs=s_poly, b_pos=0.98 * planck[:, -1:],
                                BDRF_Fourier_modes=[0.02]),
                    tau_eval=lev, phi_eval=None, outputs=("flux",),
                    compact=dict(Leg_coeffs_all=HenyeyGreenstein(g, NQuad + 1), s_poly_coeffs=LevelSource(planck)))
    if name == "ha":
        L, NQuad = 100, 32
        U = _uniform_block(SEEDS[name], first, ncol, 3 * L)
        tau = np.cumsum(0.005 + 0.045 * U[:, :L], axis=1)
        omega = 0.3 + 0.699 * U[:, L:2 * L]
        g = 0.5 + 0.3 * U[:, 2 * L:3 * L]
        Leg = g[:, :, None] ** np.arange(NQuad + 1)[None, None, :]
        mu0 = 0.6
        modes = hapke_fourier_tables(NQuad // 2, NQuad, mu0)
        return dict(name=name, B=ncol, L=L, NQuad=NQuad,
                    args=(tau, omega, NQuad, Leg, np.full(ncol, mu0), np.full(ncol, pi / mu0), 0.0),
                    kwargs=dict(BDRF_Fourier_modes=modes),
                    tau_eval=np.concatenate([np.zeros((ncol, 1)), tau], axis=1),
                    phi_eval=np.linspace(0, pi, 5), outputs=("flux", "u"),
                    mu_user=np.array([-0.9, -0.5, -0.1, 0.1, 0.5, 0.9]),
                    compact=dict(Leg_coeffs_all=HenyeyGreenstein(g, NQuad + 1)))
    if name in ("tp9c16", "tp9c"):
        # DISORT test problem 9c (pydisotest/9_test.py:175-236); "16" = the NQuad=16 variant
        NQuad = 16 if name == "tp9c16" else 8
        tau = np.array([1.0, 3, 6, 10, 15, 21])
        omega = 0.6 + np.arange(1, 7) * 0.05
        Leg = np.vstack([(l / 7) ** np.arange(NQuad + 1) for l in range(1, 7)])
        s_poly = sub.generate_s_poly_coeffs(tau, 600 + np.arange(7) * 10, 999, 1000)
        kw = dict(b_pos=float(sub.blackbody_contrib_to_BCs(700, 999, 1000)) * 0.5,
                  b_neg=float(sub.blackbody_contrib_to_BCs(550, 999, 1000)) + 1,
                  s_poly_coeffs=np.tile(s_poly, (ncol, 1, 1)), BDRF_Fourier_modes=[0.5])
        rep = lambda a: np.tile(a, (ncol,) + (1,) * np.ndim(a))
        return dict(name=name, B=ncol, L=6, NQuad=NQuad,
                    args=(rep(tau), rep(omega), NQuad, rep(Leg), np.full(ncol, 0.5), np.full(ncol, pi), 0.0),
                    kwargs=kw, tau_eval=rep(np.concatenate([[0.0], tau])),
                    phi_eval=np.array([0.0, pi / 2, pi]), outputs=("flux", "u"))
    if name == "tp1":
        # DISORT test problem 1a-1f stacked (pydisotest/1_test.py); B must be a multiple of 6
        cases = [(0.03125, 0.2), (0.03125, 1 - 1e-6), (0.03125, 0.99), (32, 0.2), (32, 1 - 1e-6), (32, 0.99)]
        idx = (first + np.arange(ncol)) % 6
        tau = np.array([cases[i][0] for i in idx], dtype=float)[:, None]
        omega = np.array([cases[i][1] for i in idx], dtype=float)[:, None]
        Leg = np.zeros((ncol, 1, 17))
        Leg[:, :, 0] = 1
        frac = np.array([0, 0.25, 0.5, 0.75, 1.0])
        return dict(name=name, B=ncol, L=1, NQuad=16,
                    args=(tau, omega, 16, Leg, np.full(ncol, 0.1), np.full(ncol, pi / 0.1), pi),
                    kwargs={}, tau_eval=tau * frac[None, :], phi_eval=np.array([0.0, pi / 2, pi]),
                    outputs=("flux", "u"))
    raise ValueError(f"unknown ensemble {name!r}")


def column_call(ens, b):
    """(args, kwargs) of the single-column, reference-style call for column ``b``."""
    B = ens["B"]

    def pick(v):
        if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == B:
            out = v[b]
            return float(out) if np.ndim(out) == 0 else (float(out[0]) if out.shape == (1,) else out)
        return v

    args = tuple(pick(a) for a in ens["args"])
    kwargs = {}
    for k, v in ens["kwargs"].items():
        if k == "BDRF_Fourier_modes":
            kwargs[k] = [pick(np.asarray(m)) if isinstance(m, np.ndarray) else m for m in v]
        else:
            kwargs[k] = pick(v)
    return args, kwargs


def run_reference_like(pydisort_fn, ens, columns=None, at_user_mu=False):
    """Evaluate a one-column-per-call implementation with the reference's
    signature (the reference itself, or the oracle) on the ensemble's level
    grid.  Returns arrays stacked over columns.  ``at_user_mu``: ensembles that
    name user polar angles (``mu_user``) get ``u`` there through the host-side
    ``interpolate`` (what a user of the reference would call)."""
    cols = range(ens["B"]) if columns is None else columns
    want_u = "u" in ens["outputs"]
    Fp, Fmd, Fdir, U0, Uu = [], [], [], [], []
    for b in cols:
        args, kwargs = column_call(ens, b)
        out = pydisort_fn(*args, **kwargs)
        t = ens["tau_eval"][b]
        Fp.append(out[1](t))
        dn = out[2](t)
        Fmd.append(dn[0])
        Fdir.append(np.broadcast_to(dn[1], np.shape(dn[0])))
        U0.append(out[3](t))
        if want_u and at_user_mu and ens.get("mu_user") is not None:
            Uu.append(sub.interpolate(out[4])(ens["mu_user"], t, ens["phi_eval"]))
        elif want_u:
            Uu.append(out[4](t, ens["phi_eval"]))
    res = dict(flux_up=np.array(Fp), flux_down_diffuse=np.array(Fmd), flux_down_direct=np.array(Fdir),
               u0=np.array(U0))
    if want_u:
        res["u"] = np.array(Uu)
    return res
