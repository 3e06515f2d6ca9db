"""CPU oracle for the PythonicDISORT solver hot path -- TEST INFRASTRUCTURE ONLY.

This module is a NumPy/SciPy restatement of the reference algorithm (one
column per call, FP64).  It exists to *check* the CUDA path and to serve as
the CPU baseline leg of ``bench.py``; nothing under ``pythonic_disort_b200/``
imports it and the product never falls back to it.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the unmodified
reference (``/root/reference/src/PythonicDISORT``) in the build container and
stores its outputs; ``tests/test_oracle_vs_golden.py`` checks this restatement
against those vectors (and against the Stamnes DISORT 4.0.99 fixtures carried
over from ``pydisotest/Stamnes_results``).

Reference map (all paths relative to /root/reference/src/PythonicDISORT):
  prepare_column      <- pydisort.py:184-372      (defaults, delta-M, rescale)
  eigen_and_particular<- _solve_for_gen_and_part_sols.py:63-243
  thermal_poly        <- subroutines.py:746-862   (_mathscr_v)
  boundary_solve      <- _solve_for_coeffs.py:79-390
  OracleColumn.*      <- _assemble_intensity_and_fluxes.py:170-613
  _tms / _ims         <- pydisort.py:409-639      (Nakajima-Tanaka)
"""
from math import pi, comb

import numpy as np
import scipy.linalg
import scipy.special


# --------------------------------------------------------------------------
# quadrature (subroutines.py:116-138)
# --------------------------------------------------------------------------
def double_gauss_nodes(N):
    """Gauss-Legendre nodes/weights on [0, 1] (one hemisphere)."""
    x, w = np.polynomial.legendre.leggauss(int(N))
    return (x + 1.0) / 2.0, w / 2.0


def scattering_cosine(mu, phi, mu_p, phi_p):
    """cos(scattering angle) for outgoing (mu, phi) and a single incoming
    direction (mu_p, phi_p); shape (len(mu), len(phi)).  subroutines.py:85-112"""
    mu = np.atleast_1d(mu)[:, None]
    phi = np.atleast_1d(phi)[None, :]
    return mu_p * mu + np.sqrt(1 - mu_p**2) * np.sqrt(1 - mu**2) * np.cos(phi_p - phi)


def shift_poly(coeffs, a, b):
    """Re-express sum_j c_j x^j in y = a x + b, per row.  subroutines.py:574-610"""
    nrow, ncoef = coeffs.shape
    out = np.zeros_like(coeffs, dtype=float)
    for j in range(ncoef):
        for i in range(j + 1):
            out[:, i] += comb(j, i) * (1.0 / a) ** j * (-b) ** (j - i) * coeffs[:, j]
    return out


# --------------------------------------------------------------------------
# prologue: pydisort.py:184-372
# --------------------------------------------------------------------------
class Prepared:
    pass


def prepare_column(tau_arr, omega_arr, NQuad, Leg_coeffs_all, mu0, I0, phi0,
                   NLeg=None, NFourier=None, b_pos=0, b_neg=0, only_flux=False,
                   f_arr=0, NT_cor=False, BDRF_Fourier_modes=(), s_poly_coeffs=None):
    p = Prepared()
    tau_arr = np.atleast_1d(np.asarray(tau_arr, dtype=float))
    omega_arr = np.atleast_1d(np.asarray(omega_arr, dtype=float))
    Leg_all = np.atleast_2d(np.asarray(Leg_coeffs_all, dtype=float)).copy()
    if s_poly_coeffs is None:
        s_poly_coeffs = np.array([[]])
    s_poly = np.atleast_2d(np.asarray(s_poly_coeffs, dtype=float))
    f_arr = np.atleast_1d(np.asarray(f_arr, dtype=float))

    if NLeg is None:
        NLeg = NQuad
    if only_flux:
        NFourier = 1
    elif NFourier is None:
        NFourier = NQuad
    b_pos = np.asarray(b_pos, dtype=float)
    b_neg = np.asarray(b_neg, dtype=float)
    if np.all(b_pos == 0):
        b_pos = np.asarray(0.0)
    if np.all(b_neg == 0):
        b_neg = np.asarray(0.0)
    Ns = 0 if (s_poly.size == 0 or np.all(s_poly == 0)) else s_poly.shape[1]

    L = len(tau_arr)
    N = NQuad // 2
    Leg_all[:, 0] = 1.0  # pydisort.py:246-248
    NLeg_all = Leg_all.shape[1]
    thick = np.diff(tau_arr, prepend=0.0)
    mu_pos, W = double_gauss_nodes(N)

    if np.any(f_arr > 0):  # delta-M, pydisort.py:316-329
        f = np.broadcast_to(f_arr, (L,)).astype(float)
        scale_tau = 1 - omega_arr * f
        thick_s = scale_tau * thick
        taus = np.concatenate([[0.0], np.cumsum(thick_s)])
        leg_s = (Leg_all[:, :NLeg] - f[:, None]) / (1 - f)[:, None]
        omega_s = (1 - f) / scale_tau * omega_arr
        if Ns > 0:
            shifts = taus[:-1] - scale_tau * np.concatenate([[0.0], tau_arr[:-1]])
            s_s = (shift_poly(s_poly, scale_tau, shifts) / scale_tau[:, None]) * (1 - omega_arr)[:, None]
        else:
            s_s = np.zeros((L, 0))
    else:  # pydisort.py:331-338
        f = np.zeros(L)
        scale_tau = np.ones(L)
        thick_s = thick
        taus = np.concatenate([[0.0], tau_arr])
        leg_s = Leg_all[:, :NLeg]
        omega_s = omega_arr
        s_s = s_poly * (1 - omega_arr)[:, None] if Ns > 0 else np.zeros((L, 0))
    wleg = leg_s * (2 * np.arange(NLeg) + 1)[None, :]

    # source rescale, pydisort.py:351-372
    if Ns > 0:
        rescale = max(float(I0), float(np.max(b_pos)), float(np.max(b_neg)), float(s_s[0, 0]),
                      float(s_s[-1, :] @ (taus[-1] ** np.arange(Ns))))
        I0s, bp, bn, s_s = I0 / rescale, b_pos / rescale, b_neg / rescale, s_s / rescale
    else:
        rescale = max(float(I0), float(np.max(b_pos)), float(np.max(b_neg)))
        if rescale != 0:
            I0s, bp, bn = I0 / rescale, b_pos / rescale, b_neg / rescale
        else:
            I0s, bp, bn = I0, b_pos, b_neg

    p.tau, p.omega, p.Leg_all, p.f = tau_arr, omega_arr, Leg_all, f
    p.f_any = bool(np.any(f_arr > 0))
    p.L, p.N, p.NQuad, p.NLeg, p.NLeg_all, p.NF, p.Ns = L, N, NQuad, NLeg, NLeg_all, NFourier, Ns
    p.mu_pos, p.W = mu_pos, W
    p.mu = np.concatenate([mu_pos, -mu_pos])
    p.scale_tau, p.thick_s, p.taus, p.omega_s, p.wleg, p.s_s = scale_tau, thick_s, taus, omega_s, wleg, s_s
    p.rescale, p.I0, p.b_pos, p.b_neg = rescale, float(I0s), np.asarray(bp), np.asarray(bn)
    p.mu0, p.phi0 = float(mu0), float(phi0)
    p.beam = bool(I0 > 0)
    p.iso = Ns > 0
    p.only_flux = bool(only_flux)
    p.bdrf = list(BDRF_Fourier_modes)
    # pydisort.py:375 -- when do the NT corrections actually switch on?
    p.nt = bool(NT_cor and not only_flux and p.beam and p.f_any and NLeg < NLeg_all
                and np.any(omega_arr > 0))
    return p


# --------------------------------------------------------------------------
# stage 1: eigenpairs and particular solutions
# (_solve_for_gen_and_part_sols.py:63-243)
# --------------------------------------------------------------------------
def eigen_and_particular(p):
    L, N, NQ, NLeg, NF = p.L, p.N, p.NQuad, p.NLeg, p.NF
    Minv = 1.0 / p.mu_pos
    G = np.zeros((NF, L, NQ, NQ))
    K = np.zeros((NF, L, NQ))
    Bv = np.zeros((NF, L, NQ)) if p.beam else None
    Ginv0 = np.zeros((L, NQ, NQ)) if p.iso else None
    eyeN = np.eye(N)
    trivial = np.zeros((NQ, NQ))
    trivial[:N, N:] = eyeN
    trivial[N:, :N] = eyeN

    for m in range(NF):
        ells = np.arange(m, NLeg)
        poch = scipy.special.poch(ells + m + 1, -2.0 * m)           # (l-m)!/(l+m)!
        P = scipy.special.lpmv(m, ells[:, None], p.mu_pos[None, :])  # (n, N)
        sgn = np.where((ells - m) % 2 == 0, 1.0, -1.0)
        c = 0.5 * p.omega_s[:, None] * p.wleg[:, m:]                 # (L, n)
        act = np.any(np.abs(c) > 1e-8, axis=1)                       # :119
        # shortcut layers (:162-168)
        G[m, ~act] = trivial
        K[m, ~act, :N] = -Minv
        K[m, ~act, N:] = Minv
        if p.iso and m == 0:
            Ginv0[~act] = trivial
        if not np.any(act):
            continue
        ca = c[act] * poch[None, :]
        Dp = np.einsum("ln,ni,nj->lij", ca, P, P)
        Dm = np.einsum("ln,ni,nj->lij", ca * sgn[None, :], P, P)
        alpha = Minv[None, :, None] * (Dp * p.W[None, None, :] - eyeN[None])
        beta = Minv[None, :, None] * Dm * p.W[None, None, :]
        apb, amb = alpha + beta, alpha - beta
        k2, V = np.linalg.eig(amb @ apb)                             # :181
        k = np.sqrt(k2)
        V = V / 2
        U = apb @ (V / k[:, None, :])
        Gm = np.empty((len(k), NQ, NQ))
        Gm[:, :N, :N] = V + U
        Gm[:, N:, N:] = V + U
        Gm[:, :N, N:] = V - U
        Gm[:, N:, :N] = V - U
        Km = np.concatenate([-k, k], axis=1)
        G[m, act] = Gm
        K[m, act] = Km
        if p.iso and m == 0:
            Ginv0[act] = np.linalg.inv(Gm)                           # :203
        if p.beam:
            Pm0 = scipy.special.lpmv(m, ells, -p.mu0)
            x = (p.I0 / (4 * pi)) * (2 - (m == 0)) * poch * Pm0      # (n,)
            xt = x[None, :] * p.omega_s[act, None] * p.wleg[act, m:]
            X = np.concatenate([Minv[None, :] * (xt @ P), -Minv[None, :] * ((xt * sgn[None, :]) @ P)], axis=1)
            if p.iso and m == 0:                                     # :213-222
                y = np.einsum("lij,lj->li", Ginv0[act], X)
                Bv[m, act] = np.einsum("lij,lj->li", Gm, y / (1 / p.mu0 + Km))
            else:                                                    # :225-231
                A = np.zeros((len(k), NQ, NQ))
                A[:, :N, :N] = -alpha
                A[:, :N, N:] = -beta
                A[:, N:, :N] = beta
                A[:, N:, N:] = alpha
                Bv[m, act] = np.linalg.solve(np.eye(NQ)[None] / p.mu0 + A, X[:, :, None])[:, :, 0]
    return G, K, Bv, Ginv0


# --------------------------------------------------------------------------
# thermal particular solution (subroutines.py:746-862)
# --------------------------------------------------------------------------
def thermal_poly(G0, K0, Ginv_mu, s_s):
    """Per layer, the stream-space polynomial coefficients d[l, :, q] such that
    v_l(tau*) = sum_q d[l, :, q] tau*^q.   (coefficient of tau^q:
    b_q(k) = sum_{r>=q} s_r r!/q! K^-(r-q+1), then d_q = G (b_q * G^-1 mu^-1).)"""
    L, NQ = K0.shape
    Ns = s_s.shape[1]
    d = np.zeros((L, NQ, Ns))
    Kinv = 1.0 / K0
    fact = [float(scipy.special.factorial(i, exact=True)) for i in range(Ns)]
    for q in range(Ns):
        b = np.zeros((L, NQ))
        for r in range(q, Ns):
            b += s_s[:, r:r + 1] * (fact[r] / fact[q]) * Kinv ** (r - q + 1)
        d[:, :, q] = np.einsum("lij,lj->li", G0, b * Ginv_mu)
    return d


# --------------------------------------------------------------------------
# stage 2: boundary-condition solve (_solve_for_coeffs.py:79-390)
# --------------------------------------------------------------------------
def _bc_vector(b, m, N, NF):
    if b.ndim == 0:
        return np.full(N, float(b)) if m == 0 else np.zeros(N)
    if b.ndim == 1:
        return b.astype(float) if m == 0 else np.zeros(N)
    return b[:, m].astype(float)


def bdrf_tables(p, m):
    """R = (1+delta_m0) q^m(mu_i, mu_j) mu_j w_j and the beam reflection vector.
    _solve_for_coeffs.py:121-134"""
    N = p.N
    fm = p.bdrf[m]
    if np.isscalar(fm):
        q = np.full((N, N), float(fm))
        q0 = np.full(N, float(fm))
    else:
        q = np.asarray(fm(p.mu_pos, p.mu_pos), dtype=float)
        q0 = np.asarray(fm(p.mu_pos, np.array([p.mu0])), dtype=float)[:, 0] if p.beam else np.zeros(N)
    R = (1 + (m == 0)) * q * (p.mu_pos * p.W)[None, :]
    Xs = (p.mu0 * p.I0 / pi) * q0
    return R, Xs


def boundary_solve(p, G, K, Bv, dth, banded_from=10):
    L, N, NQ, NF = p.L, p.N, p.NQuad, p.NF
    taus = p.taus
    dim = L * NQ
    C = np.zeros((NF, L, NQ))
    for m in range(NF):
        Gm, Km = G[m], K[m]
        has_bdrf = len(p.bdrf) > m
        if has_bdrf:
            R, Xs = bdrf_tables(p, m)
        E = np.exp(-Km[:, N:] * np.diff(taus)[:, None])             # exp(-k dtau*)  (L, N)

        rhs = np.zeros(dim)
        rhs[:N] = _bc_vector(p.b_neg, m, N, NF)
        rhs[-N:] = _bc_vector(p.b_pos, m, N, NF)
        if m == 0 and p.iso:                                         # :168-232
            pw = lambda t: t ** np.arange(p.Ns)
            rhs[:N] -= (dth[0] @ pw(taus[0]))[N:]
            if L > 1:
                vt = np.stack([dth[l + 1] @ pw(taus[l + 1]) - dth[l] @ pw(taus[l + 1]) for l in range(L - 1)])
                rhs[N:-N] += vt.ravel()
            vL = dth[-1] @ pw(taus[-1])
            rhs[-N:] -= vL[:N]
            if has_bdrf:
                rhs[-N:] += R @ vL[N:]
        if p.beam:                                                   # :234-254
            Bm = Bv[m]
            rhs[:N] -= Bm[0, N:]
            if L > 1:
                rhs[N:-N] += ((Bm[1:] - Bm[:-1]) * np.exp(-taus[1:-1] / p.mu0)[:, None]).ravel()
            att = np.exp(-taus[-1] / p.mu0)
            if has_bdrf:
                rhs[-N:] += (Xs + R @ Bm[-1, N:] - Bm[-1, :N]) * att
            else:
                rhs[-N:] -= Bm[-1, :N] * att

        # rows: [top BC | interface 0 | ... | interface L-2 | bottom BC]
        top = np.concatenate([Gm[0][N:, :N], Gm[0][N:, N:] * E[0][None, :]], axis=1)       # (N, NQ)
        GL = Gm[-1]
        if has_bdrf:
            bot = np.concatenate([(GL[:N, :N] - R @ GL[N:, :N]) * E[-1][None, :], GL[:N, N:] - R @ GL[N:, N:]], axis=1)
        else:
            bot = np.concatenate([GL[:N, :N] * E[-1][None, :], GL[:N, N:]], axis=1)
        if L > 1:
            mid = np.concatenate([Gm[:-1, :, :N] * E[:-1, None, :], Gm[:-1, :, N:],
                                  -Gm[1:, :, :N], -Gm[1:, :, N:] * E[1:, None, :]], axis=2)  # (L-1, NQ, 2NQ)

        if L >= banded_from:                                         # :276-333
            kl = ku = 3 * N - 1
            ab = np.zeros((kl + ku + 1, dim))

            def put(i, j, blk):
                ab[ku + i - j, j] = blk

            r = np.arange(N)[:, None]
            c = np.arange(NQ)[None, :]
            put(r, c, top)
            put(dim - N + r, dim - NQ + c, bot)
            li = np.arange(L - 1)[:, None, None]
            ri = np.arange(NQ)[None, :, None]
            ci = np.arange(2 * NQ)[None, None, :]
            put(N + li * NQ + ri, li * NQ + ci, mid)
            Cm = scipy.linalg.solve_banded((kl, ku), ab, rhs)
        else:                                                        # :337-383
            A = np.zeros((dim, dim))
            A[:N, :NQ] = top
            A[-N:, -NQ:] = bot
            for l in range(L - 1):
                A[N + l * NQ:N + (l + 1) * NQ, l * NQ:(l + 2) * NQ] = mid[l]
            Cm = np.linalg.solve(A, rhs)
        C[m] = Cm.reshape(L, NQ)
    return G * C[:, :, None, :]


# --------------------------------------------------------------------------
# the solved column and its output functions
# (_assemble_intensity_and_fluxes.py:170-613, pydisort.py:409-694)
# --------------------------------------------------------------------------
class OracleColumn:
    def __init__(self, p):
        self.p = p
        G, K, Bv, Ginv0 = eigen_and_particular(p)
        N = p.N
        if p.iso:
            Ginv_mu = np.einsum("lij,j->li", Ginv0, np.concatenate([1 / p.mu_pos, -1 / p.mu_pos]))
            self.dth = thermal_poly(G[0], K[0], Ginv_mu, p.s_s)
        else:
            self.dth = None
        self.GC = boundary_solve(p, G, K, Bv, self.dth)
        self.G, self.K, self.Bv = G, K, Bv
        self.mu_arr = p.mu

    # -- shared pieces ------------------------------------------------------
    def _locate(self, tau):
        p = self.p
        tau = np.atleast_1d(np.asarray(tau, dtype=float))
        if np.any(tau < 0) or np.any(tau > p.tau[-1]):
            raise ValueError("tau input outside the tau range given for the atmosphere (check `tau_arr`).")
        l = np.argmax(tau[:, None] <= p.tau[None, :], axis=1)        # interface -> upper layer
        if p.f_any:
            ts = p.taus[1:][l] - (p.tau[l] - tau) * p.scale_tau[l]
        else:
            ts = tau
        return tau, l, ts

    def _modes(self, modes, l, ts, anti):
        """u^m(tau) for the requested modes -> (len(modes), NQ, Ntau)."""
        p = self.p
        N = p.N
        K = self.K[modes][:, l, :]                                   # (M, T, NQ)
        ex = np.concatenate([K[:, :, :N] * (ts - p.taus[l])[None, :, None],
                             K[:, :, N:] * (ts - p.taus[1:][l])[None, :, None]], axis=2)
        w = np.exp(ex)
        if anti:
            w = w / (p.scale_tau[l][None, :, None] * K)
        um = np.einsum("mtij,mtj->mit", self.GC[modes][:, l], w)
        if p.beam:
            bb = self.Bv[modes][:, l, :] * np.exp(-ts / p.mu0)[None, :, None]
            if anti:
                bb = bb / (-p.scale_tau[l] / p.mu0)[None, :, None]
            um = um + bb.transpose(0, 2, 1)
        return um

    def _thermal(self, l, ts, anti):
        p = self.p
        q = np.arange(p.Ns)
        if anti:  # intended semantics: divide by scale_tau of the layer of each point
            poly = ts[:, None] ** (q + 1)[None, :] / ((q + 1)[None, :] * p.scale_tau[l][:, None])
        else:
            poly = ts[:, None] ** q[None, :]
        return np.einsum("tiq,tq->it", self.dth[l], poly)

    # -- fluxes ---------------------------------------------------------------
    def u0(self, tau, is_antiderivative_wrt_tau=False):
        tau, l, ts = self._locate(tau)
        u = self._modes([0], l, ts, is_antiderivative_wrt_tau)[0]
        if self.p.iso:
            u = u + self._thermal(l, ts, is_antiderivative_wrt_tau)
        return self.p.rescale * np.squeeze(u)

    def flux_up(self, tau, is_antiderivative_wrt_tau=False):
        p = self.p
        tau, l, ts = self._locate(tau)
        u = self._modes([0], l, ts, is_antiderivative_wrt_tau)[0]
        if p.iso:
            u = u + self._thermal(l, ts, is_antiderivative_wrt_tau)
        return p.rescale * np.squeeze(2 * pi * (p.mu_pos * p.W) @ u[:p.N])[()]

    def flux_down(self, tau, is_antiderivative_wrt_tau=False):
        p = self.p
        anti = is_antiderivative_wrt_tau
        tau, l, ts = self._locate(tau)
        u = self._modes([0], l, ts, anti)[0]
        if p.iso:
            u = u + self._thermal(l, ts, anti)
        direct = direct_s = 0.0
        if p.beam:
            if anti:
                direct = p.I0 * p.mu0 * np.exp(-tau / p.mu0) * -p.mu0
                direct_s = p.I0 * p.mu0 * np.exp(-ts / p.mu0) / (-p.scale_tau[l] / p.mu0)
            else:
                direct = p.I0 * p.mu0 * np.exp(-tau / p.mu0)
                direct_s = p.I0 * p.mu0 * np.exp(-ts / p.mu0)
        diffuse = 2 * pi * (p.mu_pos * p.W) @ u[p.N:] + direct_s - direct
        direct = np.broadcast_to(direct, diffuse.shape) if np.ndim(direct) == 0 else direct
        return p.rescale * np.squeeze(diffuse)[()], p.rescale * np.squeeze(direct)[()]

    def actinic(self, tau):
        """(up, down-diffuse) actinic fluxes; subroutines.py:258-318."""
        p = self.p
        tau, l, ts = self._locate(tau)
        u = self._modes([0], l, ts, False)[0]
        if p.iso:
            u = u + self._thermal(l, ts, False)
        u = p.rescale * u
        recl = (p.I0 * np.exp(-ts / p.mu0) - p.I0 * np.exp(-tau / p.mu0)) if p.f_any else 0.0
        # NB: the reference adds the *un-rescaled* reclassification term (I0 was
        # already divided by rescale_factor when u0 captured it).
        return 2 * pi * p.W @ u[:p.N], 2 * pi * p.W @ u[p.N:] + recl

    # -- intensity ------------------------------------------------------------
    def u(self, tau, phi, is_antiderivative_wrt_tau=False, return_Fourier_error=False):
        p = self.p
        if p.only_flux:
            raise ValueError("intensity was not requested (only_flux=True)")
        anti = is_antiderivative_wrt_tau
        tau, l, ts = self._locate(tau)
        phi = np.atleast_1d(np.asarray(phi, dtype=float))
        um = self._modes(np.arange(p.NF), l, ts, anti)               # (NF, NQ, T)
        if p.iso:
            um[0] += self._thermal(l, ts, anti)
        cosm = np.cos(np.arange(p.NF)[:, None] * (p.phi0 - phi)[None, :])
        u = np.einsum("mit,mp->itp", um, cosm)
        out = u
        if p.nt:
            corr = self._tms(tau, phi, l, ts, anti)
            corr[p.N:] += self._ims(tau, phi, anti)
            out = u + corr
        res = p.rescale * np.squeeze(out)
        if return_Fourier_error:
            ulast = um[-1][:, :, None] * np.cos((p.NF - 1) * (p.phi0 - phi))[None, None, :]
            ua = np.abs(u)
            err = np.max(np.divide(np.abs(ulast), ua, out=np.zeros_like(ua), where=ua > 1e-8))
            return res, err
        return res

    # -- Nakajima-Tanaka: TMS (pydisort.py:409-597) ---------------------------
    def _tms(self, tau, phi, l, ts, anti):
        p = self.p
        N, L, mu0 = p.N, p.L, p.mu0
        Minv = 1 / p.mu_pos
        nu = scattering_cosine(p.mu, phi, -mu0, p.phi0)                       # (NQ, Nphi)
        leg = np.polynomial.legendre.legval
        wall = (2 * np.arange(p.NLeg_all) + 1) * p.Leg_all
        p_true = np.stack([leg(nu, wall[j]) for j in range(L)], axis=1)       # (NQ, L, Nphi)
        p_trun = np.stack([leg(nu, p.wleg[j]) for j in range(L)], axis=1)
        Bsc = ((p.omega_s * p.I0 / (4 * pi))[None, :, None] * (mu0 / (mu0 + p.mu))[:, None, None]
               * (p_true / (1 - p.f)[None, :, None] - p_trun))
        t_bot, t_top = p.taus[1:][l], p.taus[l]
        e0 = np.exp(-ts / mu0)
        if anti:
            sc = p.scale_tau[l]
            own_pos = (e0 / (-sc / mu0))[None, :] - np.exp((ts - t_bot)[None, :] * Minv[:, None] - t_bot[None, :] / mu0) / (sc[None, :] * Minv[:, None])
            own_neg = (e0 / (-sc / mu0))[None, :] + np.exp((t_top - ts)[None, :] * Minv[:, None] - t_top[None, :] / mu0) / (sc[None, :] * Minv[:, None])
        else:
            own_pos = e0[None, :] - np.exp((ts - t_bot)[None, :] * Minv[:, None] - t_bot[None, :] / mu0)
            own_neg = e0[None, :] - np.exp((t_top - ts)[None, :] * Minv[:, None] - t_top[None, :] / mu0)
        sol = Bsc[:, l, :] * np.concatenate([own_pos, own_neg], axis=0)[:, :, None]
        if L == 1:
            return sol
        # other layers: stable recurrences equivalent to pydisort.py:495-589
        dt = p.thick_s
        top, bot = p.taus[:-1], p.taus[1:]
        fac = (p.mu_pos[:, None] / p.scale_tau[None, :]) if anti else 1.0
        # upward streams: light singly scattered in the layers below
        term_pos = fac * -np.expm1(-dt[None, :] * (Minv + 1 / mu0)[:, None]) * np.exp(-top / mu0)[None, :]
        # Rpos[j] = sum_{r>j} term[r] exp(-(top_r - bot_j)/mu)
        Rpos = np.zeros((N, L))
        for j in range(L - 2, -1, -1):
            Rpos[:, j] = term_pos[:, j + 1] + Rpos[:, j + 1] * np.exp(-dt[j + 1] * Minv)
        sol[:N] += (Rpos[:, l] * np.exp(Minv[:, None] * (ts - t_bot)[None, :]))[:, :, None] * Bsc[:N][:, l, :]
        # downward streams: light singly scattered in the layers above
        a = dt[None, :] * (Minv - 1 / mu0)[:, None]
        em1 = np.expm1(-np.abs(a))
        term_neg = np.where(a >= 0, -em1 * np.exp(-bot / mu0)[None, :],
                            em1 * np.exp(-dt[None, :] * Minv[:, None]) * np.exp(-top / mu0)[None, :])
        if anti:
            term_neg = -fac * term_neg
        Rneg = np.zeros((N, L))
        for j in range(1, L):
            Rneg[:, j] = (Rneg[:, j - 1] * np.exp(-dt[j - 1] * Minv) + term_neg[:, j - 1])
        # Rneg[j] = sum_{r<j} term[r] exp(-(top_j - bot_r)/mu)
        sol[N:] += (Rneg[:, l] * np.exp(Minv[:, None] * (t_top - ts)[None, :]))[:, :, None] * Bsc[N:][:, l, :]
        return sol

    # -- Nakajima-Tanaka: IMS (pydisort.py:601-639) ---------------------------
    def _ims(self, tau, phi, anti):
        p = self.p
        mu0 = p.mu0
        s1 = np.sum(p.omega * p.tau)
        w_avg = s1 / np.sum(p.tau)
        s2 = np.sum(p.f * p.omega * p.tau)
        f_avg = s2 / s1
        res = p.Leg_all.copy()
        res[:, :p.NLeg] = p.f[:, None]
        res_avg = np.sum(res * (p.omega * p.tau)[:, None], axis=0) / s2
        mu0s = mu0 / (1 - w_avg * f_avg)
        nu = scattering_cosine(-p.mu_pos, phi, -mu0, p.phi0)                  # (N, Nphi)
        x = 1 / p.mu_pos - 1 / mu0s
        if anti:
            chi = ((mu0s - x[:, None] * mu0s * (mu0s + tau)[None, :]) * np.exp(-tau / mu0s)[None, :]
                   - p.mu_pos[:, None] * np.exp(-tau[None, :] / p.mu_pos[:, None])) / (p.mu_pos * mu0s * x**2)[:, None]
        else:
            chi = ((tau[None, :] - 1 / x[:, None]) * np.exp(-tau / mu0s)[None, :]
                   + np.exp(-tau[None, :] / p.mu_pos[:, None]) / x[:, None]) / (p.mu_pos * mu0s * x)[:, None]
        series = np.polynomial.legendre.legval(nu, (2 * np.arange(p.NLeg_all) + 1) * (2 * res_avg - res_avg**2))
        amp = p.I0 / (4 * pi) * (w_avg * f_avg) ** 2 / (1 - w_avg * f_avg)
        return (amp * series)[:, None, :] * chi[:, :, None]


def solve_column(*args, **kwargs):
    """Same positional/keyword inputs as the reference ``pydisort`` (one column)."""
    kwargs.pop("use_banded_solver_NLayers", None)
    return OracleColumn(prepare_column(*args, **kwargs))


def pydisort(*args, **kwargs):
    """Reference-shaped return: (mu_arr, flux_up, flux_down, u0[, u])."""
    col = solve_column(*args, **kwargs)
    out = (col.mu_arr, col.flux_up, col.flux_down, col.u0)
    return out if col.p.only_flux else out + (col.u,)
