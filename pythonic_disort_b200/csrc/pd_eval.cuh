// pd_eval.cuh -- evaluation of the solved column at user optical depths:
//   u^m(tau) = G_l (C_l * exp(K_l (tau* - tau*_{l-1 or l}))) + B_l exp(-tau*/mu0) + delta_m0 v_l(tau*)
// (_assemble_intensity_and_fluxes.py:170-613), and the Nakajima-Tanaka
// TMS / IMS intensity corrections (pydisort.py:409-694).
#pragma once
#include "pd_common.cuh"
#include "pd_stage_b.cuh"  // pd_thermal_at

struct PdEval {
    int B, L, N, NF, Ns, NLeg, NLeg_all;
    int beam, iso;
    pd_state st;
    const double* tau_q;  // [B][ntau]
    int ntau;
    int anti;
};

// first layer whose lower boundary is >= t  (np.argmax(tau <= tau_arr), :185)
PD_HD int pd_locate(const double* tau_col, int L, double t) {
    int lo = 0, hi = L - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (t <= tau_col[mid]) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}

// The same index found by a lane group together: the lanes count the boundaries above the point, every load is
// independent (one memory latency instead of log2(L) dependent ones).  Pays for groups of 16 lanes or more
// (measured: HA intensities 59.6 -> 57.5 ms; with 4 lanes per point the 15 loads per lane cost more than the search).
template <class Grp>
PD_HD int pd_locate_group(const Grp& g, const double* tau_col, int L, double t) {
    if (Grp::size < 16) return pd_locate(tau_col, L, t);
    int cnt = 0;
    for (int j = g.lane(); j < L - 1; j += Grp::size) cnt += (tau_col[j] < t) ? 1 : 0;  // strictly increasing tau
    double tot = 0.0;
    for (int s = 0; s < Grp::size; ++s) tot += g.shfl((double)cnt, s);
    return (int)tot;
}

// scaled optical depth of a query point (:189-195); dm = column uses delta-M scaling
PD_HD double pd_scaled_tau(const PdEval& a, int b, int l, double t) {
    const double dm = a.st.colp[(long)b * PD_NCOLP + PD_COL_DM];
    if (dm == 0.0) return t;
    const double* tau = a.st.tau + (long)b * a.L;
    const double* taus = a.st.taus + (long)b * (a.L + 1);
    return taus[l + 1] - (tau[l] - t) * a.st.scale_tau[(long)b * a.L + l];
}

// Interface hit: a query point that IS a layer interface (0, or bit for bit one of the column's tau[l]) takes the
// radiances the boundary-condition sweep stored for that interface (pd_state.Uif; pd_stage_b*.cuh) instead of
// G_l (C_l * e) + particular -- the same quantity to rounding, without reading G.  Returns the interface index or -1.
PD_HD int pd_interface_level(const PdEval& a, int b, int l, double tq) {
    if (!a.st.Uif || a.anti) return -1;
    if (tq == a.st.tau[(long)b * a.L + l]) return l + 1;  // bottom of the point's layer (pd_locate: first l with tq <= tau[l])
    return (tq == 0.0) ? 0 : -1;
}

// u^m of interface `lev` into uv[2n] (shared)
template <class Grp>
PD_HD void pd_mode_at_interface(const Grp& g, const PdEval& a, int b, int m, int lev, double* uv) {
    const int n2 = 2 * a.N;
    const double* us = a.st.Uif + pd_uif_index(b, lev, m, a.L, a.NF, n2);
    for (int i = g.lane(); i < n2; i += Grp::size) uv[i] = us[i];
    g.sync();
}

// u^m at one point into uv[2n] (shared); ev[2n] is scratch.  Includes the beam
// and (m = 0) thermal particular solutions; NOT multiplied by rescale_factor.
// beam attenuation factor of a query point: exp(-tau* / mu0) (its tau-antiderivative if a.anti), 0 without a beam
PD_HD double pd_beam_factor(const PdEval& a, int b, int l, double ts) {
    if (!(a.beam && a.st.colp[(long)b * PD_NCOLP + PD_COL_I0] > 0.0)) return 0.0;
    const double mu0 = a.st.colp[(long)b * PD_NCOLP + PD_COL_MU0];
    double eb = exp(-ts / mu0);
    if (a.anti) eb /= -(a.st.scale_tau[(long)b * a.L + l] / mu0);
    return eb;
}

// `eb` = pd_beam_factor of the point (the same for every mode: callers that loop over the modes compute it once)
template <class Grp, int NC = 0>
PD_HD void pd_mode_at(const Grp& g, const PdEval& a, int b, int m, int l, double ts, double eb, double* ev, double* uv) {
    const int lane = g.lane();
    const int n = NC > 0 ? NC : a.N, n2 = 2 * n;
    const long item = ((long)b * a.NF + m) * a.L + l;
    const double* K = a.st.K + item * n;
    const double* Gi = a.st.G + pd_g_base(item, n);  // layout: pd_common.cuh
    const double* C = a.st.C + item * n2;
    const double* taus = a.st.taus + (long)b * (a.L + 1);
    const double sc = a.st.scale_tau[(long)b * a.L + l];
    const double dtop = ts - taus[l], dbot = ts - taus[l + 1];
    for (int j = lane; j < n; j += Grp::size) {
        const double k = K[j];
        double em = exp(-k * dtop) * C[j], ep = exp(k * dbot) * C[n + j];
        if (a.anti) {
            em /= -(sc * k);
            ep /= (sc * k);
        }
        ev[j] = em;
        ev[n + j] = ep;
    }
    g.sync();
    const bool beam = a.beam && a.st.colp[(long)b * PD_NCOLP + PD_COL_I0] > 0.0;
    const double* Bv = beam ? a.st.Bv + item * n2 : nullptr;
    const double* dth = (a.iso && m == 0) ? a.st.dth + ((long)b * a.L + l) * a.Ns * n2 : nullptr;
    for (int i = lane; i < n; i += Grp::size) {
        double top = 0.0, bot = 0.0;
        if (NC > 0 && NC % 4 == 0) {  // G is read exactly once per query point: 256-bit streaming loads of the two rows
#pragma unroll
            for (int j = 0; j < (NC > 0 ? NC : 4); j += 4) {
                double gp[4], gm[4];
                pd_load4_stream(Gi + pd_g_off(i * n + j, n), gp);
                pd_load4_stream(Gi + pd_g_off(n * n + i * n + j, n), gm);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    top = fma(gp[e], ev[j + e], top);
                    top = fma(gm[e], ev[n + j + e], top);
                    bot = fma(gm[e], ev[j + e], bot);
                    bot = fma(gp[e], ev[n + j + e], bot);
                }
            }
        } else {
            for (int j = 0; j < n; ++j) {
                const double gp = Gi[pd_g_off(i * n + j, n)], gm = Gi[pd_g_off(n * n + i * n + j, n)];
                top = fma(gp, ev[j], top);
                top = fma(gm, ev[n + j], top);
                bot = fma(gm, ev[j], bot);
                bot = fma(gp, ev[n + j], bot);
            }
        }
        if (Bv) {
            top = fma(Bv[i], eb, top);
            bot = fma(Bv[n + i], eb, bot);
        }
        if (dth) {
            if (!a.anti) {
                top += pd_thermal_at(dth, a.Ns, n2, i, ts);
                bot += pd_thermal_at(dth, a.Ns, n2, n + i, ts);
            } else {  // sum_q d_q tau*^(q+1) / ((q+1) scale_tau_l)
                double pt = 0.0, pb = 0.0;
                for (int q = a.Ns - 1; q >= 0; --q) {
                    pt = fma(pt, ts, dth[q * n2 + i] / (q + 1));
                    pb = fma(pb, ts, dth[q * n2 + n + i] / (q + 1));
                }
                top += pt * ts / sc;
                bot += pb * ts / sc;
            }
        }
        uv[i] = top;
        uv[n + i] = bot;
    }
    g.sync();
}

// flux outputs of one point from the hemispheric sums up = sum_i mu_i w_i u_i, dn = sum_i mu_i w_i u_{n+i}  (:446-613)
PD_HD void pd_flux_store(const PdEval& a, int b, int t, int l, double tq, double ts, double up, double dn, double* Fup,
                         double* Fdn, double* Fdir) {
    const double* cp = a.st.colp + (long)b * PD_NCOLP;
    const double resc = cp[PD_COL_RESCALE];
    double direct = 0.0, direct_s = 0.0;
    if (a.beam && cp[PD_COL_I0] > 0.0) {
        const double mu0 = cp[PD_COL_MU0], I0 = cp[PD_COL_I0];
        if (a.anti) {
            const double sc = a.st.scale_tau[(long)b * a.L + l];
            direct = I0 * mu0 * exp(-tq / mu0) * -mu0;
            direct_s = I0 * mu0 * exp(-ts / mu0) / (-sc / mu0);
        } else {
            direct = I0 * mu0 * exp(-tq / mu0);
            direct_s = I0 * mu0 * exp(-ts / mu0);
        }
    }
    const long o = (long)b * a.ntau + t;
    Fup[o] = resc * (2.0 * PD_PI * up);
    Fdn[o] = resc * (2.0 * PD_PI * dn + direct_s - direct);
    Fdir[o] = resc * direct;
}

// fluxes at one point (:446-613).  sm: 4n doubles.
template <class Grp, int NC = 0>
PD_HD void pd_flux_point(const Grp& g, const PdEval& a, int b, int t, double* sm, double* Fup, double* Fdn,
                         double* Fdir) {
    const int n = a.N;
    const double tq = a.tau_q[(long)b * a.ntau + t];
    const int l = pd_locate_group(g, a.st.tau + (long)b * a.L, a.L, tq);
    const double ts = pd_scaled_tau(a, b, l, tq);
    double* ev = sm;
    double* uv = sm + 2 * n;
    const int lev = pd_interface_level(a, b, l, tq);
    if (lev >= 0) pd_mode_at_interface(g, a, b, 0, lev, uv);
    else pd_mode_at<Grp, NC>(g, a, b, 0, l, ts, pd_beam_factor(a, b, l, ts), ev, uv);
    if (g.lane() == 0) {
        double up = 0.0, dn = 0.0;
        for (int i = 0; i < n; ++i) {
            const double mw = a.st.mu_nodes[i] * a.st.w_nodes[i];
            up = fma(mw, uv[i], up);
            dn = fma(mw, uv[n + i], dn);
        }
        pd_flux_store(a, b, t, l, tq, ts, up, dn, Fup, Fdn, Fdir);
    }
    g.sync();
}

// The same for a point that is a layer interface, by ONE thread (no scratch): the 2n radiances of mode 0 are one
// contiguous record of Uif, and consecutive levels of a column are consecutive records.  Returns false (nothing
// written) if the point is not an interface; sums in the order of pd_flux_point, so both give the same bits.
template <int NC = 0>
PD_HD bool pd_flux_point_interface(const PdEval& a, int b, int t, double* Fup, double* Fdn, double* Fdir) {
    const int n = NC > 0 ? NC : a.N;
    const double tq = a.tau_q[(long)b * a.ntau + t];
    const double* tau = a.st.tau + (long)b * a.L;
    // a level grid that mirrors the layers (ntau = L + 1: 0, tau_0, .., tau_{L-1}) is found without a search
    int l = (t > 0 && t <= a.L && tau[t - 1] == tq) ? t - 1 : pd_locate(tau, a.L, tq);
    const int lev = pd_interface_level(a, b, l, tq);
    if (lev < 0) return false;
    const double* us = a.st.Uif + pd_uif_index(b, lev, 0, a.L, a.NF, 2 * n);
    double up = 0.0, dn = 0.0;
    if (NC > 0 && NC % 2 == 0) {
#pragma unroll
        for (int i = 0; i < (NC > 0 ? NC : 2); i += 2) {
            const pd_d2 p2 = *reinterpret_cast<const pd_d2*>(us + i), m2 = *reinterpret_cast<const pd_d2*>(us + n + i);
            const double w0 = a.st.mu_nodes[i] * a.st.w_nodes[i], w1 = a.st.mu_nodes[i + 1] * a.st.w_nodes[i + 1];
            up = fma(w0, p2.x, up);
            dn = fma(w0, m2.x, dn);
            up = fma(w1, p2.y, up);
            dn = fma(w1, m2.y, dn);
        }
    } else {
        for (int i = 0; i < n; ++i) {
            const double mw = a.st.mu_nodes[i] * a.st.w_nodes[i];
            up = fma(mw, us[i], up);
            dn = fma(mw, us[n + i], dn);
        }
    }
    pd_flux_store(a, b, t, l, tq, pd_scaled_tau(a, b, l, tq), up, dn, Fup, Fdn, Fdir);
    return true;
}

// u0 at one point (:334-433); recl = actinic delta-scaling reclassification term (:360-371)
template <class Grp, int NC = 0>
PD_HD void pd_u0_point(const Grp& g, const PdEval& a, int b, int t, double* sm, double* u0, double* recl) {
    const int n = a.N, n2 = 2 * n;
    const double tq = a.tau_q[(long)b * a.ntau + t];
    const int l = pd_locate_group(g, a.st.tau + (long)b * a.L, a.L, tq);
    const double ts = pd_scaled_tau(a, b, l, tq);
    double* ev = sm;
    double* uv = sm + n2;
    const int lev = pd_interface_level(a, b, l, tq);
    if (lev >= 0) pd_mode_at_interface(g, a, b, 0, lev, uv);
    else pd_mode_at<Grp, NC>(g, a, b, 0, l, ts, pd_beam_factor(a, b, l, ts), ev, uv);
    const double* cp = a.st.colp + (long)b * PD_NCOLP;
    const double resc = cp[PD_COL_RESCALE];
    for (int i = g.lane(); i < n2; i += Grp::size) u0[((long)b * n2 + i) * a.ntau + t] = resc * uv[i];
    if (recl && g.lane() == 0) {
        double r = 0.0;
        if (cp[PD_COL_DM] != 0.0) {
            const double mu0 = cp[PD_COL_MU0], I0 = cp[PD_COL_I0];
            if (a.anti) {
                const double sc = a.st.scale_tau[(long)b * a.L + l];
                r = I0 * exp(-ts / mu0) / (-sc / mu0) - I0 * exp(-tq / mu0) * -mu0;
            } else {
                r = I0 * exp(-ts / mu0) - I0 * exp(-tq / mu0);
            }
        }
        recl[(long)b * a.ntau + t] = r;
    }
    g.sync();
}

// all Fourier modes at one point into um[NF][2n] (shared); ev: 2n scratch
// (lev >= 0: the point is interface `lev`, pd_interface_level -- its NF * 2n radiances are contiguous in Uif)
template <class Grp, int NC = 0>
PD_HD void pd_all_modes_point(const Grp& g, const PdEval& a, int b, int l, int lev, double ts, double* ev, double* um) {
    if (lev >= 0) {
        const int tot = a.NF * 2 * a.N;
        const double* us = a.st.Uif + pd_uif_index(b, lev, 0, a.L, a.NF, 2 * a.N);
        for (int i = g.lane(); i < tot; i += Grp::size) um[i] = us[i];
        g.sync();
        return;
    }
    const double eb = pd_beam_factor(a, b, l, ts);
    for (int m = 0; m < a.NF; ++m) pd_mode_at<Grp, NC>(g, a, b, m, l, ts, eb, ev, um + m * 2 * a.N);
}

// sum_m um[m * stride] cos(m dphi)  (:256-260).  cos(m dphi) by the Chebyshev recurrence
// c_{m+1} = 2 cos(dphi) c_m - c_{m-1}: one cosine per (stream, azimuth) instead of NFourier;
// its rounding error grows like m * eps, far below the parity tolerance for NFourier <= 64+.
PD_HD double pd_azimuth_sum(const double* um, int stride, int NF, double dphi) {
    const double c1 = cos(dphi), two_c1 = 2.0 * c1;
    double cm1 = 1.0, cm = c1;
    double s = um[0];
    if (NF > 1) s = fma(um[stride], c1, s);
    for (int m = 2; m < NF; ++m) {
        const double cn = fma(two_c1, cm, -cm1);
        s = fma(um[(long)m * stride], cn, s);
        cm1 = cm;
        cm = cn;
    }
    return s;
}

// ---------------------------------------------------------------------------
// Nakajima-Tanaka corrections.
// ---------------------------------------------------------------------------
struct PdNT {
    const double* omega;     // [B][L]  unscaled
    const double* f;         // [B][L]
    const double* leg_all;   // [B][L][NLeg_all]  (coefficient 0 already forced to 1)
    const double* omega_s;   // [B][L]
    const double* wleg;      // [B][L][NLeg]
};

// sum_{l<nc} coef[l] P_l(x) by upward recurrence (coefficients already carry their (2l+1) weights)
// rinv[l] = 1/l (l >= 1): a table, because an FP64 division per term would dominate the series
PD_HD double pd_legendre_series(const double* coef, int nc, double x, const double* rinv) {
    double p0 = 1.0, p1 = x;
    double s = coef[0];
    if (nc > 1) s = fma(coef[1], x, s);
    double lf = 1.0;
    for (int l = 1; l + 1 < nc; ++l, lf += 1.0) {
        const double p2 = (fma(2.0, lf, 1.0) * x * p1 - lf * p0) * rinv[l + 1];
        s = fma(coef[l + 1], p2, s);
        p0 = p1;
        p1 = p2;
    }
    return s;
}

// sum_{l<nc} (2l+1) g[l] P_l(x) for unweighted phase-function moments, g[0] taken as 1 (pydisort.py:246-248)
PD_HD double pd_legendre_series_raw(const double* gl, int nc, double x, const double* rinv) {
    double p0 = 1.0, p1 = x;
    double s = 1.0;
    if (nc > 1) s = fma(3.0 * gl[1], x, s);
    double lf = 1.0;
    for (int l = 1; l + 1 < nc; ++l, lf += 1.0) {
        const double p2 = (fma(2.0, lf, 1.0) * x * p1 - lf * p0) * rinv[l + 1];
        s = fma(fma(2.0, lf, 3.0) * gl[l + 1], p2, s);
        p0 = p1;
        p1 = p2;
    }
    return s;
}

// cosine of the scattering angle between stream i at azimuth phi and the beam (-mu0, phi0)  (subroutines.py:85-112)
PD_HD double pd_nt_nu(const PdEval& a, int b, int i, double phi) {
    const double* cp = a.st.colp + (long)b * PD_NCOLP;
    const double mu0 = cp[PD_COL_MU0];
    const double mua = a.st.mu_nodes[i < a.N ? i : i - a.N];
    const double mus = (i < a.N) ? mua : -mua;
    return -mu0 * mus + sqrt(1.0 - mu0 * mu0) * sqrt(1.0 - mus * mus) * cos(cp[PD_COL_PHI0] - phi);
}

// Per column pre-computation for the TMS correction, multi-layer part
// (pydisort.py:495-589): Rpos[n][L], Rneg[n][L] by stable recurrences.
//   Rpos[i][l] = sum_{r>l} (1-exp(-dt_r (1/mu_i+1/mu0))) exp(-top_r/mu0) exp(-(top_r - bot_l)/mu_i)
//   Rneg[i][l] = sum_{r<l} term_neg[i][r] exp(-(top_l - bot_r)/mu_i)
template <class Grp>
PD_HD void pd_tms_scans(const Grp& g, const PdEval& a, int b, double* Rpos, double* Rneg) {
    const int n = a.N, L = a.L;
    const double* taus = a.st.taus + (long)b * (L + 1);
    const double* scl = a.st.scale_tau + (long)b * L;
    const double mu0 = a.st.colp[(long)b * PD_NCOLP + PD_COL_MU0];
    for (int i = g.lane(); i < n; i += Grp::size) {
        const double mi = 1.0 / a.st.mu_nodes[i];
        const double mu = a.st.mu_nodes[i];
        double acc = 0.0;
        Rpos[i * L + (L - 1)] = 0.0;
        for (int l = L - 2; l >= 0; --l) {
            const int r = l + 1;
            const double dt = taus[r + 1] - taus[r];
            double term = -expm1(-dt * (mi + 1.0 / mu0)) * exp(-taus[r] / mu0);
            if (a.anti) term *= mu / scl[r];
            // contributions of layers below r, attenuated across layer r
            acc = (l == L - 2) ? term : fma(acc, exp(-dt * mi), term);
            Rpos[i * L + l] = acc;
        }
        acc = 0.0;
        Rneg[i * L] = 0.0;
        for (int l = 1; l < L; ++l) {
            const int r = l - 1;
            const double dt = taus[r + 1] - taus[r];
            const double th = dt * (mi - 1.0 / mu0);
            const double em1 = expm1(-fabs(th));
            double term = (th >= 0.0) ? -em1 * exp(-taus[r + 1] / mu0) : em1 * exp(-dt * mi) * exp(-taus[r] / mu0);
            if (a.anti) term *= -mu / scl[r];
            acc = (l == 1) ? term : fma(acc, exp(-dt * mi), term);
            Rneg[i * L + l] = acc;
        }
    }
    g.sync();
}

// Per column IMS constants (pydisort.py:601-611): coefficient array
// imsc[NLeg_all] = (2l+1)(2 g~_l - g~_l^2), and scalars out[0] = amplitude
// I0/(4 pi) (w f)^2/(1 - w f), out[1] = scaled mu0.
template <class Grp>
PD_HD void pd_ims_setup(const Grp& g, const PdEval& a, const PdNT& nt, int b, double* imsc, double* out) {
    const int L = a.L;
    const double* tau = a.st.tau + (long)b * L;
    const double* om = nt.omega + (long)b * L;
    const double* f = nt.f + (long)b * L;
    double s1 = 0.0, st = 0.0, s2 = 0.0;
    for (int l = 0; l < L; ++l) {
        s1 += om[l] * tau[l];
        st += tau[l];
        s2 += f[l] * om[l] * tau[l];
    }
    const double wavg = s1 / st, favg = s2 / s1;
    for (int k = g.lane(); k < a.NLeg_all; k += Grp::size) {
        double acc = 0.0;
        for (int l = 0; l < L; ++l) {
            const double res = (k < a.NLeg) ? f[l] : nt.leg_all[((long)b * L + l) * a.NLeg_all + k];
            acc += res * om[l] * tau[l];
        }
        const double r = acc / s2;
        imsc[k] = (2 * k + 1) * (2.0 * r - r * r);
    }
    if (g.lane() == 0) {
        const double* cp = a.st.colp + (long)b * PD_NCOLP;
        out[0] = cp[PD_COL_I0] / (4.0 * PD_PI) * (wavg * favg) * (wavg * favg) / (1.0 - wavg * favg);
        out[1] = cp[PD_COL_MU0] / (1.0 - wavg * favg);
    }
    g.sync();
}

// TMS + IMS correction of stream i at query point (tq in layer l, scaled ts) and azimuth phi.
// wall_l: the unweighted phase-function moments g_l[NLeg_all] of layer l.
PD_HD double pd_nt_value(const PdEval& a, const PdNT& nt, int b, int i, int l, double tq, double ts, double phi,
                         const double* Rpos, const double* Rneg, const double* imsc, const double* imsv,
                         const double* wall_l, const double* rinv) {
    const int n = a.N, L = a.L;
    const double* cp = a.st.colp + (long)b * PD_NCOLP;
    const double mu0 = cp[PD_COL_MU0], I0 = cp[PD_COL_I0];
    const double* taus = a.st.taus + (long)b * (L + 1);
    const bool up = i < n;
    const int ii = up ? i : i - n;
    const double mua = a.st.mu_nodes[ii];
    const double mus = up ? mua : -mua;  // signed stream cosine
    const double mi = 1.0 / mua;
    const double fl = nt.f[(long)b * L + l];
    const double* wtr = nt.wleg + ((long)b * L + l) * a.NLeg;
    const double nu = pd_nt_nu(a, b, i, phi);
    const double ptrue = pd_legendre_series_raw(wall_l, a.NLeg_all, nu, rinv);
    const double ptrun = pd_legendre_series(wtr, a.NLeg, nu, rinv);
    const double Bsc = nt.omega_s[(long)b * L + l] * (I0 / (4.0 * PD_PI)) * (mu0 / (mu0 + mus)) * (ptrue / (1.0 - fl) - ptrun);
    const double sc = a.st.scale_tau[(long)b * L + l];
    const double ttop = taus[l], tbot = taus[l + 1];
    double e0 = exp(-ts / mu0);
    double own, other;
    if (up) {
        const double ex = exp((ts - tbot) * mi - tbot / mu0);
        own = a.anti ? e0 / (-sc / mu0) - ex / (sc * mi) : e0 - ex;
        other = (L > 1) ? Rpos[ii * L + l] * exp(mi * (ts - tbot)) : 0.0;
    } else {
        const double ex = exp((ttop - ts) * mi - ttop / mu0);
        own = a.anti ? e0 / (-sc / mu0) + ex / (sc * mi) : e0 - ex;
        other = (L > 1) ? Rneg[ii * L + l] * exp(mi * (ttop - ts)) : 0.0;
    }
    double val = Bsc * (own + other);
    if (!up) {  // IMS, downward streams only (:613-638)
        const double mu0s = imsv[1];
        const double x = mi - 1.0 / mu0s;
        double chi;
        if (a.anti)
            chi = ((mu0s - x * mu0s * (mu0s + tq)) * exp(-tq / mu0s) - mua * exp(-tq * mi)) / (mua * mu0s * x * x);
        else
            chi = ((tq - 1.0 / x) * exp(-tq / mu0s) + exp(-tq * mi) / x) / (mua * mu0s * x);
        // (for a downward stream the IMS scattering angle equals nu)
        val += imsv[0] * pd_legendre_series(imsc, a.NLeg_all, nu, rinv) * chi;
    }
    return val;
}
