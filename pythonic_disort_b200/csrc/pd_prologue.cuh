// pd_prologue.cuh -- per-column prologue of pydisort (pydisort.py:223-372):
// value checks, delta-M scaling, Kirchhoff weighting and affine transform of
// the thermal source polynomials, source rescale; plus the normalised
// associated Legendre functions P~_l^m(-mu0) needed by the beam source
// (_solve_for_gen_and_part_sols.py:80,101-103).
#pragma once
#include "pd_common.cuh"

struct PdPrologue {
    int B, L, N, NLeg, NLeg_all, NF, Ns, NFb;
    int nt_requested;
    const double *tau, *omega, *leg_all, *f, *s_poly, *mu0, *I0, *phi0, *b_pos, *b_neg, *mu_nodes;
    double *taus, *omega_s, *wleg, *scale_tau, *s_s, *colp, *bpos_s, *bneg_s, *pmu0;
    int32_t* checks;
};

// normalised associated Legendre functions P~_l^m(x), l = m..nleg-1, into out[l] (out[l<m] = 0)
//   P~_m^m = -sqrt((2m-1)/(2m)) sqrt(1-x^2) P~_{m-1}^{m-1},
//   P~_{l+1}^m = ((2l+1) x P~_l^m - sqrt((l+m)(l-m)) P~_{l-1}^m) / sqrt((l+1-m)(l+1+m))
PD_HD void pd_norm_assoc_legendre(int m, int nleg, double x, double* out) {
    const double s = sqrt(fmax(0.0, 1.0 - x * x));
    double pmm = 1.0;
    for (int k = 1; k <= m; ++k) pmm *= -sqrt((2.0 * k - 1.0) / (2.0 * k)) * s;
    for (int l = 0; l < m && l < nleg; ++l) out[l] = 0.0;
    if (m >= nleg) return;
    out[m] = pmm;
    double pm1 = 0.0, p = pmm;
    for (int l = m; l + 1 < nleg; ++l) {
        const double nxt = ((2.0 * l + 1.0) * x * p - sqrt((double)(l + m) * (double)(l - m)) * pm1) /
                           sqrt((double)(l + 1 - m) * (double)(l + 1 + m));
        out[l + 1] = nxt;
        pm1 = p;
        p = nxt;
    }
}

template <class Grp>
PD_HD int pd_prologue_column(const Grp& g, const PdPrologue& a, int b) {
    const int lane = g.lane();
    const int L = a.L, n = a.N;
    const double* tau = a.tau + (long)b * L;
    const double* om = a.omega + (long)b * L;
    const double* leg = a.leg_all + (long)b * L * a.NLeg_all;
    const double* f = a.f ? a.f + (long)b * L : nullptr;
    const double* sp = (a.Ns > 0) ? a.s_poly + (long)b * L * a.Ns : nullptr;
    const double mu0 = a.mu0[b], I0 = a.I0[b], phi0 = a.phi0[b];
    double* taus = a.taus + (long)b * (L + 1);
    double* scl = a.scale_tau + (long)b * L;
    double* oms = a.omega_s + (long)b * L;
    double* wleg = a.wleg + (long)b * L * a.NLeg;
    double* ss = (a.Ns > 0) ? a.s_s + (long)b * L * a.Ns : nullptr;
    int chk = 0;

    // ---- checks that need every element (each lane looks at a strided share) ----
    bool fany = false, wany = false;
    for (int l = lane; l < L; l += Grp::size) {
        fany = fany || (f && f[l] > 0.0);
        wany = wany || (om[l] > 0.0);
    }
    fany = g.any(fany);  // uniform across lanes
    wany = g.any(wany);
    for (int l = lane; l < L; l += Grp::size) {
        const double t = tau[l], th = t - (l ? tau[l - 1] : 0.0);
        if (!(t > 0.0)) chk |= PD_CHK_TAU_POS;
        if (!(th > 0.0)) chk |= PD_CHK_THICK_POS;
        if (!(om[l] >= 0.0 && om[l] < 1.0)) chk |= PD_CHK_OMEGA_RANGE;
        if (f && !(f[l] >= 0.0 && f[l] <= 1.0)) chk |= PD_CHK_F_RANGE;
        if (!(om[l] * leg[(long)l * a.NLeg_all] == om[l])) chk |= PD_CHK_LEG0_FIXED;
    }
    for (long idx = lane; idx < (long)L * a.NLeg_all; idx += Grp::size) {
        const int k = (int)(idx % a.NLeg_all);
        if (k > 0 && !(leg[idx] > -1.0 && leg[idx] < 1.0)) chk |= PD_CHK_LEG_RANGE;
    }
    if (I0 < 0.0) chk |= PD_CHK_I0_NEG;
    if (I0 > 0.0) {
        if (!(mu0 > 0.0 && mu0 <= 1.0)) chk |= PD_CHK_MU0_RANGE;
        if (!(phi0 >= 0.0 && phi0 < 2.0 * PD_PI)) chk |= PD_CHK_PHI0_RANGE;
    }
    if (a.nt_requested)
        for (int i = 0; i < n; ++i)
            if (fabs(a.mu_nodes[i] - mu0) < 1e-8) chk |= PD_CHK_MU0_AT_NODE;

    // ---- delta-M scaling (:316-338) ----
    for (int l = lane; l < L; l += Grp::size) {
        const double fl = fany ? f[l] : 0.0;
        const double sc = fany ? 1.0 - om[l] * fl : 1.0;
        scl[l] = sc;
        const double os = fany ? (1.0 - fl) / sc * om[l] : om[l];
        oms[l] = os;
        if (os > 1.0 - 1e-6) chk |= PD_CHK_OMEGA_NEAR1;
        for (int k = 0; k < a.NLeg; ++k) {
            const double gk = (k == 0) ? 1.0 : leg[(long)l * a.NLeg_all + k];
            const double gs = fany ? (gk - fl) / (1.0 - fl) : gk;
            if (k > 0 && (gs < -0.95 || gs > 0.95)) chk |= PD_CHK_LEG_NEAR1;
            wleg[(long)l * a.NLeg + k] = gs * (2 * k + 1);
        }
    }
    // cumulative scaled optical depth, summed top-down like np.cumsum: every lane brings one term of a chunk, the
    // terms go round by shuffles and every lane adds them up in the same order (no serial chain of memory round trips)
    if (lane == 0) taus[0] = 0.0;
    if (fany) {
        double acc = 0.0;
        for (int base = 0; base < L; base += Grp::size) {
            const int l = base + lane;
            double term = 0.0;
            if (l < L) {
                const double fl = f[l];
                term = (1.0 - om[l] * fl) * (tau[l] - (l ? tau[l - 1] : 0.0));  // scl[l] as computed above
            }
            double mine = 0.0;
            const int cnt = (L - base < Grp::size) ? L - base : Grp::size;
            for (int i = 0; i < cnt; ++i) {
                acc += g.shfl(term, i);
                if (lane == i) mine = acc;
            }
            if (l < L) taus[l + 1] = mine;
        }
    } else {
        for (int l = lane; l < L; l += Grp::size) taus[l + 1] = tau[l];
    }
    g.sync();
    // ---- thermal source polynomials: Kirchhoff weighting and, under delta-M, the
    //      change of variable tau -> tau* (:325-329, subroutines.py:574-610) ----
    for (int l = lane; l < L && a.Ns > 0; l += Grp::size) {
        const double emis = 1.0 - om[l];
        if (!fany) {
            for (int q = 0; q < a.Ns; ++q) ss[(long)l * a.Ns + q] = sp[(long)l * a.Ns + q] * emis;
        } else {
            const double sc = scl[l], ainv = 1.0 / sc;
            const double shift = taus[l] - sc * (l ? tau[l - 1] : 0.0);
            for (int i = 0; i < a.Ns; ++i) {
                // D_i = sum_{j>=i} C(j,i) a^-j (-shift)^(j-i) c_j
                double acc = 0.0, binom = 1.0, apow = 1.0, spow = 1.0;
                for (int j = 0; j < i; ++j) apow *= ainv;
                for (int j = i; j < a.Ns; ++j) {
                    if (j > i) {
                        binom = binom * j / (j - i);
                        apow *= ainv;
                        spow *= -shift;
                    }
                    acc += binom * apow * spow * sp[(long)l * a.Ns + j];
                }
                ss[(long)l * a.Ns + i] = acc / sc * emis;
            }
        }
    }
    g.sync();
    // ---- source rescale (:351-372) ----
    const int nb = a.NFb * n;
    const double* bp = a.b_pos + (long)b * nb;
    const double* bn = a.b_neg + (long)b * nb;
    double resc = I0;
    for (int i = 0; i < nb; ++i) {
        resc = fmax(resc, bp[i]);
        resc = fmax(resc, bn[i]);
    }
    if (a.Ns > 0) {
        resc = fmax(resc, ss[0]);
        double v = 0.0;
        for (int q = a.Ns - 1; q >= 0; --q) v = fma(v, taus[L], ss[(long)(L - 1) * a.Ns + q]);
        resc = fmax(resc, v);
    }
    const bool divide = (resc != 0.0);
    g.sync();  // everyone has read the un-rescaled ss
    const double div = divide ? resc : 1.0;
    for (long idx = lane; idx < (long)L * a.Ns; idx += Grp::size) ss[idx] /= div;
    for (int i = lane; i < nb; i += Grp::size) {
        a.bpos_s[(long)b * nb + i] = bp[i] / div;
        a.bneg_s[(long)b * nb + i] = bn[i] / div;
    }
    if (lane == 0) {
        double* cp = a.colp + (long)b * PD_NCOLP;
        cp[PD_COL_MU0] = mu0;
        cp[PD_COL_I0] = I0 / div;
        cp[PD_COL_RESCALE] = resc;
        cp[PD_COL_PHI0] = phi0;
        cp[PD_COL_I0_RAW] = I0;
        cp[PD_COL_DM] = fany ? 1.0 : 0.0;
        cp[PD_COL_NT] = (I0 > 0.0 && fany && wany) ? 1.0 : 0.0;
        cp[7] = 0.0;
    }
    // ---- P~_l^m(-mu0) ----
    if (I0 > 0.0)
        for (int m = lane; m < a.NF; m += Grp::size)
            pd_norm_assoc_legendre(m, a.NLeg, -mu0, a.pmu0 + ((long)b * a.NF + m) * a.NLeg);
    return chk;
}
