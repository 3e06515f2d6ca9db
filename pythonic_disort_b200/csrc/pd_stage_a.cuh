// pd_stage_a.cuh -- per (column, Fourier mode, layer) work item:
//   * assemble the symmetrised reduced matrices (alpha-beta), (alpha+beta)
//     (_solve_for_gen_and_part_sols.py:114-135),
//   * eigen-decompose (alpha-beta)(alpha+beta) and build the G blocks and K (:179-200),
//   * beam particular solution B (:141-152, :209-231),
//   * thermal particular solution coefficients (m = 0; :201-205 and subroutines.py:746-862).
//
// Everything is done in the similarity-scaled ("hatted") basis x^ = D x with
// D = diag(sqrt(w_i mu_i)), in which both reduced matrices are symmetric:
//   (alpha -/+ beta)^ = sum_{l-m odd/even} omega* (2l+1) g*_l Q_l(mu_i) Q_l(mu_j) - delta_ij / mu_i,
//   Q_l(mu_i) = P~_l^m(mu_i) sqrt(w_i / mu_i),   P~ = sqrt((l-m)!/(l+m)!) P_l^m.
// The scaling doubles as the balancing step of the eigen-solver; results are
// mapped back with D^-1 before they are stored.
#pragma once
#include "pd_linalg.cuh"

struct PdStageA {
    int B, L, N, NLeg, NF, Ns;
    int beam, iso;
    int only_flagged;       // recompute only items whose K[item][0] is NaN (fallback pass after the symmetric kernel)
    int32_t* nflagged;      // device counter of such items (null: unknown); the fallback pass returns at once if it is 0
    const double* omega_s;  // [B][L]
    const double* wleg;     // [B][L][NLeg]
    const double* s_s;      // [B][L][Ns]
    const double* colp;     // [B][PD_NCOLP]
    const double* pmu0;     // [B][NF][NLeg]
    const double* mu;       // [N]
    const double* w;        // [N]
    double* K;              // [B][NF][L][N]
    double* G;              // [B][NF][L][2][N][N]
    double* Bv;             // [B][NF][L][2N]
    double* dth;            // [B][L][Ns][2N]
    int32_t* status;        // [B]
};

// doubles of shared memory one item needs
PD_HD int pd_stage_a_item_doubles(int N, int NLeg) { return 4 * N * pd_ld(N) + 9 * N + NLeg; }

PD_HD void pd_atomic_or(int32_t* p, int v) {
#if defined(__CUDA_ARCH__)
    atomicOr(p, v);
#else
    *p |= v;
#endif
}

// Q: this mode's scaled Legendre table [NLeg - m][N] (shared by the CTA); sm: per-item scratch.
template <class Grp, int NC = 0>
PD_HD void pd_stage_a_item(const Grp& g, const PdStageA& a, int b, int m, int l, const double* Q, double* sm) {
    const int lane = g.lane();
    const int n = NC > 0 ? NC : a.N, ld = pd_ld(n), nm = a.NLeg - m;
    double* A1 = sm;            // (alpha-beta)^, later scratch Y of the eigen-solver, later Gp
    double* A2 = A1 + n * ld;   // (alpha+beta)^
    double* H = A2 + n * ld;    // product -> Schur factor -> U^
    double* Z = H + n * ld;     // LU copy for the beam solve -> eigenvectors -> Gm
    double* cs = Z + n * ld;    // [2n] rotation pairs (16-byte aligned: 4 n ld is even)
    double* wr = cs + 2 * n;    // [n] eigenvalues, then k
    double* vec = wr + n;       // [n]
    double* x1 = vec + n;       // [n]
    double* x2 = x1 + n;        // [n]
    double* rhs = x2 + n;       // [n]
    double* y1 = rhs + n;       // [n]
    double* dinv = y1 + n;      // [n] 1/sqrt(w mu)
    double* cw = dinv + n;      // [nm] omega* (2l+1) g*_l

    const long item = ((long)b * a.NF + m) * a.L + l;
    const double omega = a.omega_s[(long)b * a.L + l];
    const double* wl = a.wleg + ((long)b * a.L + l) * a.NLeg + m;
    double* Kout = a.K + item * n;
    double* Gout = a.G + pd_g_base(item, n);
    double* Bout = a.beam ? a.Bv + item * 2 * n : nullptr;
    const bool thermal = a.iso && m == 0;
    const bool beam = a.beam && a.colp[(long)b * PD_NCOLP + PD_COL_I0] > 0.0;  // per column (pydisort.py:215)
    if (a.only_flagged && Kout[0] == Kout[0]) return;  // not NaN: already solved (group-uniform)

    bool active = false;  // _solve_for_gen_and_part_sols.py:119
    for (int t = 0; t < nm; ++t) active = active || (fabs((omega / 2) * wl[t]) > 1e-8);

    for (int i = lane; i < n; i += Grp::size) dinv[i] = pd_rsqrt(a.w[i] * a.mu[i]);

    if (!active) {  // :162-168
        for (int idx = lane; idx < n * n; idx += Grp::size) {
            const int i = idx / n, j = idx - i * n;
            Gout[pd_g_off(idx, n)] = 0.0;
            Gout[pd_g_off(n * n + idx, n)] = (i == j) ? 1.0 : 0.0;
            if (thermal) {
                A1[i * ld + j] = 0.0;
                Z[i * ld + j] = (i == j) ? 1.0 : 0.0;
            }
        }
        for (int i = lane; i < n; i += Grp::size) {
            Kout[i] = 1.0 / a.mu[i];
            if (a.beam) {
                Bout[i] = 0.0;
                Bout[n + i] = 0.0;
            }
            if (thermal) {
                wr[i] = 1.0 / a.mu[i];
                y1[i] = -1.0 / a.mu[i];
            }
        }
        g.sync();
    } else {
        for (int t = lane; t < nm; t += Grp::size) cw[t] = omega * wl[t];
        g.sync();
        // symmetrised reduced matrices
        for (int idx = lane; idx < n * n; idx += Grp::size) {
            const int i = idx / n, j = idx - i * n;
            double se = 0.0, so = 0.0;
            for (int t = 0; t < nm; t += 2) {
                se = fma(cw[t] * Q[t * n + i], Q[t * n + j], se);
                if (t + 1 < nm) so = fma(cw[t + 1] * Q[(t + 1) * n + i], Q[(t + 1) * n + j], so);
            }
            const double dg = (i == j) ? 1.0 / a.mu[i] : 0.0;
            A1[i * ld + j] = so - dg;
            A2[i * ld + j] = se - dg;
        }
        g.sync();
        for (int idx = lane; idx < n * n; idx += Grp::size) {
            const int i = idx / n, j = idx - i * n;
            double s = 0.0;
            for (int t = 0; t < n; ++t) s = fma(A1[i * ld + t], A2[t * ld + j], s);
            H[i * ld + j] = s;
        }
        g.sync();

        if (a.beam && !beam && lane == 0)
            for (int i = 0; i < 2 * n; ++i) Bout[i] = 0.0;
        if (beam) {  // reduced N x N form of the 2N x 2N system of :225-231
            const double mu0 = a.colp[(long)b * PD_NCOLP + PD_COL_MU0];
            const double fac = a.colp[(long)b * PD_NCOLP + PD_COL_I0] / (4.0 * PD_PI) * ((m == 0) ? 1.0 : 2.0);
            const double* pm0 = a.pmu0 + ((long)b * a.NF + m) * a.NLeg + m;
            for (int i = lane; i < n; i += Grp::size) {
                double s1 = 0.0, s2 = 0.0;
                for (int t = 0; t < nm; ++t) {
                    const double v = fac * cw[t] * pm0[t] * Q[t * n + i];
                    s1 += v;
                    s2 += (t & 1) ? -v : v;
                }
                x1[i] = s1;   // (M^-1 X+)^
                x2[i] = -s2;  // -(M^-1 X-)^
            }
            for (int idx = lane; idx < n * n; idx += Grp::size) {
                const int i = idx / n, j = idx - i * n;
                Z[i * ld + j] = ((i == j) ? 1.0 / (mu0 * mu0) : 0.0) - H[i * ld + j];
            }
            g.sync();
            for (int i = lane; i < n; i += Grp::size) {
                double s = (x1[i] + x2[i]) / mu0;
                for (int j = 0; j < n; ++j) s = fma(A1[i * ld + j], x1[j] - x2[j], s);
                rhs[i] = s;
            }
            g.sync();
            const int st = pd_lu_solve<Grp, NC>(g, n, ld, Z, rhs);  // rhs <- p^ = (B+ + B-)^
            if (st && lane == 0) pd_atomic_or(a.status + b, st);
            for (int i = lane; i < n; i += Grp::size) {
                double s = x1[i] - x2[i];
                for (int j = 0; j < n; ++j) s = fma(A2[i * ld + j], rhs[j], s);
                const double qh = mu0 * s;  // (B+ - B-)^
                Bout[i] = 0.5 * (rhs[i] + qh) * dinv[i];
                Bout[n + i] = 0.5 * (rhs[i] - qh) * dinv[i];
            }
            g.sync();
        }

        int st = pd_eig_real<Grp, NC>(g, n, ld, H, Z, A1, wr, cs, vec);
        for (int j = 0; j < n; ++j)
            if (!(wr[j] > 0.0)) st |= PD_ST_BAD_EIGEN;
        if (st && lane == 0) pd_atomic_or(a.status + b, st);
        g.sync();
        for (int j = lane; j < n; j += Grp::size) {
            const double k = sqrt(wr[j]);
            wr[j] = k;
            Kout[j] = k;
        }
        g.sync();
        // U^ = (alpha+beta)^ V^ diag(1/k)
        for (int idx = lane; idx < n * n; idx += Grp::size) {
            const int i = idx / n, j = idx - i * n;
            double s = 0.0;
            for (int t = 0; t < n; ++t) s = fma(A2[i * ld + t], Z[t * ld + j], s);
            H[i * ld + j] = s / wr[j];
        }
        g.sync();
        for (int idx = lane; idx < n * n; idx += Grp::size) {
            const int i = idx / n, j = idx - i * n;
            const double v = Z[i * ld + j], u = H[i * ld + j];
            const double gp = 0.5 * (v + u) * dinv[i], gm = 0.5 * (v - u) * dinv[i];
            Gout[pd_g_off(idx, n)] = gp;
            Gout[pd_g_off(n * n + idx, n)] = gm;
            if (thermal) {
                A1[i * ld + j] = gp;
                Z[i * ld + j] = gm;
            }
        }
        if (thermal) {  // G^-1 [1/mu; -1/mu] = [y1; -y1],  U^ y1 = D (1/mu)
            for (int i = lane; i < n; i += Grp::size) y1[i] = 1.0 / (dinv[i] * a.mu[i]);
            g.sync();
            const int st2 = pd_lu_solve<Grp, NC>(g, n, ld, H, y1);
            if (st2 && lane == 0) pd_atomic_or(a.status + b, st2);
        }
        g.sync();
    }

    if (thermal) {
        const double* s = a.s_s + ((long)b * a.L + l) * a.Ns;
        double* dout = a.dth + ((long)b * a.L + l) * a.Ns * 2 * n;
        for (int q = 0; q < a.Ns; ++q) {
            for (int j = lane; j < n; j += Grp::size) {  // b_q(-k) y1 and b_q(+k) y1
                const double kinv = 1.0 / wr[j];
                double ratio = 1.0, pw = kinv, bp = 0.0, bn = 0.0;
                for (int r = q; r < a.Ns; ++r) {
                    if (r > q) {
                        ratio *= (double)r;
                        pw *= kinv;
                    }
                    bp = fma(s[r] * ratio, pw, bp);
                    bn = fma(s[r] * ratio, ((r - q) & 1) ? pw : -pw, bn);
                }
                x1[j] = bn * y1[j];
                x2[j] = bp * y1[j];
            }
            g.sync();
            for (int i = lane; i < n; i += Grp::size) {
                double top = 0.0, bot = 0.0;
                for (int j = 0; j < n; ++j) {
                    const double gp = A1[i * ld + j], gm = Z[i * ld + j];
                    top += gp * x1[j] - gm * x2[j];
                    bot += gm * x1[j] - gp * x2[j];
                }
                dout[q * 2 * n + i] = top;
                dout[q * 2 * n + n + i] = bot;
            }
            g.sync();
        }
    }
}
