// pd_stage_b_tps.cuh -- boundary-condition solve of one (column, Fourier mode) system by ONE THREAD (N = 2, 4: the
// longwave / two- and four-stream shapes).  Same elimination over interface radiances as pd_stage_b_add.cuh (see its
// header for the algebra: layer operators R^, T^ from the symmetric eigenvectors, forward sweep over the stack
// reflection Rup and source S, surface, back sweep, coefficients from the homogeneous parts at the two interfaces),
// but with every N x N matrix in the thread's registers: no shared memory, no operand broadcasts, no
// synchronisation -- for a 4 x 4 block the lane-group version spends most of its time on those.  The history
// (Q, Rup Q, q, Rup q + S per layer) is the only scratch; its elements are interleaved over the resident threads
// (`hs` = stride between consecutive doubles of one system) so that a warp's accesses are contiguous.
// tools/proto_adding.py is the NumPy prototype of the algebra.
#pragma once
#include "pd_stage_b.cuh"

template <int N>
struct PdStageBTps {
    static constexpr int NN = N * N, N2 = 2 * N;
    static constexpr long HIST_PER_LAYER = 2 * NN + N2;
};

// X <- X^-1 by Gauss-Jordan without pivoting (X = I + positive semidefinite: every pivot is a Schur complement >= 1)
template <int N>
PD_HD void pd_tps_inverse(double (&X)[N][N]) {
#pragma unroll
    for (int s = 0; s < N; ++s) {
        const double p = pd_rcp(X[s][s]);
        double f[N];
#pragma unroll
        for (int i = 0; i < N; ++i) f[i] = X[i][s];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const double r = ((j == s) ? 1.0 : X[s][j]) * p;
#pragma unroll
            for (int i = 0; i < N; ++i)
                if (i != s) X[i][j] = fma(-f[i], r, (j == s) ? 0.0 : X[i][j]);
            X[s][j] = r;
        }
    }
}

template <int N>
PD_HD bool pd_stage_b_tps(const PdStageB& A, int b, int m, double* hist, long hs) {
    using F = PdStageBTps<N>;
    constexpr int NN = F::NN, N2 = F::N2;
    const int L = A.L;
    const long sys = (long)b * A.NF + m;
    const double* taus = A.taus + (long)b * (L + 1);
    const double* Kc = A.K + sys * L * N;
    const long item0 = sys * L;  // G item of layer l: item0 + l, sector-interleaved over 32 items (pd_common.cuh)
    const double* Bc = A.beam ? A.Bv + sys * L * N2 : nullptr;
    const double* dthc = (A.iso && m == 0) ? A.dth + (long)b * L * A.Ns * N2 : nullptr;
    const double mu0 = A.colp[(long)b * PD_NCOLP + PD_COL_MU0];
    const double I0 = A.colp[(long)b * PD_NCOLP + PD_COL_I0];
    const bool beam = A.beam && I0 > 0.0;
    const bool has_bdrf = A.NBDRF > m;
    const bool have_b = (m == 0) || (A.NFb > 1);
    const double* bpos = A.bpos + ((long)b * A.NFb + (A.NFb > 1 ? m : 0)) * N;
    const double* bneg = A.bneg + ((long)b * A.NFb + (A.NFb > 1 ? m : 0)) * N;
    const double rmu0 = beam ? 1.0 / mu0 : 0.0;
    bool bad = false;

    double D[N];
#pragma unroll
    for (int i = 0; i < N; ++i) D[i] = sqrt(A.w[i] * A.mu[i]);

    // V^ = D (Gp + Gm) / 2, U^ = D (Gp - Gm) / 2 of layer l
    auto eigvecs = [&](int l, double (&V)[N][N], double (&U)[N][N]) {
        const double* Gl = A.G + pd_g_base(item0 + l, N);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double h = 0.5 * D[i];
#pragma unroll
            for (int k = 0; k < N; k += 2) {
                const pd_d2 gp = *reinterpret_cast<const pd_d2*>(Gl + pd_g_off(i * N + k, N));
                const pd_d2 gm = *reinterpret_cast<const pd_d2*>(Gl + pd_g_off(NN + i * N + k, N));
                V[i][k] = h * (gp.x + gm.x);
                V[i][k + 1] = h * (gp.y + gm.y);
                U[i][k] = h * (gp.x - gm.x);
                U[i][k + 1] = h * (gp.y - gm.y);
            }
        }
    };
    // particular solution of layer l (hat basis) at its top (attenuation at) and bottom (ab)
    auto particular = [&](int l, double t0, double t1, double at, double ab, double (&ptp)[N], double (&ptm)[N], double (&pbp)[N],
                          double (&pbm)[N]) {
        double top[N2], bot[N2];
#pragma unroll
        for (int i = 0; i < N2; ++i) top[i] = bot[i] = 0.0;
        if (beam) {
            const double* Bl = Bc + (long)l * N2;
#pragma unroll
            for (int i = 0; i < N2; i += 2) {
                const pd_d2 v = *reinterpret_cast<const pd_d2*>(Bl + i);
                top[i] = v.x * at;
                top[i + 1] = v.y * at;
                bot[i] = v.x * ab;
                bot[i + 1] = v.y * ab;
            }
        }
        if (dthc) {
            const double* dl = dthc + (long)l * A.Ns * N2;
            if (A.Ns == 2) {  // source linear in tau (the usual case): both coefficient vectors with vector loads
#pragma unroll
                for (int i = 0; i < N2; i += 2) {
                    const pd_d2 c0 = *reinterpret_cast<const pd_d2*>(dl + i);
                    const pd_d2 c1 = *reinterpret_cast<const pd_d2*>(dl + N2 + i);
                    top[i] += fma(c1.x, t0, c0.x);
                    top[i + 1] += fma(c1.y, t0, c0.y);
                    bot[i] += fma(c1.x, t1, c0.x);
                    bot[i + 1] += fma(c1.y, t1, c0.y);
                }
            } else {
#pragma unroll
                for (int i = 0; i < N2; ++i) {
                    top[i] += pd_thermal_at(dl, A.Ns, N2, i, t0);
                    bot[i] += pd_thermal_at(dl, A.Ns, N2, i, t1);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            ptp[i] = D[i] * top[i];
            ptm[i] = D[i] * top[N + i];
            pbp[i] = D[i] * bot[i];
            pbm[i] = D[i] * bot[N + i];
        }
    };
    // pull the inputs of layer l into L1 / L2 while the current layer is being worked on (the thread is the only
    // reader of its 2 N^2 + 5 N + ... doubles per layer, so nothing else hides that latency)
    auto prefetch_layer = [&](int l) {
#if defined(__CUDA_ARCH__)
        const double* gp = A.G + pd_g_base(item0 + l, N);
#pragma unroll
        for (int e = 0; e < 2 * NN; e += 16) asm volatile("prefetch.global.L1 [%0];" ::"l"(gp + pd_g_off(e, N)));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(Kc + (long)l * N));
        if (Bc) asm volatile("prefetch.global.L1 [%0];" ::"l"(Bc + (long)l * N2));
        if (dthc) asm volatile("prefetch.global.L1 [%0];" ::"l"(dthc + (long)l * A.Ns * N2));
#endif
    };

    // ---------------------------------------------------------------- forward sweep
    double Rup[N][N], S[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int j = 0; j < N; ++j) Rup[i][j] = 0.0;
        S[i] = D[i] * (have_b ? bneg[i] : 0.0);
    }
    double att_t = 1.0;  // exp(-tau*_l / mu0), tau*_0 = 0
    double t_top = taus[0], t_bot = taus[1];  // optical depths of the layer's interfaces; the next one is loaded a layer ahead
    for (int l = 0; l < L; ++l) {
        if (l + 1 < L) prefetch_layer(l + 1);
        const double t_ahead = taus[(l + 2 <= L) ? l + 2 : L];
        const double dtau = t_bot - t_top;
        const double att_b = beam ? exp(-t_bot * rmu0) : 0.0;
        double R[N][N], T[N][N];
        {
            double V[N][N], U[N][N], dk[N];
            eigvecs(l, V, U);
#pragma unroll
            for (int k = 0; k < N; ++k) {
                double gk = 0.0;
#pragma unroll
                for (int i = 0; i < N; ++i) gk = fma(V[i][k], U[i][k], gk);
                const double em = expm1(-Kc[(long)l * N + k] * dtau);  // tanh(x / 2) = -expm1(-x) / (2 + expm1(-x))
                dk[k] = (em / (2.0 + em)) / gk;
                bad = bad || !(gk < 0.0) || !(dk[k] >= 0.0);
            }
            double X1[N][N], X2[N][N];  // I + U^ d U^T, I + V^ d V^T
#pragma unroll
            for (int i = 0; i < N; ++i) {
                double tu[N], tv[N];
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    tu[k] = dk[k] * U[i][k];
                    tv[k] = dk[k] * V[i][k];
                }
#pragma unroll
                for (int j = i; j < N; ++j) {
                    double s1 = (i == j) ? 1.0 : 0.0, s2 = s1;
#pragma unroll
                    for (int k = 0; k < N; ++k) {
                        s1 = fma(tu[k], U[j][k], s1);
                        s2 = fma(tv[k], V[j][k], s2);
                    }
                    X1[i][j] = X1[j][i] = s1;
                    X2[i][j] = X2[j][i] = s2;
                }
            }
            pd_tps_inverse<N>(X1);
            pd_tps_inverse<N>(X2);
#pragma unroll
            for (int i = 0; i < N; ++i)
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    R[i][j] = X1[i][j] - X2[i][j];
                    T[i][j] = X1[i][j] + X2[i][j] - ((i == j) ? 1.0 : 0.0);
                }
        }
        double ptp[N], ptm[N], pbp[N], pbm[N];
        particular(l, t_top, t_bot, att_t, att_b, ptp, ptm, pbp, pbm);
        double sminus[N], vq[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double sp = ptp[i], sm = pbm[i], rs = 0.0;
#pragma unroll
            for (int k = 0; k < N; ++k) {
                sp = fma(-R[i][k], ptm[k], fma(-T[i][k], pbp[k], sp));
                sm = fma(-T[i][k], ptm[k], fma(-R[i][k], pbp[k], sm));
                rs = fma(R[i][k], S[k], rs);
            }
            sminus[i] = sm;
            vq[i] = sp + rs;
        }
        // (I - R^ Rup) [Q | q] = [T^ | R^ S + s+], Gauss-Jordan without pivoting (rows)
        double Mc[N][N], Q[N][N];
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) {
                double s = (i == j) ? 1.0 : 0.0;
#pragma unroll
                for (int k = 0; k < N; ++k) s = fma(-R[i][k], Rup[k][j], s);
                Mc[i][j] = s;
                Q[i][j] = T[i][j];
            }
        double minpiv = 1e300;
#pragma unroll
        for (int s = 0; s < N; ++s) {
            const double piv = Mc[s][s];
            minpiv = (piv < minpiv) ? piv : minpiv;
            const double p = pd_rcp(piv);
#pragma unroll
            for (int j = 0; j < N; ++j) {
                if (j > s) Mc[s][j] *= p;
                Q[s][j] *= p;
            }
            vq[s] *= p;
#pragma unroll
            for (int i = 0; i < N; ++i)
                if (i != s) {
                    const double f = Mc[i][s];
#pragma unroll
                    for (int j = 0; j < N; ++j) {
                        if (j > s) Mc[i][j] = fma(-f, Mc[s][j], Mc[i][j]);
                        Q[i][j] = fma(-f, Q[s][j], Q[i][j]);
                    }
                    vq[i] = fma(-f, vq[s], vq[i]);
                }
        }
        bad = bad || !(minpiv > 0.05);
        // Rup Q, Rup q + S; history; next Rup, S
        double RQ[N][N], rsv[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double r = S[i];
#pragma unroll
            for (int j = 0; j < N; ++j) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < N; ++k) s = fma(Rup[i][k], Q[k][j], s);
                RQ[i][j] = s;
                r = fma(Rup[i][j], vq[j], r);
            }
            rsv[i] = r;
        }
        double* hl = hist + (long)l * F::HIST_PER_LAYER * hs;
#pragma unroll
        for (int i = 0; i < N; ++i) {
#pragma unroll
            for (int j = 0; j < N; ++j) {
                hl[(long)(i * N + j) * hs] = Q[i][j];
                hl[(long)(NN + i * N + j) * hs] = RQ[i][j];
            }
            hl[(long)(2 * NN + i) * hs] = vq[i];
            hl[(long)(2 * NN + N + i) * hs] = rsv[i];
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double s2 = sminus[i];
#pragma unroll
            for (int j = 0; j < N; ++j) {
                double s = R[i][j];
#pragma unroll
                for (int k = 0; k < N; ++k) s = fma(T[i][k], RQ[k][j], s);
                Rup[i][j] = s;
                s2 = fma(T[i][j], rsv[j], s2);
            }
            S[i] = s2;
        }
        att_t = att_b;
        t_top = t_bot;
        t_bot = t_ahead;
    }

    // ---------------------------------------------------------------- surface  (_solve_for_coeffs.py:121-134, :163, :248-254)
    double ubp[N], ubm[N];  // u^+ and u^- at the bottom interface of the current layer
    {
        double bs[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double v = have_b ? bpos[i] : 0.0;
            if (beam && has_bdrf) {
                const double* q0 = A.bdrf_q0 + ((A.bdrf_percol ? (long)b * A.NBDRF : 0) + m) * N;
                v = fma((mu0 * I0 / PD_PI) * q0[i], att_t, v);
            }
            bs[i] = D[i] * v;
        }
        if (has_bdrf) {
            const double* qm = A.bdrf_q + ((A.bdrf_percol ? (long)b * A.NBDRF : 0) + m) * NN;
            const double fac = (m == 0) ? 2.0 : 1.0;
            double Rs[N][N], Mc[N][N], rhs[N];
#pragma unroll
            for (int i = 0; i < N; ++i)
#pragma unroll
                for (int k = 0; k < N; ++k) Rs[i][k] = fac * qm[i * N + k] * D[i] * D[k];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                double r = bs[i];
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    double s = (i == j) ? 1.0 : 0.0;
#pragma unroll
                    for (int k = 0; k < N; ++k) s = fma(-Rs[i][k], Rup[k][j], s);
                    Mc[i][j] = s;
                    r = fma(Rs[i][j], S[j], r);
                }
                rhs[i] = r;
            }
            double minpiv = 1e300;
#pragma unroll
            for (int s = 0; s < N; ++s) {
                const double piv = Mc[s][s];
                minpiv = (piv < minpiv) ? piv : minpiv;
                const double p = pd_rcp(piv);
#pragma unroll
                for (int j = s + 1; j < N; ++j) Mc[s][j] *= p;
                rhs[s] *= p;
#pragma unroll
                for (int i = 0; i < N; ++i)
                    if (i != s) {
                        const double f = Mc[i][s];
#pragma unroll
                        for (int j = s + 1; j < N; ++j) Mc[i][j] = fma(-f, Mc[s][j], Mc[i][j]);
                        rhs[i] = fma(-f, rhs[s], rhs[i]);
                    }
            }
            bad = bad || !(minpiv > 0.01);
#pragma unroll
            for (int i = 0; i < N; ++i) ubp[i] = rhs[i];
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) ubp[i] = bs[i];
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double s = S[i];
#pragma unroll
            for (int j = 0; j < N; ++j) s = fma(Rup[i][j], ubp[j], s);
            ubm[i] = s;
        }
    }

    // ---------------------------------------------------------------- back sweep
    double* Cout = A.C + sys * L * N2;
    // the interface radiances themselves (back in the unscaled basis) for the evaluation kernels
    double rD[N];
#pragma unroll
    for (int i = 0; i < N; ++i) rD[i] = 1.0 / D[i];
    auto store_interface = [&](int lev, const double (&up)[N], const double (&um)[N]) {
        if (!A.Uif) return;
        double* uo = A.Uif + pd_uif_index(b, lev, m, L, A.NF, N2);
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            pd_d2 p2, m2;
            p2.x = up[i] * rD[i]; p2.y = up[i + 1] * rD[i + 1];
            m2.x = um[i] * rD[i]; m2.y = um[i + 1] * rD[i + 1];
            *reinterpret_cast<pd_d2*>(uo + i) = p2;
            *reinterpret_cast<pd_d2*>(uo + N + i) = m2;
        }
    };
    store_interface(L, ubp, ubm);
    double att_b = att_t;  // exp(-tau*_L / mu0)
    t_bot = taus[L];
    t_top = taus[L - 1];
    for (int l = L - 1; l >= 0; --l) {
        if (l > 0) prefetch_layer(l - 1);
        const double t_ahead = taus[l > 0 ? l - 1 : 0];
        const double dtau = t_bot - t_top;
        const double at = beam ? exp(-t_top * rmu0) : 0.0;
        const double* hl = hist + (long)l * F::HIST_PER_LAYER * hs;
        double utp[N], utm[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double a = hl[(long)(2 * NN + i) * hs], c = hl[(long)(2 * NN + N + i) * hs];
#pragma unroll
            for (int j = 0; j < N; ++j) {
                a = fma(hl[(long)(i * N + j) * hs], ubp[j], a);
                c = fma(hl[(long)(NN + i * N + j) * hs], ubp[j], c);
            }
            utp[i] = a;
            utm[i] = c;
        }
        store_interface(l, utp, utm);
        double V[N][N], U[N][N];
        eigvecs(l, V, U);
        double ptp[N], ptm[N], pbp[N], pbm[N];
        particular(l, t_top, t_bot, at, att_b, ptp, ptm, pbp, pbm);
        double ps[N], ds[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double htp = utp[i] - ptp[i], htm = utm[i] - ptm[i];
            const double hbp = ubp[i] - pbp[i], hbm = ubm[i] - pbm[i];
            ps[i] = (htp + htm) + (hbp + hbm);
            ds[i] = (htp - htm) + (hbp - hbm);
        }
#pragma unroll
        for (int k = 0; k < N; ++k) {
            double gk = 0.0, s = 0.0, t = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                gk = fma(V[i][k], U[i][k], gk);
                s = fma(U[i][k], ps[i], s);
                t = fma(V[i][k], ds[i], t);
            }
            const double El = exp(-Kc[(long)l * N + k] * dtau);
            const double sc = 0.25 / (gk * (1.0 + El));  // (1 / (2 g (1 + E))) / 2
            const double ss = s * sc, tt = t * sc;
            Cout[(long)l * N2 + k] = ss + tt;
            Cout[(long)l * N2 + N + k] = ss - tt;
            bad = bad || !(fabs(ss) + fabs(tt) < 1e300);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            ubp[i] = utp[i];
            ubm[i] = utm[i];
        }
        att_b = at;
        t_bot = t_top;
        t_top = t_ahead;
    }
    return !bad;
}
