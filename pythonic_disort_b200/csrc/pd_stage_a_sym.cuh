// pd_stage_a_sym.cuh -- stage A for N = 4 and 8: ONE THREAD per (column, mode, layer) item,
// everything in registers, fixed control flow (production path for the shortwave / longwave
// shapes; pd_stage_a.cuh + pd_linalg.cuh stay the general path and the fallback).
//
// In the similarity-scaled basis x^ = D x, D = diag(sqrt(w mu)), the reduced matrices are symmetric:
//     X' = -(alpha - beta)^ = diag(1/mu) - sum_{l-m odd}  omega* (2l+1) g*_l Q_l Q_l^T
//     S  = -(alpha + beta)^ = diag(1/mu) - sum_{l-m even} omega* (2l+1) g*_l Q_l Q_l^T
// and for omega* < 1 both are positive definite, so the eigenproblem of
// (alpha-beta)(alpha+beta) = X' S (_solve_for_gen_and_part_sols.py:179-183) becomes a symmetric one:
//     S = L L^T (Cholesky),  T = L^T X' L,  T W = W diag(k^2),  W orthogonal,
//     eigenvectors V^ = L^-T W,   U^ = (alpha+beta)^ V^ / k = -L W / k = -L L^T V^ / k,   (V^)^-1 = W^T L^T = V^T L L^T.
// T is diagonalised by cyclic Jacobi rotations (odd-even transposition ordering): no data-dependent branches or indices, so 32 items
// run in lock-step in one warp with all matrices in registers, and Jacobi delivers the small
// eigenvalues (omega* -> 1) to high relative accuracy.  The particular solutions reuse the
// decomposition instead of LU factorisations:
//     beam    (:209-231): p^ = V^ diag(1/(1/mu0^2 - k^2)) V^T L L^T r,  q^ = mu0 [(x1-x2) - L L^T p^]
//     thermal (:201-205): G^-1 [1/mu; -1/mu] = [y; -y],  y = (U^)^-1 D/mu = -k * W^T L^-1 (D/mu) = -k * V^T (D/mu)
// (after the rotations W is overwritten by V^ row by row while the G blocks are written, so only V^ and L are kept)
// If S is not numerically positive definite or Jacobi does not converge, the item is flagged
// (K[item][0] = NaN) and the general Hessenberg-QR kernel recomputes it.
#pragma once
#include "pd_stage_a.cuh"

#define PD_JACOBI_MAX_SWEEPS 12
#ifndef PD_JACOBI_GROUP
#define PD_JACOBI_GROUP 4  // rotations of a step whose parameters are computed together (N = 8: all of them; measured
                           // 70.4 -> 68.7 ms per SW step against groups of two, and fewer spills)
#endif

template <int N>
struct PdSym {
    static constexpr int NP = N * (N + 1) / 2;      // packed symmetric / triangular size
    static constexpr int PARK = NP + N + 2 * N;     // per-thread parked doubles: L, 1/diag(L), r, x1-x2
    PD_HD static constexpr int idx(int i, int j) {  // upper-packed index of (i, j), any order
        return (i <= j) ? (i * N - i * (i - 1) / 2 + (j - i)) : (j * N - j * (j - 1) / 2 + (i - j));
    }
};

// sm: per-thread parking area with stride `ps` between consecutive doubles of one thread
// QQ: [nm][NP] products Q_t[i] Q_t[j] (i <= j), Qs: [nm][N] scaled Legendre table, both shared by the CTA
// Returns true if the item is done, false if the general solver must recompute it.
template <int N>
PD_HD bool pd_stage_a_sym_item(const PdStageA& a, int b, int m, int l, const double* QQ, const double* Qs, double* sm,
                               int ps) {
    using P = PdSym<N>;
    constexpr int NP = P::NP;
    const int nm = a.NLeg - m;
    const long item = ((long)b * a.NF + m) * a.L + l;
    const double omega = a.omega_s[(long)b * a.L + l];
    const double* wl = a.wleg + ((long)b * a.L + l) * a.NLeg + m;
    const bool thermal = a.iso && m == 0;
    const bool beam = a.beam && a.colp[(long)b * PD_NCOLP + PD_COL_I0] > 0.0;

    // the beam coefficients of this (column, mode) are read in the assembly loop below: start that miss now
#if defined(__CUDA_ARCH__)
    if (beam) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.pmu0 + ((long)b * a.NF + m) * a.NLeg + m));
#endif
    bool active = false;  // _solve_for_gen_and_part_sols.py:119  (no short circuit: the loads go out back to back)
    for (int t = 0; t < nm; ++t) active |= (fabs((omega / 2) * wl[t]) > 1e-8);
    if (!active) {  // shortcut (:162-168): G = [[0, I], [I, 0]], K = 1/mu, B = 0
        double* Kout = a.K + item * N;
        double* Gout = a.G + pd_g_base(item, N);
        double* Bout = a.beam ? a.Bv + item * 2 * N : nullptr;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            Kout[i] = 1.0 / a.mu[i];
#pragma unroll
            for (int j = 0; j < N; ++j) {
                Gout[pd_g_off(i * N + j, N)] = 0.0;
                Gout[pd_g_off(N * N + i * N + j, N)] = (i == j) ? 1.0 : 0.0;
            }
            if (a.beam) {
                Bout[i] = 0.0;
                Bout[N + i] = 0.0;
            }
        }
        if (thermal) {  // y = -1/mu, k = 1/mu:  top = b_q(+k)/mu, bottom = -b_q(-k)/mu
            const double* sc = a.s_s + ((long)b * a.L + l) * a.Ns;
            double* dout = a.dth + ((long)b * a.L + l) * a.Ns * 2 * N;
            for (int q = 0; q < a.Ns; ++q)
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const double kin = a.mu[i];
                    double ratio = 1.0, pw = kin, bp = 0.0, bn = 0.0;
                    for (int r = q; r < a.Ns; ++r) {
                        if (r > q) {
                            ratio *= (double)r;
                            pw *= kin;
                        }
                        bp = fma(sc[r] * ratio, pw, bp);
                        bn = fma(sc[r] * ratio, ((r - q) & 1) ? pw : -pw, bn);
                    }
                    dout[q * 2 * N + i] = bp / a.mu[i];
                    dout[q * 2 * N + N + i] = -bn / a.mu[i];
                }
        }
        return true;
    }

    double* Lp = sm;                   // [NP] Cholesky factor, packed: L(i,j), i >= j, at idx(j,i)
    double* Li = sm + (long)NP * ps;   // [N] 1 / L(i,i)
    double* rp = Li + (long)N * ps;    // [N] beam right-hand side r
    double* xd = rp + (long)N * ps;    // [N] x1 - x2

    // ---- X' and S (packed symmetric), beam source vectors ----
    double Xp[NP], S[NP];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = i; j < N; ++j) {
            const double dg = (i == j) ? 1.0 / a.mu[i] : 0.0;
            Xp[P::idx(i, j)] = dg;
            S[P::idx(i, j)] = dg;
        }
    double x1[N], x2[N];
#pragma unroll
    for (int i = 0; i < N; ++i) x1[i] = x2[i] = 0.0;
    const double mu0 = a.colp[(long)b * PD_NCOLP + PD_COL_MU0];
    const double fac = beam ? a.colp[(long)b * PD_NCOLP + PD_COL_I0] / (4.0 * PD_PI) * ((m == 0) ? 1.0 : 2.0) : 0.0;
    const double* pm0 = a.pmu0 + ((long)b * a.NF + m) * a.NLeg + m;
    for (int t = 0; t < nm; ++t) {
        const double c = omega * wl[t];
        const double* qq = QQ + t * NP;
        if (t & 1) {
#pragma unroll
            for (int e = 0; e < NP; ++e) Xp[e] = fma(-c, qq[e], Xp[e]);
        } else {
#pragma unroll
            for (int e = 0; e < NP; ++e) S[e] = fma(-c, qq[e], S[e]);
        }
        if (beam) {
            const double cb = fac * c * pm0[t];
            const double* q = Qs + t * N;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double v = cb * q[i];
                x1[i] += v;
                x2[i] += (t & 1) ? v : -v;  // x2 = -(M^-1 X-)^
            }
        }
    }
    if (beam) {  // r = (x1 + x2)/mu0 + (alpha-beta)^ (x1 - x2) = (x1 + x2)/mu0 - X' (x1 - x2)
        double d[N];
#pragma unroll
        for (int i = 0; i < N; ++i) d[i] = x1[i] - x2[i];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double s = (x1[i] + x2[i]) / mu0;
#pragma unroll
            for (int j = 0; j < N; ++j) s = fma(-Xp[P::idx(i, j)], d[j], s);
            rp[(long)i * ps] = s;
            xd[(long)i * ps] = d[i];
        }
    }

    // ---- Cholesky S = L L^T (in place), parked in shared memory ----
    bool ok = true;
    double linv[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        double d = S[P::idx(j, j)];
#pragma unroll
        for (int k = 0; k < j; ++k) d = fma(-S[P::idx(k, j)], S[P::idx(k, j)], d);  // L(j,k) lives at idx(k,j)
        ok = ok && (d > 0.0);
        const double ri = pd_rsqrt(d > 0.0 ? d : 1.0);
        linv[j] = ri;
        S[P::idx(j, j)] = d * ri;
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            double s = S[P::idx(j, i)];
#pragma unroll
            for (int k = 0; k < j; ++k) s = fma(-S[P::idx(k, i)], S[P::idx(k, j)], s);
            S[P::idx(j, i)] = s * ri;  // L(i,j)
        }
    }
#pragma unroll
    for (int e = 0; e < NP; ++e) Lp[(long)e * ps] = S[e];
#pragma unroll
    for (int j = 0; j < N; ++j) Li[(long)j * ps] = linv[j];
#define PD_L(i, j) Lp[(long)P::idx(j, i) * ps] /* L(i,j), i >= j */

    // ---- T = L^T X' L (packed symmetric); L is read back from its parking place to keep registers free ----
    double T[NP];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        double p[N];  // column j of X' L
#pragma unroll
        for (int r = 0; r < N; ++r) {
            double s = 0.0;
#pragma unroll
            for (int c = j; c < N; ++c) s = fma(Xp[P::idx(r, c)], PD_L(c, j), s);
            p[r] = s;
        }
#pragma unroll
        for (int i = 0; i <= j; ++i) {
            double s = 0.0;
#pragma unroll
            for (int r = i; r < N; ++r) s = fma(PD_L(r, i), p[r], s);
            T[P::idx(i, j)] = s;
        }
    }

    // ---- cyclic Jacobi: T -> diag(k^2), W accumulates the rotations ----
    double W[N * N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) W[i * N + j] = (i == j) ? 1.0 : 0.0;
    constexpr int GRP = (PD_JACOBI_GROUP < N / 2) ? PD_JACOBI_GROUP : N / 2;
    bool converged = false;
    for (int sweep = 0; sweep < PD_JACOBI_MAX_SWEEPS; ++sweep) {
        double off = 0.0, diag = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            diag += fabs(T[P::idx(i, i)]);
#pragma unroll
            for (int j = i + 1; j < N; ++j) off += fabs(T[P::idx(i, j)]);
        }
        converged = (off <= 1e-17 * diag);
#if defined(__CUDA_ARCH__)
        if (__all_sync(__activemask(), converged || !ok)) break;
#else
        if (converged || !ok) break;
#endif
        // Odd-even transposition ordering (Brent-Luk): even steps rotate the index pairs (0,1), (2,3), ..., odd steps
        // (1,2), (3,4), ...; after its rotation a pair trades places, so that in N steps every index has met every
        // other one exactly once (N (N-1) / 2 rotations, one sweep).  The exchange costs nothing -- it is the same pair
        // update with the two results written to each other's place -- and the pairs of a step are the same in
        // every step of that parity: the code of a sweep is ONE even + ONE odd step in a loop that stays rolled
        // (N = 8: 7 rotations, 12 KB, instead of 28 rotations unrolled over the 7 rounds of a round-robin tournament,
        // whose 150 KB per sweep went through the instruction cache once per sweep: 1.1 fetch stalls per issue).
        // The rotations of a step act on disjoint pairs, so the parameters of GRP = min(PD_JACOBI_GROUP, N/2) of them
        // are computed together from the same T -- their rsqrt / rcp chains (about 200 cycles each, far longer than
        // the 14 row/column pair updates they feed) overlap instead of serialising.
        // An item that has converged is frozen: its "rotation" is the quarter turn, which with the exchange leaves
        // both indices in place and flips one sign (a similarity transform; all later arithmetic is sign-symmetric),
        // so its results do not depend on which other items share its warp.
#pragma unroll 1
        for (int step2 = 0; step2 < N / 2; ++step2)
#pragma unroll
            for (int par = 0; par < 2; ++par)
#pragma unroll
                for (int g0 = 0; g0 < N / 2 - par; g0 += GRP) {
                    double cc[GRP], ss[GRP], tt[GRP];
                    bool tny[GRP];
#pragma unroll
                    for (int gi = 0; gi < GRP; ++gi) {
                        if (g0 + gi >= N / 2 - par) continue;
                        const int p = 2 * (g0 + gi) + par, q = p + 1;
                        const double apq = T[P::idx(p, q)], app = T[P::idx(p, p)], aqq = T[P::idx(q, q)];
                        // rotation that annihilates T(p,q); identity if it is already negligible
                        // t = tan(rotation angle) = sgn(theta) / (|theta| + sqrt(theta^2 + 1)), theta = (aqq - app) / (2 apq),
                        // written without the division by apq:  t = sgn(d) 2 apq / (|d| + sqrt(d^2 + 4 apq^2)).
                        const bool tiny = (apq * apq <= 1e-36 * fabs(app * aqq));
                        const double d = aqq - app, a2 = 2.0 * apq;
                        const double h = fma(d, d, a2 * a2);
                        const double den = fabs(d) + h * pd_rsqrt(h > 0.0 ? h : 1.0);
                        const double tn = tiny ? 0.0 : ((d >= 0.0) ? a2 : -a2) * pd_rcp(den > 0.0 ? den : 1.0);
                        const double c = pd_rsqrt(fma(tn, tn, 1.0));
                        cc[gi] = converged ? 0.0 : c;
                        ss[gi] = converged ? 1.0 : tn * c;
                        tt[gi] = tn;
                        tny[gi] = tiny;
                    }
#pragma unroll
                    for (int gi = 0; gi < GRP; ++gi) {
                        if (g0 + gi >= N / 2 - par) continue;
                        const int p = 2 * (g0 + gi) + par, q = p + 1;
                        const double c = cc[gi], s = ss[gi], tn = tt[gi];
                        const double apq = T[P::idx(p, q)], app = T[P::idx(p, p)], aqq = T[P::idx(q, q)];
                        T[P::idx(p, p)] = converged ? app : fma(tn, apq, aqq);   // rotated, then exchanged
                        T[P::idx(q, q)] = converged ? aqq : fma(-tn, apq, app);
                        T[P::idx(p, q)] = converged ? -apq : (tny[gi] ? apq : 0.0);
#pragma unroll
                        for (int r = 0; r < N; ++r) {
                            if (r != p && r != q) {
                                const double trp = T[P::idx(r, p)], trq = T[P::idx(r, q)];
                                T[P::idx(r, p)] = fma(s, trp, c * trq);
                                T[P::idx(r, q)] = fma(c, trp, -s * trq);
                            }
                            const double wp = W[r * N + p], wq = W[r * N + q];
                            W[r * N + p] = fma(s, wp, c * wq);
                            W[r * N + q] = fma(c, wp, -s * wq);
                        }
                    }
                }
    }
    double k[N], kinv[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const double lam = T[P::idx(j, j)];
        ok = ok && (lam > 0.0);
        kinv[j] = pd_rsqrt(lam > 0.0 ? lam : 1.0);
        k[j] = lam * kinv[j];
    }
    if (!(ok && converged)) return false;

    // ---- K, G blocks:  V^ = L^-T W,  U^ = -L W / k;  Gp = (V^ + U^)/(2 D),  Gm = (V^ - U^)/(2 D) ----
    // (output pointers are formed only now: nothing but `item` has to stay live across the Jacobi sweeps)
    double* Kout = a.K + item * N;
    double* Gout = a.G + pd_g_base(item, N);
    double* Bout = a.beam ? a.Bv + item * 2 * N : nullptr;
    double dinv[N];
#pragma unroll
    for (int i = 0; i < N; ++i) dinv[i] = 0.5 * pd_rsqrt(a.w[i] * a.mu[i]);
#pragma unroll
    for (int i = 0; i < N; i += 2) {
        pd_d2 k2;
        k2.x = k[i];
        k2.y = k[i + 1];
        *reinterpret_cast<pd_d2*>(Kout + i) = k2;
    }
    // Row by row from the bottom, V^ = L^-T W written over W in place: row r of U^ needs rows <= r of W (still
    // untouched), row r of V^ needs rows > r of V^ (already in place).  A finished row of Gp / Gm is 8 N bytes of
    // the lane's own block and leaves as 256-bit stores: the lanes of a warp are 16 N^2 bytes apart, so it is the
    // number of store transactions, not their width, that the L1 -> L2 path sees (profiles/: 128 x 8-byte stores
    // per lane made this phase 30 % of the kernel).
#pragma unroll
    for (int r = N - 1; r >= 0; --r) {
        double ur[N], vr[N];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            ur[j] = 0.0;
            vr[j] = W[r * N + j];
        }
#pragma unroll
        for (int c = 0; c <= r; ++c) {
            const double lrc = PD_L(r, c);
#pragma unroll
            for (int j = 0; j < N; ++j) ur[j] = fma(lrc, W[c * N + j], ur[j]);
        }
#pragma unroll
        for (int c = r + 1; c < N; ++c) {
            const double lcr = PD_L(c, r);
#pragma unroll
            for (int j = 0; j < N; ++j) vr[j] = fma(-lcr, W[c * N + j], vr[j]);
        }
        const double li = Li[(long)r * ps];
        double gp[N], gm[N];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const double v = vr[j] * li, u = -ur[j] * kinv[j];
            W[r * N + j] = v;
            gp[j] = (v + u) * dinv[r];
            gm[j] = (v - u) * dinv[r];
        }
#pragma unroll
        for (int j = 0; j < N; j += 4) {
            pd_store4(Gout + pd_g_off(r * N + j, N), gp[j], gp[j + 1], gp[j + 2], gp[j + 3]);
            pd_store4(Gout + pd_g_off(N * N + r * N + j, N), gm[j], gm[j + 1], gm[j + 2], gm[j + 3]);
        }
    }

    // helper products with the factors (W now holds V^):  V^ x,   U^ x = -L L^T V^ (x / k)
    auto apply_V = [&](const double (&x)[N], double (&out)[N]) {
#pragma unroll
        for (int r = 0; r < N; ++r) {
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < N; ++c) s = fma(W[r * N + c], x[c], s);
            out[r] = s;
        }
    };
    auto apply_U = [&](const double (&x)[N], double (&out)[N]) {
        double y[N], z[N];
#pragma unroll
        for (int r = 0; r < N; ++r) {
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < N; ++c) s = fma(W[r * N + c], x[c] * kinv[c], s);
            y[r] = s;
        }
#pragma unroll
        for (int r = 0; r < N; ++r) {  // z = L^T y
            double s = 0.0;
#pragma unroll
            for (int c = r; c < N; ++c) s = fma(PD_L(c, r), y[c], s);
            z[r] = s;
        }
#pragma unroll
        for (int r = 0; r < N; ++r) {
            double s = 0.0;
#pragma unroll
            for (int c = 0; c <= r; ++c) s = fma(PD_L(r, c), z[c], s);
            out[r] = -s;
        }
    };

    if (a.beam && !beam) {
#pragma unroll
        for (int i = 0; i < 2 * N; ++i) Bout[i] = 0.0;
    }
    if (beam) {
        double z[N], cvec[N], ph[N];
#pragma unroll
        for (int r = 0; r < N; ++r) {  // z = L^T r
            double s = 0.0;
#pragma unroll
            for (int c = r; c < N; ++c) s = fma(PD_L(c, r), rp[(long)c * ps], s);
            z[r] = s;
        }
        double zz[N];
#pragma unroll
        for (int r = 0; r < N; ++r) {  // zz = L z = S r
            double s = 0.0;
#pragma unroll
            for (int c = 0; c <= r; ++c) s = fma(PD_L(r, c), z[c], s);
            zz[r] = s;
        }
        const double mu0b = a.colp[(long)b * PD_NCOLP + PD_COL_MU0];
        const double m2 = 1.0 / (mu0b * mu0b);
#pragma unroll
        for (int j = 0; j < N; ++j) {  // c = diag(1/(1/mu0^2 - k^2)) (V^)^-1 r,  (V^)^-1 = W^T L^T = V^T L L^T
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < N; ++r) s = fma(W[r * N + j], zz[r], s);
            cvec[j] = s / (m2 - k[j] * k[j]);
        }
        apply_V(cvec, ph);  // p^ = V^ c
        // q^ = mu0 [ (x1 - x2) - L L^T p^ ]
        double lt[N];
#pragma unroll
        for (int r = 0; r < N; ++r) {
            double s = 0.0;
#pragma unroll
            for (int c = r; c < N; ++c) s = fma(PD_L(c, r), ph[c], s);
            lt[r] = s;
        }
        double bt[N], bb[N];
#pragma unroll
        for (int r = 0; r < N; ++r) {
            double s = 0.0;
#pragma unroll
            for (int c = 0; c <= r; ++c) s = fma(PD_L(r, c), lt[c], s);
            const double qh = mu0b * (xd[(long)r * ps] - s);
            bt[r] = (ph[r] + qh) * dinv[r];       // dinv carries the factor 1/2
            bb[r] = (ph[r] - qh) * dinv[r];
        }
#pragma unroll
        for (int r = 0; r < N; r += 2) {
            pd_d2 t2, b2;
            t2.x = bt[r]; t2.y = bt[r + 1];
            b2.x = bb[r]; b2.y = bb[r + 1];
            *reinterpret_cast<pd_d2*>(Bout + r) = t2;
            *reinterpret_cast<pd_d2*>(Bout + N + r) = b2;
        }
    }

    if (thermal) {
        // y = -k * W^T L^-1 (D / mu) = -k * V^T (D / mu)   (W^T = V^T L),  D_i / mu_i = sqrt(w_i / mu_i)
        double f[N], y1[N];
#pragma unroll
        for (int r = 0; r < N; ++r) f[r] = sqrt(a.w[r] / a.mu[r]);
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < N; ++r) s = fma(W[r * N + j], f[r], s);
            y1[j] = -k[j] * s;
        }
        const double* sc = a.s_s + ((long)b * a.L + l) * a.Ns;
        double* dout = a.dth + ((long)b * a.L + l) * a.Ns * 2 * N;
        for (int q = 0; q < a.Ns; ++q) {
            double dm[N], sp[N];  // t- - t+ and t- + t+  (t-+ = b_q(-+k) y1)
#pragma unroll
            for (int j = 0; j < N; ++j) {
                double ratio = 1.0, pw = kinv[j], bp = 0.0, bn = 0.0;
                for (int r = q; r < a.Ns; ++r) {
                    if (r > q) {
                        ratio *= (double)r;
                        pw *= kinv[j];
                    }
                    bp = fma(sc[r] * ratio, pw, bp);
                    bn = fma(sc[r] * ratio, ((r - q) & 1) ? pw : -pw, bn);
                }
                dm[j] = (bn - bp) * y1[j];
                sp[j] = (bn + bp) * y1[j];
            }
            double vv[N], uu[N];
            apply_V(dm, vv);
            apply_U(sp, uu);
#pragma unroll
            for (int i = 0; i < N; i += 2) {
                pd_d2 t2, b2;
                t2.x = (vv[i] + uu[i]) * dinv[i];              // Gp t- - Gm t+
                t2.y = (vv[i + 1] + uu[i + 1]) * dinv[i + 1];
                b2.x = (vv[i] - uu[i]) * dinv[i];              // Gm t- - Gp t+
                b2.y = (vv[i + 1] - uu[i + 1]) * dinv[i + 1];
                *reinterpret_cast<pd_d2*>(dout + q * 2 * N + i) = t2;
                *reinterpret_cast<pd_d2*>(dout + q * 2 * N + N + i) = b2;
            }
        }
    }
#undef PD_L
    return true;
}
