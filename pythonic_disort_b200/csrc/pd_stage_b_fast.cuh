// pd_stage_b_fast.cuh -- production version of the boundary-condition solve
// (same mathematics and the same partial-pivoting order as pd_stage_b.cuh, which
// stays as the size-generic path and as the host-debuggable statement of the
// algorithm) for compile-time N, device only.
//
// One group of LS lanes owns one (column, mode) system; lane t owns the panel
// columns {t, t+LS, ...} (the right-hand side is column 4N and lands on lane 0),
// so swaps and rank-1 updates touch each lane's own addresses only.  Per pivot
// step the group does one shuffle arg-max, one row swap and one column sweep;
// the multipliers are recomputed on the fly instead of being stored.
//
// Instead of the raw rows of U, the history keeps per layer
//   M_l = -U11^-1 U12  (2N x 2N)   and   z_l = U11^-1 y  (2N),
// obtained by one triangular solve per panel column (lanes work independently,
// no synchronisation), so the backward sweep is x_l = z_l + M_l x_{l+1}: half
// the scratch traffic of storing U and no sequential dependency chain.
#pragma once
#include "pd_stage_b.cuh"

#if defined(__CUDACC__)

template <int N>
struct PdStageBFast {
    static constexpr int N2 = 2 * N, NR = 3 * N, RC = 4 * N, LDP = 4 * N + 1, HROW = 2 * N + 1;
    static constexpr int SMEM_DOUBLES = NR * LDP + 3 * N2 + N * N;
    static constexpr long HIST_PER_LAYER = (long)N2 * HROW;
};

// arg-max of |v| over the group by two integer warp reductions: the IEEE-754 pattern of a
// non-negative double orders like an unsigned integer; the low 6 mantissa bits are replaced by
// (63 - candidate slot) so that ties (and values closer than 64 ulp) resolve to the first row,
// like LAPACK's idamax.  Returns the winning slot (0..63) and whether the maximum is exactly zero.
template <int LS>
__device__ __forceinline__ int pd_group_argmax_slot(unsigned mask, double absval, int slot, bool valid, bool& zero) {
    unsigned long long key = (unsigned long long)__double_as_longlong(absval);
    key = (key & ~63ull) | (unsigned long long)(63 - slot);
    unsigned hi = valid ? (unsigned)(key >> 32) : 0u, lo = (unsigned)key;
    const unsigned mhi = __reduce_max_sync(mask, hi);
    const unsigned mlo = __reduce_max_sync(mask, (valid && hi == mhi) ? lo : 0u);
    zero = (mhi == 0u) && ((mlo & ~63u) == 0u);
    return 63 - (int)(mlo & 63u);
}

// 1/x to within an ulp: hardware seed (MUFU.RCP64H) + two Newton steps, no slow path
__device__ __forceinline__ double pd_fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
}

template <int N, int LS>
__device__ void pd_stage_b_fast(const SubWarp<LS>& g, const PdStageB& a, int b, int m, double* sm, double* hist) {
    using F = PdStageBFast<N>;
    static_assert(LS >= 2 * N && (4 * N) % LS == 0, "a group needs at least 2N lanes and 4N must be a multiple of LS");
    constexpr int N2 = F::N2, RC = F::RC, LDP = F::LDP, HROW = F::HROW;
    const int lane = g.lane();
    const int L = a.L;
    double* P = sm;               // [3N][LDP]
    double* E = P + F::NR * LDP;  // [2N] exp(-k dtau*): layer l in [0,N), layer l+1 in [N,2N)
    double* xs = E + N2;          // [2N]
    double* vt = xs + N2;         // [2N]
    double* R = vt + N2;          // [N][N]

    const long sys = (long)b * a.NF + m;
    const double* taus = a.taus + (long)b * (L + 1);
    const double* Kc = a.K + sys * L * N;
    const double* Gc = a.G + sys * L * 2 * N * N;
    const double* Bc = a.beam ? a.Bv + sys * L * N2 : nullptr;
    const double* dthc = (a.iso && m == 0) ? a.dth + (long)b * L * a.Ns * N2 : nullptr;
    const double mu0 = a.colp[(long)b * PD_NCOLP + PD_COL_MU0];
    const double I0 = a.colp[(long)b * PD_NCOLP + PD_COL_I0];
    const bool beam = a.beam && I0 > 0.0;
    const bool has_bdrf = a.NBDRF > m;
    const bool have_b = (m == 0) || (a.NFb > 1);
    const double* bpos = a.bpos + ((long)b * a.NFb + (a.NFb > 1 ? m : 0)) * N;
    const double* bneg = a.bneg + ((long)b * a.NFb + (a.NFb > 1 ? m : 0)) * N;
    int status = 0;

    // element (r, cc) of G_l, r, cc in [0, 2N): blocks [[Gp, Gm], [Gm, Gp]]
    auto Gat = [&](int l, int r, int cc) -> double {
        const int rb = r >= N, cb = cc >= N;
        return Gc[((long)l * 2 + (rb ^ cb)) * N * N + (r - rb * N) * N + (cc - cb * N)];
    };

    if (has_bdrf) {
        const double* q = a.bdrf_q + ((a.bdrf_percol ? (long)b * a.NBDRF : 0) + m) * N * N;
        for (int idx = lane; idx < N * N; idx += LS)
            R[idx] = ((m == 0) ? 2.0 : 1.0) * q[idx] * a.mu[idx % N] * a.w[idx % N];
    }
    for (int i = lane; i < N; i += LS) E[N + i] = exp(-Kc[i] * (taus[1] - taus[0]));
    g.sync();

    // top boundary rows -> carry rows 0..N-1
#pragma unroll 1
    for (int c = lane; c <= RC; c += LS) {
        if (c < N2) {
            const double e = (c >= N) ? E[c] : 1.0;
#pragma unroll
            for (int r = 0; r < N; ++r) P[r * LDP + c] = Gat(0, N + r, c) * e;
        } else if (c == RC) {
            for (int r = 0; r < N; ++r) {
                double v = have_b ? bneg[r] : 0.0;
                if (beam) v -= Bc[N + r];
                if (dthc) v -= pd_thermal_at(dthc, a.Ns, N2, N + r, taus[0]);
                P[r * LDP + RC] = v;
            }
        }
    }
    g.sync();

    for (int l = 0; l < L; ++l) {
        const bool last = (l == L - 1);
        const int nrows = last ? N2 : F::NR;
        // ---- exponentials of layer l and l+1 ----
        double e_new = 0.0;
        if (lane < N) {
            e_new = E[N + lane];
        } else if (lane < N2 && !last) {
            e_new = exp(-Kc[(l + 1) * N + (lane - N)] * (taus[l + 2] - taus[l + 1]));
        }
        g.sync();
        if (lane < N2) E[lane] = e_new;
        g.sync();

        // ---- fill the panel: carry rows get zeros in the C_{l+1} columns, then the new rows ----
        if (!last) {
            const double att = beam ? exp(-taus[l + 1] / mu0) : 0.0;
#pragma unroll 1
            for (int c = lane; c <= RC; c += LS) {
                if (c >= N2 && c < RC) {
#pragma unroll
                    for (int r = 0; r < N; ++r) P[r * LDP + c] = 0.0;
                }
                if (c < RC) {
                    const int lay = (c >= N2) ? l + 1 : l;
                    const int cc = (c >= N2) ? c - N2 : c;
                    // columns multiplied by an exponential: C^- of layer l (cc < N) and C^+ of layer l+1 (cc >= N)
                    const bool scaled = (c < N) || (c >= 3 * N);
                    const double f = (scaled ? E[c >= N2 ? c - N2 : c] : 1.0) * ((c >= N2) ? -1.0 : 1.0);
                    const int cb = cc >= N;
                    const double* g0 = Gc + ((long)lay * 2) * N * N + (cc - cb * N);
#pragma unroll 8
                    for (int r = 0; r < N2; ++r) {
                        const int rb = r >= N;
                        P[(N + r) * LDP + c] = g0[(rb ^ cb) * N * N + (r - rb * N) * N] * f;
                    }
                } else {
                    for (int r = 0; r < N2; ++r) {
                        double v = 0.0;
                        if (beam) v = (Bc[(l + 1) * N2 + r] - Bc[l * N2 + r]) * att;
                        if (dthc)
                            v += pd_thermal_at(dthc + (long)(l + 1) * a.Ns * N2, a.Ns, N2, r, taus[l + 1]) -
                                 pd_thermal_at(dthc + (long)l * a.Ns * N2, a.Ns, N2, r, taus[l + 1]);
                        P[(N + r) * LDP + RC] = v;
                    }
                }
            }
        } else {
            const double att = beam ? exp(-taus[L] / mu0) : 0.0;
            if (dthc)
                for (int i = lane; i < N2; i += LS) vt[i] = pd_thermal_at(dthc + (long)l * a.Ns * N2, a.Ns, N2, i, taus[L]);
            g.sync();
            for (int c = lane; c <= RC; c += LS) {
                if (c < N2) {
                    for (int r = 0; r < N; ++r) {
                        double v = Gat(l, r, c);
                        if (has_bdrf)
                            for (int j = 0; j < N; ++j) v = fma(-R[r * N + j], Gat(l, N + j, c), v);
                        if (c < N) v *= E[c];
                        P[(N + r) * LDP + c] = v;
                    }
                } else if (c == RC) {
                    for (int r = 0; r < N; ++r) {
                        double v = have_b ? bpos[r] : 0.0;
                        if (dthc) {
                            v -= vt[r];
                            if (has_bdrf)
                                for (int j = 0; j < N; ++j) v = fma(R[r * N + j], vt[N + j], v);
                        }
                        if (beam) {
                            double s = -Bc[l * N2 + r];
                            if (has_bdrf) {
                                const double* q0 = a.bdrf_q0 + ((a.bdrf_percol ? (long)b * a.NBDRF : 0) + m) * N;
                                s += (mu0 * I0 / PD_PI) * q0[r];
                                for (int j = 0; j < N; ++j) s = fma(R[r * N + j], Bc[l * N2 + N + j], s);
                            }
                            v = fma(s, att, v);
                        }
                        P[(N + r) * LDP + RC] = v;
                    }
                } else {  // unused C_{l+1} columns of the last stage
                    for (int r = 0; r < N2; ++r) P[r * LDP + c] = 0.0;
                }
            }
        }
        g.sync();

        // ---- eliminate the 2N columns of C_l, partial pivoting over the panel rows ----
        for (int j = 0; j < N2; ++j) {
            // candidate rows j .. nrows-1, up to two per lane (slot = row - j)
            int p;
            {
                const int r0 = j + lane, r1 = j + lane + LS;
                double v0 = (r0 < nrows) ? fabs(P[r0 * LDP + j]) : 0.0;
                int slot = lane;
                bool valid = r0 < nrows;
                if (F::NR > LS && r1 < nrows) {
                    const double v1 = fabs(P[r1 * LDP + j]);
                    if (v1 > v0) {
                        v0 = v1;
                        slot = lane + LS;
                    }
                }
                bool zero;
                p = j + pd_group_argmax_slot<LS>(g.mask, v0, slot, valid, zero);
                if (zero) status |= PD_ST_ZERO_PIVOT;
            }
            // own column to the right of j (lanes whose columns are finished help with nothing;
            // lane 0 always has the right-hand side, column 4N)
            const int cfirst = (lane > j) ? lane : lane + ((j - lane) / LS + 1) * LS;
            if (p != j) {
                for (int c = cfirst; c <= RC; c += LS) {
                    const double t = P[j * LDP + c];
                    P[j * LDP + c] = P[p * LDP + c];
                    P[p * LDP + c] = t;
                }
                if (lane == 1 % LS) {  // column j itself (needed for the multipliers below)
                    const double t = P[j * LDP + j];
                    P[j * LDP + j] = P[p * LDP + j];
                    P[p * LDP + j] = t;
                }
                g.sync();
            }
            const double pinv = pd_fast_rcp(P[j * LDP + j]);
            for (int c = cfirst; c <= RC; c += LS) {
                const double u = P[j * LDP + c];
                const double* mcol = P + (j + 1) * LDP + j;
                double* acol = P + (j + 1) * LDP + c;
#pragma unroll 4
                for (int r = j + 1; r < nrows; ++r, mcol += LDP, acol += LDP) *acol = fma(-(*mcol * pinv), u, *acol);
            }
            g.sync();
            if (lane == 0) P[j * LDP + j] = pinv;
        }
        g.sync();

        if (!last) {
            // ---- M_l = -U11^-1 U12, z_l = U11^-1 y: one triangular solve per panel column, lanes independent ----
            double* h = hist + (long)l * F::HIST_PER_LAYER;
#pragma unroll 1
            for (int c = lane; c <= N2; c += LS) {
                const int pc = (c < N2) ? N2 + c : RC;
                const double sgn = (c < N2) ? -1.0 : 1.0;
                if constexpr (N2 <= 16) {  // column in registers
                    double x[N2];
#pragma unroll
                    for (int r = 0; r < N2; ++r) x[r] = P[r * LDP + pc];
#pragma unroll
                    for (int r = N2 - 1; r >= 0; --r) {
                        x[r] *= P[r * LDP + r];
#pragma unroll
                        for (int k = 0; k < r; ++k) x[k] = fma(-P[k * LDP + r], x[r], x[k]);
                    }
#pragma unroll
                    for (int r = 0; r < N2; ++r) h[r * HROW + c] = sgn * x[r];
                } else {  // in place in the panel
                    double* col = P + pc;
                    for (int r = N2 - 1; r >= 0; --r) {
                        const double xr = col[r * LDP] * P[r * LDP + r];
                        h[r * HROW + c] = sgn * xr;
#pragma unroll 4
                        for (int k = 0; k < r; ++k) col[k * LDP] = fma(-P[k * LDP + r], xr, col[k * LDP]);
                    }
                }
            }
            g.sync();
            // ---- remaining N rows (C_{l+1} columns and rhs) become the next carry ----
#pragma unroll 1
            for (int c = lane; c <= N2; c += LS) {
                const int src = (c < N2) ? N2 + c : RC, dst = (c < N2) ? c : RC;
#pragma unroll
                for (int r = 0; r < N; ++r) P[r * LDP + dst] = P[(N2 + r) * LDP + src];
            }
            g.sync();
        }
    }

    // ---- last layer: x = U11^-1 y (one lane), then sweep upwards x_l = z_l + M_l x_{l+1} ----
    double* Cout = a.C + sys * L * N2;
    if (lane == 0) {
        double* col = P + RC;
        for (int r = N2 - 1; r >= 0; --r) {
            const double xr = col[r * LDP] * P[r * LDP + r];
            xs[r] = xr;
#pragma unroll 4
            for (int k = 0; k < r; ++k) col[k * LDP] = fma(-P[k * LDP + r], xr, col[k * LDP]);
        }
    }
    g.sync();
    for (int i = lane; i < N2; i += LS) Cout[(long)(L - 1) * N2 + i] = xs[i];
    for (int l = L - 2; l >= 0; --l) {
        const double* h = hist + (long)l * F::HIST_PER_LAYER;
        for (int r = lane; r < N2; r += LS) {
            double s = h[r * HROW + N2];
#pragma unroll
            for (int c = 0; c < N2; ++c) s = fma(h[r * HROW + c], xs[c], s);
            vt[r] = s;
        }
        g.sync();
        for (int i = lane; i < N2; i += LS) {
            const double v = vt[i];
            xs[i] = v;
            Cout[(long)l * N2 + i] = v;
        }
        g.sync();
    }
    if (status && lane == 0) atomicOr(a.status + b, status);
}

#endif  // __CUDACC__
