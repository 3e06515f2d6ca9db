// pd_stage_b_row.cuh -- register-resident boundary-condition solve for N = 4 and 8
// (production path; pd_stage_b_fast.cuh covers N = 16, pd_stage_b.cuh any N and the host build).
//
// Same per-layer panel as the other variants (N carried rows + 2N new rows, 4N
// columns + right-hand side, partial pivoting over the rows still in play), but
//   * one group of LS >= 3N lanes per (column, mode) system and LANE = PANEL ROW:
//     every lane keeps its row (4N+1 doubles) in registers, statically indexed;
//   * pivot search = two integer warp reductions on |a[j]| of the rows in play;
//   * the pivot lane publishes its row once to shared memory (128-bit stores,
//     double-buffered, one __syncwarp per step); every lane reads it back with
//     128-bit broadcast loads and updates its own row:  a[c] -= (a[j]/piv) * u[c];
//   * rows are never swapped: a pivot row simply leaves the game (`active = false`);
//   * Gauss-Jordan: earlier pivot rows keep being updated (in SIMT those lanes would
//     idle anyway), so after the 2N steps pivot row r holds the rows of
//     U11^-1 [U12 | y] directly -- no triangular solves, no U in shared memory:
//     M_l = -row * (1/piv), z_l = rhs * (1/piv) go straight to the history buffer;
//   * the N rows that were never pivots are the next stage's carry; shifting them by
//     2N columns is a static register move, the freed lanes load the next interface.
//
// Shared memory per system: two row buffers + a few 2N-vectors (~1 KB instead of ~7 KB);
// per eliminated panel row the LSU sees 2 x 128-bit broadcast wavefronts per 4 FMAs
// instead of 5 wavefronts per FMA (the shared-memory panel version was LSU bound,
// profiles/r1_summary*.md).
#pragma once
#include <type_traits>

#include "pd_stage_b_fast.cuh"

#if defined(__CUDACC__)

template <int N>
struct PdStageBRow {
    static constexpr int N2 = 2 * N, NR = 3 * N, RC = 4 * N, NCOL = 4 * N + 1, LDB = 4 * N + 2, HROW = 2 * N + 1;
    static constexpr int SMEM_FIXED = 2 * LDB + 3 * N2 + N * N;
    PD_HD static int smem_doubles(int L) { return SMEM_FIXED + L * N + L + 1; }  // + exp(-k dtau*)[L][N], exp(-tau*/mu0)[L+1]
    static constexpr long HIST_PER_LAYER = (long)N2 * HROW;
};

template <int LO, int HI, class F>
__device__ __forceinline__ void pd_static_for(F&& f) {
    if constexpr (LO < HI) {
        f(std::integral_constant<int, LO>{});
        pd_static_for<LO + 1, HI>(f);
    }
}

template <int N, int LS, bool SHFL_BCAST>
__device__ void pd_stage_b_row(const SubWarp<LS>& g, const PdStageB& A, int b, int m, double* sm, double* hist) {
    using F = PdStageBRow<N>;
    static_assert(LS >= 3 * N, "one lane per panel row");
    constexpr int N2 = F::N2, NR = F::NR, RC = F::RC, NCOL = F::NCOL, LDB = F::LDB;
    const int lane = g.lane();
    const int L = A.L;
    double* buf = sm;            // [2][LDB] published pivot row (double buffered), 16-byte aligned
    double* xs = buf + 2 * LDB + N2;  // [2N]
    double* vt = xs + N2;        // [2N]
    double* R = vt + N2;         // [N][N]
    double* Eall = R + N * N;    // [L][N]  exp(-k_l dtau*_l), all layers, computed once per system
    double* att = Eall + (long)A.L * N;  // [L+1] exp(-tau*_l / mu0)

    const long sys = (long)b * A.NF + m;
    const double* taus = A.taus + (long)b * (L + 1);
    const double* Kc = A.K + sys * L * N;
    const double* Gc = A.G + sys * L * 2 * N * N;
    const double* Bc = A.beam ? A.Bv + sys * L * N2 : nullptr;
    const double* dthc = (A.iso && m == 0) ? A.dth + (long)b * L * A.Ns * N2 : nullptr;
    const double mu0 = A.colp[(long)b * PD_NCOLP + PD_COL_MU0];
    const double I0 = A.colp[(long)b * PD_NCOLP + PD_COL_I0];
    const bool beam = A.beam && I0 > 0.0;
    const bool has_bdrf = A.NBDRF > m;
    const bool have_b = (m == 0) || (A.NFb > 1);
    const double* bpos = A.bpos + ((long)b * A.NFb + (A.NFb > 1 ? m : 0)) * N;
    const double* bneg = A.bneg + ((long)b * A.NFb + (A.NFb > 1 ? m : 0)) * N;
    const unsigned gbase = (threadIdx.x & 31) - lane;  // first warp lane of this group
    int status = 0;

    // row r of G_l as stored: G_l[r][cc] = blk[(r >= N) ^ (cc >= N)][r mod N][cc mod N]
    auto Grow = [&](int l, int r, int half) -> const double* {  // the N entries G_l[r][half*N .. half*N+N)
        const int rb = r >= N;
        return Gc + ((long)l * 2 + (rb ^ half)) * N * N + (r - rb * N) * N;
    };

    if (has_bdrf) {
        const double* q = A.bdrf_q + ((A.bdrf_percol ? (long)b * A.NBDRF : 0) + m) * N * N;
        for (int idx = lane; idx < N * N; idx += LS)
            R[idx] = ((m == 0) ? 2.0 : 1.0) * q[idx] * A.mu[idx % N] * A.w[idx % N];
    }
    // all exponentials of this system up front: independent, so their latency overlaps, and the layer loop
    // has no transcendental on its critical path
    for (int idx = lane; idx < L * N; idx += LS) {
        const int ll = idx / N;
        Eall[idx] = exp(-Kc[idx] * (taus[ll + 1] - taus[ll]));
    }
    if (beam)
        for (int ll = lane; ll <= L; ll += LS) att[ll] = exp(-taus[ll] / mu0);
    g.sync();
    const double* E = Eall;

    double a[NCOL];        // this lane's panel row; a[RC] is the right-hand side
    bool active = false;   // row still a pivot candidate
    bool hasrow = false;   // lane holds a row of the current panel
    int myj = -1;          // pivot step at which this row was used in the current stage (-1: not a pivot row)
    double mypinv = 0.0;
#pragma unroll
    for (int c = 0; c < NCOL; ++c) a[c] = 0.0;

    // ---- top boundary rows: lanes 0..N-1 (they play the role of the carry of stage 0) ----
    if (lane < N) {
        const int r = N + lane;  // downward streams
        const double* g0 = Grow(0, r, 0);
        const double* g1 = Grow(0, r, 1);
#pragma unroll
        for (int c = 0; c < N; ++c) {
            a[c] = g0[c];
            a[N + c] = g1[c] * Eall[c];
        }
        double v = have_b ? bneg[lane] : 0.0;
        if (beam) v -= Bc[r];
        if (dthc) v -= pd_thermal_at(dthc, A.Ns, N2, r, taus[0]);
        a[RC] = v;
        active = true;
        hasrow = true;
    } else if (lane < NR) {
        myj = 0;  // "freed": takes a new row at stage 0
    }

    for (int l = 0; l < L; ++l) {
        const bool last = (l == L - 1);
        E = Eall + (long)l * N;  // E[c]: layer l, E[N + c]: layer l + 1
        // ---- pull the G block the NEXT stage will read (G_{l+2}) towards L1 while this stage computes ----
#if defined(__CUDA_ARCH__)
        if (l + 2 < L && lane * 16 < 2 * N * N)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(Gc + ((long)(l + 2) * 2) * N * N + lane * 16));
#endif
        // ---- carry rows: their C_{l} part is what used to be the C_{l+1} part ----
        if (l > 0 && active) {
#pragma unroll
            for (int c = 0; c < N2; ++c) {
                a[c] = a[N2 + c];
                a[N2 + c] = 0.0;
            }
        }
        // ---- freed lanes take the new rows ----
        const unsigned freed = (__ballot_sync(g.mask, myj >= 0) >> gbase) & ((LS == 32) ? 0xffffffffu : ((1u << (LS & 31)) - 1u));
        if (myj >= 0) {
            const int idx = __popc(freed & ((1u << lane) - 1u));
            myj = -1;
            if (!last) {  // continuity row `idx` of interface l  (_solve_for_coeffs.py:317-323, :242-245, :184-205)
                const double* g0 = Grow(l, idx, 0);
                const double* g1 = Grow(l, idx, 1);
                const double* h0 = Grow(l + 1, idx, 0);
                const double* h1 = Grow(l + 1, idx, 1);
#pragma unroll
                for (int c = 0; c < N; c += 2) {  // rows of G are 16-byte aligned: 128-bit loads
                    const pd_d2 v0 = *reinterpret_cast<const pd_d2*>(g0 + c), v1 = *reinterpret_cast<const pd_d2*>(g1 + c);
                    const pd_d2 w0 = *reinterpret_cast<const pd_d2*>(h0 + c), w1 = *reinterpret_cast<const pd_d2*>(h1 + c);
                    a[c] = v0.x * E[c];
                    a[c + 1] = v0.y * E[c + 1];
                    a[N + c] = v1.x;
                    a[N + c + 1] = v1.y;
                    a[N2 + c] = -w0.x;
                    a[N2 + c + 1] = -w0.y;
                    a[3 * N + c] = -w1.x * E[N + c];
                    a[3 * N + c + 1] = -w1.y * E[N + c + 1];
                }
                double v = 0.0;
                if (beam) v = (Bc[(l + 1) * N2 + idx] - Bc[l * N2 + idx]) * att[l + 1];
                if (dthc)
                    v += pd_thermal_at(dthc + (long)(l + 1) * A.Ns * N2, A.Ns, N2, idx, taus[l + 1]) -
                         pd_thermal_at(dthc + (long)l * A.Ns * N2, A.Ns, N2, idx, taus[l + 1]);
                a[RC] = v;
                active = hasrow = true;
            } else if (idx < N) {  // bottom boundary row `idx`  (:163, :208-232, :248-254, :289-293)
                const double* g0 = Grow(l, idx, 0);
                const double* g1 = Grow(l, idx, 1);
#pragma unroll
                for (int c = 0; c < N; ++c) {
                    double v0 = g0[c], v1 = g1[c];
                    if (has_bdrf)
                        for (int j = 0; j < N; ++j) {
                            v0 = fma(-R[idx * N + j], Grow(l, N + j, 0)[c], v0);
                            v1 = fma(-R[idx * N + j], Grow(l, N + j, 1)[c], v1);
                        }
                    a[c] = v0 * E[c];
                    a[N + c] = v1;
                    a[N2 + c] = 0.0;
                    a[3 * N + c] = 0.0;
                }
                double v = have_b ? bpos[idx] : 0.0;
                if (dthc) {
                    const double* dl = dthc + (long)l * A.Ns * N2;
                    v -= pd_thermal_at(dl, A.Ns, N2, idx, taus[L]);
                    if (has_bdrf)
                        for (int j = 0; j < N; ++j) v = fma(R[idx * N + j], pd_thermal_at(dl, A.Ns, N2, N + j, taus[L]), v);
                }
                if (beam) {
                    double s = -Bc[l * N2 + idx];
                    if (has_bdrf) {
                        const double* q0 = A.bdrf_q0 + ((A.bdrf_percol ? (long)b * A.NBDRF : 0) + m) * N;
                        s += (mu0 * I0 / PD_PI) * q0[idx];
                        for (int j = 0; j < N; ++j) s = fma(R[idx * N + j], Bc[l * N2 + N + j], s);
                    }
                    v = fma(s, att[L], v);
                }
                a[RC] = v;
                active = hasrow = true;
            } else {  // last stage needs only N new rows
#pragma unroll
                for (int c = 0; c < NCOL; ++c) a[c] = 0.0;
                active = hasrow = false;
            }
        }

        // ---- Gauss-Jordan elimination of the 2N columns of C_l, partial pivoting over the rows in play ----
        // Software pipelining: the pivot search of step j+1 (two warp reductions) is issued right after
        // column j+1 has been updated in step j, so its latency hides behind the rest of the row update; the
        // pivot lane publishes 1/pivot with its row, so the other lanes do not wait for a reciprocal.
        bool zero;
        int p = pd_group_argmax_slot<LS>(g.mask, fabs(a[0]), lane, active, zero);
        if (zero) status |= PD_ST_ZERO_PIVOT;
        pd_static_for<0, N2>([&](auto JI) {
            constexpr int j = decltype(JI)::value;
            if constexpr (SHFL_BCAST) {
                // variant: the pivot row is broadcast with warp shuffles straight from the pivot lane's
                // registers (no shared-memory round trip, no __syncwarp; twice the LSU-pipe instructions)
                const double piv = __shfl_sync(g.mask, a[j], p, LS);
                const double pinv = pd_rcp(piv);
                double mneg = 0.0;
                if (lane == p) {
                    active = false;
                    myj = j;
                    mypinv = pinv;
                } else if (hasrow) {
                    mneg = -a[j] * pinv;
                }
                const int pcur = p;
                if constexpr (j + 1 < N2) {
                    a[j + 1] = fma(mneg, __shfl_sync(g.mask, a[j + 1], pcur, LS), a[j + 1]);
                    bool z2;
                    p = pd_group_argmax_slot<LS>(g.mask, fabs(a[j + 1]), lane, active, z2);
                    if (z2) status |= PD_ST_ZERO_PIVOT;
                }
                constexpr int cs = (j + 1 < N2) ? j + 2 : j + 1;
#pragma unroll
                for (int c = cs; c < NCOL; ++c) a[c] = fma(mneg, __shfl_sync(g.mask, a[c], pcur, LS), a[c]);
            } else {
            double* pb = buf + (j & 1) * LDB;
            const double myrcp = pd_fast_rcp(a[j]);
            if (lane == p) {  // publish the pivot row (columns j..4N, 16-byte chunks) and 1/pivot
                constexpr int c0 = j & ~1;
#pragma unroll
                for (int c = c0; c < NCOL; c += 2) {
                    double2 v2;
                    v2.x = a[c];
                    v2.y = (c + 1 < NCOL) ? a[c + 1] : myrcp;
                    *reinterpret_cast<double2*>(pb + c) = v2;
                }
            }
            g.sync();
            double mneg = 0.0;
            if (lane == p) {
                active = false;
                myj = j;
                mypinv = myrcp;
            } else if (hasrow) {
                mneg = -a[j] * pb[NCOL];
            }
            if constexpr (j + 1 < N2) {  // column j+1 first, then start looking for the next pivot
                a[j + 1] = fma(mneg, pb[j + 1], a[j + 1]);
                bool z2;
                p = pd_group_argmax_slot<LS>(g.mask, fabs(a[j + 1]), lane, active, z2);
                if (z2) status |= PD_ST_ZERO_PIVOT;
            }
            constexpr int cs = (j + 1 < N2) ? j + 2 : j + 1;  // first column still to update
            constexpr int c1 = cs & ~1;
            constexpr int NPAIR = (NCOL - c1 + 1) / 2;
            pd_d2 ur[NPAIR];  // fetch the whole published row first, then the FMAs: exposes one load latency per step
#pragma unroll
            for (int q = 0; q < NPAIR; ++q) ur[q] = *reinterpret_cast<const pd_d2*>(pb + c1 + 2 * q);
#pragma unroll
            for (int q = 0; q < NPAIR; ++q) {
                const int c = c1 + 2 * q;
                if (c >= cs) a[c] = fma(mneg, ur[q].x, a[c]);
                if (c + 1 < NCOL) a[c + 1] = fma(mneg, ur[q].y, a[c + 1]);
            }
            }
        });

        if (!last) {
            // pivot row of step r now holds row r of U11^-1 [U12 | y] up to its pivot: write [M_l | z_l]
            // (history layout [column][row]: the 2N pivot lanes write 2N consecutive doubles per instruction)
            if (myj >= 0) {
                double* h = hist + (long)l * F::HIST_PER_LAYER + myj;
#pragma unroll
                for (int c = 0; c < N2; ++c) h[c * N2] = -a[N2 + c] * mypinv;
                h[N2 * N2] = a[RC] * mypinv;
                hasrow = false;
            }
        } else {
            if (myj >= 0) xs[myj] = a[RC] * mypinv;
        }
    }
    g.sync();

    // ---- back sweep: x_l = z_l + M_l x_{l+1} ----
    double* Cout = A.C + sys * L * N2;
    if (lane < N2) Cout[(long)(L - 1) * N2 + lane] = xs[lane];
    for (int l = L - 2; l >= 0; --l) {
        const double* h = hist + (long)l * F::HIST_PER_LAYER;
        double s = 0.0;
        if (lane < N2) {
            s = h[N2 * N2 + lane];
#pragma unroll
            for (int c = 0; c < N2; ++c) s = fma(h[c * N2 + lane], xs[c], s);
        }
        g.sync();
        if (lane < N2) {
            xs[lane] = s;
            Cout[(long)l * N2 + lane] = s;
        }
        g.sync();
    }
    if (status && lane == 0) atomicOr(A.status + b, status);
}

#endif  // __CUDACC__
