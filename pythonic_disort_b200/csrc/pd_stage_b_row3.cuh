// pd_stage_b_row3.cuh -- register-resident boundary-condition solve, THREE panel rows per lane
// (production path for N = 4 and 8; pd_stage_b_row.cuh is the one-row-per-lane predecessor).
//
// The one-row-per-lane kernel is bound by the shuffle pipe: every broadcast pivot-row entry
// costs two SHFL (64 bit) and feeds a single DFMA per lane, and SHFL issues at half the DFMA
// rate (profiles/r1_summary_final.md).  Here a system is owned by N lanes and every lane keeps
// three of the 3N panel rows in registers, so that
//   * lanes are 100 % occupied (3N rows on N lanes) and a warp works on 32/N systems at once;
//   * one broadcast pivot-row entry feeds three DFMAs per lane: the lane that owns the pivot row
//     publishes it to shared memory with 128-bit stores (three predicated copies of the store loop,
//     one per register row, so no register-row selects), the other lanes read it back with
//     128-bit broadcast loads -- about 0.5 load-store instructions per DFMA instead of 2 shuffles;
//   * pivot search: per-lane maximum of a 32-bit key (upper word of |a|, low 5 bits = 31 - slot)
//     then a log2(N)-step butterfly -- partial pivoting up to a relative 2^-15 in the magnitude
//     comparison, which is all the stability argument needs; every lane also prepares 1/candidate
//     so that the reciprocal of the pivot is published with the row;
//   * all warp-level primitives use the full mask: the systems of a warp run in lock step (same L,
//     same N; the kernel loop is warp-uniform and pads the last warp with a repeated system whose
//     stores are suppressed), which avoids the MATCH-based sub-mask barrier;
//   * as before rows never move: a pivot row leaves the game, Gauss-Jordan keeps updating it, and at
//     the end of the stage it holds its row of U11^-1 [U12 | y] up to 1/pivot; the N rows that were
//     never pivots are the carry of the next stage, the 2N freed slots load the next interface.
// Panel, right-hand sides and boundary rows follow _solve_for_coeffs.py:139-323 exactly as in
// pd_stage_b.cuh (the size-generic version that the host build tests).
#pragma once
#include "pd_stage_b_row.cuh"

#if defined(__CUDACC__)

template <int N>
struct PdStageBRow3 {
    static constexpr int LPS = N, RPL = 3;  // lanes per system, rows per lane
    static constexpr int N2 = 2 * N, RC = 4 * N, NCOL = 4 * N + 1, HROW = 2 * N + 1;
    static constexpr int LDB = 4 * N + 2;   // published pivot row (even length: 128-bit accesses)
    static constexpr int SMEM_FIXED = 2 * LDB + 2 * N2 + N * N;
    // buf[2][LDB], xs, 1/pivot of the stage, R, a three-layer ring of exp(-k dtau*)[N] (computed one stage ahead, one
    // value per lane: the full [L][N] table used to set the number of resident systems), exp(-tau*/mu0)[L+1]; the
    // per-system stride is 2 (mod 16) doubles so that the 32/N systems of a warp read their pivot rows from disjoint
    // 16-byte bank groups
    PD_HD static int smem_doubles(int L) { return ((SMEM_FIXED + 3 * N + L + 1 + 13) / 16) * 16 + 2; }
    static constexpr long HIST_PER_LAYER = (long)N2 * HROW;
};

template <int LPS>
__device__ __forceinline__ unsigned pd_pivot_key(double v, bool act, int slot) {
    const unsigned hi = (unsigned)__double2hiint(fabs(v));
    return act ? ((hi & ~31u) | (unsigned)(31 - slot)) : 0u;
}

template <int N>
__device__ void pd_stage_b_row3(const SubWarp<N>& g, const PdStageB& A, int b, int m, bool store, double* sm, double* hist) {
    using F = PdStageBRow3<N>;
    constexpr int LPS = F::LPS, N2 = F::N2, RC = F::RC, NCOL = F::NCOL;
    static_assert(N == 4 || N == 8, "three rows per lane: N = 4 or 8");
    const int lane = g.lane();
    const int L = A.L;
    double* buf = sm;                    // [2][LDB] published pivot row, double buffered (shared-memory broadcast variant)
    double* xs = buf + 2 * F::LDB;       // [2N]
    double* pinvs = xs + N2;             // [2N] 1/pivot of every step of the current stage
    double* R = pinvs + N2;              // [N][N]
    double* Er = R + N * N;              // [3][N]  exp(-k_l dtau*_l) of layers l, l+1 (in use) and l+2 (being computed)
    double* att = Er + 3 * N;            // [L+1]   exp(-tau*_l / mu0)

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
    const unsigned gbase = (threadIdx.x & 31) - lane;
    const unsigned below = (1u << lane) - 1u;
    int status = 0;

    auto Grow = [&](int l, int r, int half) -> const double* {  // the N entries G_l[r][half*N .. half*N+N)
        const int rb = r >= N;
        return Gc + ((long)l * 2 + (rb ^ half)) * N * N + (r - rb * N) * N;
    };

    if (has_bdrf) {
        const double* q = A.bdrf_q + ((A.bdrf_percol ? (long)b * A.NBDRF : 0) + m) * N * N;
        for (int idx = lane; idx < N * N; idx += LPS)
            R[idx] = ((m == 0) ? 2.0 : 1.0) * q[idx] * A.mu[idx % N] * A.w[idx % N];
    }
    auto stage_E = [&](int ll) {  // LPS == N: one value per lane
        Er[(ll % 3) * N + lane] = exp(-Kc[(long)ll * N + lane] * (taus[ll + 1] - taus[ll]));
    };
    stage_E(0);
    if (L > 1) stage_E(1);
    if (beam)
        for (int ll = lane; ll <= L; ll += LPS) att[ll] = exp(-taus[ll] / mu0);
    __syncwarp();

    double a[3][NCOL];   // slot (lane, r) = panel row r * LPS + lane; a[r][RC] is the right-hand side
    bool active[3];      // row still a pivot candidate
    int myj[3];          // pivot step at which the row was used in this stage (-1: not a pivot row)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < NCOL; ++c) a[r][c] = 0.0;
        active[r] = false;
        myj[r] = (r == 0) ? -1 : 0;  // register rows 1 and 2 are free at stage 0
    }

    // ---- top boundary rows: register row 0 of every lane (the carry of stage 0) ----
    {
        const int r = N + lane;  // downward streams
        const double* g0 = Grow(0, r, 0);
        const double* g1 = Grow(0, r, 1);
#pragma unroll
        for (int c = 0; c < N; ++c) {
            a[0][c] = g0[c];
            a[0][N + c] = g1[c] * Er[c];
        }
        double v = have_b ? bneg[lane] : 0.0;
        if (beam) v -= Bc[r];
        if (dthc) v -= pd_thermal_at(dthc, A.Ns, N2, r, taus[0]);
        a[0][RC] = v;
        active[0] = true;
    }

    for (int l = 0; l < L; ++l) {
        const bool last = (l == L - 1);
        const double* E = Er + (l % 3) * N;          // E[c]: layer l
        const double* E1 = Er + ((l + 1) % 3) * N;   // E1[c]: layer l + 1
        if (l + 2 < L) stage_E(l + 2);  // read two stages from now; its slot was last read in stage l - 1
        if (l + 2 < L && lane * 16 < 2 * N * N)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(Gc + ((long)(l + 2) * 2) * N * N + lane * 16));

        // ---- carry rows shift by 2N columns; freed slots take the new rows ----
        unsigned fr[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) fr[r] = (__ballot_sync(0xffffffffu, myj[r] >= 0) >> gbase) & ((1u << LPS) - 1u);
        pd_static_for<0, 3>([&](auto RI) {
            constexpr int r = decltype(RI)::value;
            if (l > 0 && active[r]) {
#pragma unroll
                for (int c = 0; c < N2; ++c) {
                    a[r][c] = a[r][N2 + c];
                    a[r][N2 + c] = 0.0;
                }
            }
            if (myj[r] >= 0) {
                int idx = __popc(fr[r] & below);
                if constexpr (r >= 1) idx += __popc(fr[0]);
                if constexpr (r >= 2) idx += __popc(fr[1]);
                myj[r] = -1;
                if (!last) {  // continuity row `idx` of interface l  (_solve_for_coeffs.py:317-323, :242-245, :184-205)
                    const double* g0 = Grow(l, idx, 0);
                    const double* g1 = Grow(l, idx, 1);
                    const double* h0 = Grow(l + 1, idx, 0);
                    const double* h1 = Grow(l + 1, idx, 1);
#pragma unroll
                    for (int c = 0; c < N; c += 2) {
                        const pd_d2 v0 = *reinterpret_cast<const pd_d2*>(g0 + c), v1 = *reinterpret_cast<const pd_d2*>(g1 + c);
                        const pd_d2 w0 = *reinterpret_cast<const pd_d2*>(h0 + c), w1 = *reinterpret_cast<const pd_d2*>(h1 + c);
                        a[r][c] = v0.x * E[c];
                        a[r][c + 1] = v0.y * E[c + 1];
                        a[r][N + c] = v1.x;
                        a[r][N + c + 1] = v1.y;
                        a[r][N2 + c] = -w0.x;
                        a[r][N2 + c + 1] = -w0.y;
                        a[r][3 * N + c] = -w1.x * E1[c];
                        a[r][3 * N + c + 1] = -w1.y * E1[c + 1];
                    }
                    double v = 0.0;
                    if (beam) v = (Bc[(l + 1) * N2 + idx] - Bc[l * N2 + idx]) * att[l + 1];
                    if (dthc)
                        v += pd_thermal_at(dthc + (long)(l + 1) * A.Ns * N2, A.Ns, N2, idx, taus[l + 1]) -
                             pd_thermal_at(dthc + (long)l * A.Ns * N2, A.Ns, N2, idx, taus[l + 1]);
                    a[r][RC] = v;
                    active[r] = true;
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
                        a[r][c] = v0 * E[c];
                        a[r][N + c] = v1;
                        a[r][N2 + c] = 0.0;
                        a[r][3 * N + c] = 0.0;
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
                    a[r][RC] = v;
                    active[r] = true;
                } else {  // the last stage has only N new rows
#pragma unroll
                    for (int c = 0; c < NCOL; ++c) a[r][c] = 0.0;
                    active[r] = false;
                }
            }
        });

        // ---- Gauss-Jordan elimination of the 2N columns of C_l, partial pivoting over the rows in play ----
        // Per step: the pivot lane publishes its pivot row and the reciprocal of the pivot (128-bit stores, double
        // buffered), one warp barrier, every lane reads the row back with broadcast loads and updates its three rows.
        // The search for the NEXT pivot (key per row, per-lane maximum, log2(N) butterfly steps) and the reciprocal of
        // each lane's own candidate are started as soon as column j+1 has been updated and are interleaved with
        // the remaining columns, so that neither the shuffle nor the MUFU/Newton latency sits on the critical path.
        unsigned kbest;   // this lane's best key for the coming step
        double rcand;     // 1 / (this lane's candidate pivot)
        int rbest;        // register row of this lane's candidate
        auto local_best = [&](auto JI) {
            constexpr int j = decltype(JI)::value;
            const unsigned k0 = pd_pivot_key<LPS>(a[0][j], active[0], lane);
            const unsigned k1 = pd_pivot_key<LPS>(a[1][j], active[1], LPS + lane);
            const unsigned k2 = pd_pivot_key<LPS>(a[2][j], active[2], 2 * LPS + lane);
            kbest = max(k0, max(k1, k2));
            rbest = (k0 == kbest) ? 0 : ((k1 == kbest) ? 1 : 2);
            rcand = pd_rcp((rbest == 0) ? a[0][j] : ((rbest == 1) ? a[1][j] : a[2][j]));
        };
        local_best(std::integral_constant<int, 0>{});
        unsigned kall = kbest;
#pragma unroll
        for (int o = LPS / 2; o > 0; o >>= 1) kall = max(kall, __shfl_xor_sync(0xffffffffu, kall, o, LPS));
        pd_static_for<0, N2>([&](auto JI) {
            constexpr int j = decltype(JI)::value;
            if ((kall & ~31u) == 0u) status |= PD_ST_ZERO_PIVOT;
            const bool mine = (kall == kbest);  // keys are unique (slot bits): exactly one lane of the system owns the pivot
            double* pb = buf + (j & 1) * F::LDB;
            constexpr int c0 = j & ~1;
            if (mine) {
                pd_static_for<0, 3>([&](auto RI) {
                    constexpr int r = decltype(RI)::value;
                    if (rbest == r) {
#pragma unroll
                        for (int c = c0; c < NCOL; c += 2) {
                            pd_d2 v2;
                            v2.x = a[r][c];
                            v2.y = (c + 1 < NCOL) ? a[r][c + 1] : rcand;
                            *reinterpret_cast<pd_d2*>(pb + c) = v2;
                        }
                    }
                });
            }
            __syncwarp();
            const pd_d2 tail = *reinterpret_cast<const pd_d2*>(pb + RC);  // (right-hand side, 1/pivot)
            const double pinv = tail.y;
            double mneg[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                mneg[r] = -a[r][j] * pinv;
                if (mine && rbest == r) {
                    active[r] = false;
                    myj[r] = j;
                    mneg[r] = 0.0;
                }
            }
            if (lane == 0) pinvs[j] = pinv;
            constexpr int cs = (j + 1 < N2) ? j + 2 : j + 1;  // first column still to update after column j+1
            constexpr int c1 = cs & ~1;
            constexpr int NPAIR = (RC - c1) / 2;             // full pairs below the (rhs, 1/pivot) pair
            constexpr int NST = (LPS == 8) ? 3 : 2;          // butterfly steps
            constexpr int CH = (NPAIR + NST) / (NST + 1);    // pairs per chunk
            unsigned knew = 0u, kt = 0u;
            if constexpr (j + 1 < N2) {
                const double u = pb[j + 1];
#pragma unroll
                for (int r = 0; r < 3; ++r) a[r][j + 1] = fma(mneg[r], u, a[r][j + 1]);
                local_best(std::integral_constant<int, (j + 1 < N2) ? j + 1 : j>{});
                knew = kbest;
                kt = __shfl_xor_sync(0xffffffffu, knew, LPS / 2, LPS);
            }
            pd_static_for<0, NST + 1>([&](auto QI) {
                constexpr int q = decltype(QI)::value;
                constexpr int lo = q * CH, hi = (q + 1) * CH < NPAIR ? (q + 1) * CH : NPAIR;
                if constexpr (hi > lo) {
                    pd_d2 ur[hi - lo];
#pragma unroll
                    for (int i = lo; i < hi; ++i) ur[i - lo] = *reinterpret_cast<const pd_d2*>(pb + c1 + 2 * i);
#pragma unroll
                    for (int i = lo; i < hi; ++i) {
                        const int c = c1 + 2 * i;
                        if (c >= cs) {
#pragma unroll
                            for (int r = 0; r < 3; ++r) a[r][c] = fma(mneg[r], ur[i - lo].x, a[r][c]);
                        }
#pragma unroll
                        for (int r = 0; r < 3; ++r) a[r][c + 1] = fma(mneg[r], ur[i - lo].y, a[r][c + 1]);
                    }
                }
                if constexpr (j + 1 < N2 && q < NST) {
                    knew = max(knew, kt);
                    if constexpr (q + 1 < NST) kt = __shfl_xor_sync(0xffffffffu, knew, LPS >> (q + 2), LPS);
                }
            });
#pragma unroll
            for (int r = 0; r < 3; ++r) a[r][RC] = fma(mneg[r], tail.x, a[r][RC]);
            kall = knew;
        });

        // ---- pivot row of step j holds row j of U11^-1 [U12 | y] up to 1/pivot: history [column][row] ----
        __syncwarp();  // pinvs of the last step
        pd_static_for<0, 3>([&](auto RI) {
            constexpr int r = decltype(RI)::value;
            if (myj[r] >= 0) {
                const double mypinv = pinvs[myj[r]];
                if (!last) {
                    double* h = hist + (long)l * F::HIST_PER_LAYER + myj[r];
#pragma unroll
                    for (int c = 0; c < N2; ++c) h[c * N2] = -a[r][N2 + c] * mypinv;
                    h[N2 * N2] = a[r][RC] * mypinv;
                } else {
                    xs[myj[r]] = a[r][RC] * mypinv;
                }
            }
        });
    }
    __syncwarp();

    // ---- back sweep: x_l = z_l + M_l x_{l+1}; every lane owns entries lane and lane + N ----
    double* Cout = A.C + sys * L * N2;
    if (store) {
        Cout[(long)(L - 1) * N2 + lane] = xs[lane];
        Cout[(long)(L - 1) * N2 + LPS + lane] = xs[LPS + lane];
    }
    double h0[N2 + 1], h1[N2 + 1];
    if (L >= 2) {
        const double* h = hist + (long)(L - 2) * F::HIST_PER_LAYER;
#pragma unroll
        for (int c = 0; c <= N2; ++c) {
            h0[c] = h[c * N2 + lane];
            h1[c] = h[c * N2 + LPS + lane];
        }
    }
    for (int l = L - 2; l >= 0; --l) {
        double s0 = h0[N2], s1 = h1[N2];
#pragma unroll
        for (int c = 0; c < N2; ++c) {
            const double x = xs[c];
            s0 = fma(h0[c], x, s0);
            s1 = fma(h1[c], x, s1);
        }
        if (l > 0) {  // next layer's history rows are independent of x: fetch them before the barrier
            const double* h = hist + (long)(l - 1) * F::HIST_PER_LAYER;
#pragma unroll
            for (int c = 0; c <= N2; ++c) {
                h0[c] = h[c * N2 + lane];
                h1[c] = h[c * N2 + LPS + lane];
            }
        }
        __syncwarp();
        xs[lane] = s0;
        xs[LPS + lane] = s1;
        if (store) {
            Cout[(long)l * N2 + lane] = s0;
            Cout[(long)l * N2 + LPS + lane] = s1;
        }
        __syncwarp();
    }
    if (status && lane == 0 && store) atomicOr(A.status + b, status);
}

#endif  // __CUDACC__
