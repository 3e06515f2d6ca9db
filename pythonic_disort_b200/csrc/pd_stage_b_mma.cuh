// pd_stage_b_mma.cuh -- boundary-condition solve on the FP64 tensor cores (production path for N = 8, 16).
//
// Measured on B200 (tools/ubench/dmma_rate.cu, profiles/r1_dmma_ubench.txt): mma.sync.m8n8k4.f64 reaches
// 37.1 TFLOP/s with ONE warp per scheduler and a single dependent chain, i.e. the vector-DFMA peak (34.1) at an
// eighth of the issue slots and with no operand broadcasts.  The register kernels (pd_stage_b_row*.cuh) are bound by
// exactly those: issue slots and the shuffle / shared-memory broadcasts of the pivot row.  So here the elimination
// of a stage is done in blocks of four pivot columns and every block's update of the rest of the panel is one
// rank-4 product on the tensor cores.
//
// One warp owns one (column, mode) system.  The 3N x 4N panel of stage l (N carried rows + 2N rows of interface l;
// _solve_for_coeffs.py:139-323 as in pd_stage_b.cuh) lives in registers as RT x CT accumulator tiles of
// mma.m8n8k4 (tile row = lane / 4, tile columns = 2 (lane % 4) + {0, 1}).  Per block of four columns J = j0 .. j0 + 3:
//   1. the lanes that hold columns J publish them to shared memory; every lane picks up the four J-values of "its"
//      row (one lane per row; the right-hand side lives in this layout and is eliminated like a fifth column);
//   2. the four pivots are found one after the other with partial pivoting over the rows still in play (32-bit
//      magnitude key + REDUX.MAX; the pivot lane's block values, multipliers, right-hand side and -1/pivot are
//      broadcast with shuffles -- shorter than a shared-memory round trip, and this chain is what bounds the kernel).  The elimination is Gauss-Jordan without normalisation (earlier pivot
//      rows keep being reduced) and is applied to the four block columns only; every row accumulates its multipliers
//      g with respect to the ORIGINAL four pivot rows:   new_row_i = row_i + sum_k g[i][k] * row_{p_k};
//   3. A = g (3N x 4) and B = the four pivot rows as they were at the start of the block (4 x 4N, published by the
//      quads that hold them) go through shared memory into operand layout and every live tile gets one DMMA:
//      C = A B + C.  (The kernel is bound by the shared-memory / shuffle pipe, profiles/: only the block columns,
//      the multipliers and the four pivot rows travel through it, not the panel.)
// After the 2N/4 blocks the pivot row of column j holds pivot_j times row j of U11^-1 [U12 | y]; M_l = -U12 and z_l
// go to the history buffer, the N rows that were never pivots are the next stage's carry (their columns shift by
// 2N: a move of whole tiles), and the freed rows load interface l + 1.  Back substitution x_l = z_l + M_l x_{l+1}.
// Rows never move; pivot choice equals dgbsv's up to ties and a relative 2^-14 in the magnitude comparison.
#pragma once
#include <type_traits>

#include "pd_stage_b_row.cuh"

#if defined(__CUDACC__)

#ifndef PD_MMA_BCAST_SHFL
#define PD_MMA_BCAST_SHFL 0  // pivot packet by warp shuffles (1) or through shared memory (0)
#endif
#ifndef PD_MMA_PFD
#define PD_MMA_PFD 12        // L2 prefetch distance (layers) of the history in the back sweep
#endif
#ifndef PD_MMA_PF
#define PD_MMA_PF 2          // layers of history in flight in the back sweep
#endif

template <int N>
struct PdStageBMma {
    static_assert(N % 8 == 0 && N <= 16, "tensor-core stage B: N = 8 or 16");
    static constexpr int N2 = 2 * N, NR = 3 * N, RC = 4 * N, RT = NR / 8, CT = RC / 8, HT = CT / 2, NB = N2 / 4;
    static constexpr int RPL = (NR + 31) / 32;  // rows per lane in the one-lane-per-row layout
    static constexpr int NT8 = N / 8;           // tiles per N columns
    static constexpr int PS = RC + 4;           // stride of a published pivot row: 4 (mod 16) doubles, so that the
                                                // operand loads of a half warp touch 16 distinct bank pairs
    static constexpr int BLK = NR * 4, PROW = 4 * PS, PKT = 2 * 8;
    static constexpr int RING = 3;              // layers of G / E / beam vector staged in shared memory (cp.async)
    static constexpr int GL = 2 * N * N;        // doubles of one layer's two G blocks
    static constexpr int SMEM_FIXED = 2 * BLK + PROW + PKT + N * N + N2 + NR + NR / 2 + NR / 2 + RING * (GL + N + N2);
    PD_HD static int smem_doubles(int L) { return (SMEM_FIXED + L + 1 + 1) & ~1; }
    static constexpr long HIST_PER_LAYER = (long)N2 * N2 + N2;  // M_l [2N][2N] row-major, z_l [2N]
    // per-slot scratch in global memory: the history of all layers, then exp(-k_l dtau*_l) [L][N] (kept out of shared
    // memory: the footprint per warp decides how many systems an SM keeps in flight)
    PD_HD static long scratch_doubles(int L) { return (long)L * HIST_PER_LAYER + (long)L * N; }
    using mask_t = typename std::conditional<(NR > 32), unsigned long long, unsigned>::type;
};

__device__ __forceinline__ int pd_popc(unsigned v) { return __popc(v); }
__device__ __forceinline__ int pd_popc(unsigned long long v) { return __popcll(v); }

// 1/x to about an ulp with a three-deep dependent chain after the hardware seed (cubic correction: e + e^2)
__device__ __forceinline__ double pd_rcp3(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    r = fma(r, fma(e, e, e), r);
    const double e2 = fma(-x, r, 1.0);  // off the chain when the seed is good to 2^-20; keeps full accuracy if it is not
    return fma(r, e2, r);
}

// 16 bytes global -> shared without passing through registers (L2 only: the source may have been written by this kernel)
__device__ __forceinline__ void pd_cp_async16(double* dst_smem, const double* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void pd_cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// D = A (8x4, row) * B (4x8, col) + D on the FP64 tensor cores
__device__ __forceinline__ void pd_dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int N>
__device__ void pd_stage_b_mma(const PdStageB& A, int b, int m, double* sm, double* hist) {
    using F = PdStageBMma<N>;
    using mask_t = typename F::mask_t;
    constexpr int N2 = F::N2, NR = F::NR, RT = F::RT, CT = F::CT, HT = F::HT, NB = F::NB, RPL = F::RPL;
    constexpr int NT8 = F::NT8, PS = F::PS;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, q = lane >> 2, t = lane & 3;
    const int L = A.L;
    double* blkV = sm;                   // [NR][4] the block's four columns, one row per slot
    double* blkA = blkV + F::BLK;        // [NR][4] A operand: the multipliers g
    double* prow = blkA + F::BLK;        // [4][PS] B operand: the block's four pivot rows as they were at its start
    double* pkt = prow + F::PROW;        // [2][8] pivot packet, double buffered
    double* R = pkt + F::PKT;            // [N][N]
    double* xs = R + N * N;              // [2N]
    double* npOf = xs + N2;              // [NR] -1/pivot of a row slot that was a pivot row in this stage
    int* colOf = reinterpret_cast<int*>(npOf + NR);  // [NR] its pivot column, -1: none
    int* kOf = colOf + NR;               // [NR] index (0..3) of a row slot among the current block's pivots, -1: none
    double* gring = npOf + NR + NR / 2 + NR / 2;  // [RING][2][N][N] G blocks of layers l, l+1 (in use) and l+2 (in flight)
    double* ering = gring + F::RING * F::GL;      // [RING][N]  exp(-k dtau*) of those layers
    double* bring = ering + F::RING * N;          // [RING][2N] beam particular solution of those layers
    double* att = bring + F::RING * N2;           // [L+1]   exp(-tau*_l / mu0)
    double* Eall = hist + (long)L * F::HIST_PER_LAYER;  // [L][N]  exp(-k_l dtau*_l), global scratch of this slot

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

    // the N entries G_l[r][half*N .. half*N+N), from the shared-memory ring (layer l must be staged)
    auto Grow = [&](int l, int r, int half) -> const double* {
        const int rb = r >= N;
        return gring + (l % F::RING) * F::GL + (rb ^ half) * N * N + (r - rb * N) * N;
    };
    // asynchronous copy of layer ll (G blocks, exp(-k dtau*), beam vector) into its ring slot
    auto stage_layer = [&](int ll) {
        const int sl = ll % F::RING;
        const double* gsrc = Gc + (long)ll * F::GL;
#pragma unroll
        for (int ch = lane; ch < F::GL / 2; ch += 32) pd_cp_async16(gring + sl * F::GL + 2 * ch, gsrc + 2 * ch);
        if (lane < N / 2) pd_cp_async16(ering + sl * N + 2 * lane, Eall + (long)ll * N + 2 * lane);
        else if (Bc && lane >= 16 && lane < 16 + N) pd_cp_async16(bring + sl * N2 + 2 * (lane - 16), Bc + (long)ll * N2 + 2 * (lane - 16));
    };
    auto bit = [](mask_t mk, int s) -> bool { return (mk >> s) & (mask_t)1; };
    auto below = [](int s) -> mask_t { return ((mask_t)1 << s) - (mask_t)1; };
    constexpr mask_t ALL = (NR == 8 * (int)sizeof(mask_t)) ? ~(mask_t)0 : (((mask_t)1 << NR) - (mask_t)1);

    if (has_bdrf) {
        const double* qm = A.bdrf_q + ((A.bdrf_percol ? (long)b * A.NBDRF : 0) + m) * N * N;
        for (int idx = lane; idx < N * N; idx += 32) R[idx] = ((m == 0) ? 2.0 : 1.0) * qm[idx] * A.mu[idx % N] * A.w[idx % N];
    }
    for (int idx = lane; idx < L * N; idx += 32) {
        const int ll = idx / N;
        Eall[idx] = exp(-Kc[idx] * (taus[ll + 1] - taus[ll]));
    }
    if (beam)
        for (int ll = lane; ll <= L; ll += 32) att[ll] = exp(-taus[ll] / mu0);
    __syncwarp();
    stage_layer(0);
    if (L > 1) stage_layer(1);
    pd_cp_async_wait_all();
    __syncwarp();

    double c[RT][CT][2];  // accumulator tiles: row slot rt*8 + q, columns ct*8 + 2t + {0, 1}
    double rhs[RPL];      // right-hand side of row slot lane + 32 r
    double mynp[RPL];     // -1/pivot of that slot if it was a pivot row in the current stage
    int pj[RPL];          // its pivot column (-1: none)
    bool act[RPL];        // still a pivot candidate
#pragma unroll
    for (int rt = 0; rt < RT; ++rt)
#pragma unroll
        for (int ct = 0; ct < CT; ++ct) c[rt][ct][0] = c[rt][ct][1] = 0.0;

    // ---- top boundary rows: slots 0 .. N-1 (the carry of stage 0) ----
#pragma unroll
    for (int rt = 0; rt < NT8; ++rt) {
        const int r = N + rt * 8 + q;  // downward streams
#pragma unroll
        for (int ct = 0; ct < HT; ++ct) {
            const int nb = ct / NT8, cc0 = (ct % NT8) * 8 + 2 * t;
            pd_d2 v = *reinterpret_cast<const pd_d2*>(Grow(0, r, nb) + cc0);
            if (nb == 1) {
                v.x *= ering[cc0];
                v.y *= ering[cc0 + 1];
            }
            c[rt][ct][0] = v.x;
            c[rt][ct][1] = v.y;
        }
    }
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int slot = lane + 32 * r;
        double v = 0.0;
        if (slot < N) {
            v = have_b ? bneg[slot] : 0.0;
            if (beam) v -= bring[N + slot];
            if (dthc) v -= pd_thermal_at(dthc, A.Ns, N2, N + slot, taus[0]);
        }
        rhs[r] = v;
        mynp[r] = 0.0;
        pj[r] = -1;
    }
    mask_t freem = ALL & ~below(N);  // slots that take new rows at the start of the coming stage
    unsigned kmin = 0xffffffffu;     // smallest pivot key seen (zero-pivot detection)

    for (int l = 0; l < L; ++l) {
        const bool last = (l == L - 1);
        // layers l and l+1 are in the ring (staged during the previous stages); layer l+2 starts its way in now and
        // has the whole elimination of this stage to arrive; farther layers are pulled from HBM into L2
        if (l > 0) {
            pd_cp_async_wait_all();
            __syncwarp();
        }
        if (l + 2 < L) stage_layer(l + 2);
        if (l + 6 < L && lane * 16 < F::GL)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(Gc + (long)(l + 6) * F::GL + lane * 16));
        const double* E = ering + (l % F::RING) * N;         // exp(-k dtau*) of layer l
        const double* E1 = ering + ((l + 1) % F::RING) * N;  // ... of layer l + 1
        const double* Bl = bring + (l % F::RING) * N2;       // beam particular solution of layer l
        const double* Bl1 = bring + ((l + 1) % F::RING) * N2;

        // ---- carry rows shift by 2N columns; freed slots load the new rows ----
        pd_static_for<0, RT>([&](auto RI) {
            constexpr int rt = decltype(RI)::value;
            const int slot = rt * 8 + q;
            if (!bit(freem, slot)) {
                if (l > 0) {
#pragma unroll
                    for (int ct = 0; ct < HT; ++ct) {
                        c[rt][ct][0] = c[rt][ct + HT][0];
                        c[rt][ct][1] = c[rt][ct + HT][1];
                        c[rt][ct + HT][0] = 0.0;
                        c[rt][ct + HT][1] = 0.0;
                    }
                }
            } else {
                const int idx = pd_popc(freem & below(slot));
                if (!last) {  // continuity row `idx` of interface l  (_solve_for_coeffs.py:317-323)
#pragma unroll
                    for (int ct = 0; ct < CT; ++ct) {
                        const int nb = ct / NT8, cc0 = (ct % NT8) * 8 + 2 * t;
                        pd_d2 v = *reinterpret_cast<const pd_d2*>(Grow(l + (nb >> 1), idx, nb & 1) + cc0);
                        if (nb == 0) {
                            const pd_d2 e = *reinterpret_cast<const pd_d2*>(E + cc0);
                            v.x *= e.x;
                            v.y *= e.y;
                        } else if (nb == 2) {
                            v.x = -v.x;
                            v.y = -v.y;
                        } else if (nb == 3) {
                            const pd_d2 e = *reinterpret_cast<const pd_d2*>(E1 + cc0);
                            v.x *= -e.x;
                            v.y *= -e.y;
                        }
                        c[rt][ct][0] = v.x;
                        c[rt][ct][1] = v.y;
                    }
                } else if (idx < N) {  // bottom boundary row `idx`  (:163, :289-293)
#pragma unroll
                    for (int ct = 0; ct < HT; ++ct) {
                        const int nb = ct / NT8, cc0 = (ct % NT8) * 8 + 2 * t;
                        const double* g0 = Grow(l, idx, nb) + cc0;
                        double v0 = g0[0], v1 = g0[1];
                        if (has_bdrf)
                            for (int j = 0; j < N; ++j) {
                                const double* gj = Grow(l, N + j, nb) + cc0;
                                v0 = fma(-R[idx * N + j], gj[0], v0);
                                v1 = fma(-R[idx * N + j], gj[1], v1);
                            }
                        if (nb == 0) {
                            v0 *= E[cc0];
                            v1 *= E[cc0 + 1];
                        }
                        c[rt][ct][0] = v0;
                        c[rt][ct][1] = v1;
                        c[rt][ct + HT][0] = 0.0;
                        c[rt][ct + HT][1] = 0.0;
                    }
                } else {  // unused slot of the last stage
#pragma unroll
                    for (int ct = 0; ct < CT; ++ct) c[rt][ct][0] = c[rt][ct][1] = 0.0;
                }
            }
        });
        // right-hand sides of the new rows  (:184-254); which rows are in play
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const int slot = lane + 32 * r;
            pj[r] = -1;
            act[r] = slot < NR;
            if (slot < NR && bit(freem, slot)) {
                const int idx = pd_popc(freem & below(slot));
                double v = 0.0;
                if (!last) {
                    if (beam) v = (Bl1[idx] - Bl[idx]) * att[l + 1];
                    if (dthc)
                        v += pd_thermal_at(dthc + (long)(l + 1) * A.Ns * N2, A.Ns, N2, idx, taus[l + 1]) -
                             pd_thermal_at(dthc + (long)l * A.Ns * N2, A.Ns, N2, idx, taus[l + 1]);
                } else if (idx < N) {
                    v = have_b ? bpos[idx] : 0.0;
                    if (dthc) {
                        const double* dl = dthc + (long)l * A.Ns * N2;
                        v -= pd_thermal_at(dl, A.Ns, N2, idx, taus[L]);
                        if (has_bdrf)
                            for (int j = 0; j < N; ++j) v = fma(R[idx * N + j], pd_thermal_at(dl, A.Ns, N2, N + j, taus[L]), v);
                    }
                    if (beam) {
                        double s = -Bl[idx];
                        if (has_bdrf) {
                            const double* q0 = A.bdrf_q0 + ((A.bdrf_percol ? (long)b * A.NBDRF : 0) + m) * N;
                            s += (mu0 * I0 / PD_PI) * q0[idx];
                            for (int j = 0; j < N; ++j) s = fma(R[idx * N + j], Bl[N + j], s);
                        }
                        v = fma(s, att[L], v);
                    }
                } else {
                    act[r] = false;  // the last stage has only N new rows
                }
                rhs[r] = v;
            }
        }

        // ---- elimination of the 2N columns of C_l in blocks of four ----
        pd_static_for<0, NB>([&](auto KB) {
            constexpr int kb = decltype(KB)::value;
            constexpr int j0 = 4 * kb, ctJ = j0 / 8, t0 = (j0 % 8) / 2, ctmin = (j0 + 4) / 8;
            // 1. the lanes that hold the block's columns publish them; one row per lane picks them up
            if ((t >> 1) == (t0 >> 1)) {
#pragma unroll
                for (int rt = 0; rt < RT; ++rt) {
                    pd_d2 v2;
                    v2.x = c[rt][ctJ][0];
                    v2.y = c[rt][ctJ][1];
                    *reinterpret_cast<pd_d2*>(blkV + (rt * 8 + q) * 4 + 2 * (t - t0)) = v2;
                }
            }
            __syncwarp();
            double v[RPL][4], g[RPL][4];
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const int slot = lane + 32 * r;
                pd_d2 lo, hi;
                lo.x = lo.y = hi.x = hi.y = 0.0;
                if (slot < NR) {
                    lo = *reinterpret_cast<const pd_d2*>(blkV + slot * 4);
                    hi = *reinterpret_cast<const pd_d2*>(blkV + slot * 4 + 2);
                }
                v[r][0] = lo.x; v[r][1] = lo.y; v[r][2] = hi.x; v[r][3] = hi.y;
            }
            // 2. four pivots, partial pivoting over the rows in play; g: new_row = row + sum_k g[k] * (original pivot row k)
            int p[4], myk[RPL];
#pragma unroll
            for (int r = 0; r < RPL; ++r) myk[r] = -1;
            pd_static_for<0, 4>([&](auto KI) {
                constexpr int k = decltype(KI)::value;
                unsigned key[RPL], kbest = 0u;
#pragma unroll
                for (int r = 0; r < RPL; ++r) {
                    const int slot = lane + 32 * r;
                    const unsigned hi = (unsigned)__double2hiint(fabs(v[r][k]));
                    key[r] = act[r] ? ((hi & ~63u) | (unsigned)(63 - slot)) : 0u;
                    kbest = max(kbest, key[r]);
                }
                int mr = 0;  // register row of this lane's candidate
                if constexpr (RPL > 1) mr = (key[RPL - 1] > key[0]) ? RPL - 1 : 0;
                // this lane's packet in case it wins: multipliers g[< k], block values v[> k], right-hand side, -1/candidate
                // (the reciprocal is computed by every lane in the shadow of the reduction)
                double pk[4], prhs, cinv;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (i < k) pk[i] = (RPL > 1 && mr) ? g[RPL - 1][i] : g[0][i];
                    else pk[i] = (RPL > 1 && mr) ? v[RPL - 1][i] : v[0][i];
                }
                prhs = (RPL > 1 && mr) ? rhs[RPL - 1] : rhs[0];
                cinv = pd_rcp3(-pk[k]);
                const unsigned kall = __reduce_max_sync(FULL, kbest);
                kmin = min(kmin, kall);
                p[k] = 63 - (int)(kall & 63u);
                double x[4], xr, npinv;
#if PD_MMA_BCAST_SHFL
                const int pl = p[k] & 31;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (i != k) x[i] = __shfl_sync(FULL, pk[i], pl);
                xr = __shfl_sync(FULL, prhs, pl);
                npinv = __shfl_sync(FULL, cinv, pl);
#else
                {
                    double* pb = pkt + (k & 1) * 8;
                    if (kbest == kall) {  // keys are unique: this lane owns the pivot row
                        pd_d2 a0, a1, a2;
                        a0.x = pk[0]; a0.y = pk[1]; a1.x = pk[2]; a1.y = pk[3]; a2.x = prhs; a2.y = cinv;
                        *reinterpret_cast<pd_d2*>(pb) = a0;
                        *reinterpret_cast<pd_d2*>(pb + 2) = a1;
                        *reinterpret_cast<pd_d2*>(pb + 4) = a2;
                    }
                    __syncwarp();
                    const pd_d2 x01 = *reinterpret_cast<const pd_d2*>(pb), x23 = *reinterpret_cast<const pd_d2*>(pb + 2);
                    const pd_d2 x45 = *reinterpret_cast<const pd_d2*>(pb + 4);
                    x[0] = x01.x; x[1] = x01.y; x[2] = x23.x; x[3] = x23.y;
                    xr = x45.x;
                    npinv = x45.y;
                }
#endif
#pragma unroll
                for (int r = 0; r < RPL; ++r) {
                    const bool mine = key[r] == kall;
                    const double f = mine ? 0.0 : v[r][k] * npinv;  // row += f * (pivot row)
#pragma unroll
                    for (int i = k + 1; i < 4; ++i) v[r][i] = fma(f, x[i], v[r][i]);
                    rhs[r] = fma(f, xr, rhs[r]);
#pragma unroll
                    for (int i = 0; i < k; ++i) g[r][i] = fma(f, x[i], g[r][i]);
                    g[r][k] = f;
                    if (mine) {
                        act[r] = false;
                        pj[r] = j0 + k;
                        myk[r] = k;
                        mynp[r] = npinv;
                    }
                }
            });
            // 3. operands, then the rank-4 update of every live tile
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const int slot = lane + 32 * r;
                if (slot < NR) {
                    pd_d2 lo, hi;
                    lo.x = g[r][0]; lo.y = g[r][1]; hi.x = g[r][2]; hi.y = g[r][3];
                    *reinterpret_cast<pd_d2*>(blkA + slot * 4) = lo;
                    *reinterpret_cast<pd_d2*>(blkA + slot * 4 + 2) = hi;
                    kOf[slot] = myk[r];
                }
            }
            // the quads that hold this block's pivot rows publish them (values as of the start of the block); which
            // slots those are travels through a small table written by the lanes that own the rows
            __syncwarp();
#pragma unroll
            for (int rt = 0; rt < RT; ++rt) {
                const int kidx = kOf[rt * 8 + q];
                if (kidx >= 0) {
                    double* dst = prow + kidx * PS + 2 * t;
#pragma unroll
                    for (int ct = ctmin; ct < CT; ++ct) {
                        pd_d2 v2;
                        v2.x = c[rt][ct][0];
                        v2.y = c[rt][ct][1];
                        *reinterpret_cast<pd_d2*>(dst + ct * 8) = v2;
                    }
                }
            }
            __syncwarp();
            double af[RT], bf[CT];
#pragma unroll
            for (int rt = 0; rt < RT; ++rt) af[rt] = blkA[(rt * 8 + q) * 4 + t];
#pragma unroll
            for (int ct = ctmin; ct < CT; ++ct) bf[ct] = prow[t * PS + ct * 8 + q];
#pragma unroll
            for (int rt = 0; rt < RT; ++rt)
#pragma unroll
                for (int ct = ctmin; ct < CT; ++ct) pd_dmma(c[rt][ct][0], c[rt][ct][1], af[rt], bf[ct]);
        });

        // ---- the pivot row of column j holds pivot_j * (row j of U11^-1 [U12 | y]) ----
        mask_t pivm;
        {
            const unsigned b0 = __ballot_sync(FULL, pj[0] >= 0);
            pivm = (mask_t)b0;
            if constexpr (RPL > 1) pivm |= (mask_t)__ballot_sync(FULL, pj[RPL - 1] >= 0) << 32 * (RPL - 1);
        }
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const int slot = lane + 32 * r;
            if (slot < NR) {
                colOf[slot] = pj[r];
                npOf[slot] = mynp[r];
            }
        }
        __syncwarp();
        double* hl = hist + (long)l * F::HIST_PER_LAYER;
        if (!last) {
#pragma unroll
            for (int rt = 0; rt < RT; ++rt) {
                const int j = colOf[rt * 8 + q];
                if (j >= 0) {
                    const double np = npOf[rt * 8 + q];
#pragma unroll
                    for (int ct = HT; ct < CT; ++ct) {
                        pd_d2 v2;
                        v2.x = c[rt][ct][0] * np;
                        v2.y = c[rt][ct][1] * np;
                        *reinterpret_cast<pd_d2*>(hl + (long)j * N2 + (ct - HT) * 8 + 2 * t) = v2;
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < RPL; ++r)
                if (pj[r] >= 0) hl[N2 * N2 + pj[r]] = -rhs[r] * mynp[r];
        } else {
#pragma unroll
            for (int r = 0; r < RPL; ++r)
                if (pj[r] >= 0) xs[pj[r]] = -rhs[r] * mynp[r];
        }
        freem = pivm;
        __syncwarp();
    }

    // ---- back sweep: x_l = z_l + M_l x_{l+1}; HL lanes per row ----
    constexpr int HL = 32 / N2, SEG = N2 / HL;
    const int jr = lane / HL, hh = lane % HL;
    double* Cout = A.C + sys * L * N2;
    if (hh == 0) Cout[(long)(L - 1) * N2 + jr] = xs[jr];
    constexpr int PF = PD_MMA_PF;  // layers of history in flight: the loads do not depend on x
    double mrow[PF][SEG], zj[PF];
    const double* hbase = hist + (long)jr * N2 + hh * SEG;
    for (int l0 = L - 2; l0 >= 0; l0 -= PF) {
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const int l = l0 - u;
            if (l >= 0) {
                const double* h = hbase + (long)l * F::HIST_PER_LAYER;
#pragma unroll
                for (int cc = 0; cc < SEG; cc += 4) {  // SEG = 8 or 32 doubles, 32-byte aligned: 256-bit loads
                    double v4[4];
                    pd_load4(h + cc, v4);
                    mrow[u][cc] = v4[0];
                    mrow[u][cc + 1] = v4[1];
                    mrow[u][cc + 2] = v4[2];
                    mrow[u][cc + 3] = v4[3];
                }
                zj[u] = (hh == 0) ? hist[(long)l * F::HIST_PER_LAYER + N2 * N2 + jr] : 0.0;
                // the history was written a whole forward sweep ago and has left L2 (the slots of one wave hold ~300 MB):
                // pull the block of the layer PD_MMA_PFD steps ahead from HBM into L2, one 128-byte line per lane
                if (l - PD_MMA_PFD >= 0) {
                    const char* pf = reinterpret_cast<const char*>(hist + (long)(l - PD_MMA_PFD) * F::HIST_PER_LAYER);
                    for (int ln = lane * 128; ln < (int)F::HIST_PER_LAYER * 8; ln += 32 * 128)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + ln));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const int l = l0 - u;
            if (l >= 0) {
                double s0 = zj[u], s1 = 0.0;
#pragma unroll
                for (int cc = 0; cc < SEG; cc += 2) {
                    const pd_d2 xv = *reinterpret_cast<const pd_d2*>(xs + hh * SEG + cc);
                    s0 = fma(mrow[u][cc], xv.x, s0);
                    s1 = fma(mrow[u][cc + 1], xv.y, s1);
                }
                double s = s0 + s1;
                if constexpr (HL == 2) s += __shfl_xor_sync(FULL, s, 1);
                __syncwarp();
                if (hh == 0) {
                    xs[jr] = s;
                    Cout[(long)l * N2 + jr] = s;
                }
                __syncwarp();
            }
        }
    }
    if ((kmin & ~63u) == 0u && lane == 0) atomicOr(A.status + b, PD_ST_ZERO_PIVOT);
}

#endif  // __CUDACC__
