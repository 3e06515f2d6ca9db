// pd_stage_b_add.cuh -- boundary-condition solve of one (column, Fourier mode) system by block elimination over
// N x N layer operators (production path for N = 2, 4, 8, 16).
//
// The block-banded system of _solve_for_coeffs.py:139-383 couples the 2N coefficients C_l of neighbouring layers
// through the continuity of the 2N stream radiances at every interface.  Here the same equations are eliminated in
// the variables "radiances at the interfaces", where the symmetry of the layer eigenvectors
// G_l = [[V+U, V-U], [V-U, V+U]] halves the block size and removes the need for pivoting.  In the basis scaled by
// D = diag(sqrt(w_i mu_i)) (hats):
//      g_j = v^_j . u^_j  (< 0),      U^T V^ = diag(g)   =>   V^-1 = diag(1/g) U^T,   U^-1 = diag(1/g) V^T
//      d_j = -tanh(k_j dtau / 2) / g_j
//      A1 = (I + U^ d U^T)^-1,  A2 = (I + V^ d V^T)^-1          (both I + positive semidefinite)
//      R^ = A1 - A2,   T^ = A1 + A2 - I                          (reflection / transmission of the layer, symmetric)
//      u+_top = R^ u-_top + T^ u+_bot + s+,    u-_bot = T^ u-_top + R^ u+_bot + s-      (s: particular solutions)
// Forward sweep (top down), stack above interface i:  u-_i = Rup_i u+_i + S_i,  Rup_0 = 0, S_0 = b_neg:
//      (I - R^ Rup_i) [Q | q] = [T^ | R^ S_i + s+]            (Gauss-Jordan WITHOUT pivoting: energy conservation
//      Rup_{i+1} = R^ + T^ Rup_i Q                              makes I - R^ Rup column diagonally dominant in the
//      S_{i+1}   = T^ (Rup_i q + S_i) + s-                     flux basis; measured pivots >= 0.88 on every golden)
// Surface:  (I - R^_s Rup_L) u+_L = R^_s S_L + b_s.   Back sweep:  u+_i = Q u+_{i+1} + q,  u-_i = Rup_i Q u+_{i+1} + (Rup_i q + S_i).
// Coefficients of layer l from the homogeneous parts h = u - particular at its two interfaces, through the sums that
// stay well conditioned for thin and near-conservative layers (C- + C+ and C- - C+ separately):
//      C- + C+ = U^T (p^_top + p^_bot) / (2 g (1 + E)),   C- - C+ = V^T (d^_top + d^_bot) / (2 g (1 + E)),
//      p = h+ + h-,  d = h+ - h-,  E = exp(-k dtau).
// Work per layer ~ 10 N^3 flops instead of ~ 80 N^3 for the pivoted band elimination, no pivot search, and every
// operation is a column update with all lanes busy.  The identities above hold to rounding for eigenvectors from
// the symmetric (Cholesky + Jacobi) eigen stage; a system whose pivots or norms look wrong returns false and is
// redone by the pivoted band solver (pd_stage_b.cuh).  tools/proto_adding.py is the NumPy prototype of this file.
//
// Mapping: LS = Grp::size lanes per system, lane j owns column j (+ LS, ...) of every N x N matrix and component j
// of every vector; operands other lanes need travel through the system's shared-memory block as column-major
// copies (broadcast reads).  The host build (tests/hostsim) runs the same code with one lane.
#pragma once
#include "pd_stage_b.cuh"

template <int N, int LS = 1>
struct PdStageBAdd {
    static_assert(N >= 2 && N % 2 == 0, "N must be even");
    static constexpr int N2 = 2 * N, NN = N * N;
    // Every N x N matrix in shared memory is stored by lines of N doubles (a line = a column of V^, U^, R^, T^, Rup
    // or a row of the staged G blocks).  A lane reads and writes ITS OWN line j with 128-bit accesses; at N = 8 the
    // lines j and j + 2 start in the same bank, at N = 16 all lines do.  Position p of line l therefore lives at
    // p ^ sw(l): the four / eight lanes of a quarter warp (what one 128-bit wavefront serves) then touch distinct
    // banks, and a broadcast read of line l applies the same compile-time sw(l) (ncu on the unswizzled kernel: 36 %
    // of the shared-memory wavefronts were bank conflicts, all of them on own-line accesses, the staging copies and
    // the pivot-column publication).
    PD_HD static constexpr int sw(int l) {
        return (LS == 4 && N == 8) ? ((l & 2) << 1) : (LS == 8 && N == 8) ? (((l >> 1) & 3) << 1) : (LS > 1 && N == 16) ? ((l & 7) << 1) : 0;
    }
    static constexpr int LDM = N;
    static constexpr int MAT = N * LDM;
    static constexpr int LAYER = 2 * MAT + N + N2;         // one staged layer: G blocks [2][N][N], k [N], beam vector [2N]
    static constexpr int NVEC = 6;
    static constexpr int OFF_RING = 0;                     // [LAYER]: the layer in use; the next one is fetched into the
                                                           // same place as soon as this one has been unpacked
    static constexpr int OFF_VC = OFF_RING + LAYER;        // V^ column-major; later the columns of R^ (or of R^_s)
    static constexpr int OFF_UC = OFF_VC + MAT;            // U^ column-major; later the columns of T^
    static constexpr int OFF_WC = OFF_UC + MAT;            // columns of Rup
    static constexpr int OFF_CB = OFF_WC + MAT;            // [2][2N] pivot columns (double buffered)
    static constexpr int OFF_VEC = OFF_CB + 4 * N;         // [NVEC][N] vectors every lane needs
    static constexpr int RAW = OFF_VEC + NVEC * N;
    // stride between the systems of a warp.  N = 8 (four lanes per system, two systems per quarter warp): 2 (mod 16)
    // doubles = 16 bytes, so that the swizzled own-line accesses of the two systems interleave and the eight single
    // lanes that publish a pivot column hit eight different 16-byte slots.  Otherwise LS (mod 16) doubles: a 64-bit
    // access of LS consecutive doubles per system then touches every bank once per half warp.
    static constexpr int SD = ((RAW + 15) & ~15) + ((N == 8 && LS == 4) ? 2 : ((LS < 2 ? 2 : LS) % 16));
    static constexpr long HIST_PER_LAYER = 2 * NN + N2;    // Q^T, (Rup Q)^T, q, Rup q + S
};

#define PD_FOR_OWN(jj, j) for (int jj = 0, j = lane; jj < NJ; ++jj, j += LS)

// Gauss-Jordan without pivoting on the columns a lane owns: M <- I, Y <- M^-1 Y (if WITH_Y), v <- M^-1 v.
// `cbuf` [2][2N] shared; returns the smallest pivot seen.
template <class Grp, int N, bool WITH_Y>
PD_HD double pd_add_gj_solve(const Grp& g, int lane, double (&Mc)[N / Grp::size][N], double (&Y)[N / Grp::size][N],
                             double (&vq)[N / Grp::size], double* cbuf) {
    constexpr int LS = Grp::size, NJ = N / LS;
    double minpiv = 1e300;
#pragma unroll
    for (int s = 0; s < N; ++s) {
        double* cb = cbuf + (s & 1) * 2 * N;
#pragma unroll
        PD_FOR_OWN(jj, j)
            if (j == s) {
#pragma unroll
                for (int i = 0; i < N; i += 2) {
                    pd_d2 c2;
                    c2.x = Mc[jj][i]; c2.y = Mc[jj][i + 1];
                    *reinterpret_cast<pd_d2*>(cb + i) = c2;
                }
                cb[N] = vq[jj];
            }
        g.sync();
        double f[N];
#pragma unroll
        for (int i = 0; i < N; ++i) f[i] = cb[i];
        const double piv = f[s];
        minpiv = (piv < minpiv) ? piv : minpiv;  // a NaN pivot fails the caller's test through the results
        const double p = pd_rcp(piv);
        const double vs = cb[N] * p;
#pragma unroll
        PD_FOR_OWN(jj, j) {
            const double r = Mc[jj][s] * p;
            double y = 0.0;
            if (WITH_Y) y = Y[jj][s] * p;
#pragma unroll
            for (int i = 0; i < N; ++i)
                if (i != s) {
                    Mc[jj][i] = fma(-f[i], r, Mc[jj][i]);
                    if (WITH_Y) Y[jj][i] = fma(-f[i], y, Y[jj][i]);
                }
            Mc[jj][s] = r;
            if (WITH_Y) Y[jj][s] = y;
            vq[jj] = (j == s) ? vs : fma(-cb[j], vs, vq[jj]);
        }
    }
    return minpiv;
}

// SPLIT: the layer operators R^, T^ come precomputed (A.RT, packed symmetric, pd_layer_ops.cuh) instead of being
// formed in the sweep.
template <class Grp, int N, bool SPLIT = false>
PD_HD bool pd_stage_b_add(const Grp& g, const PdStageB& A, int b, int m, double* sm, double* hist) {
    using F = PdStageBAdd<N, Grp::size>;
    constexpr int LS = Grp::size, NJ = N / LS, N2 = 2 * N, NN = N * N, LD = F::LDM, MAT = F::MAT;
    static_assert(N % LS == 0, "lanes per system must divide N");
    const int lane = g.lane();
    const int L = A.L;
    double* ring = sm + F::OFF_RING;
    double* Vc = sm + F::OFF_VC;
    double* Uc = sm + F::OFF_UC;
    double* Wc = sm + F::OFF_WC;
    double* cbuf = sm + F::OFF_CB;
    double* vec = sm + F::OFF_VEC;
    double* Rc = Vc;  // aliases: the layer operators replace the eigenvector copies once these are used up
    double* Tc = Uc;

    const long sys = (long)b * A.NF + m;
    const double* taus = A.taus + (long)b * (L + 1);
    const double* Kc = A.K + sys * L * N;
    const long item0 = sys * L;  // G item of layer l: item0 + l (layout: pd_common.cuh)
    constexpr int NP = N * (N + 1) / 2;
    const double* RTc = SPLIT ? A.RT + item0 * 2 * NP : nullptr;
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

    // layer ll (G blocks or packed R^, T^; k; beam vector) -> the staging block (asynchronously on the GPU)
    auto stage_any = [&](int ll, bool ops) {
        double* dst = ring;
        const double* ks = Kc + (long)ll * N;
        const double* gs = A.G + pd_g_base(item0 + ll, N);
#if defined(__CUDA_ARCH__)
        if (ops) {
            const double* rs = RTc + (long)ll * 2 * NP;
            for (int ch = lane; ch < NP; ch += LS) pd_cp_async16(dst + 2 * ch, rs + 2 * ch);
        } else {
            for (int ch = lane; ch < NN; ch += LS) {  // 16-byte chunk ch = (row ch / (N/2), column pair ch % (N/2)) of [2N][N]
                const int r = ch / (N / 2), c2 = 2 * (ch % (N / 2));
                pd_cp_async16(dst + r * LD + (c2 ^ F::sw(r)), gs + pd_g_off(2 * ch, N));
            }
        }
        for (int ch = lane; ch < N / 2; ch += LS) pd_cp_async16(dst + 2 * MAT + 2 * ch, ks + 2 * ch);
        if (Bc)
            for (int ch = lane; ch < N; ch += LS) pd_cp_async16(dst + 2 * MAT + N + 2 * ch, Bc + (long)ll * N2 + 2 * ch);
#else
        if (ops) {
            for (int i = 0; i < 2 * NP; ++i) dst[i] = RTc[(long)ll * 2 * NP + i];
        } else {
            for (int i = 0; i < 2 * N; ++i)
                for (int k = 0; k < N; ++k) dst[i * LD + (k ^ F::sw(i))] = gs[pd_g_off(i * N + k, N)];
        }
        for (int i = 0; i < N; ++i) dst[2 * MAT + i] = ks[i];
        if (Bc)
            for (int i = 0; i < N2; ++i) dst[2 * MAT + N + i] = Bc[(long)ll * N2 + i];
#endif
    };
    auto stage = [&](int ll) { stage_any(ll, false); };          // eigenvectors (back sweep; forward sweep if !SPLIT)
    auto stage_fwd = [&](int ll) { stage_any(ll, SPLIT); };      // what the forward sweep consumes
    auto stage_wait = [&]() {
#if defined(__CUDA_ARCH__)
        pd_cp_async_wait_all();
#endif
        g.sync();
    };
    // rows j of V^ = D (Gp + Gm) / 2 and U^ = D (Gp - Gm) / 2 -> column-major copies in shared memory (+ sync)
    double Dj[NJ];
#pragma unroll
    PD_FOR_OWN(jj, j) Dj[jj] = sqrt(A.w[j] * A.mu[j]);
    auto eigvec_columns = [&](const double* Gl, double (&vrow)[NJ][N], double (&urow)[NJ][N]) {
#pragma unroll
        PD_FOR_OWN(jj, j) {
            const double h = 0.5 * Dj[jj];
#pragma unroll
            for (int k = 0; k < N; k += 2) {
                const pd_d2 gp = *reinterpret_cast<const pd_d2*>(Gl + j * LD + (k ^ F::sw(j)));
                const pd_d2 gm = *reinterpret_cast<const pd_d2*>(Gl + MAT + j * LD + (k ^ F::sw(N + j)));
                vrow[jj][k] = h * (gp.x + gm.x);
                vrow[jj][k + 1] = h * (gp.y + gm.y);
                urow[jj][k] = h * (gp.x - gm.x);
                urow[jj][k + 1] = h * (gp.y - gm.y);
            }
#pragma unroll
            for (int k = 0; k < N; ++k) {
                Vc[k * LD + (j ^ F::sw(k))] = vrow[jj][k];
                Uc[k * LD + (j ^ F::sw(k))] = urow[jj][k];
            }
        }
        g.sync();
    };
    // y_j = sum_i col[i] x[i], x broadcast from shared memory
    auto dot_own = [&](const double (&col)[N], const double* x) -> double {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            const pd_d2 xv = *reinterpret_cast<const pd_d2*>(x + i);
            s0 = fma(col[i], xv.x, s0);
            s1 = fma(col[i + 1], xv.y, s1);
        }
        return s0 + s1;
    };
    // acc[:] += sum_k Mc[:, k] * coef[k], columns of Mc broadcast from shared memory (column-major)
    auto add_cols = [&](double (&acc)[N], const double* Mcm, const double (&coef)[N]) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const double c = coef[k];
#pragma unroll
            for (int i = 0; i < N; i += 2) {
                const pd_d2 mv = *reinterpret_cast<const pd_d2*>(Mcm + k * LD + (i ^ F::sw(k)));
                acc[i] = fma(mv.x, c, acc[i]);
                acc[i + 1] = fma(mv.y, c, acc[i + 1]);
            }
        }
    };
    // particular solution of layer l (hat basis) at the top (tau*_l, attenuation at) and bottom (tau*_{l+1}, ab)
    auto particular = [&](int l, double t_top, double t_bot, const double (&blp)[NJ], const double (&blm)[NJ], double at,
                          double ab, double (&ptp)[NJ], double (&ptm)[NJ], double (&pbp)[NJ], double (&pbm)[NJ]) {
#pragma unroll
        PD_FOR_OWN(ii, i) {
            double a = 0.0, c = 0.0, d = 0.0, e = 0.0;
            if (beam) {
                a = blp[ii] * at;
                c = blm[ii] * at;
                d = blp[ii] * ab;
                e = blm[ii] * ab;
            }
            if (dthc) {
                const double* dl = dthc + (long)l * A.Ns * N2;
                a += pd_thermal_at(dl, A.Ns, N2, i, t_top);
                c += pd_thermal_at(dl, A.Ns, N2, N + i, t_top);
                d += pd_thermal_at(dl, A.Ns, N2, i, t_bot);
                e += pd_thermal_at(dl, A.Ns, N2, N + i, t_bot);
            }
            ptp[ii] = Dj[ii] * a;
            ptm[ii] = Dj[ii] * c;
            pbp[ii] = Dj[ii] * d;
            pbm[ii] = Dj[ii] * e;
        }
    };

    // ---------------------------------------------------------------- forward sweep
    double Rup[NJ][N], S[NJ];
#pragma unroll
    PD_FOR_OWN(jj, j) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            Rup[jj][i] = 0.0;
            Wc[j * LD + (i ^ F::sw(j))] = 0.0;
        }
        S[jj] = Dj[jj] * (have_b ? bneg[j] : 0.0);
    }
    int rt_off[SPLIT ? NJ : 1][SPLIT ? N : 1];  // packed index of (i, j) for the columns j this lane owns
    if (SPLIT) {
#pragma unroll
        PD_FOR_OWN(jj, j)
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const int lo = i < j ? i : j, hi = i < j ? j : i;
                rt_off[SPLIT ? jj : 0][SPLIT ? i : 0] = lo * N - lo * (lo - 1) / 2 + (hi - lo);
            }
    }
    stage_fwd(0);
    double att_t = 1.0;  // exp(-tau*_l / mu0), tau*_0 = 0
    // the optical depths of the layer's two interfaces travel in registers, the next one is fetched a layer ahead (the
    // dependent load at the top of every layer was a quarter of the kernel's global-memory stalls)
    double t_top = taus[0], t_bot = taus[1];
    for (int l = 0; l < L; ++l) {
        const double t_ahead = taus[(l + 2 <= L) ? l + 2 : L];
        stage_wait();
        const double* Gl = ring;
        const double* Kl = Gl + 2 * MAT;
        const double* Bl = Kl + N;
        const double dtau = t_bot - t_top;
        const double att_b = beam ? exp(-t_bot * rmu0) : 0.0;

        double Rh[NJ][N], Th[NJ][N], blp[NJ], blm[NJ];
        if (SPLIT) {
            // columns j of the symmetric R^, T^ out of their packed copies; column-major copies for the other lanes
#pragma unroll
            PD_FOR_OWN(jj, j) {
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const int o = rt_off[SPLIT ? jj : 0][SPLIT ? i : 0];
                    Rh[jj][i] = Gl[o];
                    Th[jj][i] = Gl[NP + o];
                }
                blp[jj] = beam ? Bl[j] : 0.0;
                blm[jj] = beam ? Bl[N + j] : 0.0;
            }
            bad = bad || !(Gl[0] == Gl[0]);  // NaN mark of pd_layer_ops_item
            g.sync();  // every lane has unpacked the staged layer: refill it behind the arithmetic
            if (l + 1 < L) stage_fwd(l + 1);
#pragma unroll
            PD_FOR_OWN(jj, j) {
#pragma unroll
                for (int i = 0; i < N; i += 2) {
                    pd_d2 r2, t2;
                    r2.x = Rh[jj][i]; r2.y = Rh[jj][i + 1]; t2.x = Th[jj][i]; t2.y = Th[jj][i + 1];
                    *reinterpret_cast<pd_d2*>(Rc + j * LD + (i ^ F::sw(j))) = r2;
                    *reinterpret_cast<pd_d2*>(Tc + j * LD + (i ^ F::sw(j))) = t2;
                }
            }
        } else {
            double vrow[NJ][N], urow[NJ][N];
            eigvec_columns(Gl, vrow, urow);
            // d_k = -tanh(k dtau / 2) / g_k by the lane that owns k
#pragma unroll
            PD_FOR_OWN(kk, k) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int i = 0; i < N; i += 2) {
                    const pd_d2 v2 = *reinterpret_cast<const pd_d2*>(Vc + k * LD + (i ^ F::sw(k)));
                    const pd_d2 u2 = *reinterpret_cast<const pd_d2*>(Uc + k * LD + (i ^ F::sw(k)));
                    s0 = fma(v2.x, u2.x, s0);
                    s1 = fma(v2.y, u2.y, s1);
                }
                const double gk = s0 + s1;
                const double em = expm1(-Kl[k] * dtau);  // tanh(x / 2) = -expm1(-x) / (2 + expm1(-x))
                const double dk = (em / (2.0 + em)) / gk;
                bad = bad || !(gk < 0.0) || !(dk >= 0.0);
                vec[k] = dk;
                blp[kk] = beam ? Bl[k] : 0.0;
                blm[kk] = beam ? Bl[N + k] : 0.0;
            }
            g.sync();
            if (l + 1 < L) stage_fwd(l + 1);  // the staged copy of layer l is used up: fetch the next one behind the arithmetic
            // columns j of I + U^ d U^T and I + V^ d V^T
            double X1[NJ][N], X2[NJ][N];
#pragma unroll
            PD_FOR_OWN(jj, j) {
                double c1[N], c2[N];
#pragma unroll
                for (int k = 0; k < N; k += 2) {
                    const pd_d2 d2 = *reinterpret_cast<const pd_d2*>(vec + k);
                    c1[k] = d2.x * urow[jj][k];
                    c1[k + 1] = d2.y * urow[jj][k + 1];
                    c2[k] = d2.x * vrow[jj][k];
                    c2[k + 1] = d2.y * vrow[jj][k + 1];
                }
#pragma unroll
                for (int i = 0; i < N; ++i) X1[jj][i] = X2[jj][i] = (i == j) ? 1.0 : 0.0;
                add_cols(X1[jj], Uc, c1);
                add_cols(X2[jj], Vc, c2);
            }
            // both inverses in place, interleaved (Gauss-Jordan; the pivots of I + PSD are >= 1)
#pragma unroll
            for (int s = 0; s < N; ++s) {
                double* cb = cbuf + (s & 1) * 2 * N;
#pragma unroll
                PD_FOR_OWN(jj, j)
                    if (j == s) {
#pragma unroll
                        for (int i = 0; i < N; i += 2) {
                            pd_d2 a2, b2;
                            a2.x = X1[jj][i]; a2.y = X1[jj][i + 1]; b2.x = X2[jj][i]; b2.y = X2[jj][i + 1];
                            *reinterpret_cast<pd_d2*>(cb + i) = a2;
                            *reinterpret_cast<pd_d2*>(cb + N + i) = b2;
                        }
                    }
                g.sync();
                double f1[N], f2[N];
#pragma unroll
                for (int i = 0; i < N; i += 2) {
                    const pd_d2 a2 = *reinterpret_cast<const pd_d2*>(cb + i);
                    const pd_d2 b2 = *reinterpret_cast<const pd_d2*>(cb + N + i);
                    f1[i] = a2.x; f1[i + 1] = a2.y; f2[i] = b2.x; f2[i + 1] = b2.y;
                }
                const double p1 = pd_rcp(f1[s]), p2 = pd_rcp(f2[s]);
#pragma unroll
                PD_FOR_OWN(jj, j) {
                    const bool own = (j == s);  // the eliminated column is replaced by that of the inverse
                    const double r1 = (own ? 1.0 : X1[jj][s]) * p1;
                    const double r2 = (own ? 1.0 : X2[jj][s]) * p2;
#pragma unroll
                    for (int i = 0; i < N; ++i)
                        if (i != s) {
                            X1[jj][i] = fma(-f1[i], r1, own ? 0.0 : X1[jj][i]);
                            X2[jj][i] = fma(-f2[i], r2, own ? 0.0 : X2[jj][i]);
                        }
                    X1[jj][s] = r1;
                    X2[jj][s] = r2;
                }
            }
            // R^ = A1 - A2, T^ = A1 + A2 - I; their columns replace the eigenvector copies
#pragma unroll
            PD_FOR_OWN(jj, j) {
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    Rh[jj][i] = X1[jj][i] - X2[jj][i];
                    Th[jj][i] = X1[jj][i] + X2[jj][i] - ((i == j) ? 1.0 : 0.0);
                }
#pragma unroll
                for (int i = 0; i < N; i += 2) {
                    pd_d2 r2, t2;
                    r2.x = Rh[jj][i]; r2.y = Rh[jj][i + 1]; t2.x = Th[jj][i]; t2.y = Th[jj][i + 1];
                    *reinterpret_cast<pd_d2*>(Rc + j * LD + (i ^ F::sw(j))) = r2;
                    *reinterpret_cast<pd_d2*>(Tc + j * LD + (i ^ F::sw(j))) = t2;
                }
            }
        }
        // source terms
        double ptp[NJ], ptm[NJ], pbp[NJ], pbm[NJ];
        particular(l, t_top, t_bot, blp, blm, att_t, att_b, ptp, ptm, pbp, pbm);
#pragma unroll
        PD_FOR_OWN(ii, i) {
            vec[N + i] = ptm[ii];
            vec[2 * N + i] = pbp[ii];
            vec[3 * N + i] = S[ii];
        }
        g.sync();
        double sminus[NJ], vq[NJ], Mc[NJ][N], Y[NJ][N];
#pragma unroll
        PD_FOR_OWN(jj, j) {
            // rows j of the symmetric R^, T^ are the columns this lane holds
            const double splus = ptp[jj] - dot_own(Rh[jj], vec + N) - dot_own(Th[jj], vec + 2 * N);
            sminus[jj] = pbm[jj] - dot_own(Th[jj], vec + N) - dot_own(Rh[jj], vec + 2 * N);
            vq[jj] = splus + dot_own(Rh[jj], vec + 3 * N);
            double coef[N];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                Mc[jj][i] = (i == j) ? 1.0 : 0.0;
                Y[jj][i] = Th[jj][i];
                coef[i] = -Rup[jj][i];
            }
            add_cols(Mc[jj], Rc, coef);  // column j of I - R^ Rup
        }
        const double mp = pd_add_gj_solve<Grp, N, true>(g, lane, Mc, Y, vq, cbuf);
        bad = bad || !(mp > 0.05);
        // Y = Q (columns), vq = q
#pragma unroll
        PD_FOR_OWN(ii, i) vec[4 * N + i] = vq[ii];
        g.sync();
        double RQ[NJ][N], rs[NJ];
#pragma unroll
        PD_FOR_OWN(jj, j) {
#pragma unroll
            for (int i = 0; i < N; ++i) RQ[jj][i] = 0.0;
            add_cols(RQ[jj], Wc, Y[jj]);
            rs[jj] = S[jj] + dot_own(Rup[jj], vec + 4 * N);
            vec[5 * N + j] = rs[jj];
        }
        g.sync();  // also: every lane is done with the columns of Rup in Wc
        double* hl = hist + (long)l * F::HIST_PER_LAYER;
#pragma unroll
        PD_FOR_OWN(jj, j) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                hl[j * N + i] = Y[jj][i];
                hl[NN + j * N + i] = RQ[jj][i];
            }
            hl[2 * NN + j] = vq[jj];
            hl[2 * NN + N + j] = rs[jj];
#pragma unroll
            for (int i = 0; i < N; ++i) Rup[jj][i] = Rh[jj][i];
            add_cols(Rup[jj], Tc, RQ[jj]);
            S[jj] = sminus[jj] + dot_own(Th[jj], vec + 5 * N);
#pragma unroll
            for (int i = 0; i < N; i += 2) {
                pd_d2 w2;
                w2.x = Rup[jj][i]; w2.y = Rup[jj][i + 1];
                *reinterpret_cast<pd_d2*>(Wc + j * LD + (i ^ F::sw(j))) = w2;
            }
        }
        att_t = att_b;
        t_top = t_bot;
        t_bot = t_ahead;
    }
    g.sync();
    if (SPLIT) stage(L - 1);  // the back sweep starts from the eigenvectors of the last layer

    // ---------------------------------------------------------------- surface  (_solve_for_coeffs.py:121-134, :163, :248-254)
    double ubp[NJ], ubm[NJ];  // u^+ and u^- at the bottom interface of the current layer
    {
        double bs[NJ];
#pragma unroll
        PD_FOR_OWN(ii, i) {
            double v = have_b ? bpos[i] : 0.0;
            if (beam && has_bdrf) {
                const double* q0 = A.bdrf_q0 + ((A.bdrf_percol ? (long)b * A.NBDRF : 0) + m) * N;
                v = fma((mu0 * I0 / PD_PI) * q0[i], att_t, v);
            }
            bs[ii] = Dj[ii] * v;
        }
        if (has_bdrf) {
            // R^_s[i][k] = (1 + delta_m0) q[i][k] D_i D_k, stored column-major where R^ was
            const double* qm = A.bdrf_q + ((A.bdrf_percol ? (long)b * A.NBDRF : 0) + m) * NN;
            const double fac = (m == 0) ? 2.0 : 1.0;
            for (int idx = lane; idx < NN; idx += LS) {
                const int k = idx / N, i = idx - k * N;
                Rc[k * LD + (i ^ F::sw(k))] = fac * qm[i * N + k] * sqrt(A.w[i] * A.mu[i] * A.w[k] * A.mu[k]);
            }
#pragma unroll
            PD_FOR_OWN(ii, i) vec[i] = S[ii];
            g.sync();
            double Mc[NJ][N], Y[NJ][N], vq[NJ];
#pragma unroll
            PD_FOR_OWN(jj, j) {
                double coef[N];
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    Mc[jj][i] = (i == j) ? 1.0 : 0.0;
                    Y[jj][i] = 0.0;
                    coef[i] = -Rup[jj][i];
                }
                add_cols(Mc[jj], Rc, coef);  // column j of I - R^_s Rup
                double s = bs[jj];           // row j of R^_s times S
#pragma unroll
                for (int k = 0; k < N; ++k) s = fma(Rc[k * LD + (j ^ F::sw(k))], vec[k], s);
                vq[jj] = s;
            }
            const double mp = pd_add_gj_solve<Grp, N, false>(g, lane, Mc, Y, vq, cbuf);
            bad = bad || !(mp > 0.01);
#pragma unroll
            PD_FOR_OWN(ii, i) ubp[ii] = vq[ii];
        } else {
#pragma unroll
            PD_FOR_OWN(ii, i) ubp[ii] = bs[ii];
        }
#pragma unroll
        PD_FOR_OWN(ii, i) vec[N + i] = ubp[ii];
        g.sync();
#pragma unroll
        PD_FOR_OWN(ii, i) ubm[ii] = S[ii] + dot_own(Rup[ii], vec + N);
    }

    // ---------------------------------------------------------------- back sweep
    double* Cout = A.C + sys * L * N2;
    // the interface radiances themselves (back in the unscaled basis) for the evaluation kernels
    double rDj[NJ];
#pragma unroll
    PD_FOR_OWN(jj, j) rDj[jj] = 1.0 / Dj[jj];
    auto store_interface = [&](int lev, const double (&up)[NJ], const double (&um)[NJ]) {
        if (!A.Uif) return;
        double* uo = A.Uif + pd_uif_index(b, lev, m, L, A.NF, N2);
#pragma unroll
        PD_FOR_OWN(ii, i) {
            uo[i] = up[ii] * rDj[ii];
            uo[N + i] = um[ii] * rDj[ii];
        }
    };
    store_interface(L, ubp, ubm);
    double Qr[NJ][N], RQr[NJ][N], qi[NJ], rsi[NJ];
    auto load_history = [&](int l) {
        const double* hl = hist + (long)l * F::HIST_PER_LAYER;
#pragma unroll
        PD_FOR_OWN(ii, i) {
#pragma unroll
            for (int j = 0; j < N; ++j) {
                Qr[ii][j] = hl[j * N + i];
                RQr[ii][j] = hl[NN + j * N + i];
            }
            qi[ii] = hl[2 * NN + i];
            rsi[ii] = hl[2 * NN + N + i];
        }
    };
    load_history(L - 1);
    double att_b = att_t;  // exp(-tau*_L / mu0); the staging block holds (SPLIT: is being refilled with) layer L - 1
    t_bot = taus[L];
    t_top = taus[L - 1];
    for (int l = L - 1; l >= 0; --l) {
        const double t_ahead = taus[l > 0 ? l - 1 : 0];
        const double dtau = t_bot - t_top;
        const double at = beam ? exp(-t_top * rmu0) : 0.0;
#pragma unroll
        PD_FOR_OWN(ii, i) vec[i] = ubp[ii];
        stage_wait();
        const double* Gl = ring;
        const double* Kl = Gl + 2 * MAT;
        const double* Bl = Kl + N;
        double utp[NJ], utm[NJ];
#pragma unroll
        PD_FOR_OWN(ii, i) {
            utp[ii] = qi[ii] + dot_own(Qr[ii], vec);
            utm[ii] = rsi[ii] + dot_own(RQr[ii], vec);
        }
        store_interface(l, utp, utm);
        if (l > 0) load_history(l - 1);  // in flight while this layer's coefficients are recovered
        double vrow[NJ][N], urow[NJ][N], blp[NJ], blm[NJ], El[NJ];
        eigvec_columns(Gl, vrow, urow);
#pragma unroll
        PD_FOR_OWN(kk, k) {
            blp[kk] = beam ? Bl[k] : 0.0;
            blm[kk] = beam ? Bl[N + k] : 0.0;
            El[kk] = exp(-Kl[k] * dtau);
        }
        g.sync();
        if (l > 0) stage(l - 1);
        double ptp[NJ], ptm[NJ], pbp[NJ], pbm[NJ];
        particular(l, t_top, t_bot, blp, blm, at, att_b, ptp, ptm, pbp, pbm);
#pragma unroll
        PD_FOR_OWN(ii, i) {
            const double htp = utp[ii] - ptp[ii], htm = utm[ii] - ptm[ii];
            const double hbp = ubp[ii] - pbp[ii], hbm = ubm[ii] - pbm[ii];
            vec[N + i] = (htp + htm) + (hbp + hbm);
            vec[2 * N + i] = (htp - htm) + (hbp - hbm);
        }
        g.sync();
#pragma unroll
        PD_FOR_OWN(kk, k) {
            double g0 = 0.0, g1 = 0.0, s0 = 0.0, s1 = 0.0, t0 = 0.0, t1 = 0.0;
#pragma unroll
            for (int i = 0; i < N; i += 2) {
                const pd_d2 v2 = *reinterpret_cast<const pd_d2*>(Vc + k * LD + (i ^ F::sw(k)));
                const pd_d2 u2 = *reinterpret_cast<const pd_d2*>(Uc + k * LD + (i ^ F::sw(k)));
                const pd_d2 ps = *reinterpret_cast<const pd_d2*>(vec + N + i);
                const pd_d2 ds = *reinterpret_cast<const pd_d2*>(vec + 2 * N + i);
                g0 = fma(v2.x, u2.x, g0);
                g1 = fma(v2.y, u2.y, g1);
                s0 = fma(u2.x, ps.x, s0);
                s1 = fma(u2.y, ps.y, s1);
                t0 = fma(v2.x, ds.x, t0);
                t1 = fma(v2.y, ds.y, t1);
            }
            const double sc = 0.25 / ((g0 + g1) * (1.0 + El[kk]));  // (1 / (2 g (1 + E))) / 2
            const double ss = (s0 + s1) * sc, tt = (t0 + t1) * sc;
            Cout[(long)l * N2 + k] = ss + tt;
            Cout[(long)l * N2 + N + k] = ss - tt;
            bad = bad || !(fabs(ss) + fabs(tt) < 1e300);
        }
#pragma unroll
        PD_FOR_OWN(ii, i) {
            ubp[ii] = utp[ii];
            ubm[ii] = utm[ii];
        }
        att_b = at;
        t_bot = t_top;
        t_top = t_ahead;
    }
    return !g.any(bad);
}
