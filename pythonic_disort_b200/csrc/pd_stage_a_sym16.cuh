// pd_stage_a_sym16.cuh -- stage A for N = 16 (NQuad = 32): EIGHT LANES per (column, mode, layer) item, matrices in
// registers, fixed control flow (production path of the high-accuracy shape; pd_stage_a.cuh + pd_linalg.cuh stay the
// general path and the fallback for flagged items).
//
// Same mathematics as pd_stage_a_sym.cuh (_solve_for_gen_and_part_sols.py:114-231 in the similarity-scaled basis):
//     X' = -(alpha-beta)^,  S = -(alpha+beta)^ = L L^T,  T = L^T X' L,  T W = W diag(k^2),
//     V^ = L^-T W,  U^ = -L W / k,  beam:  p^ = V^ diag(1/(1/mu0^2 - k^2)) W^T L^T r,  q^ = mu0 [(x1 - x2) - L L^T p^].
// A 16 x 16 matrix does not fit one thread's registers, so the item is spread over 8 lanes, lane `lam` holding two
// COLUMNS of every matrix.  T is diagonalised by two-sided Jacobi rotations in the odd-even transposition ordering
// (16 steps per sweep: even steps rotate the column pairs (2i, 2i+1), odd steps the pairs (2i+1, 2i+2), and the two
// columns of a pair swap places after their rotation, so every pair meets once per sweep).  A lane always rotates the
// two columns it holds; after every step it keeps its second column and takes the first column of its right neighbour
// (32 doubles with W, by shuffles), which turns the even pairing into the odd one and back.  Row indices are kept RELATIVE to the
// lane's first column, rho = (row - first column) mod 16: the pivot block is then always rows 0 and 1, the row pairs
// are always (2j, 2j+1) with the rotation of lane lam + j, and every register index is a compile-time constant --
// frame changes are renamings of registers.  Jacobi keeps the small eigenvalues of near-conservative layers to high
// relative accuracy; a converged item is frozen, so its result does not depend on its warp neighbours.
// tools/proto_jacobi16.py is the NumPy model of the rotation schedule.
//
// Shared memory per item: X' (16 x 16, later P = X' L), L (16 x 16) and four vectors; the Cholesky factorisation
// broadcasts each finished column through L's buffer, the products T = L^T (X' L) read one operand from there in the
// lane's own row order (address arithmetic instead of register indices).
#pragma once
#include "pd_stage_a.cuh"

#if defined(__CUDACC__)

#define PD_J16_MAX_SWEEPS 14

struct PdJ16 {
    static constexpr int N = 16, NL = 8;
    static constexpr int OFF_X = 0;       // [16][16] X'(r, c) at c*16 + r (symmetric); later P(i, c) at i*16 + c
    static constexpr int OFF_L = 256;     // [16][16] L(i, j) at j*16 + i (column-major, zeros above the diagonal)
    static constexpr int OFF_LINV = 512;  // [16] 1 / L(i, i)
    static constexpr int OFF_V0 = 528;    // [16] x1 - x2, later t = L^T r
    static constexpr int OFF_V1 = 544;    // [16] r, later L^T p^
    static constexpr int ITEM = 568;      // doubles per item (stride = 8 mod 16: the two items of a half warp do not collide)
};

__device__ __forceinline__ void pd_ld2(const double* p, double& a, double& b) {
    const pd_d2 v = *reinterpret_cast<const pd_d2*>(p);
    a = v.x;
    b = v.y;
}
__device__ __forceinline__ void pd_st2(double* p, double a, double b) {
    pd_d2 v;
    v.x = a;
    v.y = b;
    *reinterpret_cast<pd_d2*>(p) = v;
}
// v[rho] = base[(2 lam + rho) mod 16], rho = 0 .. 15: a 16-vector read in the lane's relative order (128-bit loads)
__device__ __forceinline__ void pd_j16_load_rel(const double* base, int lam, double (&v)[16]) {
#pragma unroll
    for (int r = 0; r < 16; r += 2) pd_ld2(base + ((2 * lam + r) & 15), v[r], v[r + 1]);
}
__device__ __forceinline__ void pd_j16_load_abs(const double* base, double (&v)[16]) {
#pragma unroll
    for (int r = 0; r < 16; r += 2) pd_ld2(base + r, v[r], v[r + 1]);
}

// One Jacobi step on the pair a lane holds: A, B = the pair's columns of T (relative rows, pivot block at rows 0, 1),
// Wa, Wb = the same columns of the accumulated rotations (absolute rows).  `idle`: in an odd step one lane holds the
// two end columns of the line (15 and 0), which are not a pair; it "rotates" by a quarter turn (c = 0, s = 1), which
// together with the swap leaves both columns where they are and flips the sign of one of them -- a similarity
// transform like any other (an eigenvector changes sign), so no lane and no row pair needs a special case.
__device__ __forceinline__ void pd_j16_step(double (&A)[16], double (&B)[16], double (&Wa)[16], double (&Wb)[16], int lam,
                                            int grp, bool frozen, bool idle) {
    const double app = A[0], aqq = B[1], apq = B[0];
    // t = tan(rotation angle) = sgn(d) 2 apq / (|d| + sqrt(d^2 + 4 apq^2)), d = aqq - app  (as in pd_stage_a_sym.cuh)
    const bool tiny = frozen || (apq * apq <= 1e-36 * fabs(app * aqq));
    const double d = aqq - app, a2 = 2.0 * apq;
    const double h = fma(d, d, a2 * a2);
    const double den = fabs(d) + h * pd_rsqrt(h > 0.0 ? h : 1.0);
    const double tn = tiny ? 0.0 : ((d >= 0.0) ? a2 : -a2) * pd_rcp(den > 0.0 ? den : 1.0);
    const double c0 = pd_rsqrt(fma(tn, tn, 1.0));
    const double c = idle ? 0.0 : c0, s = idle ? 1.0 : tn * c0;
    // columns: new_p = c p - s q, new_q = s p + c q; the rotated columns swap places (new_q first)
#pragma unroll
    for (int r = 2; r < 16; ++r) {
        const double p = A[r], q = B[r];
        A[r] = fma(s, p, c * q);
        B[r] = fma(c, p, -s * q);
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const double p = Wa[r], q = Wb[r];
        Wa[r] = fma(s, p, c * q);
        Wb[r] = fma(c, p, -s * q);
    }
    {  // pivot block, exact: the first slot now holds column q
        const double tpq = idle ? -apq : (tiny ? apq : 0.0);
        A[0] = idle ? app : fma(tn, apq, aqq);
        A[1] = tpq;
        B[0] = tpq;
        B[1] = idle ? aqq : fma(-tn, apq, app);
    }
    // rows (2j, 2j+1) are the pair of lane lam + j: same rotation, same swap
#pragma unroll
    for (int j = 1; j < 8; ++j) {
        const int src = grp | ((lam + j) & 7);
        const double cj = __shfl_sync(0xffffffffu, c, src), sj = __shfl_sync(0xffffffffu, s, src);
        {
            const double p = A[2 * j], q = A[2 * j + 1];
            A[2 * j] = fma(sj, p, cj * q);
            A[2 * j + 1] = fma(cj, p, -sj * q);
        }
        {
            const double p = B[2 * j], q = B[2 * j + 1];
            B[2 * j] = fma(sj, p, cj * q);
            B[2 * j + 1] = fma(cj, p, -sj * q);
        }
    }
}

// After every step the frame moves on by one column: the lane keeps its second column as the new first one and takes
// the first column of its right neighbour (lane 7 from lane 0: the line of columns is handled as a ring, the pair
// that closes it is the idle one).  After n shifts lane lam holds columns 2 lam + n and 2 lam + n + 1 (mod 16); 16
// shifts -- one sweep -- bring every column home.
__device__ __forceinline__ void pd_j16_shift(double (&A)[16], double (&B)[16], double (&Wa)[16], double (&Wb)[16], int lam,
                                             int grp) {
    const int src = grp | ((lam + 1) & 7);
    double nA[16], nB[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        nA[r] = B[(r + 1) & 15];
        nB[r] = __shfl_sync(0xffffffffu, A[(r + 15) & 15], src);
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        A[r] = nA[r];
        B[r] = nB[r];
        const double t = __shfl_sync(0xffffffffu, Wa[r], src);
        Wa[r] = Wb[r];
        Wb[r] = t;
    }
}

// sum over the 8 lanes of an item
__device__ __forceinline__ double pd_j16_sum8(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}

// Qs: [nm][16] scaled Legendre table of this mode, tab: [0..15] 0.5 / sqrt(w mu), [16..31] 1 / mu, [32..47] sqrt(w / mu) (shared by the
// CTA); sm: this item's scratch.  All 32 lanes of a warp must call this together (`store` = false for the padding
// items of the last CTA).  Returns false if the general solver must redo the item.
__device__ __forceinline__ bool pd_stage_a_j16_item(const PdStageA& a, int b, int m, int l, bool store, const double* Qs,
                                                    const double* tab, double* sm) {
    using P = PdJ16;
    const int lane = threadIdx.x & 31, lam = lane & 7, grp = lane & 24;
    const int nm = a.NLeg - m;
    const long item = ((long)b * a.NF + m) * a.L + l;
    const double omega = a.omega_s[(long)b * a.L + l];
    const double* wl = a.wleg + ((long)b * a.L + l) * a.NLeg + m;
    const bool thermal = a.iso && m == 0;
    const bool beam = a.beam && a.colp[(long)b * PD_NCOLP + PD_COL_I0] > 0.0;
    const double* dinvs = tab;
    const double* rmu = tab + 16;
    double* XS = sm + P::OFF_X;
    double* LL = sm + P::OFF_L;
    double* Linv = sm + P::OFF_LINV;
    double* V0 = sm + P::OFF_V0;
    double* V1 = sm + P::OFF_V1;

    // omega* (2l+1) g*_l and the beam coefficients of this item, fetched by the 8 lanes together (L's buffer is free
    // until the factorisation starts)
    double* CW = LL;
    double* CB = LL + 32;
    const double mu0 = a.colp[(long)b * PD_NCOLP + PD_COL_MU0];
    const double fac = beam ? a.colp[(long)b * PD_NCOLP + PD_COL_I0] / (4.0 * PD_PI) * ((m == 0) ? 1.0 : 2.0) : 0.0;
    bool active = false;  // _solve_for_gen_and_part_sols.py:119
    {
        const double* pm0 = a.pmu0 + ((long)b * a.NF + m) * a.NLeg + m;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int t = k * 8 + lam;
            if (t < nm) {
                const double w = wl[t];
                active |= (fabs((omega / 2) * w) > 1e-8);
                CW[t] = omega * w;
                CB[t] = beam ? fac * (omega * w) * pm0[t] : 0.0;
            }
        }
        active = (__ballot_sync(0xffffffffu, active) & (0xffu << grp)) != 0u;  // also orders the stores above
        __syncwarp();
    }

    // ---- X' and S, two columns each (relative rows), beam source vectors for the lane's two streams ----
    double x1a = 0.0, x1b = 0.0, x2a = 0.0, x2b = 0.0, ra = 0.0, rb = 0.0, da = 0.0, db = 0.0;
    double TA[16], TB[16];
    bool ok = true;
    {
        double XA[16], XB[16], SA[16], SB[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) XA[r] = XB[r] = SA[r] = SB[r] = 0.0;
        {
            double m0, m1;
            pd_ld2(rmu + 2 * lam, m0, m1);
            XA[0] = SA[0] = m0;
            XB[1] = SB[1] = m1;
        }
        for (int t = 0; t < nm; ++t) {
            const double c = CW[t];
            double q[16];
            pd_j16_load_rel(Qs + t * 16, lam, q);
            const double ca = -c * q[0], cb = -c * q[1];
            if (t & 1) {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    XA[r] = fma(ca, q[r], XA[r]);
                    XB[r] = fma(cb, q[r], XB[r]);
                }
            } else {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    SA[r] = fma(ca, q[r], SA[r]);
                    SB[r] = fma(cb, q[r], SB[r]);
                }
            }
            if (beam) {
                const double cbm = CB[t];
                const double va = cbm * q[0], vb = cbm * q[1];
                x1a += va;
                x1b += vb;
                x2a += (t & 1) ? va : -va;  // x2 = -(M^-1 X-)^
                x2b += (t & 1) ? vb : -vb;
            }
        }
        // X' to shared memory (column c = row c at XS + 16 c, absolute rows)
#pragma unroll
        for (int r = 0; r < 16; r += 2) {
            pd_st2(XS + (2 * lam) * 16 + ((2 * lam + r) & 15), XA[r], XA[r + 1]);
            pd_st2(XS + (2 * lam + 1) * 16 + ((2 * lam + r) & 15), XB[r], XB[r + 1]);
        }
        if (beam) {  // r = (x1 + x2) / mu0 - X' (x1 - x2), rows 2 lam and 2 lam + 1
            da = x1a - x2a;
            db = x1b - x2b;
            pd_st2(V0 + 2 * lam, da, db);
            __syncwarp();
            double dv[16];
            pd_j16_load_rel(V0, lam, dv);
            double s0 = (x1a + x2a) / mu0, s1 = (x1b + x2b) / mu0;
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                s0 = fma(-XA[r], dv[r], s0);
                s1 = fma(-XB[r], dv[r], s1);
            }
            ra = s0;
            rb = s1;
            pd_st2(V1 + 2 * lam, ra, rb);
        }

        __syncwarp();  // the coefficient lists in L's buffer have been read
        // ---- Cholesky S = L L^T, right-looking: the owner of column j publishes it, everybody updates ----
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int o = j >> 1;
            if (lam == o) {
                const double dj = (j & 1) ? SB[1] : SA[0];
                ok = ok && (dj > 0.0);
                const double ri = pd_rsqrt(dj > 0.0 ? dj : 1.0);
                Linv[j] = ri;
#pragma unroll
                for (int r = 0; r < 16; r += 2) {
                    const int i = (2 * o + r) & 15;  // rows i, i + 1 (compile-time for the owner)
                    const double v0 = (j & 1) ? SB[r] : SA[r], v1 = (j & 1) ? SB[r + 1] : SA[r + 1];
                    pd_st2(LL + j * 16 + i, (i >= j) ? v0 * ri : 0.0, (i + 1 >= j) ? v1 * ri : 0.0);
                }
            }
            __syncwarp();
            double Lj[16];
            pd_j16_load_rel(LL + j * 16, lam, Lj);
            if (2 * lam > j) {
#pragma unroll
                for (int r = 0; r < 16; ++r) SA[r] = fma(-Lj[r], Lj[0], SA[r]);
            }
            if (2 * lam + 1 > j) {
#pragma unroll
                for (int r = 0; r < 16; ++r) SB[r] = fma(-Lj[r], Lj[1], SB[r]);
            }
        }
    }
    __syncwarp();

    // ---- T = L^T X' L:  P = X' L by columns, published row-major, then T(:, c) = sum_i P(i, :)^T L(i, c) ----
    double ta = 0.0, tb = 0.0;  // (L^T r) for the lane's two columns
    {
        double LcA[16], LcB[16];  // the lane's two columns of L, absolute rows
        pd_j16_load_abs(LL + (2 * lam) * 16, LcA);
        pd_j16_load_abs(LL + (2 * lam + 1) * 16, LcB);
        double PA[16], PB[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) PA[r] = PB[r] = 0.0;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            double xr[16];
            pd_j16_load_rel(XS + r * 16, lam, xr);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                PA[i] = fma(xr[i], LcA[r], PA[i]);
                PB[i] = fma(xr[i], LcB[r], PB[i]);
            }
        }
        if (beam) {
            double rv[16];
            pd_j16_load_abs(V1, rv);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                ta = fma(LcA[i], rv[i], ta);
                tb = fma(LcB[i], rv[i], tb);
            }
        }
        __syncwarp();  // everybody has read X' and r
#pragma unroll
        for (int r = 0; r < 16; ++r) pd_st2(XS + ((2 * lam + r) & 15) * 16 + 2 * lam, PA[r], PB[r]);
        if (beam) pd_st2(V0 + 2 * lam, ta, tb);
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 16; ++r) TA[r] = TB[r] = 0.0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            double pi[16];
            pd_j16_load_rel(XS + i * 16, lam, pi);
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                TA[r] = fma(pi[r], LcA[i], TA[r]);
                TB[r] = fma(pi[r], LcB[i], TB[r]);
            }
        }
    }
    ok = __all_sync(0xffu << grp, ok);  // the 8 lanes of this item
    if (!ok) {  // keep the warp in step on a harmless matrix; the item is handed to the general solver
#pragma unroll
        for (int r = 0; r < 16; ++r) TA[r] = TB[r] = 0.0;
        TA[0] = 1.0;
        TB[1] = 2.0;
    }

    // ---- Jacobi: T -> diag(k^2), W accumulates the rotations ----
    double Wa[16], Wb[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        Wa[r] = (r == 2 * lam) ? 1.0 : 0.0;
        Wb[r] = (r == 2 * lam + 1) ? 1.0 : 0.0;
    }
    bool converged = false;
    for (int sweep = 0; sweep < PD_J16_MAX_SWEEPS; ++sweep) {
        double off = fabs(TA[1]) + fabs(TB[0]), dg = fabs(TA[0]) + fabs(TB[1]);
#pragma unroll
        for (int r = 2; r < 16; ++r) off += fabs(TA[r]) + fabs(TB[r]);
        off = pd_j16_sum8(off);  // every off-diagonal entry is counted twice
        dg = pd_j16_sum8(dg);
        converged = (off <= 2e-17 * dg);
        if (__all_sync(0xffffffffu, converged)) break;
#pragma unroll 1
        for (int st = 0; st < 16; ++st) {  // even steps: pairs (2i, 2i+1); odd steps: (2i+1, 2i+2), one lane idle
            pd_j16_step(TA, TB, Wa, Wb, lam, grp, converged, (st & 1) && lam == ((7 - (st >> 1)) & 7));
            pd_j16_shift(TA, TB, Wa, Wb, lam, grp);
        }
    }
    const double lama = TA[0], lamb = TB[1];
    ok = ok && converged && lama > 0.0 && lamb > 0.0;
    ok = __all_sync(0xffu << grp, ok);
    const double kia = pd_rsqrt(lama > 0.0 ? lama : 1.0), kib = pd_rsqrt(lamb > 0.0 ? lamb : 1.0);
    const double ka = lama * kia, kb = lamb * kib;

    // ---- beam coefficients  c = diag(1 / (1/mu0^2 - k^2)) W^T (L^T r)  (before W is overwritten) ----
    double ca = 0.0, cb = 0.0;
    const double mu0b = a.colp[(long)b * PD_NCOLP + PD_COL_MU0];
    if (beam) {
        double tv[16];
        pd_j16_load_abs(V0, tv);
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            s0 = fma(Wa[r], tv[r], s0);
            s1 = fma(Wb[r], tv[r], s1);
        }
        const double m2 = 1.0 / (mu0b * mu0b);
        ca = s0 / (m2 - ka * ka);
        cb = s1 / (m2 - kb * kb);
    }

    // ---- V^ = L^-T W (in place, back substitution) and U^ = -L W / k, one pass over the columns of L ----
    double Ua[16], Ub[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) Ua[r] = Ub[r] = 0.0;
#pragma unroll
    for (int i = 15; i >= 0; --i) {
        double Lc[16];
#pragma unroll
        for (int r = (i & ~1); r < 16; r += 2) pd_ld2(LL + i * 16 + r, Lc[r], Lc[r + 1]);
        const double wa = Wa[i], wb = Wb[i];
        double sa = wa, sb = wb;
#pragma unroll
        for (int r = i + 1; r < 16; ++r) {
            sa = fma(-Lc[r], Wa[r], sa);
            sb = fma(-Lc[r], Wb[r], sb);
        }
        const double li = Linv[i];
        Wa[i] = sa * li;
        Wb[i] = sb * li;
#pragma unroll
        for (int r = i; r < 16; ++r) {
            Ua[r] = fma(Lc[r], wa, Ua[r]);
            Ub[r] = fma(Lc[r], wb, Ub[r]);
        }
    }

    // ---- K and the G blocks: Gp = (V^ + U^) / (2 D), Gm = (V^ - U^) / (2 D); shortcut layers (:162-168) ----
    double* Kout = a.K + item * 16;
    double* Gout = a.G + pd_g_base(item, 16);
    if (store) {
        if (active) {
            pd_st2(Kout + 2 * lam, ka, kb);
        } else {
            double m0, m1;
            pd_ld2(rmu + 2 * lam, m0, m1);
            pd_st2(Kout + 2 * lam, m0, m1);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const double di = dinvs[i];
            const double ua = -Ua[i] * kia, ub = -Ub[i] * kib;
            double gpa = (Wa[i] + ua) * di, gpb = (Wb[i] + ub) * di, gma = (Wa[i] - ua) * di, gmb = (Wb[i] - ub) * di;
            if (!active) {
                gpa = gpb = 0.0;
                gma = (i == 2 * lam) ? 1.0 : 0.0;
                gmb = (i == 2 * lam + 1) ? 1.0 : 0.0;
            }
            pd_st2(Gout + pd_g_off(i * 16 + 2 * lam, 16), gpa, gpb);
            pd_st2(Gout + pd_g_off(256 + i * 16 + 2 * lam, 16), gma, gmb);
        }
    }

    if (a.beam) {
        double bta = 0.0, btb = 0.0, bba = 0.0, bbb = 0.0;
        // p^ = V^ c, summed over the lanes; every lane ends up with the whole vector
        double ph[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) ph[r] = pd_j16_sum8(fma(Wa[r], ca, Wb[r] * cb));
        // L^T p^ for the lane's two columns
        double LcA[16], LcB[16];
        pd_j16_load_abs(LL + (2 * lam) * 16, LcA);
        pd_j16_load_abs(LL + (2 * lam + 1) * 16, LcB);
        double t2a = 0.0, t2b = 0.0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            t2a = fma(LcA[i], ph[i], t2a);
            t2b = fma(LcB[i], ph[i], t2b);
        }
        pd_st2(V1 + 2 * lam, t2a, t2b);
        __syncwarp();
        double t2[16];
        pd_j16_load_abs(V1, t2);
        double spa = 0.0, spb = 0.0;  // rows 2 lam, 2 lam + 1 of L (L^T p^)
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            double l0, l1;
            pd_ld2(LL + c * 16 + 2 * lam, l0, l1);
            spa = fma(l0, t2[c], spa);
            spb = fma(l1, t2[c], spb);
        }
        double pa = 0.0, pb = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (lam == k) {
                pa = ph[2 * k];
                pb = ph[2 * k + 1];
            }
        double d0, d1;
        pd_ld2(dinvs + 2 * lam, d0, d1);
        const double qa = mu0b * (da - spa), qb = mu0b * (db - spb);
        if (beam && active) {
            bta = (pa + qa) * d0;  // dinvs carries the factor 1/2
            btb = (pb + qb) * d1;
            bba = (pa - qa) * d0;
            bbb = (pb - qb) * d1;
        }
        if (store) {
            double* Bout = a.Bv + item * 32;
            pd_st2(Bout + 2 * lam, bta, btb);
            pd_st2(Bout + 16 + 2 * lam, bba, bbb);
        }
    }
    if (thermal) {  // mode 0 of a problem with an isotropic source (uniform over the CTA: the shuffles below are safe)
        // y = -k V^T (D / mu) (the lane's two eigen indices), then for every power q of the source polynomial
        //   t-+ = b_q(-+k) y,   d_q = G [t-; -t+]:  top = V^ (t- - t+) + U^ (t- + t+),  bottom = V^ (t- - t+) - U^ (t- + t+)
        // (subroutines.py:783-862, as in pd_stage_a_sym.cuh); U^ column j = -Ua_j / k_j
        const double* fsq = tab + 32;  // sqrt(w / mu)
        double ya = 0.0, yb = 0.0;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const double f = fsq[r];
            ya = fma(Wa[r], f, ya);
            yb = fma(Wb[r], f, yb);
        }
        ya *= -ka;
        yb *= -kb;
        const double* sc = a.s_s + ((long)b * a.L + l) * a.Ns;
        double* dout = a.dth + ((long)b * a.L + l) * a.Ns * 32;
        double m0, m1, d0, d1;
        pd_ld2(rmu + 2 * lam, m0, m1);
        pd_ld2(dinvs + 2 * lam, d0, d1);
        for (int q = 0; q < a.Ns; ++q) {
            // polynomial sums b_q(+k), b_q(-k) for the two eigen indices (active) and for k = 1 / mu_i (shortcut)
            double bpa = 0.0, bna = 0.0, bpb = 0.0, bnb = 0.0, spa = 0.0, sna = 0.0, spb = 0.0, snb = 0.0;
            double ratio = 1.0, pwa = kia, pwb = kib, pma = 1.0 / m0, pmb = 1.0 / m1;
            for (int r = q; r < a.Ns; ++r) {
                if (r > q) {
                    ratio *= (double)r;
                    pwa *= kia;
                    pwb *= kib;
                    pma *= 1.0 / m0;
                    pmb *= 1.0 / m1;
                }
                const double c = sc[r] * ratio;
                const bool odd = ((r - q) & 1) != 0;
                bpa = fma(c, pwa, bpa);
                bna = fma(c, odd ? pwa : -pwa, bna);
                bpb = fma(c, pwb, bpb);
                bnb = fma(c, odd ? pwb : -pwb, bnb);
                spa = fma(c, pma, spa);
                sna = fma(c, odd ? pma : -pma, sna);
                spb = fma(c, pmb, spb);
                snb = fma(c, odd ? pmb : -pmb, snb);
            }
            const double dma = (bna - bpa) * ya, sma = (bna + bpa) * ya, dmb = (bnb - bpb) * yb, smb = (bnb + bpb) * yb;
            const double ua_s = -kia * sma, ub_s = -kib * smb;
            double ta = 0.0, tb = 0.0, ba = 0.0, bb = 0.0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const double vv = pd_j16_sum8(fma(Wa[i], dma, Wb[i] * dmb));
                const double uu = pd_j16_sum8(fma(Ua[i], ua_s, Ub[i] * ub_s));
                if (i == 2 * lam) {
                    ta = vv + uu;
                    ba = vv - uu;
                }
                if (i == 2 * lam + 1) {
                    tb = vv + uu;
                    bb = vv - uu;
                }
            }
            ta *= d0;  // dinvs carries the factor 1/2
            ba *= d0;
            tb *= d1;
            bb *= d1;
            if (!active) {  // shortcut layer: y = -1/mu, k = 1/mu:  top = b_q(+k) / mu, bottom = -b_q(-k) / mu
                ta = spa * m0;
                tb = spb * m1;
                ba = -sna * m0;
                bb = -snb * m1;
            }
            if (store) {
                pd_st2(dout + q * 32 + 2 * lam, ta, tb);
                pd_st2(dout + q * 32 + 16 + 2 * lam, ba, bb);
            }
        }
    }
    __syncwarp();
    return !active || ok;  // the shortcut needs no decomposition
}

#endif  // __CUDACC__
