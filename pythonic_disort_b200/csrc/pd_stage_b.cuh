// pd_stage_b.cuh -- boundary-condition solve for one (column, Fourier mode):
// the block-banded system of _solve_for_coeffs.py:139-383 (Stamnes-Conklin
// scaled, so every exponential is <= 1), solved layer by layer.
//
// Unknowns are grouped per layer, C_l = [C_l^- (N); C_l^+ (N)]; equations are
// the top boundary (N rows), the 2N continuity rows of each interface and the
// bottom boundary (N rows).  At stage l the working panel holds the N rows
// carried over from above (they only involve C_l) and the 2N rows of
// interface l (C_l and C_{l+1}); partial-pivoting elimination of the 2N
// columns of C_l over those 3N rows picks exactly the pivots LAPACK's banded
// dgbsv would (all other rows are zero in these columns), leaves 2N finished
// rows of U (saved to the history buffer) and N rows in C_{l+1} only, which
// become the next stage's carry.  The right-hand side rides along as one
// extra panel column; back-substitution walks the history bottom-up.
#pragma once
#include "pd_common.cuh"

struct PdStageB {
    int B, L, N, NF, Ns, NBDRF, NFb;
    int beam, iso, bdrf_percol;
    const double* taus;    // [B][L+1]
    const double* colp;    // [B][PD_NCOLP]
    const double* bpos;    // [B][NFb][N]
    const double* bneg;    // [B][NFb][N]
    const double* mu;      // [N]
    const double* w;       // [N]
    const double* bdrf_q;  // [(B)][NBDRF][N][N]
    const double* bdrf_q0; // [(B)][NBDRF][N]
    const double* K;       // [B][NF][L][N]
    const double* G;       // [B][NF][L][2][N][N] (N = 8: sector-interleaved over 32 items, pd_common.cuh)
    const double* RT;      // [B][NF][L][2][N(N+1)/2] layer operators R^, T^, packed symmetric (k_layer_ops, N = 8), or
                           // null: the sweep forms them from G
    const double* Bv;      // [B][NF][L][2N]
    const double* dth;     // [B][L][Ns][2N]
    double* C;             // [B][NF][L][2N]
    double* Uif;           // [B][L+1][NF][2N] stream radiances of every mode at the layer interfaces, or null
    int32_t* status;       // [B]
};

// where the 2N interface radiances of (column b, interface lev, mode m) live in Uif
PD_HD long pd_uif_index(int b, int lev, int m, int L, int NF, int n2) { return (((long)b * (L + 1) + lev) * NF + m) * n2; }

// shared-memory doubles per concurrently solved system, and history doubles per system
PD_HD int pd_stage_b_doubles(int N) { return 3 * N * (4 * N + 1) + 8 * N + N * pd_ld(N); }
PD_HD long pd_stage_b_history_doubles(int N, int L) { return (long)L * 2 * N * (4 * N + 1); }

// thermal particular solution of layer l at scaled optical depth t: v[2N]
PD_HD double pd_thermal_at(const double* dth_l, int Ns, int n2, int i, double t) {
    double v = 0.0;
    for (int q = Ns - 1; q >= 0; --q) v = fma(v, t, dth_l[q * n2 + i]);
    return v;
}

template <class Grp, int NC = 0>
PD_HD void pd_stage_b_system(const Grp& g, const PdStageB& a, int b, int m, double* sm, double* hist) {
    const int lane = g.lane();
    const int n = NC > 0 ? NC : a.N, n2 = 2 * n, L = a.L;
    const int ldp = 4 * n + 1, rc = 4 * n;  // rhs column
    double* P = sm;                   // [3n][ldp]
    double* E = P + 3 * n * ldp;      // [2n]  exp(-k dtau*) of layer l (first n) and l+1 (last n)
    double* xs = E + n2;              // [2n]  solution of the layer below during back-substitution
    double* rv = xs + n2;             // [2n]
    double* vt = rv + n2;             // [2n]  scratch
    double* R = vt + n2;              // [n][ld] surface reflection matrix
    const int ldr = pd_ld(n);

    const long sys = (long)b * a.NF + m;
    const double* taus = a.taus + (long)b * (L + 1);
    const double* Kc = a.K + sys * L * n;
    const long item0 = sys * L;  // G item of layer l: item0 + l (pd_g_base / pd_g_off, pd_common.cuh)
    const double* Bc = a.beam ? a.Bv + sys * L * n2 : nullptr;  // only read when `beam`
    const double* dthc = (a.iso && m == 0) ? a.dth + (long)b * L * a.Ns * n2 : nullptr;
    const double mu0 = a.colp[(long)b * PD_NCOLP + PD_COL_MU0];
    const double I0 = a.colp[(long)b * PD_NCOLP + PD_COL_I0];
    const bool beam = a.beam && I0 > 0.0;  // per column (pydisort.py:215)
    const bool has_bdrf = a.NBDRF > m;
    const bool have_b = (m == 0) || (a.NFb > 1);
    const double* bpos = a.bpos + ((long)b * a.NFb + (a.NFb > 1 ? m : 0)) * n;
    const double* bneg = a.bneg + ((long)b * a.NFb + (a.NFb > 1 ? m : 0)) * n;
    int status = 0;

    // G_l[r][c] for r, c in [0, 2n): block structure [[Gp, Gm], [Gm, Gp]]
    auto Gat = [&](int l, int r, int c) -> double {
        const int rb = r >= n, cb = c >= n;
        return a.G[pd_g_base(item0 + l, n) + pd_g_off((rb ^ cb) * n * n + (r - rb * n) * n + (c - cb * n), n)];
    };

    if (has_bdrf) {  // R = (1 + delta_m0) q^m(mu_i, mu_j) mu_j w_j   (:121-134)
        const double* q = a.bdrf_q + ((a.bdrf_percol ? (long)b * a.NBDRF : 0) + m) * n * n;
        for (int idx = lane; idx < n * n; idx += Grp::size) {
            const int i = idx / n, j = idx - i * n;
            R[i * ldr + j] = ((m == 0) ? 2.0 : 1.0) * q[idx] * a.mu[j] * a.w[j];
        }
    }
    for (int i = lane; i < n; i += Grp::size) E[n + i] = exp(-Kc[i] * (taus[1] - taus[0]));
    g.sync();

    // top boundary rows -> carry   (:161,:170-179,:238,:284-285)
    for (int idx = lane; idx < n * (n2 + 1); idx += Grp::size) {
        const int r = idx / (n2 + 1), c = idx - r * (n2 + 1);
        if (c < n2) {
            P[r * ldp + c] = Gat(0, n + r, c) * ((c >= n) ? E[c] : 1.0);
        } else {
            double v = have_b ? bneg[r] : 0.0;
            if (beam) v -= Bc[n + r];
            if (dthc) v -= pd_thermal_at(dthc, a.Ns, n2, n + r, taus[0]);
            P[r * ldp + rc] = v;
        }
    }
    g.sync();

    for (int l = 0; l < L; ++l) {
        const bool last = (l == L - 1);
        const int nrows = last ? n2 : 3 * n;
        // exponentials: E[0..n) <- layer l (already in E[n..2n)), E[n..2n) <- layer l+1
        for (int i = lane; i < n; i += Grp::size) {
            const double el = E[n + i];
            vt[i] = el;
            if (!last) vt[n + i] = exp(-Kc[(l + 1) * n + i] * (taus[l + 2] - taus[l + 1]));
        }
        g.sync();
        for (int i = lane; i < n2; i += Grp::size) E[i] = vt[i];
        // zero the C_{l+1} part of the carry rows
        for (int idx = lane; idx < n * n2; idx += Grp::size) P[(idx / n2) * ldp + n2 + (idx % n2)] = 0.0;
        g.sync();
        if (!last) {  // continuity rows of interface l   (:242-245, :184-205, :317-323)
            const double att = beam ? exp(-taus[l + 1] / mu0) : 0.0;
            for (int idx = lane; idx < n2 * (rc + 1); idx += Grp::size) {
                const int r = idx / (rc + 1), c = idx - r * (rc + 1);
                double v;
                if (c < n) v = Gat(l, r, c) * E[c];
                else if (c < n2) v = Gat(l, r, c);
                else if (c < 3 * n) v = -Gat(l + 1, r, c - n2);
                else if (c < rc) v = -Gat(l + 1, r, c - n2) * E[c - n2];
                else {
                    v = 0.0;
                    if (beam) v = (Bc[(l + 1) * n2 + r] - Bc[l * n2 + r]) * att;
                    if (dthc)
                        v += pd_thermal_at(dthc + (long)(l + 1) * a.Ns * n2, a.Ns, n2, r, taus[l + 1]) -
                             pd_thermal_at(dthc + (long)l * a.Ns * n2, a.Ns, n2, r, taus[l + 1]);
                }
                P[(n + r) * ldp + c] = v;
            }
        } else {  // bottom boundary rows   (:163,:208-232,:248-254,:289-293)
            const double att = beam ? exp(-taus[L] / mu0) : 0.0;
            if (dthc)
                for (int i = lane; i < n2; i += Grp::size)
                    vt[i] = pd_thermal_at(dthc + (long)l * a.Ns * n2, a.Ns, n2, i, taus[L]);
            g.sync();
            for (int idx = lane; idx < n * (n2 + 1); idx += Grp::size) {
                const int r = idx / (n2 + 1), c = idx - r * (n2 + 1);
                double v;
                if (c < n2) {
                    v = Gat(l, r, c);
                    if (has_bdrf)
                        for (int j = 0; j < n; ++j) v = fma(-R[r * ldr + j], Gat(l, n + j, c), v);
                    if (c < n) v *= E[c];
                    P[(n + r) * ldp + c] = v;
                } else {
                    v = have_b ? bpos[r] : 0.0;
                    if (dthc) {
                        v -= vt[r];
                        if (has_bdrf)
                            for (int j = 0; j < n; ++j) v = fma(R[r * ldr + j], vt[n + j], v);
                    }
                    if (beam) {
                        double s = -Bc[l * n2 + r];
                        if (has_bdrf) {
                            const double* q0 = a.bdrf_q0 + ((a.bdrf_percol ? (long)b * a.NBDRF : 0) + m) * n;
                            s += (mu0 * I0 / PD_PI) * q0[r];
                            for (int j = 0; j < n; ++j) s = fma(R[r * ldr + j], Bc[l * n2 + n + j], s);
                        }
                        v = fma(s, att, v);
                    }
                    P[(n + r) * ldp + rc] = v;
                }
            }
        }
        g.sync();

        // eliminate the 2n columns of C_l with partial pivoting over the panel rows
        for (int j = 0; j < n2; ++j) {
            int p = j;
            double best = fabs(P[j * ldp + j]);
            for (int r = j + 1; r < nrows; ++r) {
                const double v = fabs(P[r * ldp + j]);
                if (v > best) {
                    best = v;
                    p = r;
                }
            }
            if (best == 0.0) status |= PD_ST_ZERO_PIVOT;
            g.sync();
            if (p != j) {
                for (int c = j + lane; c <= rc; c += Grp::size) {
                    const double t = P[j * ldp + c];
                    P[j * ldp + c] = P[p * ldp + c];
                    P[p * ldp + c] = t;
                }
                g.sync();
            }
            const double pinv = 1.0 / P[j * ldp + j];
            for (int c = j + 1 + lane; c <= rc; c += Grp::size) {  // lanes own columns
                const double u = P[j * ldp + c];
                if (u != 0.0)
                    for (int r = j + 1; r < nrows; ++r) P[r * ldp + c] = fma(-(P[r * ldp + j] * pinv), u, P[r * ldp + c]);
            }
            g.sync();
            if (lane == 0) P[j * ldp + j] = pinv;  // keep the reciprocal pivot for back-substitution
        }
        g.sync();
        if (!last) {
            // save the 2n finished rows, move the n remaining rows up as the new carry
            double* h = hist + (long)l * n2 * ldp;
            for (int idx = lane; idx < n2 * ldp; idx += Grp::size) h[idx] = P[idx];
            g.sync();
            for (int idx = lane; idx < n * (n2 + 1); idx += Grp::size) {
                const int r = idx / (n2 + 1), c = idx - r * (n2 + 1);
                const double v = P[(n2 + r) * ldp + ((c < n2) ? n2 + c : rc)];
                // rows r < n are being overwritten while rows 2n.. are read: disjoint since n2 + r >= n
                P[r * ldp + ((c < n2) ? c : rc)] = v;
            }
            g.sync();
        }
    }

    // back-substitution: last layer from the panel, the others from the history
    double* Cout = a.C + sys * L * n2;
    for (int l = L - 1; l >= 0; --l) {
        if (l < L - 1) {
            const double* h = hist + (long)l * n2 * ldp;
            for (int idx = lane; idx < n2 * ldp; idx += Grp::size) P[idx] = h[idx];
            g.sync();
            for (int r = lane; r < n2; r += Grp::size) {
                double s = P[r * ldp + rc];
                for (int c = 0; c < n2; ++c) s = fma(-P[r * ldp + n2 + c], xs[c], s);
                rv[r] = s;
            }
        } else {
            for (int r = lane; r < n2; r += Grp::size) rv[r] = P[r * ldp + rc];
        }
        g.sync();
        for (int j = n2 - 1; j >= 0; --j) {
            const double xj = rv[j] * P[j * ldp + j];
            g.sync();
            if (lane == 0) xs[j] = xj;
            for (int r = lane; r < j; r += Grp::size) rv[r] = fma(-P[r * ldp + j], xj, rv[r]);
            g.sync();
        }
        for (int i = lane; i < n2; i += Grp::size) Cout[l * n2 + i] = xs[i];
        g.sync();
    }
    // interface radiances from the coefficients: u = G_l (C_l * e) + particular at the bottom of layer lev - 1 (the top
    // of layer 0 for lev = 0) -- term by term what pd_mode_at (pd_eval.cuh) evaluates at such a point
    for (int lev = 0; a.Uif && lev <= L; ++lev) {
        const int l = lev > 0 ? lev - 1 : 0;
        const double dtau = taus[l + 1] - taus[l];
        for (int i = lane; i < n; i += Grp::size) {
            const double e = exp(-Kc[l * n + i] * dtau);
            vt[i] = Cout[l * n2 + i] * (lev > 0 ? e : 1.0);
            vt[n + i] = Cout[l * n2 + n + i] * (lev > 0 ? 1.0 : e);
        }
        g.sync();
        const double att = beam ? exp(-taus[lev] / mu0) : 0.0;
        double* uo = a.Uif + pd_uif_index(b, lev, m, L, a.NF, n2);
        for (int r = lane; r < n2; r += Grp::size) {
            double s = 0.0;
            for (int c = 0; c < n2; ++c) s = fma(Gat(l, r, c), vt[c], s);
            if (beam) s = fma(Bc[l * n2 + r], att, s);
            if (dthc) s += pd_thermal_at(dthc + (long)l * a.Ns * n2, a.Ns, n2, r, taus[lev]);
            uo[r] = s;
        }
        g.sync();
    }
    if (status && lane == 0) {
#if defined(__CUDA_ARCH__)
        atomicOr(a.status + b, status);
#else
        a.status[b] |= status;
#endif
    }
}
