// pd_linalg.cuh -- small dense FP64 linear algebra on shared-memory matrices,
// one lane group per matrix.
//
//   pd_eig_real : eigenvalues/eigenvectors of a real nonsymmetric matrix whose
//                 spectrum is known to be real (the reduced DISORT eigenproblem,
//                 _solve_for_gen_and_part_sols.py:179-183 uses LAPACK dgeev).
//                 ASYMTX-style: Householder reduction to Hessenberg form,
//                 explicit single-shift QR with deflation accumulating the
//                 Schur vectors, back-substitution on the triangular factor.
//   pd_lu_solve : A x = b by LU with partial pivoting (dgesv), one right-hand side.
#pragma once
#include "pd_common.cuh"

#define PD_QR_MAX_ITS 40

// ---------------------------------------------------------------------------
// H (n x n, leading dimension ld) is destroyed; on exit it holds the upper
// triangular Schur factor.  Z (n x n) receives the eigenvectors as columns,
// each scaled to unit 2-norm; wr[n] the eigenvalues; Y is n x n scratch,
// cs 2n scratch, vec n scratch.  Returns PD_ST_* bits.
// ---------------------------------------------------------------------------
template <class Grp, int NC = 0>
PD_HD int pd_eig_real(const Grp& g, int n_rt, int ld_rt, double* H, double* Z, double* Y, double* wr, double* cs,
                      double* vec) {
    const int n = NC > 0 ? NC : n_rt, ld = NC > 0 ? (NC | 1) : ld_rt;  // NC > 0: size known at compile time
    const int lane = g.lane();
    int status = 0;

    for (int idx = lane; idx < n * n; idx += Grp::size) {
        const int i = idx / n, j = idx - i * n;
        Z[i * ld + j] = (i == j) ? 1.0 : 0.0;
    }
    g.sync();

    // ---- Householder reduction to upper Hessenberg form, Z accumulates Q ----
    for (int k = 0; k + 2 < n; ++k) {
        const double x0 = H[(k + 1) * ld + k];
        double ss = 0.0;
        for (int i = k + 2; i < n; ++i) {
            const double t = H[i * ld + k];
            ss = fma(t, t, ss);
        }
        if (ss == 0.0) continue;  // nothing to annihilate (group-uniform)
        const double nrm = sqrt(fma(x0, x0, ss));
        const double a = (x0 >= 0.0) ? -nrm : nrm;
        const double v0 = x0 - a;
        const double beta = -1.0 / (a * v0);
        for (int i = k + 1 + lane; i < n; i += Grp::size) vec[i] = (i == k + 1) ? v0 : H[i * ld + k];
        g.sync();
        // left:  H <- (I - beta v v^T) H      (lanes own columns)
        for (int j = k + 1 + lane; j < n; j += Grp::size) {
            double s = 0.0;
            for (int i = k + 1; i < n; ++i) s = fma(vec[i], H[i * ld + j], s);
            s *= beta;
            for (int i = k + 1; i < n; ++i) H[i * ld + j] = fma(-s, vec[i], H[i * ld + j]);
        }
        if (lane == 0) {
            H[(k + 1) * ld + k] = a;
            for (int i = k + 2; i < n; ++i) H[i * ld + k] = 0.0;
        }
        g.sync();
        // right: H <- H (I - beta v v^T),  Z <- Z (I - beta v v^T)   (lanes own rows)
        for (int i = lane; i < n; i += Grp::size) {
            double s = 0.0, sz = 0.0;
            for (int j = k + 1; j < n; ++j) {
                s = fma(H[i * ld + j], vec[j], s);
                sz = fma(Z[i * ld + j], vec[j], sz);
            }
            s *= beta;
            sz *= beta;
            for (int j = k + 1; j < n; ++j) {
                H[i * ld + j] = fma(-s, vec[j], H[i * ld + j]);
                Z[i * ld + j] = fma(-sz, vec[j], Z[i * ld + j]);
            }
        }
        g.sync();
    }

    // ---- norm of the Hessenberg matrix (deflation fallback / perturbation scale) ----
    double norm = 0.0;
    for (int i = 0; i < n; ++i)
        for (int j = (i > 0 ? i - 1 : 0); j < n; ++j) norm += fabs(H[i * ld + j]);
    if (norm == 0.0) norm = 1.0;

    // ---- shifted QR iteration, real spectrum ----
    int hi = n - 1, its = 0;
    while (hi >= 0) {
        int l = hi;
        while (l > 0) {
            double s = fabs(H[(l - 1) * ld + (l - 1)]) + fabs(H[l * ld + l]);
            if (s == 0.0) s = norm;
            if (fabs(H[l * ld + (l - 1)]) <= PD_EPS * s) break;
            --l;
        }
        if (l == hi) {  // one root
            g.sync();
            if (lane == 0) {
                wr[hi] = H[hi * ld + hi];
                if (hi > 0) H[hi * ld + (hi - 1)] = 0.0;
            }
            --hi;
            its = 0;
            g.sync();
            continue;
        }
        const double a = H[(hi - 1) * ld + (hi - 1)], b = H[(hi - 1) * ld + hi];
        const double c = H[hi * ld + (hi - 1)], d = H[hi * ld + hi];
        const double p = 0.5 * (a - d);
        const double q = fma(p, p, b * c);
        if (l == hi - 1) {
            // 2x2 block: rotate it to upper triangular form (its roots are real;
            // a slightly negative discriminant can only be rounding noise)
            const double z = sqrt(fabs(q));
            const double zz = p + ((p >= 0.0) ? z : -z);
            double rp, rq;
            {
                const double s = fabs(c) + fabs(zz);
                rp = c / s;
                rq = zz / s;
                const double r = sqrt(fma(rp, rp, rq * rq));
                rp /= r;
                rq /= r;
            }
            g.sync();
            for (int j = hi - 1 + lane; j < n; j += Grp::size) {  // rows hi-1, hi
                const double t1 = H[(hi - 1) * ld + j], t2 = H[hi * ld + j];
                H[(hi - 1) * ld + j] = fma(rq, t1, rp * t2);
                H[hi * ld + j] = fma(rq, t2, -rp * t1);
            }
            g.sync();
            for (int i = lane; i < n; i += Grp::size) {  // columns hi-1, hi
                if (i <= hi) {
                    const double t1 = H[i * ld + (hi - 1)], t2 = H[i * ld + hi];
                    H[i * ld + (hi - 1)] = fma(rq, t1, rp * t2);
                    H[i * ld + hi] = fma(rq, t2, -rp * t1);
                }
                const double z1 = Z[i * ld + (hi - 1)], z2 = Z[i * ld + hi];
                Z[i * ld + (hi - 1)] = fma(rq, z1, rp * z2);
                Z[i * ld + hi] = fma(rq, z2, -rp * z1);
            }
            g.sync();
            if (lane == 0) {
                H[hi * ld + (hi - 1)] = 0.0;
                wr[hi - 1] = H[(hi - 1) * ld + (hi - 1)];
                wr[hi] = H[hi * ld + hi];
            }
            hi -= 2;
            its = 0;
            g.sync();
            continue;
        }
        if (its >= PD_QR_MAX_ITS) {  // give up on this block: report and take the diagonal
            status |= PD_ST_QR_NOCONV;
            g.sync();
            if (lane == 0) {
                for (int i = l; i <= hi; ++i) {
                    wr[i] = H[i * ld + i];
                    if (i > 0) H[i * ld + (i - 1)] = 0.0;
                }
            }
            hi = l - 1;
            its = 0;
            g.sync();
            continue;
        }
        // Wilkinson shift: root of the trailing 2x2 nearer to its last diagonal entry
        double sigma = d;
        if (q >= 0.0) {
            const double zz = p + ((p >= 0.0) ? sqrt(q) : -sqrt(q));
            if (zz != 0.0) sigma = d - b * c / zz;
        }
        if (its == 10 || its == 20 || its == 30)
            sigma = d + 0.75 * (fabs(c) + fabs(H[(hi - 1) * ld + (hi - 2)]));  // exceptional shift
        g.sync();
        for (int i = l + lane; i <= hi; i += Grp::size) H[i * ld + i] -= sigma;
        g.sync();
        // left rotations (lanes own columns): R = Q^T (H - sigma).  Column k itself (r_k on the diagonal, zero
        // below) is written after the loop, so one sync per rotation is enough.
        for (int k = l; k < hi; ++k) {
            const double x = H[k * ld + k], y = H[(k + 1) * ld + k];
            double cr = 1.0, sr = 0.0, rr = x;
            if (y != 0.0) {
                const double q2 = fma(x, x, y * y);
                const double ri = pd_rsqrt(q2);
                cr = x * ri;
                sr = y * ri;
                rr = q2 * ri;
            }
            if (lane == 0) {
                pd_d2 v;
                v.x = cr;
                v.y = sr;
                *reinterpret_cast<pd_d2*>(cs + 2 * k) = v;
                vec[k] = rr;
            }
            for (int j = k + 1 + lane; j < n; j += Grp::size) {
                const double t1 = H[k * ld + j], t2 = H[(k + 1) * ld + j];
                H[k * ld + j] = fma(cr, t1, sr * t2);
                H[(k + 1) * ld + j] = fma(cr, t2, -sr * t1);
            }
            g.sync();
        }
        for (int k = l + lane; k < hi; k += Grp::size) {
            H[k * ld + k] = vec[k];
            H[(k + 1) * ld + k] = 0.0;
        }
        g.sync();
        // right rotations + shift restore (lanes own rows): H <- R Q + sigma, Z <- Z Q.  One pass over k for
        // both matrices; the element shared by consecutive rotations stays in a register.
        for (int i = lane; i < n; i += Grp::size) {
            const bool hrow = (i <= hi);
            const int kh = (i - 1 > l) ? i - 1 : l;  // first rotation that touches row i of H
            double zc = Z[i * ld + l];
            double hc = hrow ? H[i * ld + kh] : 0.0;
            for (int k = l; k < hi; ++k) {
                const pd_d2 rot = *reinterpret_cast<const pd_d2*>(cs + 2 * k);
                const double zn = Z[i * ld + (k + 1)];
                Z[i * ld + k] = fma(rot.x, zc, rot.y * zn);
                zc = fma(rot.x, zn, -rot.y * zc);
                if (hrow && k >= kh) {
                    const double hn = H[i * ld + (k + 1)];
                    const double v = fma(rot.x, hc, rot.y * hn);
                    H[i * ld + k] = (k == i) ? v + sigma : v;
                    hc = fma(rot.x, hn, -rot.y * hc);
                }
            }
            Z[i * ld + hi] = zc;
            if (hrow) H[i * ld + hi] = (i == hi) ? hc + sigma : hc;
        }
        ++its;
        g.sync();
    }

    // ---- eigenvectors of the triangular factor by back-substitution (lane owns column j) ----
    const double tiny = PD_EPS * norm;
    for (int j = lane; j < n; j += Grp::size) {
        const double lam = wr[j];
        Y[j * ld + j] = 1.0;
        for (int i = j - 1; i >= 0; --i) {
            double s = 0.0;
            for (int t = i + 1; t <= j; ++t) s = fma(H[i * ld + t], Y[t * ld + j], s);
            double den = H[i * ld + i] - lam;
            if (fabs(den) < tiny) den = (den < 0.0) ? -tiny : tiny;
            Y[i * ld + j] = -s / den;
        }
    }
    g.sync();
    // ---- back-transform: V = Z Y (in place in Z, lane owns row i, j descending) ----
    for (int i = lane; i < n; i += Grp::size) {
        for (int j = n - 1; j >= 0; --j) {
            double s = 0.0;
            for (int t = 0; t <= j; ++t) s = fma(Z[i * ld + t], Y[t * ld + j], s);
            Z[i * ld + j] = s;
        }
    }
    g.sync();
    for (int j = lane; j < n; j += Grp::size) {
        double ss = 0.0;
        for (int i = 0; i < n; ++i) ss = fma(Z[i * ld + j], Z[i * ld + j], ss);
        const double sc = pd_rsqrt(ss);
        for (int i = 0; i < n; ++i) Z[i * ld + j] *= sc;
    }
    g.sync();
    return status;
}

// ---------------------------------------------------------------------------
// Solve A x = b (n x n, leading dimension ld) with partial pivoting; A and b
// are overwritten, x is returned in b.  Returns PD_ST_ZERO_PIVOT on breakdown.
// ---------------------------------------------------------------------------
template <class Grp, int NC = 0>
PD_HD int pd_lu_solve(const Grp& g, int n_rt, int ld_rt, double* A, double* b) {
    const int n = NC > 0 ? NC : n_rt, ld = NC > 0 ? (NC | 1) : ld_rt;
    const int lane = g.lane();
    int status = 0;
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = fabs(A[k * ld + k]);
        for (int i = k + 1; i < n; ++i) {
            const double v = fabs(A[i * ld + k]);
            if (v > best) {
                best = v;
                p = i;
            }
        }
        if (best == 0.0) status |= PD_ST_ZERO_PIVOT;
        g.sync();
        if (p != k) {
            for (int j = k + lane; j < n; j += Grp::size) {
                const double t = A[k * ld + j];
                A[k * ld + j] = A[p * ld + j];
                A[p * ld + j] = t;
            }
            if (lane == 0) {
                const double t = b[k];
                b[k] = b[p];
                b[p] = t;
            }
            g.sync();
        }
        const double pinv = 1.0 / A[k * ld + k];
        const double bk = b[k];
        for (int i = k + 1 + lane; i < n; i += Grp::size) {  // lanes own rows
            const double m = A[i * ld + k] * pinv;
            for (int j = k + 1; j < n; ++j) A[i * ld + j] = fma(-m, A[k * ld + j], A[i * ld + j]);
            b[i] = fma(-m, bk, b[i]);
        }
        g.sync();
    }
    // back-substitution, column oriented
    for (int j = n - 1; j >= 0; --j) {
        const double xj = b[j] / A[j * ld + j];
        g.sync();
        if (lane == 0) b[j] = xj;
        for (int i = lane; i < j; i += Grp::size) b[i] = fma(-A[i * ld + j], xj, b[i]);
        g.sync();
    }
    return status;
}
