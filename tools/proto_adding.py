"""NumPy prototype of the pivot-free boundary-condition solve (interaction principle / adding over layers).

Test-side experiment only: replaces oracle.disort_oracle.boundary_solve by an adding-based solve and measures
how far the outputs move on the reference goldens.  Algebra (hat = basis scaled by D = sqrt(w mu), in which the
layer operators are symmetric):
    G = [[V+U, V-U],[V-U, V+U]],  n_j = -k_j (v^_j . u^_j),  d_j = k_j tanh(k_j dtau/2) / n_j
    A1 = (I + U^ d U^T)^-1,  A2 = (I + V^ d V^T)^-1,   R^ = A1 - A2,  T^ = A1 + A2 - I
    layer: u+_top = R u-_top + T u+_bot + s+,   u-_bot = T u-_top + R u+_bot + s-
    stack above interface i: u-_i = Rup_i u+_i + S_i
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import disort_oracle as O  # noqa: E402


def nopivot_solve(M, Y):
    """Gauss-Jordan without pivoting (what the kernel does)."""
    M = M.copy()
    Y = Y.copy()
    n = M.shape[0]
    minpiv = np.inf
    for j in range(n):
        piv = M[j, j]
        minpiv = min(minpiv, abs(piv))
        inv = 1.0 / piv
        for i in range(n):
            if i == j:
                continue
            f = M[i, j] * inv
            M[i, :] -= f * M[j, :]
            Y[i, :] -= f * Y[j, :]
    return Y / np.diag(M)[:, None], minpiv


def chol_inv(A):
    L = np.linalg.cholesky(A)
    Li = np.linalg.inv(L)
    return Li.T @ Li


STATS = {"minpiv": np.inf}


def boundary_solve_adding(p, G, K, Bv, dth, banded_from=10):
    L, N, NQ, NF = p.L, p.N, p.NQuad, p.NF
    taus = p.taus
    D = np.sqrt(p.W * p.mu_pos)
    Di = 1.0 / D
    C = np.zeros((NF, L, NQ))
    I = np.eye(N)
    for m in range(NF):
        Gm_, Km = G[m], K[m]
        has_bdrf = len(p.bdrf) > m
        if has_bdrf:
            Rs, Xs = O.bdrf_tables(p, m)
        bneg = O._bc_vector(p.b_neg, m, N, NF)
        bpos = O._bc_vector(p.b_pos, m, N, NF)

        def part(l, t):  # particular solution of layer l at tau* = t, [2N]
            v = np.zeros(NQ)
            if p.beam:
                v = v + Bv[m][l] * np.exp(-t / p.mu0)
            if m == 0 and p.iso:
                v = v + dth[l] @ (t ** np.arange(p.Ns))
            return v

        Rup = np.zeros((N, N))
        S = D * bneg
        hist = []
        Rl, Tl = [], []
        for l in range(L):
            Gp, Gmm = Gm_[l][:N, :N], Gm_[l][:N, N:]
            V = D[:, None] * (Gp + Gmm) * 0.5
            U = D[:, None] * (Gp - Gmm) * 0.5
            k = Km[l][N:]
            n = -k * np.sum(V * U, axis=0)
            d = k * np.tanh(0.5 * k * (taus[l + 1] - taus[l])) / n
            A1 = chol_inv(I + (U * d) @ U.T)
            A2 = chol_inv(I + (V * d) @ V.T)
            R = A1 - A2
            T = A1 + A2 - I
            pt, pb = part(l, taus[l]), part(l, taus[l + 1])
            ptp, ptm, pbp, pbm = D * pt[:N], D * pt[N:], D * pb[:N], D * pb[N:]
            sp = ptp - R @ ptm - T @ pbp
            sm = pbm - T @ ptm - R @ pbp
            M = I - R @ Rup
            Y, mp = nopivot_solve(M, np.concatenate([T, (R @ S + sp)[:, None]], axis=1))
            STATS["minpiv"] = min(STATS["minpiv"], mp)
            Q, q = Y[:, :N], Y[:, N]
            hist.append((Q, q, Rup, S))
            RQ = Rup @ Q
            Rup_new = R + T @ RQ
            Rup_new = 0.5 * (Rup_new + Rup_new.T)
            S = T @ (Rup @ q + S) + sm
            Rup = Rup_new
        # surface
        bs = bpos.copy()
        if has_bdrf:
            if p.beam:
                bs = bs + Xs * np.exp(-taus[-1] / p.mu0)
            Rsh = D[:, None] * Rs * Di[None, :]
            up = np.linalg.solve(I - Rsh @ Rup, Rsh @ S + D * bs)
        else:
            up = D * bs
        um = Rup @ up + S
        # back sweep
        for l in range(L - 1, -1, -1):
            Q, q, Rup_l, S_l = hist[l]
            up_top = Q @ up + q
            um_top = Rup_l @ up_top + S_l
            pt, pb = part(l, taus[l]), part(l, taus[l + 1])
            ht = np.concatenate([Di * up_top, Di * um_top]) - pt
            hb = np.concatenate([Di * up, Di * um]) - pb
            # symmetric / antisymmetric recovery: s = C- + C+, t = C- - C+ from the sums over the two interfaces
            Gp, Gmm = Gm_[l][:N, :N], Gm_[l][:N, N:]
            V = D[:, None] * (Gp + Gmm) * 0.5
            U = D[:, None] * (Gp - Gmm) * 0.5
            k = Km[l][N:]
            g = np.sum(V * U, axis=0)
            E = np.exp(-k * (taus[l + 1] - taus[l]))
            psum = D * ((ht[:N] + ht[N:]) + (hb[:N] + hb[N:]))
            dsum = D * ((ht[:N] - ht[N:]) + (hb[:N] - hb[N:]))
            s_ = (U.T @ psum) / (2 * g * (1 + E))
            t_ = (V.T @ dsum) / (2 * g * (1 + E))
            C[m, l, :N] = 0.5 * (s_ + t_)
            C[m, l, N:] = 0.5 * (s_ - t_)
            up, um = up_top, um_top
    return G * C[:, :, None, :]


def main():
    import golden_io
    orig = O.boundary_solve
    worst_all = 0.0
    names = sys.argv[1:] or golden_io.suite_names()
    for name in names:
        records, _ = golden_io.load_test(name)
        worst = 0.0
        worst_o = 0.0
        for rec in records[:6]:
            res = {}
            for tag, fn in (("band", orig), ("add", boundary_solve_adding)):
                O.boundary_solve = fn
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    out = O.pydisort(*rec["args"], **rec["kwargs"])
                    res[tag] = [(c, got) for c, got in golden_io.run_calls(out, rec)]
            tol = golden_io.conditioning_tolerance(rec["args"][1])
            for (call, ga), (_, gb) in zip(res["add"], res["band"]):
                scale = golden_io.group_scale(call["outs"])
                for g, g0, r in zip(ga, gb, call["outs"]):
                    e, pw, _ = golden_io.parity(np.squeeze(np.asarray(g)), np.squeeze(r), scale=scale, floor=1e-6)
                    e0, pw0, _ = golden_io.parity(np.squeeze(np.asarray(g0)), np.squeeze(r), scale=scale, floor=1e-6)
                    worst = max(worst, e / tol)
                    worst_o = max(worst_o, e0 / tol)
        print(f"{name:14s} adding err/tol {worst:9.2e}   band(oracle) err/tol {worst_o:9.2e}   min pivot {STATS['minpiv']:.2e}")
        worst_all = max(worst_all, worst)
        STATS["minpiv"] = np.inf
    print("worst", worst_all)
    O.boundary_solve = orig


if __name__ == "__main__":
    main()
