// pd_layer_ops.cuh -- reflection / transmission operators R^, T^ of one homogeneous layer from its eigen-solution
// (G, K), ONE THREAD per (column, mode, layer) item (N = 8): the part of the boundary-condition stage that does not
// depend on the neighbouring layers -- two SPD inversions and their products, 40 % of a layer's instructions and the
// part with the densest chain of shared-memory round trips in a lane group -- taken out of the sequential sweep of
// pd_stage_b_add.cuh.  G is sector-interleaved over groups of 32 items for this shape (pd_common.cuh), so the 32
// threads of a warp read whole contiguous kilobytes.
//
// With V^ = D (Gp + Gm) / 2, U^ = D (Gp - Gm) / 2 (D = diag(sqrt(w mu))), g_k = v^_k . u^_k (< 0) and
// d_k = -tanh(k_k dtau / 2) / g_k (see the header of pd_stage_b_add.cuh):
//     A1 = (I + U^ d U^T)^-1,   A2 = (I + V^ d V^T)^-1,   R^ = A1 - A2,   T^ = A1 + A2 - I      (all symmetric).
// Both matrices are I + positive semidefinite, so they are inverted in packed symmetric storage by the sweep
// operator without pivoting (every pivot is a Schur complement >= 1).  Everything lives in registers with
// compile-time indices; A1 is parked in shared memory while A2 is formed.  An item whose eigenvectors do not satisfy
// g_k < 0 is marked by a NaN in R^(0,0); the sweep kernel then hands its system to the pivoted band solver.
#pragma once
#include "pd_common.cuh"

template <int N>
struct PdLayerOps {
    static constexpr int NP = N * (N + 1) / 2;
    PD_HD static constexpr int idx(int i, int j) {  // upper-packed index of (i, j), any order
        return (i <= j) ? (i * N - i * (i - 1) / 2 + (j - i)) : (j * N - j * (j - 1) / 2 + (i - j));
    }
};

// X <- X^-1 for a symmetric positive definite X in packed storage (sweep operator, no pivoting)
template <int N>
PD_HD void pd_sym_inverse(double (&X)[N * (N + 1) / 2]) {
    using P = PdLayerOps<N>;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double p = pd_rcp(X[P::idx(k, k)]);
        double t[N];
#pragma unroll
        for (int i = 0; i < N; ++i) t[i] = X[P::idx(i, k)] * p;
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = i; j < N; ++j)
                if (i != k && j != k) X[P::idx(i, j)] = fma(-t[i], X[P::idx(k, j)], X[P::idx(i, j)]);
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (i != k) X[P::idx(i, k)] = t[i];
        X[P::idx(k, k)] = -p;
    }
#pragma unroll
    for (int e = 0; e < P::NP; ++e) X[e] = -X[e];
}

// Gl: the item's base in G (pd_g_base; elements at pd_g_off), Kl: [N], hD: [N] D_i / 2; out: R^, T^ packed symmetric
// [2][N(N+1)/2];
// park: NP doubles of this thread, stride ps
template <int N>
PD_HD void pd_layer_ops_item(const double* Gl, const double* Kl, double dtau, const double* hD, double* out,
                             double* park, int ps) {
    using P = PdLayerOps<N>;
    constexpr int NP = P::NP, NN = N * N;
    static_assert(NP % 4 == 0, "packed size must be a multiple of 4 doubles");
    double M[N][N];  // U^ (first pass), V^ (second pass), by rows
    double gk[N], dk[N];
#pragma unroll
    for (int k = 0; k < N; ++k) gk[k] = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double h = hD[i];
#pragma unroll
        for (int k = 0; k < N; k += 4) {
            double gp[4], gm[4];
            pd_load4(Gl + pd_g_off(i * N + k, N), gp);
            pd_load4(Gl + pd_g_off(NN + i * N + k, N), gm);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double v = h * (gp[q] + gm[q]), u = h * (gp[q] - gm[q]);
                M[i][k + q] = u;
                gk[k + q] = fma(v, u, gk[k + q]);
            }
        }
    }
    bool bad = false;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double em = expm1(-Kl[k] * dtau);  // tanh(x / 2) = -expm1(-x) / (2 + expm1(-x))
        dk[k] = (em / (2.0 + em)) / gk[k];
        bad = bad || !(gk[k] < 0.0) || !(dk[k] >= 0.0);
    }
    // X = I + M d M^T (packed), then its inverse
    auto form = [&](double (&X)[NP]) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double t[N];
#pragma unroll
            for (int k = 0; k < N; ++k) t[k] = dk[k] * M[i][k];
#pragma unroll
            for (int j = i; j < N; ++j) {
                double s0 = (i == j) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
                for (int k = 0; k < N; k += 2) {
                    s0 = fma(t[k], M[j][k], s0);
                    s1 = fma(t[k + 1], M[j][k + 1], s1);
                }
                X[P::idx(i, j)] = s0 + s1;
            }
        }
    };
    {
        double X1[NP];
        form(X1);
        pd_sym_inverse<N>(X1);
#pragma unroll
        for (int e = 0; e < NP; ++e) park[(long)e * ps] = X1[e];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double h = hD[i];
#pragma unroll
        for (int k = 0; k < N; k += 4) {
            double gp[4], gm[4];
            pd_load4(Gl + pd_g_off(i * N + k, N), gp);
            pd_load4(Gl + pd_g_off(NN + i * N + k, N), gm);
#pragma unroll
            for (int q = 0; q < 4; ++q) M[i][k + q] = h * (gp[q] + gm[q]);
        }
    }
    double X2[NP];
    form(X2);
    pd_sym_inverse<N>(X2);
    double R[NP];
#pragma unroll
    for (int e = 0; e < NP; ++e) {
        const double a1 = park[(long)e * ps], a2 = X2[e];
        R[e] = a1 - a2;
        X2[e] = a1 + a2;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) X2[P::idx(i, i)] -= 1.0;
    if (bad) R[0] = nan("");
#pragma unroll
    for (int e = 0; e < NP; e += 4) {
        pd_store4(out + e, R[e], R[e + 1], R[e + 2], R[e + 3]);
        pd_store4(out + NP + e, X2[e], X2[e + 1], X2[e + 2], X2[e + 3]);
    }
}
