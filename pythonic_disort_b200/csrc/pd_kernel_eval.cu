// pd_kernel_eval.cu -- evaluation kernels (flux, u0, u, NT corrections) and their C entry points
#include <stdlib.h>

#include "pd_launch.h"
#include "pd_inputs.cuh"

// ============================================================================
// evaluation kernels
// ============================================================================
// A warp takes 32 consecutive points.  If every one of them is a layer interface (the usual level grid), each lane
// finishes its own point from the interface radiances -- one contiguous record per point, no shared memory, no
// cooperation.  Otherwise the warp's points go through the lane-group routine, LANES lanes per point.
template <int LANES, int NC>
__global__ void k_eval_flux(PdEval a, double* Fup, double* Fdn, double* Fdir) {
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long pts = (long)a.B * a.ntau;
    const long p0 = ((long)blockIdx.x * (blockDim.x >> 5) + warp) * 32;
    if (p0 >= pts) return;
    const long mine = p0 + lane;
    bool done = true;
    if (a.st.Uif && !a.anti) {
        if (mine < pts) done = pd_flux_point_interface<NC>(a, (int)(mine / a.ntau), (int)(mine % a.ntau), Fup, Fdn, Fdir);
    } else {
        done = false;
    }
    if (__all_sync(0xffffffffu, done)) return;
    SubWarp<LANES> g;
    constexpr int GPW = 32 / LANES;  // points per pass of the warp
    double* sm = smem + (long)(warp * GPW + lane / LANES) * 4 * a.N;
    for (int k = 0; k < LANES; ++k) {
        const long pt = p0 + k * GPW + lane / LANES;
        if (pt < pts) pd_flux_point<SubWarp<LANES>, NC>(g, a, (int)(pt / a.ntau), (int)(pt % a.ntau), sm, Fup, Fdn, Fdir);
    }
}

template <int LANES, int NC>
__global__ void k_eval_u0(PdEval a, double* u0, double* recl) {
    extern __shared__ double smem[];
    const int gpc = blockDim.x / LANES, gi = threadIdx.x / LANES;
    const long pt = (long)blockIdx.x * gpc + gi;
    if (pt >= (long)a.B * a.ntau) return;
    SubWarp<LANES> g;
    pd_u0_point<SubWarp<LANES>, NC>(g, a, (int)(pt / a.ntau), (int)(pt % a.ntau), smem + (long)gi * 4 * a.N, u0, recl);
}

// u(tau, phi) = sum_m u^m(tau) cos(m (phi0 - phi))   (:256-260)
template <int LANES, int NC>
__global__ void k_eval_u(PdEval a, const double* __restrict__ phi_q, int nphi, int group_doubles, double* u,
                         double* ulast) {
    extern __shared__ double smem[];
    const int gpc = blockDim.x / LANES, gi = threadIdx.x / LANES;
    const long pt = (long)blockIdx.x * gpc + gi;
    if (pt >= (long)a.B * a.ntau) return;
    SubWarp<LANES> g;
    const int b = (int)(pt / a.ntau), t = (int)(pt % a.ntau);
    const int n2 = 2 * a.N;
    double* ev = smem + (long)gi * group_doubles;
    double* um = ev + n2;
    const double tq = a.tau_q[pt];
    const int l = pd_locate_group(g, a.st.tau + (long)b * a.L, a.L, tq);
    const double ts = pd_scaled_tau(a, b, l, tq);
    pd_all_modes_point<SubWarp<LANES>, NC>(g, a, b, l, pd_interface_level(a, b, l, tq), ts, ev, um);
    const double* cp = a.st.colp + (long)b * PD_NCOLP;
    const double resc = cp[PD_COL_RESCALE], phi0 = cp[PD_COL_PHI0];
    for (int idx = g.lane(); idx < n2 * nphi; idx += LANES) {
        const int i = idx / nphi, p = idx - i * nphi;
        u[(((long)b * n2 + i) * a.ntau + t) * nphi + p] = resc * pd_azimuth_sum(um + i, n2, a.NF, phi0 - phi_q[p]);
    }
    if (ulast)
        for (int i = g.lane(); i < n2; i += LANES) ulast[((long)b * n2 + i) * a.ntau + t] = um[(a.NF - 1) * n2 + i];
}

// Nakajima-Tanaka corrections added to u in place: one CTA per column; warp 0 runs the TMS layer scans while
// warp 1 prepares the IMS constants, then all threads sweep the (level, stream, azimuth) outputs.
// (Tabulating P_l(nu) per (stream, azimuth) in shared memory and replacing the recurrences by dot products was
// measured to be slower: two loads per FMA make it LSU bound.)
__global__ void __launch_bounds__(256) k_nt(PdEval a, PdNT nt, const double* __restrict__ phi_q, int nphi, double* u) {
    extern __shared__ double smem[];
    const int b = blockIdx.x;
    const int n = a.N, n2 = 2 * n, L = a.L, NA = a.NLeg_all;
    double* Rpos = smem;                 // [n][L]
    double* Rneg = Rpos + n * L;         // [n][L]
    double* imsc = Rneg + n * L;         // [NLeg_all]
    double* imsv = imsc + NA;            // [2]
    double* rinv = imsv + 2;             // [NLeg_all + 1] 1/l
    if (a.st.colp[(long)b * PD_NCOLP + PD_COL_NT] == 0.0) return;  // gate of pydisort.py:375, column part
    for (int i = threadIdx.x; i <= NA; i += blockDim.x) rinv[i] = (i > 0) ? 1.0 / (double)i : 0.0;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        SubWarp<32> g;
        if (L > 1) pd_tms_scans(g, a, b, Rpos, Rneg);
    } else if (warp == 1) {
        SubWarp<32> g;
        pd_ims_setup(g, a, nt, b, imsc, imsv);
    }
    __syncthreads();
    const double resc = a.st.colp[(long)b * PD_NCOLP + PD_COL_RESCALE];
    const long total = (long)a.ntau * n2 * nphi;
    for (long idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int p = (int)(idx % nphi);
        const int i = (int)((idx / nphi) % n2);
        const int t = (int)(idx / ((long)nphi * n2));
        const double tq = a.tau_q[(long)b * a.ntau + t];
        const int l = pd_locate(a.st.tau + (long)b * L, L, tq);
        const double ts = pd_scaled_tau(a, b, l, tq);
        const double v = pd_nt_value(a, nt, b, i, l, tq, ts, phi_q[p], Rpos, Rneg, imsc, imsv,
                                     nt.leg_all + ((long)b * L + l) * NA, rinv);
        u[(((long)b * n2 + i) * a.ntau + t) * nphi + p] += resc * v;
    }
}

// Nakajima-Tanaka corrections, tabulated form (production path for NLeg_all <= NA): one CTA per column.
//   phase 1: TMS layer scans (warp 0), IMS constants (warp 1), per-layer series coefficients
//            c_l[k] = omega*_l [ (2k+1) g_l[k] / (1 - f_l) - w*_l[k] ]   (p_true / (1-f) - p_trunc as ONE series), level -> layer;
//   phase 2: everything that depends on (level, stream) only -- the exponentials of pd_nt_value -- once instead of per azimuth;
//   phase 3: a thread owns one (stream, azimuth) pair, keeps P_k(nu) in registers and walks over levels: one dot product with
//            broadcast 128-bit loads of c_l per output instead of two three-term recurrences.
// Same formulas as pd_nt_value (pydisort.py:409-694), which stays the reference implementation for the host build and the
// fallback for longer phase functions.
template <int NA>
__global__ void __launch_bounds__(256) k_nt_tab(PdEval a, PdNT nt, const double* __restrict__ phi_q, int nphi, double* u) {
    extern __shared__ double smem[];
    const int b = blockIdx.x;
    const int n = a.N, n2 = 2 * n, L = a.L, NAr = a.NLeg_all, ntau = a.ntau;
    if (a.st.colp[(long)b * PD_NCOLP + PD_COL_NT] == 0.0) return;  // gate of pydisort.py:375, column part
    double* Rpos = smem;                     // [n][L]
    double* Rneg = Rpos + n * L;             // [n][L]
    double* imsc = Rneg + n * L;             // [NA]
    double* imsv = imsc + NA;                // [2]
    double* rinv = imsv + 2;                 // [NA + 2] 1/l
    double* coef = rinv + NA + 2;            // [L][NA]
    double* own = coef + (long)L * NA;       // [ntau][n2]
    double* oth = own + (long)ntau * n2;     // [ntau][n2]
    double* chi = oth + (long)ntau * n2;     // [ntau][n]
    double* tsl = chi + (long)ntau * n;      // [ntau] scaled optical depth of the level
    double* termP = tsl + ntau;              // [n][L] terms and decay factors of the TMS layer scans
    double* termN = termP + n * L;
    double* decay = termN + n * L;
    int* lay = reinterpret_cast<int*>(decay + n * L);  // [ntau] layer of the level
    const double* cp = a.st.colp + (long)b * PD_NCOLP;
    const double mu0 = cp[PD_COL_MU0], I0 = cp[PD_COL_I0];
    const double* taus = a.st.taus + (long)b * (L + 1);

    for (int i = threadIdx.x; i <= NA + 1; i += blockDim.x) rinv[i] = (i > 0) ? 1.0 / (double)i : 0.0;
    for (int i = threadIdx.x; i < NA; i += blockDim.x) imsc[i] = 0.0;
    const int warp = threadIdx.x >> 5;
    __syncthreads();
    if (warp == 1) {
        SubWarp<32> g;
        pd_ims_setup(g, a, nt, b, imsc, imsv);
    } else {
        const int tid = threadIdx.x - (warp > 1 ? 32 : 0), nth = blockDim.x - 32;
        // terms of the TMS layer scans (pd_tms_scans, same expressions): the exponentials of all (stream, layer)
        // pairs in parallel, so that the sequential part below is one FMA per layer
        const double* scl = a.st.scale_tau + (long)b * L;
        for (int idx = tid; idx < n * L; idx += nth) {
            const int i = idx / L, r = idx - i * L;
            const double mu = a.st.mu_nodes[i], mi = 1.0 / mu;
            const double dt = taus[r + 1] - taus[r];
            double tp = -expm1(-dt * (mi + 1.0 / mu0)) * exp(-taus[r] / mu0);
            if (a.anti) tp *= mu / scl[r];
            const double dec = exp(-dt * mi);
            const double th = dt * (mi - 1.0 / mu0);
            const double em1 = expm1(-fabs(th));
            double tn = (th >= 0.0) ? -em1 * exp(-taus[r + 1] / mu0) : em1 * dec * exp(-taus[r] / mu0);
            if (a.anti) tn *= -mu / scl[r];
            termP[idx] = tp;
            termN[idx] = tn;
            decay[idx] = dec;
        }
        for (int idx = tid; idx < L * NA; idx += nth) {
            const int l = idx / NA, k = idx - l * NA;
            double v = 0.0;
            if (k < NAr) {
                const double gk = (k == 0) ? 1.0 : nt.leg_all[((long)b * L + l) * NAr + k];
                v = (2 * k + 1) * gk / (1.0 - nt.f[(long)b * L + l]);
                if (k < a.NLeg) v -= nt.wleg[((long)b * L + l) * a.NLeg + k];
                v *= nt.omega_s[(long)b * L + l];
            }
            coef[idx] = v;
        }
        for (int t = tid; t < ntau; t += nth) {
            const double tq = a.tau_q[(long)b * ntau + t];
            const int l = pd_locate(a.st.tau + (long)b * L, L, tq);
            lay[t] = l;
            tsl[t] = pd_scaled_tau(a, b, l, tq);
        }
    }
    __syncthreads();
    if (threadIdx.x < 2 * n && L > 1) {  // the scans themselves: Rpos by lanes 0..n-1 (bottom up), Rneg by lanes n..2n-1
        const int i = threadIdx.x % n;
        if (threadIdx.x < n) {
            double acc = 0.0;
            Rpos[i * L + (L - 1)] = 0.0;
            for (int l = L - 2; l >= 0; --l) {
                const int r = l + 1;
                acc = (l == L - 2) ? termP[i * L + r] : fma(acc, decay[i * L + r], termP[i * L + r]);
                Rpos[i * L + l] = acc;
            }
        } else {
            double acc = 0.0;
            Rneg[i * L] = 0.0;
            for (int l = 1; l < L; ++l) {
                const int r = l - 1;
                acc = (l == 1) ? termN[i * L + r] : fma(acc, decay[i * L + r], termN[i * L + r]);
                Rneg[i * L + l] = acc;
            }
        }
    }
    __syncthreads();
    // ---- phase 2: (level, stream) factors ----
    for (int idx = threadIdx.x; idx < ntau * n2; idx += blockDim.x) {
        const int t = idx / n2, i = idx - t * n2;
        const bool up = i < n;
        const int ii = up ? i : i - n;
        const double mua = a.st.mu_nodes[ii], mi = 1.0 / mua;
        const int l = lay[t];
        const double ts = tsl[t], tq = a.tau_q[(long)b * ntau + t];
        const double sc = a.st.scale_tau[(long)b * L + l];
        const double ttop = taus[l], tbot = taus[l + 1];
        const double e0 = exp(-ts / mu0);
        double o, f2;
        if (up) {
            const double ex = exp((ts - tbot) * mi - tbot / mu0);
            o = a.anti ? e0 / (-sc / mu0) - ex / (sc * mi) : e0 - ex;
            f2 = (L > 1) ? Rpos[ii * L + l] * exp(mi * (ts - tbot)) : 0.0;
        } else {
            const double ex = exp((ttop - ts) * mi - ttop / mu0);
            o = a.anti ? e0 / (-sc / mu0) + ex / (sc * mi) : e0 - ex;
            f2 = (L > 1) ? Rneg[ii * L + l] * exp(mi * (ttop - ts)) : 0.0;
            const double mu0s = imsv[1];
            const double x = mi - 1.0 / mu0s;
            double ch;
            if (a.anti)
                ch = ((mu0s - x * mu0s * (mu0s + tq)) * exp(-tq / mu0s) - mua * exp(-tq * mi)) / (mua * mu0s * x * x);
            else
                ch = ((tq - 1.0 / x) * exp(-tq / mu0s) + exp(-tq * mi) / x) / (mua * mu0s * x);
            chi[t * n + ii] = ch;
        }
        own[idx] = o + f2;
    }
    __syncthreads();
    // ---- phase 3: (stream, azimuth) pairs walk over the levels ----
    const int ncombo = n2 * nphi;
    const int slices = blockDim.x / ncombo > 0 ? blockDim.x / ncombo : 1;
    const double resc = cp[PD_COL_RESCALE];
    for (int c0 = threadIdx.x; c0 < ncombo * slices; c0 += blockDim.x) {
        const int combo = c0 % ncombo, slice = c0 / ncombo;
        const int i = combo / nphi, p = combo - i * nphi;
        const bool up = i < n;
        const int ii = up ? i : i - n;
        const double mua = a.st.mu_nodes[ii];
        const double mus = up ? mua : -mua;
        const double nu = pd_nt_nu(a, b, i, phi_q[p]);
        double P[NA];  // Legendre polynomials P_k(nu), registers
        P[0] = 1.0;
        P[1] = nu;
#pragma unroll
        for (int k = 1; k + 1 < NA; ++k) P[k + 1] = (fma(2.0, (double)k, 1.0) * nu * P[k] - (double)k * P[k - 1]) * rinv[k + 1];
        double ims = 0.0;
        if (!up) {
#pragma unroll
            for (int k = 0; k < NA; ++k) ims = fma(imsc[k], P[k], ims);
            ims *= imsv[0];
        }
        const double pref = (I0 / (4.0 * PD_PI)) * (mu0 / (mu0 + mus));
        for (int t = slice; t < ntau; t += slices) {
            const double* cl = coef + (long)lay[t] * NA;
            double d0 = 0.0, d1 = 0.0;
#pragma unroll
            for (int k = 0; k < NA; k += 2) {
                const pd_d2 c2 = *reinterpret_cast<const pd_d2*>(cl + k);
                d0 = fma(c2.x, P[k], d0);
                d1 = fma(c2.y, P[k + 1], d1);
            }
            double val = pref * (d0 + d1) * own[t * n2 + i];
            if (!up) val = fma(ims, chi[t * n + ii], val);
            u[(((long)b * n2 + i) * ntau + t) * nphi + p] += resc * val;
        }
    }
}

// u or u0 at user polar angles: contraction of the stream axis with the barycentric weight matrix
// (subroutines.py:614-705).  One thread per (column, point); NO outputs per pass; u is read once per pass, coalesced.
template <int NO>
__global__ void k_interp_mu(int n2, long M, int nmu, const double* __restrict__ wts, const double* __restrict__ u,
                            double* __restrict__ out) {
    extern __shared__ double smem[];  // [NO][n2] weights of this pass
    const long chunks = (M + blockDim.x - 1) / blockDim.x;
    const int o0 = blockIdx.y * NO;
    const long b = blockIdx.x / chunks;
    for (int idx = threadIdx.x; idx < NO * n2; idx += blockDim.x) {
        const int o = o0 + idx / n2;
        smem[idx] = (o < nmu) ? wts[(long)o * n2 + idx % n2] : 0.0;
    }
    __syncthreads();
    const long mm = (blockIdx.x % chunks) * blockDim.x + threadIdx.x;
    if (mm >= M) return;
    double acc[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) acc[o] = 0.0;
    const double* ub = u + b * n2 * M + mm;
    for (int i = 0; i < n2; ++i) {
        const double v = ub[(long)i * M];
#pragma unroll
        for (int o = 0; o < NO; ++o) acc[o] = fma(smem[o * n2 + i], v, acc[o]);
    }
#pragma unroll
    for (int o = 0; o < NO; ++o)
        if (o0 + o < nmu) out[(b * nmu + o0 + o) * M + mm] = acc[o];
}

// ---- thermal-source inputs (row f2) ----
__global__ void k_planck_band(long n, const double* __restrict__ T, double wlo, double whi, const double* __restrict__ gl,
                              double* __restrict__ out) {
    __shared__ double g[32];
    if (threadIdx.x < 32) g[threadIdx.x] = gl[threadIdx.x];
    __syncthreads();
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = pd_planck_band_value(T[i], wlo, whi, g);
}

// one CTA per column: band emission at the L + 1 levels, then the spline coefficients of the L layers
__global__ void k_s_poly_coeffs(int L, const double* __restrict__ tau, const double* __restrict__ temper, double wlo, double whi,
                                const double* __restrict__ gl, double* __restrict__ s_poly) {
    extern __shared__ double smem[];  // [32] rule, [L + 1] emission
    double* g = smem;
    double* em = smem + 32;
    const long b = blockIdx.x;
    if (threadIdx.x < 32) g[threadIdx.x] = gl[threadIdx.x];
    __syncthreads();
    for (int lv = threadIdx.x; lv <= L; lv += blockDim.x) em[lv] = pd_planck_band_value(temper[b * (L + 1) + lv], wlo, whi, g);
    __syncthreads();
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        const double x0 = (l == 0) ? 0.0 : tau[b * L + l - 1], x1 = tau[b * L + l];
        pd_linear_segment(x0, em[l], x1, em[l + 1], s_poly + (b * L + l) * 2);
    }
}

// Henyey-Greenstein moments g^k, k < NA: one thread per (layer, group of four moments), 256-bit stores.  pow() keeps
// every moment within 2 ulp of the correctly rounded power (what numpy's g ** k gives); a running product would drift
// by k / 2 ulp, which the NQuad = 32 problems amplify to 1e-11 of the radiances.
__global__ void k_hg_moments(long n, int NA, const double* __restrict__ g, double* __restrict__ out) {
    const int q4 = (NA + 3) / 4;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * q4) return;
    const long i = idx / q4;
    const int k0 = (int)(idx - i * q4) * 4;
    const double gi = g[i];
    double v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = (k0 + e == 0) ? 1.0 : ((k0 + e == 1) ? gi : pow(gi, (double)(k0 + e)));
    double* o = out + i * NA + k0;
    if ((NA & 3) == 0) pd_store4(o, v[0], v[1], v[2], v[3]);
    else
        for (int e = 0; e < 4 && k0 + e < NA; ++e) o[e] = v[e];
}

// thermal source linear in tau inside every layer, from its level values
__global__ void k_level_source(long n, int L, const double* __restrict__ tau, const double* __restrict__ lev,
                               double* __restrict__ s_poly) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const long b = idx / L;
    const int l = (int)(idx - b * L);
    const double x0 = (l == 0) ? 0.0 : tau[idx - 1], x1 = tau[idx];
    pd_linear_segment(x0, lev[b * (L + 1) + l], x1, lev[b * (L + 1) + l + 1], s_poly + idx * 2);
}

// Hapke BDRF Fourier modes (row f4): one thread per (reflection node i, incidence cosine j); out[m][i][j]
template <int NFMAX>
__global__ void k_hapke_modes(int N, long M, int NF, int npanel, const double* __restrict__ gl, const double* __restrict__ mu,
                              const double* __restrict__ mup, double B0, double HH, double W, double* __restrict__ out) {
    __shared__ double g[32];
    if (threadIdx.x < 32) g[threadIdx.x] = gl[threadIdx.x];
    __syncthreads();
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)N * M) return;
    const int i = (int)(idx / M);
    const long j = idx - (long)i * M;
    pd_hapke_modes_point<NFMAX>(mu[i], mup[j], NF, npanel, g, B0, HH, W, out + (long)i * M + j, (long)N * M);
}

extern "C" {

int pd_hapke_modes(int N, long M, int NF, int npanel, const double* gl16, const double* mu, const double* mup, double B0,
                   double HH, double W, double* out, void* stream) {
    if (N < 1 || M < 1 || NF < 1 || NF > 64 || npanel < 1 || 16 * npanel < NF || !gl16 || !mu || !mup || !out) return -60;
    const int threads = 128;
    const long blocks = ((long)N * M + threads - 1) / threads;
    if (blocks > 2147483647L) return -61;
    if (NF <= 16) k_hapke_modes<16><<<(unsigned)blocks, threads, 0, pd_stream(stream)>>>(N, M, NF, npanel, gl16, mu, mup, B0, HH, W, out);
    else if (NF <= 32) k_hapke_modes<32><<<(unsigned)blocks, threads, 0, pd_stream(stream)>>>(N, M, NF, npanel, gl16, mu, mup, B0, HH, W, out);
    else k_hapke_modes<64><<<(unsigned)blocks, threads, 0, pd_stream(stream)>>>(N, M, NF, npanel, gl16, mu, mup, B0, HH, W, out);
    return (int)cudaGetLastError();
}

int pd_hg_moments(long n, int NLeg_all, const double* g, double* out, void* stream) {
    if (n < 1 || NLeg_all < 1 || !g || !out) return -70;
    const int threads = 256;
    const long blocks = (n * ((NLeg_all + 3) / 4) + threads - 1) / threads;
    if (blocks > 2147483647L) return -71;
    k_hg_moments<<<(unsigned)blocks, threads, 0, pd_stream(stream)>>>(n, NLeg_all, g, out);
    return (int)cudaGetLastError();
}

int pd_level_source(int B, int L, const double* tau, const double* lev, double* s_poly, void* stream) {
    if (B < 1 || L < 1 || !tau || !lev || !s_poly) return -70;
    const int threads = 256;
    const long n = (long)B * L, blocks = (n + threads - 1) / threads;
    if (blocks > 2147483647L) return -71;
    k_level_source<<<(unsigned)blocks, threads, 0, pd_stream(stream)>>>(n, L, tau, lev, s_poly);
    return (int)cudaGetLastError();
}

int pd_planck_band(long n, const double* T, double wvnmlo, double wvnmhi, const double* gl16, double* out, void* stream) {
    if (n < 1 || !T || !gl16 || !out) return -50;
    const int threads = 128;
    const long blocks = (n + threads - 1) / threads;
    if (blocks > 2147483647L) return -51;
    k_planck_band<<<(unsigned)blocks, threads, 0, pd_stream(stream)>>>(n, T, wvnmlo, wvnmhi, gl16, out);
    return (int)cudaGetLastError();
}

int pd_s_poly_coeffs(int B, int L, const double* tau, const double* temper, double wvnmlo, double wvnmhi, const double* gl16,
                     double* s_poly, void* stream) {
    if (B < 1 || L < 1 || !tau || !temper || !gl16 || !s_poly) return -50;
    const int threads = (L + 1 <= 32) ? 32 : (L + 1 <= 64 ? 64 : 128);
    k_s_poly_coeffs<<<B, threads, (size_t)(32 + L + 1) * 8, pd_stream(stream)>>>(L, tau, temper, wvnmlo, wvnmhi, gl16, s_poly);
    return (int)cudaGetLastError();
}

int pd_interp_mu(int B, int n2, long M, int nmu, const double* wts, const double* u, double* out, void* stream) {
    if (B < 1 || n2 < 2 || (n2 & 1) || M < 1 || nmu < 1 || !wts || !u || !out) return -40;
    constexpr int NO = 8;
    const int threads = 256;
    const long blocks = (long)B * ((M + threads - 1) / threads);
    if (blocks > 2147483647L || (nmu + NO - 1) / NO > 65535) return -41;
    const dim3 grid((unsigned)blocks, (unsigned)((nmu + NO - 1) / NO));
    k_interp_mu<NO><<<grid, threads, (size_t)NO * n2 * 8, pd_stream(stream)>>>(n2, M, nmu, wts, u, out);
    return (int)cudaGetLastError();
}

static PdEval make_eval(const pd_config* cfg, const pd_state* st, const double* tau_q, int ntau, int anti) {
    PdEval a;
    a.B = cfg->B; a.L = cfg->L; a.N = cfg->NQuad / 2; a.NF = cfg->NFourier; a.Ns = cfg->Nscoeffs;
    a.NLeg = cfg->NLeg; a.NLeg_all = cfg->NLeg_all;
    a.beam = (cfg->flags & PD_FLAG_BEAM) != 0; a.iso = (cfg->flags & PD_FLAG_ISO) != 0;
    a.st = *st; a.tau_q = tau_q; a.ntau = ntau; a.anti = anti;
    return a;
}

int pd_eval_flux(const pd_config* cfg, const pd_state* st, const double* tau_q, int ntau, int anti, double* Fup,
                 double* Fdn_diffuse, double* Fdn_direct, void* stream) {
    if (int e = pd_check_cfg(cfg)) return e;
    if (ntau < 1) return -30;
    const PdEval a = make_eval(cfg, st, tau_q, ntau, anti);
    const int lanes = pd_lanes_for(a.N), threads = 128, gpc = threads / lanes;
    const long pts = (long)a.B * ntau;
    const size_t smem = (size_t)gpc * 4 * a.N * 8;
    const unsigned grid = (unsigned)((pts + threads - 1) / threads);  // a point per thread (k_eval_flux)
    PD_DISPATCH_N(a.N, (k_eval_flux<LN, NC><<<grid, threads, smem, pd_stream(stream)>>>(a, Fup, Fdn_diffuse, Fdn_direct)));
    return (int)cudaGetLastError();
}

int pd_eval_u0(const pd_config* cfg, const pd_state* st, const double* tau_q, int ntau, int anti, double* u0,
               double* recl, void* stream) {
    if (int e = pd_check_cfg(cfg)) return e;
    if (ntau < 1) return -30;
    const PdEval a = make_eval(cfg, st, tau_q, ntau, anti);
    const int lanes = pd_lanes_for(a.N), threads = 128, gpc = threads / lanes;
    const long pts = (long)a.B * ntau;
    const size_t smem = (size_t)gpc * 4 * a.N * 8;
    const unsigned grid = (unsigned)((pts + gpc - 1) / gpc);
    PD_DISPATCH_N(a.N, (k_eval_u0<LN, NC><<<grid, threads, smem, pd_stream(stream)>>>(a, u0, recl)));
    return (int)cudaGetLastError();
}

int pd_eval_u(const pd_config* cfg, const pd_state* st, const double* tau_q, int ntau, const double* phi_q, int nphi,
              int anti, int nt, const double* omega, const double* f, const double* leg_all, const double* omega_s,
              const double* wleg, double* u, double* ulast, void* stream) {
    if (int e = pd_check_cfg(cfg)) return e;
    if (ntau < 1 || nphi < 1) return -30;
    const PdEval a = make_eval(cfg, st, tau_q, ntau, anti);
    const int lanes = pd_lanes_for(a.N);
    const int group_doubles = (a.NF + 1) * 2 * a.N;
    int gpc = 128 / lanes;
    while (gpc > 1 && (size_t)gpc * group_doubles * 8 > 96 * 1024) gpc >>= 1;
    const size_t smem = (size_t)gpc * group_doubles * 8;
    if (smem > PD_SMEM_MAX_CTA) return -31;
    const int threads = gpc * lanes < 32 ? 32 : gpc * lanes;
    const long pts = (long)a.B * ntau;
    const unsigned grid = (unsigned)((pts + gpc - 1) / gpc);
    cudaError_t e = cudaSuccess;
    PD_DISPATCH_N(a.N, {
        e = cudaFuncSetAttribute(k_eval_u<LN, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            k_eval_u<LN, NC><<<grid, threads, smem, pd_stream(stream)>>>(a, phi_q, nphi, group_doubles, u, ulast);
    });
    if (e != cudaSuccess) return (int)e;
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    if (nt) {
        if (!omega || !f || !leg_all || !omega_s || !wleg) return -32;
        PdNT p;
        p.omega = omega; p.f = f; p.leg_all = leg_all; p.omega_s = omega_s; p.wleg = wleg;
        // tabulated kernel for NLeg_all <= 32 (registers hold P_k(nu)); recurrence kernel otherwise or with PD_FLAG_GENERIC_KERNELS
        constexpr int NA = 32;
        const size_t smt = (size_t)(5 * a.N * a.L + NA + 2 + NA + 2 + (size_t)a.L * NA + (size_t)ntau * (5 * a.N + 1) +
                                    (ntau + 1) / 2 + 2) * 8;
        if (a.NLeg_all <= NA && 2 * a.N * nphi <= 256 && smt <= PD_SMEM_MAX_CTA && !(cfg->flags & PD_FLAG_GENERIC_KERNELS)) {
            e = cudaFuncSetAttribute(k_nt_tab<NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smt);
            if (e != cudaSuccess) return (int)e;
            k_nt_tab<NA><<<a.B, 256, smt, pd_stream(stream)>>>(a, p, phi_q, nphi, u);
            return (int)cudaGetLastError();
        }
        const size_t sm2 = (size_t)(2 * a.N * a.L + 2 * a.NLeg_all + 4) * 8;
        if (sm2 > PD_SMEM_MAX_CTA) return -33;
        e = cudaFuncSetAttribute(k_nt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
        if (e != cudaSuccess) return (int)e;
        k_nt<<<a.B, 256, sm2, pd_stream(stream)>>>(a, p, phi_q, nphi, u);
        e = cudaGetLastError();
    }
    return (int)e;
}


}  // extern "C"
