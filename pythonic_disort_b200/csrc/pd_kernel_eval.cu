// pd_kernel_eval.cu -- evaluation kernels (flux, u0, u, NT corrections) and their C entry points
#include "pd_launch.h"

// ============================================================================
// evaluation kernels
// ============================================================================
template <int LANES, int NC>
__global__ void k_eval_flux(PdEval a, double* Fup, double* Fdn, double* Fdir) {
    extern __shared__ double smem[];
    const int gpc = blockDim.x / LANES, gi = threadIdx.x / LANES;
    const long pt = (long)blockIdx.x * gpc + gi;
    if (pt >= (long)a.B * a.ntau) return;
    SubWarp<LANES> g;
    pd_flux_point<SubWarp<LANES>, NC>(g, a, (int)(pt / a.ntau), (int)(pt % a.ntau), smem + (long)gi * 4 * a.N, Fup, Fdn, Fdir);
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
    const int l = pd_locate(a.st.tau + (long)b * a.L, a.L, tq);
    const double ts = pd_scaled_tau(a, b, l, tq);
    pd_all_modes_point<SubWarp<LANES>, NC>(g, a, b, l, ts, ev, um);
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

extern "C" {

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
    const unsigned grid = (unsigned)((pts + gpc - 1) / gpc);
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
