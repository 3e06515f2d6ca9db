// pd_kernel_a.cu -- stage A kernel: one lane group per (column, layer) item, one Fourier mode per blockIdx.y
#include "pd_launch.h"
#include "pd_stage_a_sym.cuh"
#include "pd_stage_a_sym16.cuh"

template <int LANES, int NC>
__global__ void k_stage_a(PdStageA a, const double* __restrict__ ptab, int items_per_cta, int item_doubles, int reps) {
    extern __shared__ double smem[];
    if (a.only_flagged && a.nflagged && *a.nflagged == 0) return;  // nothing was handed back (the usual case)
    const int m = blockIdx.y;
    const int n = NC > 0 ? NC : a.N, nm = a.NLeg - m;
    double* Q = smem;  // [nm][n] scaled Legendre table of this mode (built when the CTA has work)
    const int gi = threadIdx.x / LANES;
    const long items = (long)a.B * a.L;
    bool have_q = false;
    // reps = 1 in the normal pass.  The fallback pass after the symmetric kernel (only_flagged) gives every CTA
    // `reps` consecutive item groups and skips a group unless one of its items is flagged: almost always none is,
    // and a million CTAs that only look at one flag each cost more than the looking (1.5 ms per 16k SW columns).
    for (int r = 0; r < reps; ++r) {
        const long it = ((long)blockIdx.x * reps + r) * items_per_cta + gi;
        const bool valid = gi < items_per_cta && it < items;
        if (a.only_flagged) {
            bool flagged = false;
            if (valid && (threadIdx.x % LANES) == 0) {
                const double k0 = a.K[(((it / a.L) * a.NF + m) * a.L + it % a.L) * n];
                flagged = !(k0 == k0);
            }
            if (!__syncthreads_or(flagged)) continue;
        }
        if (!have_q) {
            for (int idx = threadIdx.x; idx < nm * n; idx += blockDim.x) {
                const int i = idx % n;
                Q[idx] = ptab[((long)m * a.NLeg + m) * n + idx] * sqrt(a.w[i] / a.mu[i]);
            }
            __syncthreads();
            have_q = true;
        }
        if (valid) {
            SubWarp<LANES> g;
            double* sm = smem + ((nm * n + 1) & ~1) + (long)gi * item_doubles;
            pd_stage_a_item<SubWarp<LANES>, NC>(g, a, (int)(it / a.L), m, (int)(it % a.L), Q, sm);
        }
    }
}

// one thread per item, symmetric (Cholesky + Jacobi) path, N = 4 or 8
#ifndef PD_SYM_THREADS
#define PD_SYM_THREADS 128
#endif
template <int N>
__global__ void __launch_bounds__(PD_SYM_THREADS, ((N <= 4) ? 4 : 2) * (128 / PD_SYM_THREADS)) k_stage_a_sym(PdStageA a, const double* __restrict__ ptab) {
    extern __shared__ double smem[];
    using P = PdSym<N>;
    const int m = blockIdx.y, nm = a.NLeg - m;
    double* Qs = smem;                                  // [nm][N]
    double* QQ = Qs + ((nm * N + 1) & ~1);              // [nm][NP]
    double* park = QQ + ((nm * P::NP + 1) & ~1);        // [PARK][threads]
    for (int idx = threadIdx.x; idx < nm * N; idx += blockDim.x) {
        const int i = idx % N;
        Qs[idx] = ptab[((long)m * a.NLeg + m) * N + idx] * sqrt(a.w[i] / a.mu[i]);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < nm * P::NP; idx += blockDim.x) {
        const int t = idx / P::NP, e = idx - t * P::NP;
        int i = 0, rem = e;  // unpack e -> (i, j), i <= j
        while (rem >= N - i) {
            rem -= N - i;
            ++i;
        }
        QQ[idx] = Qs[t * N + i] * Qs[t * N + i + rem];
    }
    __syncthreads();
    const long it = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= (long)a.B * a.L) return;
    const int b = (int)(it / a.L), l = (int)(it % a.L);
    const bool done = pd_stage_a_sym_item<N>(a, b, m, l, QQ, Qs, park + threadIdx.x, blockDim.x);
    if (!done) {
        a.K[(((long)b * a.NF + m) * a.L + l) * N] = __longlong_as_double(0x7ff8000000000000LL);
        if (a.nflagged) atomicAdd(a.nflagged, 1);
    }
}

// eight lanes per item, symmetric (Cholesky + two-sided Jacobi) path, N = 16
#ifndef PD_J16_THREADS
#define PD_J16_THREADS 64
#endif
#ifndef PD_J16_MINB
#define PD_J16_MINB 4
#endif
// 252 registers without spills at four CTAs of 64 threads per SM; capping at 168 / 200 registers for five or six CTAs
// spills and is slower (measured 79 against 84 / 133 ms per 2,048 HA columns)
__global__ void __launch_bounds__(PD_J16_THREADS, PD_J16_MINB) k_stage_a_j16(
    PdStageA a, const double* __restrict__ ptab) {
    extern __shared__ double smem[];
    const int m = blockIdx.y, nm = a.NLeg - m;
    double* Qs = smem;            // [nm][16]
    double* tab = Qs + 32 * 16;   // 0.5 / sqrt(w mu), 1 / mu, sqrt(w / mu)
    double* scratch = tab + 48;
    for (int idx = threadIdx.x; idx < nm * 16; idx += blockDim.x) {
        const int i = idx & 15;
        Qs[idx] = ptab[((long)m * a.NLeg + m) * 16 + idx] * sqrt(a.w[i] / a.mu[i]);
    }
    if (threadIdx.x < 16) {
        tab[threadIdx.x] = 0.5 * pd_rsqrt(a.w[threadIdx.x] * a.mu[threadIdx.x]);
        tab[16 + threadIdx.x] = 1.0 / a.mu[threadIdx.x];
        tab[32 + threadIdx.x] = sqrt(a.w[threadIdx.x] / a.mu[threadIdx.x]);
    }
    __syncthreads();
    const long items = (long)a.B * a.L;
    long it = (long)blockIdx.x * (PD_J16_THREADS / 8) + (threadIdx.x >> 3);
    const bool valid = it < items;
    if (!valid) it = items - 1;  // padding lanes of the last CTA run along (the warp's shuffles need them) without storing
    const int b = (int)(it / a.L), l = (int)(it % a.L);
    const bool done = pd_stage_a_j16_item(a, b, m, l, valid, Qs, tab, scratch + (long)(threadIdx.x >> 3) * PdJ16::ITEM);
    if (valid && !done && (threadIdx.x & 7) == 0) {
        a.K[(((long)b * a.NF + m) * a.L + l) * 16] = __longlong_as_double(0x7ff8000000000000LL);
        if (a.nflagged) atomicAdd(a.nflagged, 1);
    }
}

static int launch_j16(const PdStageA& a, const double* ptab, cudaStream_t st) {
    const int ipc = PD_J16_THREADS / 8;
    const size_t smem = (size_t)(32 * 16 + 48 + ipc * PdJ16::ITEM) * 8;
    cudaError_t e = cudaFuncSetAttribute(k_stage_a_j16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const long items = (long)a.B * a.L;
    dim3 grid((unsigned)((items + ipc - 1) / ipc), a.NF);
    k_stage_a_j16<<<grid, PD_J16_THREADS, smem, st>>>(a, ptab);
    return (int)cudaGetLastError();
}

template <int N>
static int launch_sym(const PdStageA& a, const double* ptab, cudaStream_t st) {
    using P = PdSym<N>;
    const size_t smem = (size_t)(((a.NLeg * N + 1) & ~1) + ((a.NLeg * P::NP + 1) & ~1) + P::PARK * PD_SYM_THREADS) * 8;
    cudaError_t e = cudaFuncSetAttribute(k_stage_a_sym<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const long items = (long)a.B * a.L;
    dim3 grid((unsigned)((items + PD_SYM_THREADS - 1) / PD_SYM_THREADS), a.NF);
    k_stage_a_sym<N><<<grid, PD_SYM_THREADS, smem, st>>>(a, ptab);
    return (int)cudaGetLastError();
}

int pd_launch_stage_a(const PdStageA& a_in, int flags, const double* ptab, cudaStream_t st) {
    PdStageA a = a_in;
    a.only_flagged = 0;
    if (a.nflagged) {
        cudaError_t e = cudaMemsetAsync(a.nflagged, 0, sizeof(int32_t), st);
        if (e != cudaSuccess) return (int)e;
    }
    if ((a.N == 4 || a.N == 8 || (a.N == 16 && a.NLeg <= 32)) && !(flags & PD_FLAG_GENERIC_KERNELS)) {  // symmetric fast path first; the general kernel then only redoes flagged items
        const int rc = (a.N == 4) ? launch_sym<4>(a, ptab, st) : (a.N == 8) ? launch_sym<8>(a, ptab, st) : launch_j16(a, ptab, st);
        if (rc) return rc;
        a.only_flagged = 1;
    }
    const int N = a.N, lanes = pd_lanes_for(N);
    const int item_doubles = (pd_stage_a_item_doubles(N, a.NLeg) + 1) & ~1;
    const size_t qbytes = (size_t)((a.NLeg * N + 1) & ~1) * 8;
    int ipc = 128 / lanes;
    const size_t cta_budget = 48 * 1024;  // shared memory per CTA: four resident CTAs pack the SM better than two when an item needs ~10 KB (N = 16; measured 439 -> 355 ms)
    while (ipc > 1 && qbytes + (size_t)ipc * item_doubles * 8 > cta_budget) ipc >>= 1;
    const int threads = ipc * lanes < 32 ? 32 : ipc * lanes;
    const size_t smem = qbytes + (size_t)ipc * item_doubles * 8;
    if (smem > PD_SMEM_MAX_CTA) return -22;
    const long items = (long)a.B * a.L;
    const int reps = a.only_flagged ? 16 : 1;
    dim3 grid((unsigned)((items + (long)ipc * reps - 1) / ((long)ipc * reps)), a.NF);
    cudaError_t e = cudaSuccess;
    PD_DISPATCH_N(N, {
        e = cudaFuncSetAttribute(k_stage_a<LN, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) k_stage_a<LN, NC><<<grid, threads, smem, st>>>(a, ptab, ipc, item_doubles, reps);
    });
    if (e != cudaSuccess) return (int)e;
    return (int)cudaGetLastError();
}
