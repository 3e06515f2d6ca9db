// pd_kernel_a.cu -- stage A kernel: one lane group per (column, layer) item, one Fourier mode per blockIdx.y
#include "pd_launch.h"

template <int LANES, int NC>
__global__ void k_stage_a(PdStageA a, const double* __restrict__ ptab, int items_per_cta, int item_doubles) {
    extern __shared__ double smem[];
    const int m = blockIdx.y;
    const int n = NC > 0 ? NC : a.N, nm = a.NLeg - m;
    double* Q = smem;  // [nm][n] scaled Legendre table of this mode
    for (int idx = threadIdx.x; idx < nm * n; idx += blockDim.x) {
        const int i = idx % n;
        Q[idx] = ptab[((long)m * a.NLeg + m) * n + idx] * sqrt(a.w[i] / a.mu[i]);
    }
    __syncthreads();
    const int gi = threadIdx.x / LANES;
    const long it = (long)blockIdx.x * items_per_cta + gi;
    if (gi >= items_per_cta || it >= (long)a.B * a.L) return;
    SubWarp<LANES> g;
    double* sm = smem + ((nm * n + 1) & ~1) + (long)gi * item_doubles;
    pd_stage_a_item<SubWarp<LANES>, NC>(g, a, (int)(it / a.L), m, (int)(it % a.L), Q, sm);
}

int pd_launch_stage_a(const PdStageA& a, const double* ptab, cudaStream_t st) {
    const int N = a.N, lanes = pd_lanes_for(N);
    const int item_doubles = (pd_stage_a_item_doubles(N, a.NLeg) + 1) & ~1;
    const size_t qbytes = (size_t)((a.NLeg * N + 1) & ~1) * 8;
    int ipc = 128 / lanes;
    while (ipc > 1 && qbytes + (size_t)ipc * item_doubles * 8 > 96 * 1024) ipc >>= 1;
    const int threads = ipc * lanes < 32 ? 32 : ipc * lanes;
    const size_t smem = qbytes + (size_t)ipc * item_doubles * 8;
    if (smem > PD_SMEM_MAX_CTA) return -22;
    const long items = (long)a.B * a.L;
    dim3 grid((unsigned)((items + ipc - 1) / ipc), a.NF);
    cudaError_t e = cudaSuccess;
    PD_DISPATCH_N(N, {
        e = cudaFuncSetAttribute(k_stage_a<LN, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) k_stage_a<LN, NC><<<grid, threads, smem, st>>>(a, ptab, ipc, item_doubles);
    });
    if (e != cudaSuccess) return (int)e;
    return (int)cudaGetLastError();
}
