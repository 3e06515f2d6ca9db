// pd_kernel_b.cu -- stage B kernels: the boundary-condition solve of every (column, Fourier mode) system.
//   k_stage_b_add<N>  production path (N = 8, 16): a lane group per system, block elimination over the interface
//                     radiances (pd_stage_b_add.cuh); persistent grid, one history slot per resident system
//   k_layer_ops<8>    N = 8: reflection / transmission operators of every layer, one thread per (column, mode, layer)
//                     item (pd_layer_ops.cuh); k_stage_b_add<8> then only runs the sweep over the layers
//   k_stage_b_tps<N>  production path (N = 2, 4): the same elimination with ONE THREAD per system, every matrix in
//                     registers (pd_stage_b_tps.cuh); persistent grid, history interleaved over the threads
//   k_stage_b<NC>     size-generic pivoted band solver (pd_stage_b.cuh), one warp per system: any N, the
//                     PD_FLAG_GENERIC_KERNELS test path, and the second pass over systems the first kernel flagged
#include "pd_launch.h"
#include "pd_layer_ops.cuh"
#include "pd_stage_b_add.cuh"
#include "pd_stage_b_tps.cuh"

template <int NC>
__global__ void k_stage_b(PdStageB a, double* hist, long hist_doubles, int sys_doubles, const int32_t* only_flagged,
                          const int32_t* nflagged) {
    extern __shared__ double smem[];
    if (only_flagged && nflagged && *nflagged == 0) return;  // the first kernel handed nothing back (the usual case)
    const int wpb = blockDim.x >> 5, w = threadIdx.x >> 5;
    const long slot = (long)blockIdx.x * wpb + w;
    const long nslots = (long)gridDim.x * wpb;
    SubWarp<32> g;
    double* sm = smem + (long)w * sys_doubles;
    double* h = hist + slot * hist_doubles;
    const long nsys = (long)a.B * a.NF;
    for (long s = slot; s < nsys; s += nslots) {
        if (only_flagged && only_flagged[s] == 0) continue;  // warp-uniform
        pd_stage_b_system<SubWarp<32>, NC>(g, a, (int)(s / a.NF), (int)(s % a.NF), sm, h);
    }
}

// Lanes per system (N = 8, 16): every N x N product needs N^2 operands per lane from other lanes (shared memory),
// however the columns are dealt out; with two columns per lane each operand feeds two FMAs, which is what the
// shared-memory pipe (one 128-byte wavefront per cycle per SM against two warp-wide DFMAs) needs to stay off the
// critical path.  N = 16 keeps one column per lane (two would spill).  Measured on the SW ensemble (N = 8): four
// lanes at 255 registers / 8 warps per SM and eight lanes at 128 registers / 16 warps per SM run within 3 % of each
// other -- half the occupancy is paid for by half the operand traffic per FMA.
#ifndef PD_ADD_LS
#define PD_ADD_LS(N) ((N) == 8 ? 4 : (N))  /* N = 8, 16 only: smaller systems take the one-thread-per-system kernel */
#endif
template <int N>
struct AddCfg {
    static constexpr int LS = PD_ADD_LS(N);              // lanes per system
    static constexpr int THREADS = 64;                   // two warps per CTA
    static constexpr int SPC = THREADS / LS;             // systems per CTA
    static constexpr int MINB = 4;                       // resident CTAs per SM the kernel is compiled for
    using F = PdStageBAdd<N, LS>;
    static constexpr size_t SMEM = (size_t)F::SD * 8 * SPC;
};

#ifndef PD_OPS_THREADS
#define PD_OPS_THREADS 128
#endif
template <int N>
__global__ void __launch_bounds__(PD_OPS_THREADS, 2) k_layer_ops(PdStageB a, double* RT) {
    extern __shared__ double smem[];
    double* hD = smem;            // [N] sqrt(w mu) / 2
    double* park = smem + 16;     // [NP][threads]
    if (threadIdx.x < N) hD[threadIdx.x] = 0.5 * sqrt(a.w[threadIdx.x] * a.mu[threadIdx.x]);
    __syncthreads();
    const long item = (long)blockIdx.x * blockDim.x + threadIdx.x;  // (column, mode, layer)
    if (item >= (long)a.B * a.NF * a.L) return;
    const int l = (int)(item % a.L);
    const int b = (int)(item / a.L / a.NF);
    const double* ts = a.taus + (long)b * (a.L + 1) + l;
    pd_layer_ops_item<N>(a.G + pd_g_base(item, N), a.K + item * N, ts[1] - ts[0], hD, RT + item * N * (N + 1),
                         park + threadIdx.x, blockDim.x);
}

template <int N>
static int launch_ops(const PdStageB& a, double* RT, cudaStream_t st) {
    const size_t smem = (size_t)(16 + PdLayerOps<N>::NP * PD_OPS_THREADS) * 8;
    cudaError_t e = cudaFuncSetAttribute(k_layer_ops<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const long items = (long)a.B * a.NF * a.L;
    k_layer_ops<N><<<(unsigned)((items + PD_OPS_THREADS - 1) / PD_OPS_THREADS), PD_OPS_THREADS, smem, st>>>(a, RT);
    return (int)cudaGetLastError();
}

template <int N, bool SPLIT>
__global__ void __launch_bounds__(AddCfg<N>::THREADS, AddCfg<N>::MINB) k_stage_b_add(PdStageB a, double* hist, long hist_doubles, int32_t* sysflag,
                                                                                     int32_t* nflagged) {
    extern __shared__ double smem[];
    using Cf = AddCfg<N>;
    const int gi = threadIdx.x / Cf::LS;
    const long slot = (long)blockIdx.x * Cf::SPC + gi;
    const long nslots = (long)gridDim.x * Cf::SPC;
    SubWarp<Cf::LS> g;
    double* sm = smem + (long)gi * Cf::F::SD;
    double* h = hist + slot * hist_doubles;
    const long nsys = (long)a.B * a.NF;
    for (long s = slot; s < nsys; s += nslots) {
        const bool ok = pd_stage_b_add<SubWarp<Cf::LS>, N, SPLIT>(g, a, (int)(s / a.NF), (int)(s % a.NF), sm, h);
        if (!ok && g.lane() == 0) {  // redone by k_stage_b
            sysflag[s] = 1;
            atomicAdd(nflagged, 1);
        }
    }
}

#ifndef PD_TPS_THREADS
#define PD_TPS_THREADS 64
#endif
#ifndef PD_TPS_MINB
#define PD_TPS_MINB 4
#endif
template <int N>
__global__ void __launch_bounds__(PD_TPS_THREADS, PD_TPS_MINB) k_stage_b_tps(PdStageB a, double* hist, int32_t* sysflag,
                                                                             int32_t* nflagged) {
    const long slot = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long nslots = (long)gridDim.x * blockDim.x;
    const long nsys = (long)a.B * a.NF;
    for (long s = slot; s < nsys; s += nslots) {
        const bool ok = pd_stage_b_tps<N>(a, (int)(s / a.NF), (int)(s % a.NF), hist + slot, nslots);
        if (!ok) {  // redone by k_stage_b
            sysflag[s] = 1;
            atomicAdd(nflagged, 1);
        }
    }
}

template <int N>
static void plan_tps(StageBPlan& p, long nsys, int L) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_stage_b_tps<N>, PD_TPS_THREADS, 0) != cudaSuccess || occ < 1) {
        cudaGetLastError();  // no device (sizing from a host-only process): plan for the compiled residency
        occ = PD_TPS_MINB;
    }
    long blocks = (nsys + PD_TPS_THREADS - 1) / PD_TPS_THREADS;
    if (blocks > (long)PD_NUM_SMS * occ) blocks = (long)PD_NUM_SMS * occ;
    p.add = 2;
    p.add_blocks = (int)blocks;
    p.add_slots = blocks * PD_TPS_THREADS;
    p.add_hist = (long)L * PdStageBTps<N>::HIST_PER_LAYER;
}

template <int N>
static int launch_tps(const PdStageB& a, const StageBPlan& p, double* hist, int32_t* sysflag, cudaStream_t st) {
    k_stage_b_tps<N><<<p.add_blocks, PD_TPS_THREADS, 0, st>>>(a, hist, sysflag, sysflag + (p.flag_bytes - 256) / 4);
    return (int)cudaGetLastError();
}

template <int N, bool SPLIT>
static void plan_add(StageBPlan& p, long nsys, int L) {
    using Cf = AddCfg<N>;
    int occ = 0;
    if (cudaFuncSetAttribute(k_stage_b_add<N, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cf::SMEM) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_stage_b_add<N, SPLIT>, Cf::THREADS, Cf::SMEM) != cudaSuccess || occ < 1) {
        cudaGetLastError();  // no device (sizing from a host-only process): plan for the compiled residency
        occ = Cf::MINB;
    }
    long blocks = (nsys + Cf::SPC - 1) / Cf::SPC;
    if (blocks > (long)PD_NUM_SMS * occ) blocks = (long)PD_NUM_SMS * occ;
    p.add = 1;
    p.add_blocks = (int)blocks;
    p.add_slots = blocks * Cf::SPC;
    p.add_hist = (long)L * Cf::F::HIST_PER_LAYER;
    p.split = SPLIT ? 1 : 0;
    p.rt_doubles = SPLIT ? (size_t)nsys * L * N * (N + 1) : 0;
}

template <int N, bool SPLIT>
static int launch_add(const PdStageB& a, const StageBPlan& p, double* hist, int32_t* sysflag, cudaStream_t st) {
    using Cf = AddCfg<N>;
    cudaError_t e = cudaFuncSetAttribute(k_stage_b_add<N, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cf::SMEM);
    if (e != cudaSuccess) return (int)e;
    k_stage_b_add<N, SPLIT><<<p.add_blocks, Cf::THREADS, Cf::SMEM, st>>>(a, hist, p.add_hist, sysflag, sysflag + (p.flag_bytes - 256) / 4);
    return (int)cudaGetLastError();
}

static bool add_supported(int N) { return N == 2 || N == 4 || N == 8 || N == 16; }

StageBPlan pd_plan_stage_b(int B, int NF, int N, int L, int flags) {
    StageBPlan p = {};
    const long nsys = (long)B * NF;
    if (add_supported(N) && !(flags & PD_FLAG_GENERIC_KERNELS)) {
        switch (N) {
            case 2: plan_tps<2>(p, nsys, L); break;
            case 4: plan_tps<4>(p, nsys, L); break;
            case 8: plan_add<8, true>(p, nsys, L); break;
            default: plan_add<16, false>(p, nsys, L); break;
        }
        p.flag_bytes = (((size_t)nsys * sizeof(int32_t) + 255) & ~(size_t)255) + 256;  // + the counter of flagged systems
    }
    // pivoted band solver: the whole job, or a small grid for the systems the first kernel flags
    p.sys_doubles = (pd_stage_b_doubles(N) + 1) & ~1;
    const size_t per_warp = (size_t)p.sys_doubles * 8;
    int wpb = 4;
    while (wpb > 1 && per_warp * wpb > 64 * 1024) wpb >>= 1;
    p.wpb = wpb;
    p.smem = per_warp * wpb;
    int ctas_per_sm = (int)(PD_SMEM_BUDGET / p.smem);
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm * wpb > 24) ctas_per_sm = 24 / wpb;
    if (p.add) ctas_per_sm = 1;
    long blocks = (nsys + wpb - 1) / wpb;
    if (blocks > (long)PD_NUM_SMS * ctas_per_sm) blocks = (long)PD_NUM_SMS * ctas_per_sm;
    p.blocks = (int)blocks;
    p.slots = blocks * wpb;
    p.hist_doubles = pd_stage_b_history_doubles(N, L);
    return p;
}

template <int NC>
static int launch_b(const PdStageB& a, const StageBPlan& pb, double* hist, const int32_t* only_flagged, cudaStream_t st) {
    const int32_t* nflagged = only_flagged ? only_flagged + (pb.flag_bytes - 256) / 4 : nullptr;
    cudaError_t e = cudaFuncSetAttribute(k_stage_b<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pb.smem);
    if (e != cudaSuccess) return (int)e;
    k_stage_b<NC><<<pb.blocks, pb.wpb * 32, pb.smem, st>>>(a, hist, pb.hist_doubles, pb.sys_doubles, only_flagged, nflagged);
    return (int)cudaGetLastError();
}

int pd_launch_stage_b(const PdStageB& a, int flags, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    const StageBPlan pb = pd_plan_stage_b(a.B, a.NF, a.N, a.L, flags);
    if (workspace_bytes < pb.bytes() || !workspace) return -20;
    if (pb.smem > PD_SMEM_MAX_CTA) return -21;
    char* ws = static_cast<char*>(workspace);
    int32_t* sysflag = nullptr;
    if (pb.add) {
        sysflag = reinterpret_cast<int32_t*>(ws);
        double* rt = reinterpret_cast<double*>(ws + pb.flag_bytes);
        double* hist = rt + pb.rt_doubles;
        cudaError_t e = cudaMemsetAsync(sysflag, 0, pb.flag_bytes, st);
        if (e != cudaSuccess) return (int)e;
        PdStageB as = a;
        as.RT = pb.split ? rt : nullptr;
        int rc = 0;
        if (pb.split) rc = launch_ops<8>(as, rt, st);
        if (rc) return rc;
        switch (a.N) {
            case 2: rc = launch_tps<2>(as, pb, hist, sysflag, st); break;
            case 4: rc = launch_tps<4>(as, pb, hist, sysflag, st); break;
            case 8: rc = launch_add<8, true>(as, pb, hist, sysflag, st); break;
            default: rc = launch_add<16, false>(as, pb, hist, sysflag, st); break;
        }
        if (rc) return rc;
        ws += pb.flag_bytes + (pb.rt_doubles + (size_t)pb.add_slots * pb.add_hist) * 8;
    }
    double* hist_b = reinterpret_cast<double*>(ws);
    switch (a.N) {
        case 4: return launch_b<4>(a, pb, hist_b, sysflag, st);
        case 8: return launch_b<8>(a, pb, hist_b, sysflag, st);
        case 16: return launch_b<16>(a, pb, hist_b, sysflag, st);
        default: return launch_b<0>(a, pb, hist_b, sysflag, st);
    }
}
