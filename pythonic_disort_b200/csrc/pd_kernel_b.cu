// pd_kernel_b.cu -- stage B kernel: one warp per (column, mode) system, persistent grid, per-warp history slot
#include <stdlib.h>

#include "pd_launch.h"
#include "pd_stage_b_fast.cuh"
#include "pd_stage_b_row.cuh"
#include "pd_stage_b_row3.cuh"
#include "pd_stage_b_mma.cuh"

template <int NC>
__global__ void k_stage_b(PdStageB a, double* hist, long hist_doubles, int sys_doubles) {
    extern __shared__ double smem[];
    const int wpb = blockDim.x >> 5, w = threadIdx.x >> 5;
    const long slot = (long)blockIdx.x * wpb + w;
    const long nslots = (long)gridDim.x * wpb;
    SubWarp<32> g;
    double* sm = smem + (long)w * sys_doubles;
    double* h = hist + slot * hist_doubles;
    const long nsys = (long)a.B * a.NF;
    for (long s = slot; s < nsys; s += nslots)
        pd_stage_b_system<SubWarp<32>, NC>(g, a, (int)(s / a.NF), (int)(s % a.NF), sm, h);
}

// production sizes: lanes per system of the fast kernel (0 = size-generic kernel, one warp per system)
static int fast_lanes(int N) {
    if (const char* e = getenv("PD_STAGE_B_GENERIC"))
        if (e[0] == '1') return 0;
    int ls = (N == 4) ? 16 : (N == 8 || N == 16) ? 32 : 0;
    if (const char* e = getenv("PD_STAGE_B_LS")) {
        const int v = atoi(e);
        if (ls && (v == 8 || v == 16 || v == 32) && v >= 2 * N && (4 * N) % v == 0) ls = v;
    }
    return ls;
}

template <int N, int LS>
__global__ void __launch_bounds__(128, (N <= 8 && LS == 32) ? 6 : 1) k_stage_b_fast(PdStageB a, double* hist, long hist_doubles) {
    extern __shared__ double smem[];
    constexpr int SD = (PdStageBFast<N>::SMEM_DOUBLES + 1) & ~1;
    const int gpb = blockDim.x / LS, gi = threadIdx.x / LS;
    const long slot = (long)blockIdx.x * gpb + gi;
    const long nslots = (long)gridDim.x * gpb;
    SubWarp<LS> g;
    double* sm = smem + (long)gi * SD;
    double* h = hist + slot * hist_doubles;
    const long nsys = (long)a.B * a.NF;
    for (long s = slot; s < nsys; s += nslots) pd_stage_b_fast<N, LS>(g, a, (int)(s / a.NF), (int)(s % a.NF), sm, h);
}

template <int N, int LS>
static int fast_ctas_per_sm(size_t smem, int threads) {
    // resident CTAs per SM from the compiled kernel's real register / shared-memory footprint
    // (a host-side query of the cubin: needs no GPU, so pd_workspace_bytes stays callable anywhere)
    int n = 0;
    if (cudaFuncSetAttribute(k_stage_b_fast<N, LS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_stage_b_fast<N, LS>, threads, smem) != cudaSuccess || n < 1) {
        cudaGetLastError();
        n = 0;
    }
    return n;
}

// register-resident variant (N = 4, 8): one lane per panel row, LS = 16 / 32 lanes per system
static bool use_reg(int N) {
    if (N != 4 && N != 8) return false;
    if (const char* e = getenv("PD_STAGE_B_SMEM"))
        if (e[0] == '1') return false;
    if (const char* e = getenv("PD_STAGE_B_GENERIC"))
        if (e[0] == '1') return false;
    return true;
}

template <int N, int MINB, bool SH>
__global__ void __launch_bounds__(128, MINB) k_stage_b_reg(PdStageB a, double* hist, long hist_doubles) {
    extern __shared__ double smem[];
    constexpr int LS = 4 * N;
    const int SD = (PdStageBRow<N>::smem_doubles(a.L) + 1) & ~1;
    const int gpb = blockDim.x / LS, gi = threadIdx.x / LS;
    const long slot = (long)blockIdx.x * gpb + gi;
    const long nslots = (long)gridDim.x * gpb;
    SubWarp<LS> g;
    double* sm = smem + (long)gi * SD;
    double* h = hist + slot * hist_doubles;
    const long nsys = (long)a.B * a.NF;
    for (long s = slot; s < nsys; s += nslots) pd_stage_b_row<N, LS, SH>(g, a, (int)(s / a.NF), (int)(s % a.NF), sm, h);
}

static int reg_minb() {  // resident CTAs per SM the register-resident kernel is compiled for (tuning knob)
    if (const char* e = getenv("PD_STAGE_B_MINB")) {
        const int v = atoi(e);
        if (v >= 3 && v <= 5) return v;
    }
    return 4;
}

template <int N, int MINB, bool SH>
static StageBPlan plan_reg(int B, int NF, int L) {
    StageBPlan p;
    constexpr int LS = 4 * N;
    p.sys_doubles = (PdStageBRow<N>::smem_doubles(L) + 1) & ~1;
    p.wpb = 4;
    const int gpb = p.wpb * (32 / LS);
    p.smem = (size_t)p.sys_doubles * 8 * gpb;
    int ctas_per_sm = (int)(PD_SMEM_BUDGET / p.smem);
    if (ctas_per_sm * p.wpb > 32) ctas_per_sm = 32 / p.wpb;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    int occ = 0;
    if (cudaFuncSetAttribute(k_stage_b_reg<N, MINB, SH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_stage_b_reg<N, MINB, SH>, p.wpb * 32, p.smem) == cudaSuccess &&
        occ > 0) {
        if (occ < ctas_per_sm) ctas_per_sm = occ;
    } else {
        cudaGetLastError();
    }
    const long nsys = (long)B * NF;
    long blocks = (nsys + gpb - 1) / gpb;
    if (blocks > (long)PD_NUM_SMS * ctas_per_sm) blocks = (long)PD_NUM_SMS * ctas_per_sm;
    p.blocks = (int)blocks;
    p.slots = blocks * gpb;
    p.hist_doubles = (long)L * PdStageBRow<N>::HIST_PER_LAYER;
    return p;
}

template <int N, int MINB, bool SH>
static int launch_reg(const PdStageB& a, const StageBPlan& pb, void* workspace, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(k_stage_b_reg<N, MINB, SH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pb.smem);
    if (e != cudaSuccess) return (int)e;
    k_stage_b_reg<N, MINB, SH><<<pb.blocks, pb.wpb * 32, pb.smem, st>>>(a, (double*)workspace, pb.hist_doubles);
    return (int)cudaGetLastError();
}

static bool reg_shfl() {  // pivot-row broadcast by warp shuffles (default, ~5 % faster on B200) or via shared memory
    const char* e = getenv("PD_STAGE_B_SHFL");
    return !(e && e[0] == '0');
}
#define PD_REG_DISPATCH(N_, CALL)                                                  \
    if (reg_shfl()) {                                                              \
        constexpr bool SH = true;                                                  \
        switch (reg_minb()) {                                                      \
            case 3: { constexpr int MB = 3; return CALL; }                         \
            case 5: { constexpr int MB = 5; return CALL; }                         \
            default: { constexpr int MB = 4; return CALL; }                        \
        }                                                                          \
    } else {                                                                       \
        constexpr bool SH = false;                                                 \
        switch (reg_minb()) {                                                      \
            case 3: { constexpr int MB = 3; return CALL; }                         \
            case 5: { constexpr int MB = 5; return CALL; }                         \
            default: { constexpr int MB = 4; return CALL; }                        \
        }                                                                          \
    }
static StageBPlan plan_reg_any(int B, int NF, int N, int L) {
    if (N == 4) { PD_REG_DISPATCH(4, (plan_reg<4, MB, SH>(B, NF, L))) }
    PD_REG_DISPATCH(8, (plan_reg<8, MB, SH>(B, NF, L)))
}
static int launch_reg_any(const PdStageB& a, const StageBPlan& pb, void* workspace, cudaStream_t st) {
    if (a.N == 4) { PD_REG_DISPATCH(4, (launch_reg<4, MB, SH>(a, pb, workspace, st))) }
    PD_REG_DISPATCH(8, (launch_reg<8, MB, SH>(a, pb, workspace, st)))
}

// three-rows-per-lane register kernel (N = 4, 8; default): N lanes per system, 32/N systems per warp
static bool use_r3(int N) {
    if (!use_reg(N)) return false;
    if (const char* e = getenv("PD_STAGE_B_ROW1"))
        if (e[0] == '1') return false;
    return true;
}

#ifndef PD_R3_MINB4
#define PD_R3_MINB4 5  // resident CTAs (of two warps) per SM the N = 4 register kernel is compiled for
#endif
template <int N>
__global__ void __launch_bounds__(64, (N == 8) ? 4 : PD_R3_MINB4) k_stage_b_r3(PdStageB a, double* hist, long hist_doubles) {
    extern __shared__ double smem[];
    const int SD = PdStageBRow3<N>::smem_doubles(a.L);
    constexpr int GPW = 32 / N;  // systems per warp
    const int gpb = blockDim.x / N, gi = threadIdx.x / N;
    const long slot = (long)blockIdx.x * gpb + gi;
    const long nslots = (long)gridDim.x * gpb;
    SubWarp<N> g;
    double* sm = smem + (long)gi * SD;
    double* h = hist + slot * hist_doubles;
    const long nsys = (long)a.B * a.NF;
    // warp-uniform trip count: the groups of a warp whose slot runs past the end repeat the last system, stores off
    for (long s0 = slot - (gi % GPW); s0 < nsys; s0 += nslots) {
        const long s = s0 + (gi % GPW);
        const bool store = s < nsys;
        const long se = store ? s : nsys - 1;
        pd_stage_b_row3<N>(g, a, (int)(se / a.NF), (int)(se % a.NF), store, sm, h);
    }
}

template <int N>
static StageBPlan plan_r3(int B, int NF, int L) {
    StageBPlan p;
    p.sys_doubles = PdStageBRow3<N>::smem_doubles(L);
    p.wpb = 2;
    const int gpb = p.wpb * (32 / N);
    p.smem = (size_t)p.sys_doubles * 8 * gpb;
    int ctas_per_sm = (int)(PD_SMEM_BUDGET / p.smem);
    if (ctas_per_sm * p.wpb > 32) ctas_per_sm = 32 / p.wpb;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    int occ = 0;
    if (cudaFuncSetAttribute(k_stage_b_r3<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_stage_b_r3<N>, p.wpb * 32, p.smem) == cudaSuccess && occ > 0) {
        if (occ < ctas_per_sm) ctas_per_sm = occ;
    } else {
        cudaGetLastError();
    }
    const long nsys = (long)B * NF;
    long blocks = (nsys + gpb - 1) / gpb;
    if (blocks > (long)PD_NUM_SMS * ctas_per_sm) blocks = (long)PD_NUM_SMS * ctas_per_sm;
    p.blocks = (int)blocks;
    p.slots = blocks * gpb;
    p.hist_doubles = (long)L * PdStageBRow3<N>::HIST_PER_LAYER;
    return p;
}

template <int N>
static int launch_r3(const PdStageB& a, const StageBPlan& pb, void* workspace, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(k_stage_b_r3<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pb.smem);
    if (e != cudaSuccess) return (int)e;
    k_stage_b_r3<N><<<pb.blocks, pb.wpb * 32, pb.smem, st>>>(a, (double*)workspace, pb.hist_doubles);
    return (int)cudaGetLastError();
}

// tensor-core kernel (N = 8, 16; default): one warp per system, panel in DMMA accumulator tiles
static bool use_mma(int N) {
    if (N != 8 && N != 16) return false;
    if (const char* e = getenv("PD_STAGE_B_MMA"))
        if (e[0] == '0') return false;
    if (const char* e = getenv("PD_STAGE_B_SMEM"))
        if (e[0] == '1') return false;
    if (const char* e = getenv("PD_STAGE_B_GENERIC"))
        if (e[0] == '1') return false;
    return true;
}

#ifndef PD_MMA_MINB
#define PD_MMA_MINB 4  // resident CTAs per SM the N = 8 tensor-core kernel is compiled for
#endif
template <int N>
__global__ void __launch_bounds__(128, (N == 8) ? PD_MMA_MINB : 2) k_stage_b_mma(PdStageB a, double* hist, long hist_doubles) {
    extern __shared__ double smem[];
    const int SD = PdStageBMma<N>::smem_doubles(a.L);
    const int wpb = blockDim.x >> 5, w = threadIdx.x >> 5;
    const long slot = (long)blockIdx.x * wpb + w;
    const long nslots = (long)gridDim.x * wpb;
    const long nsys = (long)a.B * a.NF;
    for (long s = slot; s < nsys; s += nslots)
        pd_stage_b_mma<N>(a, (int)(s / a.NF), (int)(s % a.NF), smem + (long)w * SD, hist + slot * hist_doubles);
}

template <int N>
static StageBPlan plan_mma(int B, int NF, int L) {
    StageBPlan p;
    p.sys_doubles = PdStageBMma<N>::smem_doubles(L);
    p.wpb = 4;
    p.smem = (size_t)p.sys_doubles * 8 * p.wpb;
    int ctas_per_sm = (int)((PD_SMEM_MAX_CTA + 1024) / (p.smem + 1024));  // refined by the occupancy query below
    if (ctas_per_sm * p.wpb > 32) ctas_per_sm = 32 / p.wpb;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    int occ = 0;
    if (cudaFuncSetAttribute(k_stage_b_mma<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_stage_b_mma<N>, p.wpb * 32, p.smem) == cudaSuccess && occ > 0) {
        if (occ < ctas_per_sm) ctas_per_sm = occ;
    } else {
        cudaGetLastError();
    }
    if (const char* e = getenv("PD_STAGE_B_CTAS")) {
        const int v = atoi(e);
        if (v >= 1 && v < ctas_per_sm) ctas_per_sm = v;
    }
    const long nsys = (long)B * NF;
    long blocks = (nsys + p.wpb - 1) / p.wpb;
    if (blocks > (long)PD_NUM_SMS * ctas_per_sm) blocks = (long)PD_NUM_SMS * ctas_per_sm;
    p.blocks = (int)blocks;
    p.slots = blocks * p.wpb;
    p.hist_doubles = PdStageBMma<N>::scratch_doubles(L);
    return p;
}

template <int N>
static int launch_mma(const PdStageB& a, const StageBPlan& pb, void* workspace, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(k_stage_b_mma<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pb.smem);
    if (e != cudaSuccess) return (int)e;
    k_stage_b_mma<N><<<pb.blocks, pb.wpb * 32, pb.smem, st>>>(a, (double*)workspace, pb.hist_doubles);
    return (int)cudaGetLastError();
}

template <int N>
static StageBPlan plan_fast(int B, int NF, int L, int ls) {
    StageBPlan p;
    p.sys_doubles = (PdStageBFast<N>::SMEM_DOUBLES + 1) & ~1;
    p.wpb = 4;
    const int gpw = 32 / ls, gpb = p.wpb * gpw;
    p.smem = (size_t)p.sys_doubles * 8 * gpb;
    int ctas_per_sm = (int)(PD_SMEM_BUDGET / p.smem);
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm * p.wpb > 32) ctas_per_sm = 32 / p.wpb;
    {
        int occ = 0;
        if (N == 4 && ls == 8) occ = fast_ctas_per_sm<4, 8>(p.smem, p.wpb * 32);
        if (N == 4 && ls == 16) occ = fast_ctas_per_sm<4, 16>(p.smem, p.wpb * 32);
        if (N == 8 && ls == 16) occ = fast_ctas_per_sm<8, 16>(p.smem, p.wpb * 32);
        if (N == 8 && ls == 32) occ = fast_ctas_per_sm<8, 32>(p.smem, p.wpb * 32);
        if (N == 16 && ls == 32) occ = fast_ctas_per_sm<16, 32>(p.smem, p.wpb * 32);
        if (occ > 0 && occ < ctas_per_sm) ctas_per_sm = occ;
    }
    const long nsys = (long)B * NF;
    long blocks = (nsys + gpb - 1) / gpb;
    if (blocks > (long)PD_NUM_SMS * ctas_per_sm) blocks = (long)PD_NUM_SMS * ctas_per_sm;
    p.blocks = (int)blocks;
    p.slots = blocks * gpb;
    p.hist_doubles = (long)L * PdStageBFast<N>::HIST_PER_LAYER;
    return p;
}

StageBPlan pd_plan_stage_b(int B, int NF, int N, int L) {
    if (use_mma(N)) return (N == 8) ? plan_mma<8>(B, NF, L) : plan_mma<16>(B, NF, L);
    if (use_r3(N)) return (N == 4) ? plan_r3<4>(B, NF, L) : plan_r3<8>(B, NF, L);
    if (use_reg(N)) return plan_reg_any(B, NF, N, L);
    if (const int ls = fast_lanes(N)) {
        if (N == 4) return plan_fast<4>(B, NF, L, ls);
        if (N == 8) return plan_fast<8>(B, NF, L, ls);
        return plan_fast<16>(B, NF, L, ls);
    }
    StageBPlan p;
    p.sys_doubles = (pd_stage_b_doubles(N) + 1) & ~1;
    const size_t per_warp = (size_t)p.sys_doubles * 8;
    int wpb = 4;
    while (wpb > 1 && per_warp * wpb > 64 * 1024) wpb >>= 1;
    p.wpb = wpb;
    p.smem = per_warp * wpb;
    int ctas_per_sm = (int)(PD_SMEM_BUDGET / p.smem);
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm * wpb > 24) ctas_per_sm = 24 / wpb;
    const long nsys = (long)B * NF;
    long blocks = (nsys + wpb - 1) / wpb;
    if (blocks > (long)PD_NUM_SMS * ctas_per_sm) blocks = (long)PD_NUM_SMS * ctas_per_sm;
    p.blocks = (int)blocks;
    p.slots = blocks * wpb;
    p.hist_doubles = pd_stage_b_history_doubles(N, L);
    return p;
}

template <int NC>
static int launch_b(const PdStageB& a, const StageBPlan& pb, void* workspace, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(k_stage_b<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pb.smem);
    if (e != cudaSuccess) return (int)e;
    k_stage_b<NC><<<pb.blocks, pb.wpb * 32, pb.smem, st>>>(a, (double*)workspace, pb.hist_doubles, pb.sys_doubles);
    return (int)cudaGetLastError();
}

template <int N, int LS>
static int launch_fast(const PdStageB& a, const StageBPlan& pb, void* workspace, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(k_stage_b_fast<N, LS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pb.smem);
    if (e != cudaSuccess) return (int)e;
    k_stage_b_fast<N, LS><<<pb.blocks, pb.wpb * 32, pb.smem, st>>>(a, (double*)workspace, pb.hist_doubles);
    return (int)cudaGetLastError();
}

int pd_launch_stage_b(const PdStageB& a, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    const StageBPlan pb = pd_plan_stage_b(a.B, a.NF, a.N, a.L);
    if (workspace_bytes < (size_t)pb.slots * pb.hist_doubles * 8 || !workspace) return -20;
    if (pb.smem > PD_SMEM_MAX_CTA) return -21;
    if (use_mma(a.N)) return (a.N == 8) ? launch_mma<8>(a, pb, workspace, st) : launch_mma<16>(a, pb, workspace, st);
    if (use_r3(a.N)) return (a.N == 4) ? launch_r3<4>(a, pb, workspace, st) : launch_r3<8>(a, pb, workspace, st);
    if (use_reg(a.N)) return launch_reg_any(a, pb, workspace, st);
    if (const int ls = fast_lanes(a.N)) {
        const int key = a.N * 100 + ls;
        switch (key) {
            case 408: return launch_fast<4, 8>(a, pb, workspace, st);
            case 416: return launch_fast<4, 16>(a, pb, workspace, st);
            case 816: return launch_fast<8, 16>(a, pb, workspace, st);
            case 832: return launch_fast<8, 32>(a, pb, workspace, st);
            case 1632: return launch_fast<16, 32>(a, pb, workspace, st);
            default: return -23;
        }
    }
    switch (a.N) {
        case 4: return launch_b<4>(a, pb, workspace, st);
        case 8: return launch_b<8>(a, pb, workspace, st);
        case 16: return launch_b<16>(a, pb, workspace, st);
        default: return launch_b<0>(a, pb, workspace, st);
    }
}
