// pd_launch.h -- private: host-side launchers shared by the translation units of libpydisort_b200.so
#pragma once
#include <cuda_runtime.h>

#include "pd_eval.cuh"
#include "pd_prologue.cuh"
#include "pd_stage_a.cuh"
#include "pd_stage_b.cuh"

#define PD_NUM_SMS 148               // B200
#define PD_SMEM_BUDGET (200 * 1024)  // per-SM shared memory we plan residency against
#define PD_SMEM_MAX_CTA (227 * 1024)
#define PD_WS_HEAD 256               // first bytes of the caller's workspace: counter of eigen items handed to the general kernel

static inline int pd_lanes_for(int n) {
    int l = 1;
    while (l < n && l < 32) l <<= 1;
    return l;
}
static inline cudaStream_t pd_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Stage B launch plan and workspace layout (after the PD_WS_HEAD bytes):
// [system flags int32 + counter | layer operators R^, T^ (N = 8) | history of the first kernel | history of k_stage_b]
struct StageBPlan {
    int add;                 // 1: k_stage_b_add (2: k_stage_b_tps) runs first, k_stage_b only redoes the systems it flags
    int split;               // 1: k_layer_ops forms R^, T^ of every layer first, the sweep only consumes them (N = 8)
    size_t rt_doubles;       // B * NF * L * N (N + 1) if split
    int add_blocks;
    long add_slots, add_hist;  // resident systems and history doubles per system of that first kernel
    size_t flag_bytes;
    int wpb, sys_doubles, blocks;  // k_stage_b: warps per CTA, shared doubles per system, grid
    size_t smem;
    long hist_doubles, slots;
    size_t bytes() const { return flag_bytes + (rt_doubles + (size_t)add_slots * add_hist + (size_t)slots * hist_doubles) * 8; }
};
StageBPlan pd_plan_stage_b(int B, int NF, int N, int L, int flags);
int pd_check_cfg(const pd_config* c);

// return a cudaError_t (0 = ok) or a negative argument error
int pd_launch_stage_a(const PdStageA& a, int flags, const double* ptab, cudaStream_t st);
int pd_launch_stage_b(const PdStageB& a, int flags, void* workspace, size_t workspace_bytes, cudaStream_t st);

#define PD_DISPATCH_LANES(lanes, ...)                           \
    switch (lanes) {                                            \
        case 1: { constexpr int LN = 1; __VA_ARGS__; } break;   \
        case 2: { constexpr int LN = 2; __VA_ARGS__; } break;   \
        case 4: { constexpr int LN = 4; __VA_ARGS__; } break;   \
        case 8: { constexpr int LN = 8; __VA_ARGS__; } break;   \
        case 16: { constexpr int LN = 16; __VA_ARGS__; } break; \
        default: { constexpr int LN = 32; __VA_ARGS__; } break; \
    }

// lane count AND (for the production sizes N = 4, 8, 16) a compile-time N
#define PD_DISPATCH_N(N, ...)                                                        \
    switch (N) {                                                                     \
        case 4: { constexpr int LN = 4, NC = 4; __VA_ARGS__; } break;                \
        case 8: { constexpr int LN = 8, NC = 8; __VA_ARGS__; } break;                \
        case 16: { constexpr int LN = 16, NC = 16; __VA_ARGS__; } break;             \
        default: { constexpr int NC = 0; PD_DISPATCH_LANES(pd_lanes_for(N), __VA_ARGS__); } break; \
    }
