// pd_common.cuh -- shared definitions for the pydisort_b200 CUDA kernels.
//
// All numerical routines are written against a "lane group" policy: a group of
// G::size threads cooperates on one work item whose matrices live in shared
// memory; loops are strided by lane and `sync()` orders shared-memory traffic
// inside the group.  On the GPU a group is a power-of-two slice of a warp
// (SubWarp<LANES>); SerialGroup (size 1) runs the same code in one thread and
// is also what the host-side debug build under tests/hostsim compiles.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/pydisort_b200.h"

#if defined(__CUDACC__)
#define PD_HD __host__ __device__ __forceinline__
#else
#define PD_HD inline
#endif

#define PD_EPS 2.220446049250313e-16
#define PD_PI 3.141592653589793238462643383279502884

struct SerialGroup {
    static constexpr int size = 1;
    PD_HD int lane() const { return 0; }
    PD_HD void sync() const {}
    PD_HD bool any(bool v) const { return v; }
    PD_HD double shfl(double v, int) const { return v; }
};

#if defined(__CUDACC__)
template <int LANES>
struct SubWarp {
    static_assert(LANES >= 1 && LANES <= 32 && (LANES & (LANES - 1)) == 0, "LANES must be a power of two <= 32");
    static constexpr int size = LANES;
    unsigned mask;
    int ln;
    __device__ __forceinline__ SubWarp() {
        const int l = threadIdx.x & 31;
        ln = l & (LANES - 1);
        mask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << (l - ln));
    }
    __device__ __forceinline__ int lane() const { return ln; }
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
    __device__ __forceinline__ bool any(bool v) const { return (__ballot_sync(mask, v) & mask) != 0u; }
    // value of `v` held by lane `src` of this group
    __device__ __forceinline__ double shfl(double v, int src) const { return __shfl_sync(mask, v, src, LANES); }
};

// 16 bytes global -> shared without passing through registers (L2 only: the source may have been written by this kernel)
__device__ __forceinline__ void pd_cp_async16(double* dst_smem, const double* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void pd_cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
#endif

// Layout of G in HBM.  An item = the two N x N blocks [Gp | Gm] of one (column, mode, layer), 2 N^2 doubles, element
// e = block * N^2 + row * N + column; item index = (column * NFourier + mode) * L + layer.
//   N != 8: items one after the other, offset(item, e) = item * 2N^2 + e.
//   N == 8: groups of 32 consecutive items, sector-interleaved -- the 32-byte chunk c (elements 4c .. 4c+3) of the 32
//           items of a group lie next to each other,
//               offset(item, e) = (item / 32) * 32 * 2N^2  +  (e / 4) * 128  +  (item % 32) * 4  +  e % 4.
//           The one-thread-per-item kernels of this shape (the symmetric eigen stage, which writes G, and the layer-
//           operator kernel, which reads it) then move whole contiguous kilobytes per warp instruction instead of 32
//           sectors a kilobyte apart (eigen stage 80 -> 73 ms per SW step); lane-group readers still get full 32-byte
//           sectors.  For the other shapes the readers are lane groups or one thread per SYSTEM, which lose 5-15 % on
//           the scattered sectors (measured on LW and HA), so they keep the plain layout.
// The buffer holds ceil(items / 32) * 32 items in either case.
PD_HD bool pd_g_interleaved(int n) { return n == 8; }
PD_HD long pd_g_base(long item, int n) {
    return pd_g_interleaved(n) ? (item >> 5) * 32L * (2 * n * n) + (item & 31) * 4 : item * (2L * n * n);
}
PD_HD long pd_g_off(int e, int n) { return pd_g_interleaved(n) ? (long)(e >> 2) * 128 + (e & 3) : (long)e; }

// padded leading dimension: odd, so that row- and column-wise lane access are both bank-conflict free
PD_HD int pd_ld(int n) { return n | 1; }

// 1/x to within about an ulp: hardware seed (MUFU.RCP64H) + two Newton steps, no slow path
PD_HD double pd_rcp(double x) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
#else
    return 1.0 / x;
#endif
}

struct alignas(16) pd_d2 {  // two doubles moved with one 128-bit access
    double x, y;
};

// four consecutive doubles, 32-byte aligned, in one 256-bit store (sm_100: STG.256)
PD_HD void pd_store4(double* p, double a, double b, double c, double d) {
#if defined(__CUDA_ARCH__)
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
#else
    p[0] = a;
    p[1] = b;
    p[2] = c;
    p[3] = d;
#endif
}

// four consecutive doubles, 32-byte aligned, read once (streaming: not kept in L1) with one 256-bit load
PD_HD void pd_load4_stream(const double* p, double (&v)[4]) {
#if defined(__CUDA_ARCH__)
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3])
                 : "l"(p));
#else
    v[0] = p[0];
    v[1] = p[1];
    v[2] = p[2];
    v[3] = p[3];
#endif
}

// same width through the coherent path (data written earlier by this kernel)
PD_HD void pd_load4(const double* p, double (&v)[4]) {
#if defined(__CUDA_ARCH__)
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p) : "memory");
#else
    v[0] = p[0];
    v[1] = p[1];
    v[2] = p[2];
    v[3] = p[3];
#endif
}

// 1/sqrt(x), x > 0 finite: hardware seed (MUFU.RSQ64H) + two Newton steps (about 1 ulp)
PD_HD double pd_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    y = fma(y, fma(-hx * y, y, 0.5), y);
    y = fma(y, fma(-hx * y, y, 0.5), y);
    return y;
#else
    return 1.0 / sqrt(x);
#endif
}
