// micro-benchmark: per-SM throughput of SHFL.IDX (32-bit), 64-bit LDS broadcast, REDUX and DFMA
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, int iters, long long* cyc) {
    __shared__ __align__(16) double s[1024];
    s[threadIdx.x] = threadIdx.x;
    __syncthreads();
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    unsigned u0 = threadIdx.x, u1 = u0 * 3;
    int src = threadIdx.x & 7;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) { u0 = __shfl_sync(0xffffffffu, u0, src); u1 = __shfl_sync(0xffffffffu, u1, src ^ 1); }
        } else if (MODE == 1) {
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) { a0 += s[(i + k2) & 1023]; a1 += s[(i + k2 + 7) & 1023]; }
        } else if (MODE == 2) {
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) { u0 = __reduce_max_sync(0xffffffffu, u0) + k2; u1 = __reduce_max_sync(0xffffffffu, u1) + k2; }
        } else if (MODE == 4) {  // SHFL and LDS.64 broadcast interleaved: do they share a pipe?
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) { u0 = __shfl_sync(0xffffffffu, u0, src); a0 += s[(i + k2) & 1023]; }
        } else if (MODE == 5) {  // LDS.128 broadcast
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) { const double2 v = *reinterpret_cast<const double2*>(&s[((i + k2) * 2) & 1022]); a0 += v.x; a1 += v.y; const double2 w = *reinterpret_cast<const double2*>(&s[((i + k2) * 2 + 14) & 1022]); a2 += w.x; a3 += w.y; }
        } else if (MODE == 6) {  // STS.64 by a single lane
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) { if ((threadIdx.x & 31) == 3) s[(threadIdx.x + k2 * 2 + i) & 1023] = a0; }
        } else {
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) { a0 = fma(a0, 1.0000001, 0.5); a1 = fma(a1, 1.0000001, 0.5); a2 = fma(a2, 1.0000001, 0.5); a3 = fma(a3, 1.0000001, 0.5); }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + u0 + u1;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    const int iters = 4096, threads = 1024;  // one CTA of 32 warps per SM
    const char* names[] = {"SHFL.IDX 32-bit", "LDS.64 broadcast", "REDUX.MAX", "DFMA", "SHFL+LDS.64 mixed (8+8)", "LDS.128 broadcast", "STS.64 one lane"};
    for (int mode = 0; mode < 7; ++mode) {
        long long h = 0;
        for (int rep = 0; rep < 2; ++rep) {
            if (mode == 0) k<0><<<148, threads>>>(out, iters, cyc);
            if (mode == 1) k<1><<<148, threads>>>(out, iters, cyc);
            if (mode == 2) k<2><<<148, threads>>>(out, iters, cyc);
            if (mode == 3) k<3><<<148, threads>>>(out, iters, cyc);
            if (mode == 4) k<4><<<148, threads>>>(out, iters, cyc);
            if (mode == 5) k<5><<<148, threads>>>(out, iters, cyc);
            if (mode == 6) k<6><<<148, threads>>>(out, iters, cyc);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        const double winst = 16.0 * iters * (threads / 32);  // warp-instructions of the measured kind per SM
        printf("%-18s %8.3f warp-instr/clk/SM  (%lld cycles)\n", names[mode], winst / (double)h, h);
    }
    return 0;
}
