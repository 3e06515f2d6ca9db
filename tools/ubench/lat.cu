// lat.cu -- dependent-chain latencies (cycles per op) of the instructions on stage B's critical path
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* t, double x0, int n) {
    __shared__ double sm[64];
    sm[threadIdx.x & 63] = x0;
    __syncthreads();
    double x = x0 + threadIdx.x * 1e-9, y = 1.0000001;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = fma(x, y, 1e-9);
    long long t1 = clock64();
    unsigned u = threadIdx.x;
#pragma unroll 16
    for (int i = 0; i < n; ++i) u = __shfl_xor_sync(0xffffffffu, u, 1) + 1;
    long long t2 = clock64();
    double r = x;
#pragma unroll 16
    for (int i = 0; i < n; ++i) {
        double q;
        asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(r));
        r = q;
    }
    long long t3 = clock64();
    int idx = threadIdx.x & 63;
#pragma unroll 16
    for (int i = 0; i < n; ++i) idx = (int)sm[idx & 63] & 63;
    long long t4 = clock64();
    unsigned v = u;
#pragma unroll 16
    for (int i = 0; i < n; ++i) v = max(v ^ 0x5u, v + 3u);
    long long t5 = clock64();
    if (threadIdx.x == 0) { t[0] = t1 - t0; t[1] = t2 - t1; t[2] = t3 - t2; t[3] = t4 - t3; t[4] = t5 - t4; }
    out[threadIdx.x] = x + u + r + idx + v;
}
int main() {
    double* o; long long* t; cudaMalloc(&o, 1024 * 8); cudaMalloc(&t, 64);
    const int n = 4096;
    for (int w = 1; w <= 8; w *= 2) {
        k<<<1, 32 * w>>>(o, t, 1.0, n); cudaDeviceSynchronize();
        k<<<1, 32 * w>>>(o, t, 1.0, n); cudaDeviceSynchronize();
        long long h[5]; cudaMemcpy(h, t, 40, cudaMemcpyDeviceToHost);
        printf("warps %d: DFMA %.1f  SHFL+IADD %.1f  RCP64H %.1f  LDS+cvt %.1f  IMNMX-chain(2 ops) %.1f cyc\n", w, h[0] / (double)n, h[1] / (double)n,
               h[2] / (double)n, h[3] / (double)n, h[4] / (double)n);
    }
    return 0;
}
