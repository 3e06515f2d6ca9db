// micro-benchmark: FP64 tensor-core (mma.sync.m8n8k4.f64) throughput and latency against vector DFMA on sm_100a.
// Decides the north-star question "FP64 DMMA only if it pays" with a measurement.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_rate dmma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// MODE 0: DMMA, NCH independent accumulator chains per warp; MODE 1: DFMA with 2*NCH independent chains
template <int MODE, int NCH>
__global__ void k(double* out, int iters, long long* cyc) {
    double c[NCH][2];
#pragma unroll
    for (int i = 0; i < NCH; ++i) { c[i][0] = threadIdx.x + i; c[i][1] = threadIdx.x - i; }
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                if (MODE == 0) dmma(c[i][0], c[i][1], a, b);
                else { c[i][0] = fma(c[i][0], a, b); c[i][1] = fma(c[i][1], a, b); }
            }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int NCH>
void run(const char* name, int threads, double* out, long long* cyc) {
    const int iters = 2048;
    long long h = 0;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k<MODE, NCH><<<148 * 4, threads>>>(out, iters, cyc);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        cudaEventElapsedTime(&ms, e0, e1);
    }
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double winst = 4.0 * NCH * iters * (threads / 32) * (MODE == 0 ? 1 : 2);
    const double fma_per_inst = MODE == 0 ? 256.0 : 32.0;
    const double tflops = 148.0 * 4 * winst * fma_per_inst * 2 / (ms * 1e-3) / 1e12;
    printf("%-8s chains=%d warps/CTA=%2d: %7.2f TFLOP/s   %.1f cycles per dependent step (one CTA's clock)\n", name, NCH, threads / 32,
           tflops, (double)h / (4.0 * iters));
}

int main() {
    double* out; long long* cyc; cudaMalloc(&out, 148 * 4 * 1024 * 8); cudaMalloc(&cyc, 8);
    run<0, 1>("DMMA", 32, out, cyc);
    run<0, 1>("DMMA", 128, out, cyc);
    run<0, 4>("DMMA", 128, out, cyc);
    run<0, 8>("DMMA", 256, out, cyc);
    run<0, 8>("DMMA", 512, out, cyc);
    run<1, 1>("DFMA", 32, out, cyc);
    run<1, 4>("DFMA", 128, out, cyc);
    run<1, 8>("DFMA", 256, out, cyc);
    run<1, 8>("DFMA", 512, out, cyc);
    return 0;
}
