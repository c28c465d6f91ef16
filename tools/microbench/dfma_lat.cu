// Micro-benchmark: FP64 pipe latency / throughput on one SM sub-partition.
//   (a) one warp, C independent DFMA chains  -> dependent-issue latency and single-warp rate
//   (b) W warps on one SMSP (block of 4*W warps), 1 chain each -> how many warps saturate the pipe
#include <cstdio>
#include <cuda_runtime.h>
template <int C>
__global__ void chains(double* out, long long* cyc, int iters) {
    double x[C];
    for (int c = 0; c < C; ++c) x[c] = 1.0 + c + threadIdx.x;
    const double m = 0.999999, a = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < C; ++c) x[c] = fma(x[c], m, a);
    }
    long long t1 = clock64();
    double s = 0;
    for (int c = 0; c < C; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMallocManaged(&cyc, 1024);
    const int iters = 4096;
#define RUN(C, THREADS) chains<C><<<1, THREADS>>>(out, cyc, iters); cudaDeviceSynchronize(); \
    printf("threads %4d (warps/SMSP %2d) chains/thread %d: %.2f cycles per DFMA per warp, SMSP rate %.3f DFMA/cycle\n", THREADS, (THREADS/32+3)/4, C, \
           (double)cyc[0] / iters / C, (double)C * iters * ((THREADS/32+3)/4) / cyc[0]);
    RUN(1, 32) RUN(2, 32) RUN(4, 32) RUN(8, 32)
    RUN(1, 128) RUN(1, 256) RUN(1, 512) RUN(1, 1024) RUN(2, 512) RUN(2, 1024) RUN(4, 512) RUN(4,1024)
    return 0;
}
