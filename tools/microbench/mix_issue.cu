// Micro-benchmark: can an SM sub-partition keep the FP64 pipe busy while ~half of the
// instruction stream is integer / other work?  (the shape of the rule kernels: ~50 FP64 + ~54
// other instructions per integrand sample, profiles/S1_r01.md)
//   C independent DFMA chains per thread, M independent integer ops (IMAD) per DFMA,
//   W warps per sub-partition (block of 4*W warps on one SM).
// Prints FP64 pipe utilisation = DFMA issued / (cycles / 2) and the issue-slot utilisation.
#include <cstdio>
#include <cuda_runtime.h>
template <int C, int M>
__global__ void mix(double* out, long long* cyc, int iters) {
    double x[C];
    unsigned y[4];
    for (int c = 0; c < C; ++c) x[c] = 1.0 + c + threadIdx.x;
    for (int k = 0; k < 4; ++k) y[k] = threadIdx.x * 2654435761u + k;
    const double m = 0.999999, a = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            x[c] = fma(x[c], m, a);
#pragma unroll
            for (int k = 0; k < M; ++k) y[(c * M + k) & 3] = y[(c * M + k) & 3] * 1664525u + 1013904223u;
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int c = 0; c < C; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (double)(y[0] ^ y[1] ^ y[2] ^ y[3]);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMallocManaged(&cyc, 1024);
    const int iters = 2048;
#define RUN(C, M, THREADS) mix<C, M><<<1, THREADS>>>(out, cyc, iters); cudaDeviceSynchronize(); { \
    const double W = THREADS / 128.0; const double dfma = (double)C * iters * W; \
    printf("chains %d int/dfma %d warps/SMSP %4.1f: FP64 pipe %.2f issue slots %.2f\n", C, M, W, \
           dfma * 2.0 / cyc[0], dfma * (1 + M) / cyc[0]); }
    RUN(1, 0, 512) RUN(1, 0, 1024) RUN(2, 0, 512) RUN(2, 0, 1024) RUN(4, 0, 512)
    RUN(1, 1, 512) RUN(1, 1, 1024) RUN(2, 1, 512) RUN(2, 1, 1024) RUN(4, 1, 512) RUN(4, 1, 1024)
    RUN(1, 2, 1024) RUN(2, 2, 1024) RUN(4, 2, 1024)
    return 0;
}
