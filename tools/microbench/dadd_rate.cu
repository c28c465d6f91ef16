// Micro-benchmark: DADD / DMUL / DFMA issue rate with 8 independent chains per thread, 8 warps per
// sub-partition (latency fully hidden), by operand source.
// Measured on B200: one vector-register source (others uniform / constant) -> 2.0 cycles per warp
// instruction; two or three vector-register sources as allocated HERE -> 2.9-3.2 cycles.  The
// engine's calibration kernel with the same fma(x, r, c) shape but a different register allocation
// runs at the full 2.0, so the penalty is register-bank conflicts between 64-bit operands, i.e.
// up to ptxas, not a property of the operand count.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, long long* cyc, int iters, double m_in, const double* gtab) {
    double x[8], r[8];
    for (int c = 0; c < 8; ++c) { x[c] = 1.0 + c + threadIdx.x; r[c] = gtab[c]; }
    const double m = m_in;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (MODE == 0) x[c] = x[c] + m;            // DADD reg, uniform
            else if (MODE == 1) x[c] = x[c] + r[c];    // DADD reg, reg
            else if (MODE == 2) x[c] = x[c] * m;       // DMUL reg, uniform
            else if (MODE == 3) x[c] = x[c] * r[c];    // DMUL reg, reg
            else if (MODE == 4) x[c] = fma(x[c], m, m);        // DFMA reg, uniform, uniform
            else if (MODE == 5) x[c] = fma(x[c], r[c], m);     // DFMA reg, reg, uniform
            else x[c] = fma(x[c], r[c], r[(c + 1) & 7]);       // DFMA reg, reg, reg
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int c = 0; c < 8; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double* out; long long* cyc; double* gtab;
    cudaMalloc(&out, 1 << 20); cudaMallocManaged(&cyc, 1024); cudaMallocManaged(&gtab, 64);
    for (int c = 0; c < 8; ++c) gtab[c] = 1.0 + 1e-9 * (c + 1);
    const int iters = 1024;
#define RUN(MODE, NAME) k<MODE><<<1, 1024>>>(out, cyc, iters, 0.999999, gtab); cudaDeviceSynchronize(); \
    printf("%-28s %.2f cycles per warp instruction per SMSP\n", NAME, (double)cyc[0] / (8.0 * iters * 8));
    RUN(0, "DADD reg,uniform") RUN(1, "DADD reg,reg") RUN(2, "DMUL reg,uniform") RUN(3, "DMUL reg,reg")
    RUN(4, "DFMA reg,uniform,uniform") RUN(5, "DFMA reg,reg,uniform") RUN(6, "DFMA reg,reg,reg")
    return 0;
}
