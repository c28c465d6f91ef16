// Micro-benchmark: what an FP64 instruction costs on the sm_100a FP64 pipe as a function of how
// many DISTINCT vector-register source operands it reads, with and without operand reuse, and
// whether the extra cost blocks the issue port (integer instructions mixed in).
// Every per-thread value comes from a per-thread global load so that ptxas cannot promote it to a
// uniform register; check the SASS (cuobjdump -sass) for the operand shapes named below.
// One block on one SM, W warps per scheduler (W = blockDim / 128); result = cycles per warp-level
// FP64 instruction per scheduler.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int C = 8;
template <int MODE>
__global__ void k(double* out, long long* cyc, int iters, double U, const double* g, int istep) {
    double x[C], r[C], s[C];
    const int t = threadIdx.x;
    for (int c = 0; c < C; ++c) {
        x[c] = g[t * 32 + c];
        r[c] = g[t * 32 + 8 + c];
        s[c] = g[t * 32 + 16 + c];
    }
    int y[C];
    for (int c = 0; c < C; ++c) y[c] = t + c;
    const double r0 = r[0], s0 = s[0];
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            if (MODE == 0) x[c] = fma(x[c], U, 0.5);                 // 1 vector operand
            if (MODE == 1) x[c] = fma(x[c], r[c], U);                // 2 distinct
            if (MODE == 2) x[c] = fma(x[c], r0, U);                  // 2, one shared by all 8 (reuse)
            if (MODE == 3) x[c] = fma(x[c], r[c], s[c]);             // 3 distinct
            if (MODE == 4) x[c] = fma(x[c], r0, s0);                 // 3, two shared (reuse)
            if (MODE == 5) x[c] = x[c] + r[c];                       // DADD 2 distinct
            if (MODE == 6) x[c] = x[c] * r[c];                       // DMUL 2 distinct
            if (MODE == 7) x[c] = x[c] * x[c];                       // DMUL same register twice
            if (MODE == 8) { x[c] = fma(x[c], r[c], U); y[c] = y[c] * istep + c; }                       // 2 distinct + 1 int
            if (MODE == 9) { x[c] = fma(x[c], r[c], U); y[c] = y[c] * istep + c; y[c] ^= (y[c] >> 3); }  // 2 distinct + 2..3 int
            if (MODE == 10) { x[c] = fma(x[c], U, 0.5); y[c] = y[c] * istep + c; }                       // 1 + 1 int
            if (MODE == 11) { x[c] = fma(x[c], U, 0.5); y[c] = y[c] * istep + c; y[c] ^= (y[c] >> 3); }  // 1 + 2..3 int
            if (MODE == 12) x[c] = fma(x[c], r[c >> 1], U);          // 2, shared by adjacent couples
            if (MODE == 13) x[c] = fma(x[c], r[c & 1], U);           // 2, alternating between two
            if (MODE == 14) x[c] = x[c] + U;                         // DADD 1 vector
            if (MODE == 15) x[c] = fma(r[c], s[c], x[c]);            // 3 distinct, accumulator form
        }
    }
    long long t1 = clock64();
    double sum = 0;
    int ys = 0;
    for (int c = 0; c < C; ++c) { sum += x[c]; ys += y[c]; }
    out[blockIdx.x * blockDim.x + t] = sum + ys;
    if (t == 0) cyc[0] = t1 - t0;
}
int main() {
    double* out; long long* cyc; double* g;
    cudaMalloc(&out, 1 << 20); cudaMallocManaged(&cyc, 1024); cudaMallocManaged(&g, 1024 * 32 * 8);
    for (int i = 0; i < 1024 * 32; ++i) g[i] = 1.0 + 1e-9 * ((i * 7919) % 1000);
    const int iters = 2048;
    const char* names[16] = {"DFMA 1 vec", "DFMA 2 vec distinct", "DFMA 2 vec, one shared", "DFMA 3 vec distinct",
        "DFMA 3 vec, two shared", "DADD 2 vec distinct", "DMUL 2 vec distinct", "DMUL x*x", "DFMA 2 vec + 1 int",
        "DFMA 2 vec + 3 int", "DFMA 1 vec + 1 int", "DFMA 1 vec + 3 int", "DFMA 2 vec shared by couples",
        "DFMA 2 vec alternating two", "DADD 1 vec", "DFMA 3 vec accumulate"};
    printf("%-30s", "mode \\ warps per scheduler");
    for (int w = 1; w <= 8; w *= 2) printf("%8d", w);
    printf("\n");
#define RUN(MODE) { printf("%-30s", names[MODE]); for (int w = 1; w <= 8; w *= 2) { \
        k<MODE><<<1, 128 * w>>>(out, cyc, iters, 0.999999, g, 3); cudaDeviceSynchronize(); \
        k<MODE><<<1, 128 * w>>>(out, cyc, iters, 0.999999, g, 3); cudaDeviceSynchronize(); \
        printf("%8.2f", (double)cyc[0] / ((double)iters * C * w)); } printf("\n"); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13) RUN(14) RUN(15)
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
