// Micro-benchmark: latency of a store -> dependent load chain through (a) a global slab with the
// engine's thread-interleaved layout, (b) a per-thread local array, (c) shared memory.
// Decides where the per-integral interval lists should live for the latency-bound tail.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void chain_global(double* slab, int stride, int iters, long long* out, double* sink) {
    double* d = slab + threadIdx.x;
    // warm: write 16 slots
    for (int k = 0; k < 16; ++k) d[(size_t)k * stride] = k + 1.0;
    __syncwarp();
    int idx = 0;
    double acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        double v = d[(size_t)idx * stride];      // dependent load
        acc += v;
        d[(size_t)((idx + 5) & 15) * stride] = acc;  // store to another slot
        idx = ((int)v + i) & 15;                  // next index depends on the loaded value
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0) / iters;
    sink[threadIdx.x] = acc;
}
__global__ void chain_local(int iters, long long* out, double* sink) {
    double d[16];
    for (int k = 0; k < 16; ++k) d[k] = k + 1.0;
    int idx = threadIdx.x & 15;
    double acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        double v = d[idx];
        acc += v;
        d[(idx + 5) & 15] = acc;
        idx = ((int)v + i) & 15;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[1] = (t1 - t0) / iters;
    sink[threadIdx.x] = acc;
}
__global__ void chain_shared(int iters, long long* out, double* sink) {
    __shared__ double sh[16 * 32];
    double* d = sh + threadIdx.x;
    for (int k = 0; k < 16; ++k) d[k * 32] = k + 1.0;
    int idx = 0;
    double acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        double v = d[idx * 32];
        acc += v;
        d[((idx + 5) & 15) * 32] = acc;
        idx = ((int)v + i) & 15;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[2] = (t1 - t0) / iters;
    sink[threadIdx.x] = acc;
}
int main() {
    double *slab, *sink;
    long long* out;
    cudaMalloc(&slab, 16 * 100000 * 8);
    cudaMalloc(&sink, 1024);
    cudaMallocManaged(&out, 64);
    for (int active = 32; active >= 1; active /= 32) {
        chain_global<<<1, active>>>(slab, 100000, 2000, out, sink);
        chain_local<<<1, active>>>(2000, out, sink);
        chain_shared<<<1, active>>>(2000, out, sink);
        cudaDeviceSynchronize();
        printf("active lanes %2d: cycles per store+dependent-load step: global slab %lld, local array %lld, shared %lld\n",
               active, out[0], out[1], out[2]);
    }
    // DFMA / DADD dependent latency
    return 0;
}
