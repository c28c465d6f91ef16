// Host enqueue cost of gkq_integrate_batch (device entry) without Python: CPU us per call vs GPU us per call.
#include <chrono>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/gkquad_b200.h"
int main() {
    gkq_handle_t h;
    if (gkq_create(0, &h)) { printf("create failed: %s\n", gkq_last_error(nullptr)); return 1; }
    const int64_t n = 1 << 20;
    std::vector<double> p(2 * n);
    for (int64_t i = 0; i < n; ++i) { p[i] = 0.5 + 1.5 * (i % 1000) / 1000.0; p[n + i] = 1 + 9.0 * ((i * 7) % 1000) / 1000.0; }
    double *d_p, *d_ab, *d_e, *d_d; uint32_t *d_n, *d_s;
    cudaMalloc(&d_p, 16 * n); cudaMalloc(&d_ab, 16); cudaMalloc(&d_e, 8 * n); cudaMalloc(&d_d, 8 * n);
    cudaMalloc(&d_n, 4 * n); cudaMalloc(&d_s, 4 * n);
    cudaMemcpy(d_p, p.data(), 16 * n, cudaMemcpyHostToDevice);
    double ab[2] = {0.0, 1.0};
    cudaMemcpy(d_ab, ab, 16, cudaMemcpyHostToDevice);
    gkq_config c; gkq_config_default(1, &c);
    c.algo = GKQ_ALGO_QAGS; c.tol.kind = GKQ_TOL_RELATIVE; c.tol.rel_tol = 1e-10;
    int fid = gkq_functor_lookup(1, "expcos_p");
    cudaStream_t st; cudaStreamCreate(&st);
    for (int mode = 0; mode < 2; ++mode) {
        gkq_set_deferred_join(h, mode);
        for (int k = 0; k < 12; ++k) gkq_integrate_batch(h, &c, fid, n, d_ab, d_ab + 1, 0, d_p, n, nullptr, nullptr, d_e, d_d, d_n, d_s, st);
        gkq_sync(h, st);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int K = 40;
        cudaEventRecord(e0, st);
        auto t0 = std::chrono::steady_clock::now();
        for (int k = 0; k < K; ++k) gkq_integrate_batch(h, &c, fid, n, d_ab, d_ab + 1, 0, d_p, n, nullptr, nullptr, d_e, d_d, d_n, d_s, st);
        gkq_join(h, st);
        auto t1 = std::chrono::steady_clock::now();
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("deferred=%d: GPU %.1f us/step, host enqueue %.1f us/step\n", mode, ms * 1e3 / K,
               std::chrono::duration<double, std::micro>(t1 - t0).count() / K);
        gkq_sync(h, st);
    }
    gkq_set_deferred_join(h, 0);
    gkq_destroy(h);
    return 0;
}
