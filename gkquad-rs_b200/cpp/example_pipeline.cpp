// Throughput mode from compiled host code: BASELINE config 2 (workload S1: 1 M integrals of
// exp(-p0 x) cos(p1 x) on [0, 1], QAGS, Relative(1e-10)) driven through the C ABI with `depth`
// handles on `depth` CUDA streams, batch k on lane k % depth, every lane in deferred-join mode --
// the C++ form of the Python mirror's BatchPipeline (what rayon over cloned Integrators is for the
// reference).  Device-resident inputs and outputs; timed with CUDA events; every batch is checked
// against the closed form.  Usage: example_pipeline [depth=3] [n=1048576] [steps=240]
// Build: g++ -std=c++17 -O2 example_pipeline.cpp -I$CUDA/include -L../lib -lgkquad_b200 -L$CUDA/lib64 -lcudart
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/gkquad_b200.h"

#define CU(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                      \
            return 2;                                                                          \
        }                                                                                      \
    } while (0)
#define GK(h, x)                                                                               \
    do {                                                                                       \
        int rc_ = (x);                                                                         \
        if (rc_ != GKQ_OK) {                                                                   \
            std::fprintf(stderr, "%s: %d %s\n", #x, rc_, gkq_last_error(h));                   \
            return 3;                                                                          \
        }                                                                                      \
    } while (0)

// the counter-based parameter generator of SURVEY.md section 8d (gkquad_b200/workloads.py)
static uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static double uniform(int k, uint64_t i) {
    return (double)(splitmix64(0x9E3779B97F4A7C15ull + 2ull * i + (uint64_t)k) >> 11) * 0x1.0p-53;
}

struct Set {
    double *params, *est, *dlt;
    uint32_t *nev, *st;
};

int main(int argc, char** argv) {
    const int depth = argc > 1 ? std::atoi(argv[1]) : 3;
    const int64_t n = argc > 2 ? std::atoll(argv[2]) : (1 << 20);
    const int steps = argc > 3 ? std::atoi(argv[3]) : 240;
    const int NSETS = 8;  // 8 x (16 MB in + 24 MB out) > L2
    if (depth < 1 || depth > 8 || n < 1 || steps < 1) return 1;

    std::vector<gkq_handle_t> h(depth);
    std::vector<cudaStream_t> lane(depth);
    std::vector<cudaEvent_t> done(depth);
    for (int l = 0; l < depth; ++l) {
        GK(nullptr, gkq_create(0, &h[l]));
        GK(h[l], gkq_set_deferred_join(h[l], 1));
        CU(cudaStreamCreateWithFlags(&lane[l], cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&done[l], cudaEventDisableTiming));
    }
    const int fid = gkq_functor_lookup(1, "expcos_p");
    if (fid < 0) return 4;
    gkq_config cfg;
    gkq_config_default(1, &cfg);
    cfg.algo = GKQ_ALGO_QAGS;
    cfg.tol.kind = GKQ_TOL_RELATIVE;
    cfg.tol.abs_tol = 0.0;
    cfg.tol.rel_tol = 1e-10;
    cfg.max_evals = 2000;

    std::vector<Set> sets(NSETS);
    std::vector<double> hp(2 * (size_t)n);
    for (int s = 0; s < NSETS; ++s) {
        for (int64_t i = 0; i < n; ++i) {
            const uint64_t g = (uint64_t)s * (uint64_t)n + (uint64_t)i;
            hp[i] = 0.5 + 1.5 * uniform(0, g);
            hp[n + i] = 1.0 + 9.0 * uniform(1, g);
        }
        CU(cudaMalloc(&sets[s].params, 2 * n * sizeof(double)));
        CU(cudaMalloc(&sets[s].est, n * sizeof(double)));
        CU(cudaMalloc(&sets[s].dlt, n * sizeof(double)));
        CU(cudaMalloc(&sets[s].nev, n * sizeof(uint32_t)));
        CU(cudaMalloc(&sets[s].st, n * sizeof(uint32_t)));
        CU(cudaMemcpy(sets[s].params, hp.data(), 2 * n * sizeof(double), cudaMemcpyHostToDevice));
    }
    double* range = nullptr;  // shared range [0, 1] (range_stride 0)
    const double ab[2] = {0.0, 1.0};
    CU(cudaMalloc(&range, sizeof ab));
    CU(cudaMemcpy(range, ab, sizeof ab, cudaMemcpyHostToDevice));

    auto run = [&](int k) -> int {
        const Set& S = sets[k % NSETS];
        const int l = k % depth;
        return gkq_integrate_batch(h[l], &cfg, fid, n, range, range + 1, 0, S.params, n, nullptr, nullptr,
                                   S.est, S.dlt, S.nev, S.st, lane[l]);
    };
    auto join = [&](cudaStream_t consumer) -> int {
        for (int l = 0; l < depth; ++l) {
            int rc = gkq_join(h[l], lane[l]);
            if (rc) return rc;
            if (cudaEventRecord(done[l], lane[l]) != cudaSuccess) return GKQ_ERR_CUDA;
            if (cudaStreamWaitEvent(consumer, done[l], 0) != cudaSuccess) return GKQ_ERR_CUDA;
        }
        return GKQ_OK;
    };

    cudaStream_t mainS;
    CU(cudaStreamCreateWithFlags(&mainS, cudaStreamNonBlocking));
    for (int k = 0; k < 18 * depth; ++k) GK(h[k % depth], run(k));  // warm-up: scratch at steady state
    GK(h[0], join(mainS));
    CU(cudaStreamSynchronize(mainS));

    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    CU(cudaEventRecord(e0, mainS));
    for (int l = 0; l < depth; ++l) CU(cudaStreamWaitEvent(lane[l], e0, 0));
    for (int k = 0; k < steps; ++k) GK(h[k % depth], run(k));
    GK(h[0], join(mainS));
    CU(cudaEventRecord(e1, mainS));
    CU(cudaEventSynchronize(e1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));

    // check every set against the closed form
    std::vector<double> est(n);
    std::vector<uint32_t> st(n);
    long long converged = 0, wrong = 0;
    for (int s = 0; s < NSETS; ++s) {
        CU(cudaMemcpy(est.data(), sets[s].est, n * sizeof(double), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(st.data(), sets[s].st, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < n; ++i) {
            if (st[i] != GKQ_STATUS_NONE) continue;
            ++converged;
            const uint64_t g = (uint64_t)s * (uint64_t)n + (uint64_t)i;
            const double p0 = 0.5 + 1.5 * uniform(0, g), p1 = 1.0 + 9.0 * uniform(1, g);
            const double truth = (p0 * (1 - std::exp(-p0) * std::cos(p1)) + p1 * std::exp(-p0) * std::sin(p1)) / (p0 * p0 + p1 * p1);
            if (!(std::fabs(est[i] - truth) <= 1e-10 * std::fabs(truth) + 1e-15)) ++wrong;
        }
    }
    const double per_step = (double)converged / NSETS;
    std::printf("{\"example\": \"pipeline\", \"depth\": %d, \"batch\": %lld, \"steps\": %d, \"ms_per_step\": %.5f, "
                "\"converged_integrals_per_s\": %.4g, \"converged_fraction\": %.6f, \"outside_tolerance_of_closed_form\": %lld}\n",
                depth, (long long)n, steps, ms / steps, per_step / (ms / steps * 1e-3), per_step / (double)n, wrong);
    for (int l = 0; l < depth; ++l) {
        gkq_set_deferred_join(h[l], 0);
        gkq_destroy(h[l]);
    }
    return wrong == 0 && converged > 0 ? 0 : 5;
}
