// Multi-GPU over the C ABI alone (no Python, no torch): `nranks` host threads, one handle and one
// GPU each, a static block-cyclic shard of one batch, and ONE gkq_gather of the results to rank 0
// -- what a Rust / C++ host of gkquad would bind to spread Algorithm::integrate
// (single/algorithm/mod.rs:16-23) or Algorithm2::integrate (double/algorithm/mod.rs:7-10) over the
// GPUs of a box.  Checks, on rank 0, that the gathered arrays are bit-identical to the same batch
// integrated unsharded on one GPU.
//
// Usage: example_two_ranks [nranks=min(2, #GPUs)] [n_total=300000] [workload: 1 = QAGS 1-D (S2), 2 = QAGP2 (D0)]
// Exit code 0 = identical; prints one line per check.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/gkquad_b200.h"

static uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static double uniform(int k, uint64_t i) {
    return (double)(splitmix64(0x9E3779B97F4A7C15ull + 2ull * i + (uint64_t)k) >> 11) * 0x1.0p-53;
}

struct Results {
    std::vector<double> est, dlt;
    std::vector<uint32_t> nev, st;
};

static const int64_t BLOCKSZ = 4096;

// integrate the integrals `idx` (global numbers) on `device`; results stay on the device
struct DeviceBatch {
    double *p = nullptr, *est = nullptr, *dlt = nullptr;
    uint32_t *nev = nullptr, *st = nullptr;
    int64_t n = 0;
};

static int run_shard(gkq_handle_t h, int workload, const std::vector<int64_t>& idx, DeviceBatch& B, cudaStream_t s) {
    const int64_t n = (int64_t)idx.size();
    B.n = n;
    std::vector<double> p((size_t)n);
    for (int64_t j = 0; j < n; ++j)
        p[(size_t)j] = workload == 1 ? 1.0 + 19.0 * uniform(0, (uint64_t)idx[(size_t)j])   // S2: 1/(1+p0 x^2)
                                     : 0.5 + 1.5 * uniform(0, (uint64_t)idx[(size_t)j]);  // D0: p0/(x y)
    if (cudaMalloc(&B.p, (size_t)(n ? n : 1) * 8) || cudaMalloc(&B.est, (size_t)(n ? n : 1) * 8) ||
        cudaMalloc(&B.dlt, (size_t)(n ? n : 1) * 8) || cudaMalloc(&B.nev, (size_t)(n ? n : 1) * 4) ||
        cudaMalloc(&B.st, (size_t)(n ? n : 1) * 4))
        return 2;
    if (n == 0) return 0;
    if (cudaMemcpyAsync(B.p, p.data(), (size_t)n * 8, cudaMemcpyHostToDevice, s)) return 2;
    if (cudaStreamSynchronize(s)) return 2;  // `p` leaves scope
    double r[4] = {0.0, 1.0, -1.0, 1.0};
    double* d_r = nullptr;
    if (cudaMalloc(&d_r, sizeof r) || cudaMemcpy(d_r, r, sizeof r, cudaMemcpyHostToDevice)) return 2;
    gkq_config cfg;
    int rc;
    if (workload == 1) {
        gkq_config_default(1, &cfg);
        cfg.algo = GKQ_ALGO_QAGS;
        cfg.tol.kind = GKQ_TOL_RELATIVE;
        cfg.tol.rel_tol = 1e-10;
        rc = gkq_integrate_batch(h, &cfg, gkq_functor_lookup(1, "lorentz_p"), n, d_r, d_r + 1, 0, B.p, n, nullptr,
                                 nullptr, B.est, B.dlt, B.nev, B.st, s);
    } else {
        gkq_config_default(2, &cfg);  // benches/double.rs:18-28: p0/(x y) on [-1,1]^2, point (0,0), Absolute(1.49e-8)
        cfg.tol.kind = GKQ_TOL_ABSOLUTE;
        static const double pt[2] = {0.0, 0.0};
        cfg.points = pt;
        cfg.npoints = 1;
        rc = gkq_integrate2_batch(h, &cfg, gkq_functor_lookup(2, "recip_xy_p"), gkq_functor_lookup(0, "const"), n,
                                  d_r + 2, d_r + 3, 0, B.p, n, d_r + 2, 0, nullptr, nullptr, B.est, B.dlt, B.nev,
                                  B.st, s);
    }
    if (rc) {
        std::fprintf(stderr, "integrate: %d %s\n", rc, gkq_last_error(h));
        return 3;
    }
    return 0;
}

int main(int argc, char** argv) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        std::fprintf(stderr, "no CUDA device (there is no CPU fallback)\n");
        return 1;
    }
    int nranks = argc > 1 ? std::atoi(argv[1]) : (ndev < 2 ? ndev : 2);
    const int64_t n_total = argc > 2 ? std::atoll(argv[2]) : 300000;
    const int workload = argc > 3 ? std::atoi(argv[3]) : 1;
    if (nranks < 1 || nranks > ndev || n_total < 1 || (workload != 1 && workload != 2)) {
        std::fprintf(stderr, "bad arguments (nranks %d, devices %d)\n", nranks, ndev);
        return 1;
    }
    unsigned char id[GKQ_COMM_ID_BYTES] = {0};
    if (nranks > 1 && gkq_comm_get_unique_id(id) != GKQ_OK) {
        std::fprintf(stderr, "gkq_comm_get_unique_id: %s\n", gkq_last_error(nullptr));
        return 1;
    }

    Results gathered;
    std::vector<int> rcs((size_t)nranks, 0);
    auto rank_main = [&](int rank) {
        int& rc = rcs[(size_t)rank];
        cudaSetDevice(rank);
        gkq_handle_t h = nullptr;
        if ((rc = gkq_create(rank, &h))) return;
        cudaStream_t s;
        cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
        if ((rc = gkq_comm_init(h, nranks, rank, id))) {
            std::fprintf(stderr, "rank %d gkq_comm_init: %s\n", rank, gkq_last_error(h));
            return;
        }
        const int64_t mine = gkq_shard_size(n_total, nranks, rank, BLOCKSZ);
        std::vector<int64_t> idx((size_t)mine);
        for (int64_t j = 0; j < mine; ++j) idx[(size_t)j] = gkq_shard_index(n_total, nranks, rank, BLOCKSZ, j);
        DeviceBatch B;
        if ((rc = run_shard(h, workload, idx, B, s))) return;
        double *g_est = nullptr, *g_dlt = nullptr;
        uint32_t *g_nev = nullptr, *g_st = nullptr;
        if (rank == 0) {
            cudaMalloc(&g_est, (size_t)n_total * 8);
            cudaMalloc(&g_dlt, (size_t)n_total * 8);
            cudaMalloc(&g_nev, (size_t)n_total * 4);
            cudaMalloc(&g_st, (size_t)n_total * 4);
        }
        // in two chunks (gkq_gather_range) when the shard allows it, else in one piece
        const int64_t largest = gkq_shard_size(n_total, nranks, 0, BLOCKSZ);
        const int64_t half = (largest / 2) / 4 * 4;
        if (half > 0) {
            rc = gkq_gather_range(h, n_total, BLOCKSZ, 0, 0, half, B.est, B.dlt, B.nev, B.st, g_est, g_dlt, g_nev, g_st, s);
            if (!rc)
                rc = gkq_gather_range(h, n_total, BLOCKSZ, 0, half, largest - half, B.est, B.dlt, B.nev, B.st, g_est,
                                      g_dlt, g_nev, g_st, s);
        } else {
            rc = gkq_gather(h, n_total, BLOCKSZ, 0, B.est, B.dlt, B.nev, B.st, g_est, g_dlt, g_nev, g_st, s);
        }
        if (rc) {
            std::fprintf(stderr, "rank %d gkq_gather: %s\n", rank, gkq_last_error(h));
            return;
        }
        if ((rc = gkq_sync(h, s))) return;
        if (rank == 0) {
            gathered.est.resize((size_t)n_total);
            gathered.dlt.resize((size_t)n_total);
            gathered.nev.resize((size_t)n_total);
            gathered.st.resize((size_t)n_total);
            cudaMemcpy(gathered.est.data(), g_est, (size_t)n_total * 8, cudaMemcpyDeviceToHost);
            cudaMemcpy(gathered.dlt.data(), g_dlt, (size_t)n_total * 8, cudaMemcpyDeviceToHost);
            cudaMemcpy(gathered.nev.data(), g_nev, (size_t)n_total * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(gathered.st.data(), g_st, (size_t)n_total * 4, cudaMemcpyDeviceToHost);
        }
        cudaDeviceSynchronize();
        gkq_comm_destroy(h);
        gkq_destroy(h);
    };
    std::vector<std::thread> th;
    for (int r = 0; r < nranks; ++r) th.emplace_back(rank_main, r);
    for (auto& t : th) t.join();
    for (int r = 0; r < nranks; ++r)
        if (rcs[(size_t)r]) {
            std::fprintf(stderr, "rank %d failed: %d\n", r, rcs[(size_t)r]);
            return 4;
        }

    // the same batch, unsharded, on GPU 0
    cudaSetDevice(0);
    gkq_handle_t h = nullptr;
    if (gkq_create(0, &h)) return 5;
    std::vector<int64_t> all((size_t)n_total);
    for (int64_t i = 0; i < n_total; ++i) all[(size_t)i] = i;
    DeviceBatch B;
    if (run_shard(h, workload, all, B, nullptr)) return 5;
    gkq_sync(h, nullptr);
    Results ref;
    ref.est.resize((size_t)n_total);
    ref.dlt.resize((size_t)n_total);
    ref.nev.resize((size_t)n_total);
    ref.st.resize((size_t)n_total);
    cudaMemcpy(ref.est.data(), B.est, (size_t)n_total * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(ref.dlt.data(), B.dlt, (size_t)n_total * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(ref.nev.data(), B.nev, (size_t)n_total * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(ref.st.data(), B.st, (size_t)n_total * 4, cudaMemcpyDeviceToHost);
    gkq_destroy(h);

    const bool same = std::memcmp(ref.est.data(), gathered.est.data(), (size_t)n_total * 8) == 0 &&
                      std::memcmp(ref.dlt.data(), gathered.dlt.data(), (size_t)n_total * 8) == 0 &&
                      std::memcmp(ref.nev.data(), gathered.nev.data(), (size_t)n_total * 4) == 0 &&
                      std::memcmp(ref.st.data(), gathered.st.data(), (size_t)n_total * 4) == 0;
    uint64_t evals = 0;
    for (auto v : ref.nev) evals += v;
    std::printf("example_two_ranks: workload %d, %lld integrals over %d rank(s), %llu evaluations: gathered results %s "
                "the unsharded run\n",
                workload, (long long)n_total, nranks, (unsigned long long)evals,
                same ? "are bit-identical to" : "DIFFER from");
    return same ? 0 : 6;
}
