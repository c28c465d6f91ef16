// comm.cu -- multi-GPU plumbing of the C ABI: gkq_comm_* / gkq_shard_* / gkq_gather
// (include/gkquad_b200.h, "multi-GPU").
//
// The path shards with no data-path collective: integrals are independent,
// every rank integrates a static block-cyclic shard of the batch (i -> rank (i / block) mod nranks)
// and only RESULTS move -- one gather of 20 bytes per integral over NCCL / NVLink:
//
//     estimate f64 | delta f64 | (nevals << 3 | status) u32
//
// gkq_gather packs a range of the local shard with one kernel, moves it with grouped ncclSend /
// ncclRecv to the root (or ncclAllGather when every rank wants the result) and un-permutes the
// block-cyclic order on the receiving side with one kernel, everything on the caller's stream, so a
// gather of batch k on a side stream overlaps the kernels of batch k + 1.
//
// NCCL is bound at RUN time (dlopen of libnccl.so.2 -- inside a PyTorch process that is the copy
// torch already loaded; GKQ_NCCL_LIB overrides): libgkquad_b200.so itself has no link-time
// dependency on it and single-GPU hosts never touch it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/gkquad_b200.h"
#include "engine_internal.h"

namespace {

struct NcclApi {
    void* dl = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};

NcclApi& nccl_api() {
    static NcclApi api;
    return api;
}

bool load_nccl(std::string& why) {
    NcclApi& A = nccl_api();
    if (A.dl) return true;
    const char* names[3] = {getenv("GKQ_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* dl = nullptr;
    for (const char* nm : names) {
        if (!nm || !*nm) continue;
        // a copy that the process already holds (PyTorch's bundled NCCL) is preferred
        dl = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (!dl) dl = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (dl) break;
    }
    if (!dl) {
        why = std::string("cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "not found");
        return false;
    }
    bool ok = true;
    auto sym = [&](const char* name) {
        void* p = dlsym(dl, name);
        if (!p) {
            ok = false;
            why = std::string("NCCL lacks symbol ") + name;
        }
        return p;
    };
    A.GetUniqueId = (decltype(A.GetUniqueId))sym("ncclGetUniqueId");
    A.CommInitRank = (decltype(A.CommInitRank))sym("ncclCommInitRank");
    A.CommDestroy = (decltype(A.CommDestroy))sym("ncclCommDestroy");
    A.AllGather = (decltype(A.AllGather))sym("ncclAllGather");
    A.Send = (decltype(A.Send))sym("ncclSend");
    A.Recv = (decltype(A.Recv))sym("ncclRecv");
    A.GroupStart = (decltype(A.GroupStart))sym("ncclGroupStart");
    A.GroupEnd = (decltype(A.GroupEnd))sym("ncclGroupEnd");
    A.GetErrorString = (decltype(A.GetErrorString))sym("ncclGetErrorString");
    if (!ok) return false;
    A.dl = dl;
    return true;
}

constexpr int CB = 256;

// ---- the two kernels of a gather ----------------------------------------------------------------
// Packed layout of one rank's contribution to a gather of `cnt` shard positions:
//   [ est: cnt f64 | delta: cnt f64 | (nevals << 3 | status): cnt u32 ]  = 20 cnt bytes
// (cnt is padded to a multiple of 4 by the caller, so every section is 16-byte aligned).
__global__ void __launch_bounds__(CB) k_gather_pack(long long m, long long cnt, const double* est, const double* dlt,
                                                    const uint32_t* nev, const uint32_t* st, unsigned char* out) {
    const long long j = (long long)blockIdx.x * CB + threadIdx.x;
    if (j >= m) return;
    double* pe = reinterpret_cast<double*>(out);
    double* pd = pe + cnt;
    uint32_t* pw = reinterpret_cast<uint32_t*>(pd + cnt);
    pe[j] = est[j];
    pd[j] = dlt[j];
    pw[j] = (nev[j] << 3) | (st[j] & 7u);
}

// Un-permute: position j of rank r's range [off, off + cnt) is global integral
//   ((off + j) / block * nranks + r) * block + (off + j) % block      (block-cyclic).
__global__ void __launch_bounds__(CB) k_gather_unpack(long long cnt, long long off, long long block, int nranks,
                                                      long long n_total, const unsigned char* in, double* g_est,
                                                      double* g_dlt, uint32_t* g_nev, uint32_t* g_st) {
    const long long j = (long long)blockIdx.x * CB + threadIdx.x;
    const int r = blockIdx.y;
    if (j >= cnt) return;
    const long long l = off + j;
    const long long g = ((l / block) * nranks + r) * block + l % block;
    if (g >= n_total) return;  // beyond this rank's shard
    const unsigned char* base = in + (size_t)r * (size_t)cnt * 20u;
    const double* pe = reinterpret_cast<const double*>(base);
    const double* pd = pe + cnt;
    const uint32_t* pw = reinterpret_cast<const uint32_t*>(pd + cnt);
    const uint32_t w = pw[j];
    g_est[g] = pe[j];
    g_dlt[g] = pd[j];
    g_nev[g] = w >> 3;
    g_st[g] = w & 7u;
}

}  // namespace

struct gkq_comm_s {
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    unsigned char* send = nullptr;
    unsigned char* recv = nullptr;
    size_t send_cap = 0, recv_cap = 0;
};

#define GKQ_NCCL(h, call)                                                                             \
    do {                                                                                              \
        ncclResult_t r_ = (call);                                                                     \
        if (r_ != ncclSuccess) {                                                                      \
            char buf_[384];                                                                           \
            snprintf(buf_, sizeof buf_, "%s failed: %s", #call, nccl_api().GetErrorString(r_));       \
            return gkq_internal_fail(h, GKQ_ERR_CUDA, buf_);                                          \
        }                                                                                             \
    } while (0)
#define GKQ_CU(h, call)                                                                               \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            char buf_[384];                                                                           \
            snprintf(buf_, sizeof buf_, "%s failed: %s", #call, cudaGetErrorString(e_));              \
            return gkq_internal_fail(h, e_ == cudaErrorMemoryAllocation ? GKQ_ERR_OUT_OF_MEMORY : GKQ_ERR_CUDA, buf_); \
        }                                                                                             \
    } while (0)

extern "C" {

int64_t gkq_shard_size(int64_t n_total, int nranks, int rank, int64_t block) {
    if (n_total < 0 || nranks < 1 || rank < 0 || rank >= nranks || block < 1) return GKQ_ERR_INVALID_ARGUMENT;
    const int64_t nb = (n_total + block - 1) / block;  // blocks of the batch; the last may be short
    if (nb == 0) return 0;
    int64_t mine = nb / nranks + (rank < nb % nranks ? 1 : 0);
    if (mine == 0) return 0;
    int64_t size = mine * block;
    const int64_t last_owner = (nb - 1) % nranks;
    if (rank == last_owner) size -= nb * block - n_total;
    return size;
}

int64_t gkq_shard_index(int64_t n_total, int nranks, int rank, int64_t block, int64_t j) {
    if (n_total < 0 || nranks < 1 || rank < 0 || rank >= nranks || block < 1 || j < 0) return GKQ_ERR_INVALID_ARGUMENT;
    const int64_t g = ((j / block) * nranks + rank) * block + j % block;
    return g < n_total ? g : (int64_t)GKQ_ERR_INVALID_ARGUMENT;
}

int gkq_comm_get_unique_id(void* id_out) {
    if (!id_out) return GKQ_ERR_INVALID_ARGUMENT;
    std::string why;
    if (!load_nccl(why)) return gkq_internal_fail(nullptr, GKQ_ERR_UNSUPPORTED, why.c_str());
    static_assert(sizeof(ncclUniqueId) == GKQ_COMM_ID_BYTES, "GKQ_COMM_ID_BYTES must match ncclUniqueId");
    ncclUniqueId id;
    ncclResult_t r = nccl_api().GetUniqueId(&id);
    if (r != ncclSuccess) return gkq_internal_fail(nullptr, GKQ_ERR_CUDA, nccl_api().GetErrorString(r));
    memcpy(id_out, &id, sizeof id);
    return GKQ_OK;
}

int gkq_comm_init(gkq_handle_t h, int nranks, int rank, const void* unique_id) {
    if (!h) return GKQ_ERR_INVALID_ARGUMENT;
    if (nranks < 1 || rank < 0 || rank >= nranks) return gkq_internal_fail(h, GKQ_ERR_INVALID_ARGUMENT, "bad nranks / rank");
    if (nranks > 1 && !unique_id) return gkq_internal_fail(h, GKQ_ERR_INVALID_ARGUMENT, "unique_id is NULL");
    gkq_comm_s*& slot = gkq_internal_comm_slot(h);
    if (slot) return gkq_internal_fail(h, GKQ_ERR_INVALID_ARGUMENT, "the handle already has a communicator");
    GKQ_CU(h, cudaSetDevice(gkq_internal_device(h)));
    gkq_comm_s* c = new gkq_comm_s();
    c->nranks = nranks;
    c->rank = rank;
    if (nranks > 1) {  // (one rank gathers with the two kernels alone: NCCL is not even loaded)
        std::string why;
        if (!load_nccl(why)) {
            delete c;
            return gkq_internal_fail(h, GKQ_ERR_UNSUPPORTED, why.c_str());
        }
        ncclUniqueId id;
        memcpy(&id, unique_id, sizeof id);
        ncclResult_t r = nccl_api().CommInitRank(&c->comm, nranks, id, rank);
        if (r != ncclSuccess) {
            delete c;
            return gkq_internal_fail(h, GKQ_ERR_CUDA, nccl_api().GetErrorString(r));
        }
    }
    slot = c;
    return GKQ_OK;
}

int gkq_comm_destroy(gkq_handle_t h) {
    if (!h) return GKQ_ERR_INVALID_ARGUMENT;
    gkq_comm_s*& slot = gkq_internal_comm_slot(h);
    if (!slot) return GKQ_OK;
    cudaSetDevice(gkq_internal_device(h));
    cudaDeviceSynchronize();
    if (slot->comm) nccl_api().CommDestroy(slot->comm);
    cudaFree(slot->send);
    cudaFree(slot->recv);
    delete slot;
    slot = nullptr;
    return GKQ_OK;
}

int gkq_comm_info(gkq_handle_t h, int* nranks, int* rank) {
    if (!h) return GKQ_ERR_INVALID_ARGUMENT;
    gkq_comm_s* c = gkq_internal_comm_slot(h);
    if (nranks) *nranks = c ? c->nranks : 1;
    if (rank) *rank = c ? c->rank : 0;
    return GKQ_OK;
}

int gkq_gather_range(gkq_handle_t h, int64_t n_total, int64_t block, int root, int64_t local_offset,
                     int64_t local_count, const double* est, const double* dlt, const uint32_t* nev,
                     const uint32_t* st, double* g_est, double* g_dlt, uint32_t* g_nev, uint32_t* g_st,
                     void* stream) {
    if (!h) return GKQ_ERR_INVALID_ARGUMENT;
    gkq_comm_s* c = gkq_internal_comm_slot(h);
    if (!c) return gkq_internal_fail(h, GKQ_ERR_INVALID_ARGUMENT, "gkq_comm_init has not been called on this handle");
    if (n_total < 0 || block < 1 || local_offset < 0 || local_count < 0 || root >= c->nranks)
        return gkq_internal_fail(h, GKQ_ERR_INVALID_ARGUMENT, "bad gather geometry");
    if (local_offset % 4) return gkq_internal_fail(h, GKQ_ERR_INVALID_ARGUMENT, "local_offset must be a multiple of 4");
    if (gkq_internal_max_evals_seen(h) >= (1ull << 29))
        return gkq_internal_fail(h, GKQ_ERR_UNSUPPORTED, "packed results carry nevals in 29 bits: max_evals must stay below 2^29");
    const bool receiver = root < 0 || root == c->rank;
    const int64_t mine = gkq_shard_size(n_total, c->nranks, c->rank, block);
    int64_t m = mine - local_offset;  // positions of this rank inside the range
    if (m < 0) m = 0;
    if (m > local_count) m = local_count;
    if (m > 0 && (!est || !dlt || !nev || !st)) return gkq_internal_fail(h, GKQ_ERR_INVALID_ARGUMENT, "NULL result array");
    if (receiver && (!g_est || !g_dlt || !g_nev || !g_st))
        return gkq_internal_fail(h, GKQ_ERR_INVALID_ARGUMENT, "NULL global array on a receiving rank");
    if (local_count == 0 || n_total == 0) return GKQ_OK;
    GKQ_CU(h, cudaSetDevice(gkq_internal_device(h)));
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t cnt = (local_count + 3) / 4 * 4;  // every rank uses the same stride
    const size_t bytes = (size_t)cnt * 20u;
    // staging, grown on demand (a growing call drains the stream first: an earlier gather may still
    // be reading the old buffers)
    if (receiver && (size_t)c->nranks * bytes > c->recv_cap) {
        GKQ_CU(h, cudaStreamSynchronize(s));
        if (c->recv) GKQ_CU(h, cudaFree(c->recv));
        c->recv = nullptr;
        c->recv_cap = 0;
        GKQ_CU(h, cudaMalloc(&c->recv, (size_t)c->nranks * bytes));
        c->recv_cap = (size_t)c->nranks * bytes;
    }
    if (!receiver && bytes > c->send_cap) {
        GKQ_CU(h, cudaStreamSynchronize(s));
        if (c->send) GKQ_CU(h, cudaFree(c->send));
        c->send = nullptr;
        c->send_cap = 0;
        GKQ_CU(h, cudaMalloc(&c->send, bytes));
        c->send_cap = bytes;
    }
    // a receiving rank packs straight into its own slot of the receive staging
    unsigned char* mine_packed = receiver ? c->recv + (size_t)c->rank * bytes : c->send;
    if (m > 0) {
        k_gather_pack<<<(unsigned)((m + CB - 1) / CB), CB, 0, s>>>(m, cnt, est + local_offset, dlt + local_offset,
                                                                  nev + local_offset, st + local_offset, mine_packed);
        gkq_internal_count_launch(h);
    }
    if (c->nranks > 1) {
        NcclApi& A = nccl_api();
        const size_t words = bytes / 4;
        if (root < 0) {
            GKQ_NCCL(h, A.AllGather(mine_packed, c->recv, words, ncclUint32, c->comm, s));
        } else {
            GKQ_NCCL(h, A.GroupStart());
            if (receiver) {
                for (int r = 0; r < c->nranks; ++r)
                    if (r != c->rank) GKQ_NCCL(h, A.Recv(c->recv + (size_t)r * bytes, words, ncclUint32, r, c->comm, s));
            } else {
                GKQ_NCCL(h, A.Send(c->send, words, ncclUint32, root, c->comm, s));
            }
            GKQ_NCCL(h, A.GroupEnd());
        }
    }
    if (receiver) {
        dim3 grid((unsigned)((cnt + CB - 1) / CB), (unsigned)c->nranks);
        k_gather_unpack<<<grid, CB, 0, s>>>(cnt, local_offset, block, c->nranks, n_total, c->recv, g_est, g_dlt, g_nev, g_st);
        gkq_internal_count_launch(h);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return gkq_internal_fail(h, GKQ_ERR_CUDA, cudaGetErrorString(e));
    return GKQ_OK;
}

int gkq_gather(gkq_handle_t h, int64_t n_total, int64_t block, int root, const double* est, const double* dlt,
               const uint32_t* nev, const uint32_t* st, double* g_est, double* g_dlt, uint32_t* g_nev,
               uint32_t* g_st, void* stream) {
    if (!h) return GKQ_ERR_INVALID_ARGUMENT;
    gkq_comm_s* c = gkq_internal_comm_slot(h);
    if (!c) return gkq_internal_fail(h, GKQ_ERR_INVALID_ARGUMENT, "gkq_comm_init has not been called on this handle");
    if (n_total < 0 || block < 1) return gkq_internal_fail(h, GKQ_ERR_INVALID_ARGUMENT, "bad gather geometry");
    // the largest shard (rank 0 owns the first block) defines the common range
    const int64_t largest = gkq_shard_size(n_total, c->nranks, 0, block);
    return gkq_gather_range(h, n_total, block, root, 0, largest, est, dlt, nev, st, g_est, g_dlt, g_nev, g_st, stream);
}

}  // extern "C"
