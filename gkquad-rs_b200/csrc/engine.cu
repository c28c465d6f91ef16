// engine.cu -- handle, device scratch, dispatch and the C ABI of libgkquad_b200.so
// (include/gkquad_b200.h).  No torch, no oracle, no CPU fallback: without a CUDA device
// gkq_create fails and nothing can be computed.
//
// Orchestration of one QAGS batch:
//   bulk      k_qags_init<qk17> -> { k_qags_first[0] } || { k_qags_init<qk25> -> k_qags_first[1] }
//             flat, stateless, throughput-bound kernels; the few integrals that are still
//             unfinished after their first bisection step leave as self-contained STRAGGLER
//             RECORDS in a handle-owned buffer;
//   remainder one launch of the persistent kernel k_adaptive<.., QAGS> finishes the records --
//             immediately (ordered mode), at gkq_join (deferred mode: the remainders of many
//             batches share one launch), or once at the end of a host-pipeline call.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <atomic>
#include <vector>

#include "../../include/gkquad_b200.h"
#include "engine_internal.h"
#include "kernels_1d.cuh"
#include "kernels_2d.cuh"

namespace gkq {

#define GKQ_ROW(ID, NAME, NP, FL) {NAME, NP, FL},
static const FunctorInfo F1_TABLE[F1_COUNT] = {GKQ_F1_LIST(GKQ_ROW)};
static const FunctorInfo F2_TABLE[F2_COUNT] = {GKQ_F2_LIST(GKQ_ROW)};
static const FunctorInfo YR_TABLE[YR_COUNT] = {GKQ_YR_LIST(GKQ_ROW)};
#undef GKQ_ROW

static thread_local std::string g_create_error = "";

// ---- small kernels that are not templated on a functor ----------------------------------------
__global__ void __launch_bounds__(BLOCK) k_fill_result(long long n, Outputs O, double est, double dlt,
                                                       uint32_t nev, uint32_t st) {
    const long long i = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (i < n) write_result(O, i, est, dlt, nev, st);
}

// The same over a device-built worklist (AUTO split of a CSR batch).
__global__ void __launch_bounds__(BLOCK) k_fill_result_list(const unsigned int* list, const unsigned long long* count,
                                                            Outputs O, double est, double dlt, uint32_t nev, uint32_t st) {
    const unsigned long long total = *count;
    for (unsigned long long k = (unsigned long long)blockIdx.x * BLOCK + threadIdx.x; k < total;
         k += (unsigned long long)gridDim.x * BLOCK)
        write_result(O, (long long)list[k], est, dlt, nev, st);
}

// Host entry: nevals and status of a chunk leave the device as ONE u32 (nevals << 3 | status):
// 20 instead of 24 bytes per integral over a link that is the bound of the whole call.
__global__ void __launch_bounds__(BLOCK) k_pack_nev_st(long long n, const uint32_t* nev, const uint32_t* st, uint32_t* out) {
    const long long i = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (i < n) out[i] = (nev[i] << 3) | (st[i] & 7u);
}

// GKQ_FLAG_VALIDATE_INPUTS: any NaN among the first `n` entries of up to five optional arrays?
__global__ void __launch_bounds__(BLOCK) k_any_nan(long long n, const double* a0, const double* a1, const double* a2,
                                                   const double* a3, unsigned int* flag) {
    const long long i = (long long)blockIdx.x * BLOCK + threadIdx.x;
    bool bad = false;
    if (i < n) {
        if (a0) bad = bad || a0[i] != a0[i];
        if (a1) bad = bad || a1[i] != a1[i];
        if (a2) bad = bad || a2[i] != a2[i];
        if (a3) bad = bad || a3[i] != a3[i];
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1u);
}

// AUTO over a CSR batch (auto.rs:29-33): integrals without points -> QAGS worklist, with
// points -> QAGP worklist; also records the largest per-integral point count.
__global__ void __launch_bounds__(BLOCK) k_split_auto(long long n, const long long* pts_offsets,
                                                      unsigned int* listS, unsigned int* listP,
                                                      Control* ctl, int mode) {
    const long long i = (long long)blockIdx.x * BLOCK + threadIdx.x;
    const bool valid = i < n;
    long long cnt = 0;
    if (valid) cnt = pts_offsets[i + 1] - pts_offsets[i];
    // mode 0 = AUTO split, 1 = everything to QAGP (only the max count is needed)
    if (mode == 0) {
        list_append(valid && cnt == 0, (unsigned int)i, listS, &ctl->countS);
        list_append(valid && cnt != 0, (unsigned int)i, listP, &ctl->countP);
    }
    unsigned long long m = (unsigned long long)(cnt > 0 ? cnt : 0);
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, m, o);
        m = other > m ? other : m;
    }
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(&ctl->max_pts, m);
}

// FP64 pipe calibration: 8 independent dependent-DFMA chains per thread.
__global__ void __launch_bounds__(BLOCK) k_dfma_peak(double* sink, int iters, double seed) {
    double x0 = seed + threadIdx.x, x1 = x0 + 1., x2 = x0 + 2., x3 = x0 + 3.;
    double x4 = x0 + 4., x5 = x0 + 5., x6 = x0 + 6., x7 = x0 + 7.;
    const double m = 0.999999, c = 1e-9;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, m, c); x1 = fma(x1, m, c); x2 = fma(x2, m, c); x3 = fma(x3, m, c);
        x4 = fma(x4, m, c); x5 = fma(x5, m, c); x6 = fma(x6, m, c); x7 = fma(x7, m, c);
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) sink[0] = s;  // never true: keeps the chains alive
}

}  // namespace gkq

using namespace gkq;

// ---- handle ----------------------------------------------------------------------------------
struct gkq_handle_s {
    int device = 0;
    int sm_count = 0;
    size_t total_mem = 0;
    std::string err;
    gkq_stats stats{};
    bool deferred = false;  // gkq_set_deferred_join
    gkq_comm_s* comm = nullptr;            // gkq_comm_init (comm.cu)
    unsigned long long max_evals_seen = 0;  // largest cfg->max_evals of any call (gkq_gather packs nevals in 29 bits)

    // Bulk scratch: two contexts, so the host pipeline can run chunks on two compute streams.
    struct Scratch {
        Control* ctl = nullptr;       // device
        Control* ctl_host = nullptr;  // pinned mirror (profiling read-back, CSR split)
        bool ctl_pending = false;
        size_t cap_n = 0;
        ParkRec* park = nullptr;      // result0 of loop entrants, parked between the bulk kernels
        unsigned int *list25 = nullptr, *listS = nullptr, *listP = nullptr;
        unsigned int* listLoop[2] = {nullptr, nullptr};
        cudaStream_t side = nullptr;  // chain 1 of the QAGS bulk
        cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
        // slab of the persistent QAGP / nested kernels
        size_t slab_d_elems = 0, slab_w_elems = 0;
        double* slab_d = nullptr;
        int* slab_w = nullptr;
        // shared points (device copy + pinned staging)
        size_t dev_pts_cap = 0, host_pts_cap = 0, pts_count = 0;
        double* dev_pts = nullptr;
        double* host_pts = nullptr;
    };
    Scratch sc[2];
    int cur = 0;

    // Straggler records: two buffers, so that the remainder of one group of batches can run while
    // the next batches already append to the other buffer.
    static constexpr int MAX_CALLS = 64;
    struct RecBuf {
        StragglerRec* recs = nullptr;
        size_t cap = 0;
        RecCtl* rctl = nullptr;
        RecCtl* rctl_host = nullptr;      // pinned
        Outputs* call_outs = nullptr;     // device [MAX_CALLS]
        Outputs* call_outs_pin = nullptr;  // pinned mirror
        int flushed_calls = 0;            // calls covered by the flush whose count is still to be read back
        int flushed_fid = -1;             // functor of that flush (R.fid may be reused before the count arrives)
        int ncalls = 0;                   // calls that appended since the last flush
        size_t n_bound = 0;               // upper bound of the records they may have appended
        int fid = -1;                     // configuration the pending records were produced under
        Tol tol{};
        unsigned long long max_evals = 0, ws_capacity = 0;
        cudaEvent_t ev_bulk = nullptr;    // after the last bulk kernel that appended
        cudaEvent_t ev_tail = nullptr;    // after the flush
        bool tail_pending = false;
        // slab of the persistent QAGS kernel
        size_t slab_d_elems = 0, slab_w_elems = 0;
        double* slab_d = nullptr;
        int* slab_w = nullptr;
        // patch arrays (host pipeline): [cap] idx | est | dlt | nev | st, device + pinned
        char* patch_dev = nullptr;
        char* patch_pin = nullptr;
        size_t patch_cap = 0;
    };
    RecBuf rb[2];
    int rb_cur = 0;
    // deferred mode: the remainder of a full record buffer runs on this stream, beside the caller's
    // next batches (which append to the twin buffer); gkq_join orders the caller's stream behind it
    cudaStream_t s_tail = nullptr;
    // records per call seen by completed flushes (sizes the next remainder launch: a persistent grid
    // with more warps than records only takes registers away from the bulk kernels running beside it)
    double rec_per_call = -1.0;
    int rec_est_fid = -1;

    // host-entry pipeline (gkq_integrate_batch_host)
    static constexpr int MAX_CHUNKS = 64;
    cudaStream_t s_in = nullptr, s_compute = nullptr, s_compute2 = nullptr, s_out = nullptr, s_out2 = nullptr;
    cudaEvent_t ev_in[MAX_CHUNKS] = {}, ev_done[MAX_CHUNKS] = {};
    cudaEvent_t ev_c2 = nullptr;
    bool host_records_primed = false;  // the host pipeline has reset the record counters itself
    int host_tail_fid = -1;            // functor and straggler share of the last QAGS host call
    double host_tail_share = 0.0;
    size_t stage_bytes = 0;
    char* stage_dev = nullptr;
    uint32_t* pk_pin = nullptr;  // pinned landing area of the packed (nevals << 3 | status) words
    size_t pk_cap = 0;
    cudaEvent_t ev_out[MAX_CHUNKS] = {};
    double* fixed_pin = nullptr;  // pinned copy of the shared (stride-0) range / parameters

    // optional per-kernel timing (gkq_set_profiling): one CUDA event pair per launch
    bool profiling = false;
    struct ProfEntry {
        const char* name;
        cudaEvent_t e0, e1;
    };
    std::vector<ProfEntry> prof;
    std::vector<cudaEvent_t> prof_pool;

    // level-synchronous 2-D path (kernels_2d_flat.cuh): one carved allocation + pinned round counters
    char* flat_buf = nullptr;
    size_t flat_bytes = 0;
    Ctl2* flat_ctl_host = nullptr;  // pinned [2]
    cudaEvent_t ev_flat[2] = {nullptr, nullptr};
    int occ_2df[F2_COUNT][2];

    Launch1D l1[F1_COUNT];
    int occ_loop[F1_COUNT], occ_qagp[F1_COUNT], occ_init25[F1_COUNT], occ_first[F1_COUNT];
    Launch2D l2[F2_COUNT];
    int occ_2d[F2_COUNT][2];
    int occ_2dw[F2_COUNT][2];
};

namespace {

int fail(gkq_handle_t h, int code, const std::string& msg) {
    if (h)
        h->err = msg;
    else
        g_create_error = msg;
    return code;
}
#define GKQ_CUDA(h, call)                                                                       \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            char buf_[512];                                                                     \
            snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                     __FILE__, __LINE__);                                                       \
            return fail(h, e_ == cudaErrorMemoryAllocation ? GKQ_ERR_OUT_OF_MEMORY : GKQ_ERR_CUDA, \
                        buf_);                                                                  \
        }                                                                                       \
    } while (0)

template <int ID>
struct FillL1 {
    static void go(gkq_handle_s* h) {
        h->l1[ID] = make_launch1d<ID>();
        h->occ_loop[ID] = h->occ_qagp[ID] = h->occ_init25[ID] = h->occ_first[ID] = -1;
        FillL1<ID + 1>::go(h);
    }
};
template <>
struct FillL1<F1_COUNT> {
    static void go(gkq_handle_s*) {}
};
template <int ID>
struct FillL2 {
    static void go(gkq_handle_s* h) {
        h->l2[ID] = make_launch2d<ID>();
        h->occ_2d[ID][0] = h->occ_2d[ID][1] = -1;
        h->occ_2dw[ID][0] = h->occ_2dw[ID][1] = -1;
        h->occ_2df[ID][0] = h->occ_2df[ID][1] = -1;
        FillL2<ID + 1>::go(h);
    }
};
template <>
struct FillL2<F2_COUNT> {
    static void go(gkq_handle_s*) {}
};

template <class T>
int grow(gkq_handle_t h, T*& p, size_t& have, size_t want) {
    if (want <= have) return 0;
    if (p) {
        GKQ_CUDA(h, cudaFree(p));
        h->stats.scratch_bytes -= have * sizeof(T);
        p = nullptr;
        have = 0;
    }
    size_t w = want + want / 8;
    GKQ_CUDA(h, cudaMalloc(&p, w * sizeof(T)));
    have = w;
    h->stats.scratch_bytes += w * sizeof(T);
    return 0;
}

int ensure_batch_scratch(gkq_handle_t h, size_t n) {
    gkq_handle_s::Scratch& S = h->sc[h->cur];
    if (n <= S.cap_n) return 0;
    const size_t w = n + n / 8 + 1024;
    int rc;
    size_t have;
    have = S.cap_n; if ((rc = grow(h, S.park, have, w))) return rc;
    have = S.cap_n; if ((rc = grow(h, S.list25, have, w))) return rc;
    have = S.cap_n; if ((rc = grow(h, S.listS, have, w))) return rc;
    have = S.cap_n; if ((rc = grow(h, S.listP, have, w))) return rc;
    for (int c = 0; c < 2; ++c) {
        have = S.cap_n;
        if ((rc = grow(h, S.listLoop[c], have, w))) return rc;
    }
    S.cap_n = w;  // every grow() allocated >= w elements
    return 0;
}

int slab_geom(gkq_handle_t h, double*& sd, size_t& sd_n, int*& sw, size_t& sw_n, unsigned long long threads,
              int cap, int npts_max, SlabGeom& g) {
    const unsigned long long nd = slab_doubles(cap, npts_max) * threads;
    const unsigned long long nw = slab_ints(cap) * threads;
    int rc;
    if ((rc = grow(h, sd, sd_n, (size_t)nd))) return rc;
    if ((rc = grow(h, sw, sw_n, (size_t)nw))) return rc;
    g.dbase = sd;
    g.wbase = sw;
    g.stride = threads;
    g.cap = cap;
    g.npts_max = npts_max;
    return 0;
}
// slab of the persistent QAGP / nested kernels (bulk scratch context)
int ensure_slab(gkq_handle_t h, unsigned long long threads, int cap, int npts_max, SlabGeom& g) {
    gkq_handle_s::Scratch& S = h->sc[h->cur];
    return slab_geom(h, S.slab_d, S.slab_d_elems, S.slab_w, S.slab_w_elems, threads, cap, npts_max, g);
}

int check_tol(gkq_handle_t h, const gkq_tolerance& t) {
    if (t.kind < 0 || t.kind > 3) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "unknown tolerance kind");
    const bool need_abs = t.kind != GKQ_TOL_RELATIVE, need_rel = t.kind != GKQ_TOL_ABSOLUTE;
    if ((need_abs && t.abs_tol != t.abs_tol) || (need_rel && t.rel_tol != t.rel_tol))
        return fail(h, GKQ_ERR_NAN_INPUT, "Tolerance must not contain NAN.");  // integrator.rs:71
    return 0;
}

// NaN anywhere in a host array?  (the reference panics on NaN ranges / tolerances / points:
// single/common.rs:94, single/integrator.rs:71,86-90)
bool host_has_nan(const double* p, size_t n) {
    if (!p) return false;
    bool bad = false;
    for (size_t i = 0; i < n; ++i) bad |= p[i] != p[i];
    return bad;
}

int upload_shared_points(gkq_handle_t h, const double* pts, size_t count, cudaStream_t st) {
    if (count == 0) return 0;
    gkq_handle_s::Scratch& s = h->sc[h->cur];
    // unchanged points (the usual case: one Integrator, many batches) stay resident
    if (s.dev_pts && s.host_pts && s.pts_count == count && memcmp(s.host_pts, pts, count * sizeof(double)) == 0)
        return 0;
    int rc;
    if ((rc = grow(h, s.dev_pts, s.dev_pts_cap, count))) return rc;
    // an earlier batch may still be copying out of the pinned staging buffer
    GKQ_CUDA(h, cudaStreamSynchronize(st));
    if (count > s.host_pts_cap) {
        if (s.host_pts) GKQ_CUDA(h, cudaFreeHost(s.host_pts));
        s.host_pts = nullptr;
        GKQ_CUDA(h, cudaMallocHost(&s.host_pts, (count + 64) * sizeof(double)));
        s.host_pts_cap = count + 64;
    }
    memcpy(s.host_pts, pts, count * sizeof(double));
    s.pts_count = count;
    GKQ_CUDA(h, cudaMemcpyAsync(s.dev_pts, s.host_pts, count * sizeof(double), cudaMemcpyHostToDevice, st));
    return 0;
}

// Persistent grid size: resident blocks on every SM, bounded by the slab memory budget.
int persistent_grid(gkq_handle_t h, int occ, int cap, int npts_max, int block = BLOCK) {
    if (occ < 1) occ = 1;
    unsigned long long per_thread = slab_doubles(cap, npts_max) * 8ull + slab_ints(cap) * 4ull;
    unsigned long long budget = (unsigned long long)h->total_mem / 3ull;
    unsigned long long max_threads = budget / per_thread;
    long long grid = (long long)h->sm_count * occ;
    long long max_grid = (long long)(max_threads / block);
    if (max_grid < 1) return -1;
    if (grid > max_grid) grid = max_grid;
    return (int)grid;
}

int launch_check(gkq_handle_t h, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        char buf[256];
        snprintf(buf, sizeof buf, "launch of %s failed: %s", what, cudaGetErrorString(e));
        return fail(h, GKQ_ERR_CUDA, buf);
    }
    h->stats.kernel_launches += 1;
    return 0;
}

// GKQ_FLAG_VALIDATE_INPUTS on a device entry: one small kernel + one stream synchronisation.
int validate_device_inputs(gkq_handle_t h, cudaStream_t st, long long n, long long n_range, const double* a,
                           const double* b, const double* abs_each, const double* rel_each) {
    gkq_handle_s::Scratch& S = h->sc[h->cur];
    unsigned int* flag = (unsigned int*)&S.ctl->pad[0];  // zeroed with the control block
    k_any_nan<<<(int)((n_range + BLOCK - 1) / BLOCK), BLOCK, 0, st>>>(n_range, a, b, nullptr, nullptr, flag);
    if (abs_each || rel_each)
        k_any_nan<<<(int)((n + BLOCK - 1) / BLOCK), BLOCK, 0, st>>>(n, abs_each, rel_each, nullptr, nullptr, flag);
    int rc;
    if ((rc = launch_check(h, "k_any_nan"))) return rc;
    GKQ_CUDA(h, cudaMemcpyAsync(S.ctl_host, S.ctl, sizeof(Control), cudaMemcpyDeviceToHost, st));
    GKQ_CUDA(h, cudaStreamSynchronize(st));
    if (S.ctl_host->pad[0] != 0ull)
        return fail(h, GKQ_ERR_NAN_INPUT, "NaN in a range bound or in a per-integral tolerance");
    return GKQ_OK;
}

// Bracket a kernel launch with events on its own stream when profiling is on.
void prof_begin(gkq_handle_t h, const char* name, cudaStream_t st) {
    if (!h->profiling) return;
    gkq_handle_s::ProfEntry pe;
    pe.name = name;
    auto get = [&]() {
        cudaEvent_t ev;
        if (!h->prof_pool.empty()) {
            ev = h->prof_pool.back();
            h->prof_pool.pop_back();
        } else {
            cudaEventCreate(&ev);
        }
        return ev;
    };
    pe.e0 = get();
    pe.e1 = get();
    cudaEventRecord(pe.e0, st);
    h->prof.push_back(pe);
}
void prof_end(gkq_handle_t h, cudaStream_t st) {
    if (!h->profiling || h->prof.empty()) return;
    cudaEventRecord(h->prof.back().e1, st);
}
void prof_clear(gkq_handle_t h) {
    for (auto& pe : h->prof) {
        h->prof_pool.push_back(pe.e0);
        h->prof_pool.push_back(pe.e1);
    }
    h->prof.clear();
}

// ---- straggler records ------------------------------------------------------------------------
inline size_t patch_bytes(size_t cap) { return cap * (8 + 8 + 8 + 4 + 4); }
inline PatchView patch_view(char* base, size_t cap) {
    PatchView v;
    v.idx = (unsigned long long*)base;
    v.est = (double*)(v.idx + cap);
    v.dlt = v.est + cap;
    v.nev = (unsigned int*)(v.dlt + cap);
    v.st = v.nev + cap;
    return v;
}

// Finish the pending records of the current buffer with one launch of the persistent QAGS kernel
// on `stream` and switch to the other buffer.  patch = true: results go to the buffer's patch.
int flush_records(gkq_handle_t h, cudaStream_t stream, bool patch) {
    gkq_handle_s::RecBuf& R = h->rb[h->rb_cur];
    if (R.ncalls == 0) return GKQ_OK;
    int rc;
    const int fid = R.fid;
    const Launch1D& L = h->l1[fid];
    if (h->occ_loop[fid] < 0) h->occ_loop[fid] = L.occupancy_loop();
    const unsigned long long need = (R.max_evals - 17) / 50 + 1;  // qags.rs:98-99, nevals >= 17
    if (need > (1ull << 26)) return fail(h, GKQ_ERR_OUT_OF_MEMORY, "max_evals too large for the interval slabs");
    const int cap = (int)need;  // the list never outgrows the reserve request
    int grid = persistent_grid(h, h->occ_loop[fid], cap, 0, ABLOCK);
    if (grid < 1) return fail(h, GKQ_ERR_OUT_OF_MEMORY, "max_evals too large for the interval slabs");
    const long long ablocks = ((long long)R.n_bound + ABLOCK - 1) / ABLOCK;
    if ((long long)grid > ablocks) grid = (int)(ablocks > 0 ? ablocks : 1);
    const int full_grid = grid;  // (sizes the slabs: the estimate below may grow again)
    // how many records did completed flushes of this configuration find per call?
    for (int c = 0; c < 2; ++c) {
        gkq_handle_s::RecBuf& B = h->rb[c];
        if (B.flushed_calls == 0) continue;
        const cudaError_t q = cudaEventQuery(B.ev_tail);
        if (q == cudaSuccess) {
            h->rec_per_call = (double)B.rctl_host->count / B.flushed_calls;
            h->rec_est_fid = B.flushed_fid;
            B.flushed_calls = 0;
        } else if (q == cudaErrorNotReady) {
            (void)cudaGetLastError();  // "still running" must not surface in the next launch_check
        }
    }
    if (!patch && h->deferred && h->rec_per_call >= 0.0 && h->rec_est_fid == fid) {
        // one warp per expected record (+25 %): same latency as the full grid while the queue is this
        // short, a fraction of its registers; a longer queue than expected still drains correctly
        const double warps = h->rec_per_call * R.ncalls * 1.25 + 64.0;
        const long long blocks = (long long)(warps / (ABLOCK / 32)) + 1;
        if (blocks < (long long)grid) grid = (int)blocks;
    }
    Batch1D Q;
    memset(&Q, 0, sizeof Q);
    // (the slab keeps the size of the full grid; the geometry handed to the kernel is that of this launch)
    if ((rc = slab_geom(h, R.slab_d, R.slab_d_elems, R.slab_w, R.slab_w_elems, (unsigned long long)full_grid * ABLOCK,
                        cap, 0, Q.slab)))
        return rc;
    Q.slab.stride = (unsigned long long)grid * ABLOCK;
    if (h->deferred && !patch) {
        // deferred mode ping-pongs between the two record buffers: size the twin's slab now, so that
        // its first flush (in the caller's steady state) allocates nothing
        gkq_handle_s::RecBuf& O = h->rb[h->rb_cur ^ 1];
        SlabGeom unused;
        if (!O.tail_pending && O.ncalls == 0 &&
            (rc = slab_geom(h, O.slab_d, O.slab_d_elems, O.slab_w, O.slab_w_elems,
                            (unsigned long long)full_grid * ABLOCK, cap, 0, unused)))
            return rc;
    }
    Q.tol = R.tol;
    Q.max_evals = R.max_evals;
    Q.ws_capacity = R.ws_capacity;
    Q.recs = R.recs;
    Q.rctl = R.rctl;
    Q.call_outs = R.call_outs;
    Q.ctl = h->sc[0].ctl;  // unused by the QAGS instance
    if (patch)
        Q.patch = patch_view(R.patch_dev, R.patch_cap);
    else
        GKQ_CUDA(h, cudaStreamWaitEvent(stream, R.ev_bulk, 0));
    prof_begin(h, "k_adaptive<QAGS>", stream);
    L.qags_loop(Q, grid, stream);
    prof_end(h, stream);
    if ((rc = launch_check(h, "k_adaptive<QAGS>"))) return rc;
    if (h->profiling || (h->deferred && !patch)) {
        GKQ_CUDA(h, cudaMemcpyAsync(R.rctl_host, R.rctl, sizeof(RecCtl), cudaMemcpyDeviceToHost, stream));
        R.flushed_calls = R.ncalls;
        R.flushed_fid = fid;
    }
    GKQ_CUDA(h, cudaEventRecord(R.ev_tail, stream));
    R.tail_pending = true;
    R.ncalls = 0;
    R.n_bound = 0;
    h->rb_cur ^= 1;
    return GKQ_OK;
}

// Prepare the current record buffer for a call that may append up to n records under the given
// configuration; flushes first when the pending records are incompatible or would not fit.
int records_begin(gkq_handle_t h, cudaStream_t st, int fid, const Batch1D& P, size_t n, bool patch,
                  unsigned int& slot) {
    int rc;
    {
        gkq_handle_s::RecBuf& R = h->rb[h->rb_cur];
        const bool same = R.fid == fid && R.tol.kind == P.tol.kind && R.tol.abs_tol == P.tol.abs_tol &&
                          R.tol.rel_tol == P.tol.rel_tol && R.max_evals == P.max_evals &&
                          R.ws_capacity == P.ws_capacity;
        if (R.ncalls > 0 && (!same || (!patch && R.ncalls >= gkq_handle_s::MAX_CALLS) || R.n_bound + n > R.cap))
            if ((rc = flush_records(h, (h->deferred && !patch) ? h->s_tail : st, patch))) return rc;
    }
    gkq_handle_s::RecBuf& R = h->rb[h->rb_cur];
    if (R.ncalls == 0) {
        // fresh buffer: its previous remainder must be over, then reset the counters
        if (R.tail_pending) {
            GKQ_CUDA(h, cudaStreamWaitEvent(st, R.ev_tail, 0));
            R.tail_pending = false;
        }
        // deferred mode batches the remainders of up to 8 calls into one launch (every integral of
        // a call may become a record, so room for 8 n records is kept: sizeof(StragglerRec) each)
        // (... but never more than 2^23 records, 1.5 GB, beyond what one call needs: large batches
        // do not need the amortisation, and several handles may share the device)
        size_t want = n;
        if (h->deferred && !patch && !h->profiling) {
            const size_t lim = n > ((size_t)1 << 23) ? n : ((size_t)1 << 23);
            want = 8 * n < lim ? 8 * n : lim;
        }
        if (want > R.cap) {
            GKQ_CUDA(h, cudaStreamSynchronize(st));  // nothing may still read the old buffer
            if ((rc = grow(h, R.recs, R.cap, want))) return rc;
            // grow the idle twin buffer now as well (keeps allocations out of steady state)
            gkq_handle_s::RecBuf& O = h->rb[h->rb_cur ^ 1];
            if (O.ncalls == 0 && !O.tail_pending && want > O.cap)
                if ((rc = grow(h, O.recs, O.cap, want))) return rc;
        }
        if (!h->host_records_primed) GKQ_CUDA(h, cudaMemsetAsync(R.rctl, 0, sizeof(RecCtl), st));
        R.fid = fid;
        R.tol = P.tol;
        R.max_evals = P.max_evals;
        R.ws_capacity = P.ws_capacity;
    }
    if (patch && R.patch_cap < R.cap) {
        if (R.patch_dev) GKQ_CUDA(h, cudaFree(R.patch_dev));
        if (R.patch_pin) GKQ_CUDA(h, cudaFreeHost(R.patch_pin));
        R.patch_dev = R.patch_pin = nullptr;
        GKQ_CUDA(h, cudaMalloc(&R.patch_dev, patch_bytes(R.cap)));
        GKQ_CUDA(h, cudaMallocHost(&R.patch_pin, patch_bytes(R.cap)));
        R.patch_cap = R.cap;
        h->stats.scratch_bytes += patch_bytes(R.cap);
    }
    // (device entry: k_qags_first publishes the call's output arrays in slot `slot` of R.call_outs
    // from its kernel parameters -- Batch1D::call_outs_w; patched results never look the call up)
    slot = (unsigned int)R.ncalls;
    return GKQ_OK;
}

}  // namespace

int gkq_internal_fail(gkq_handle_t h, int code, const char* msg) { return fail(h, code, msg ? msg : ""); }
gkq_comm_s*& gkq_internal_comm_slot(gkq_handle_t h) { return h->comm; }
int gkq_internal_device(gkq_handle_t h) { return h->device; }
unsigned long long gkq_internal_max_evals_seen(gkq_handle_t h) { return h->max_evals_seen; }
void gkq_internal_count_launch(gkq_handle_t h) { h->stats.kernel_launches += 1; }

// ---- C ABI -----------------------------------------------------------------------------------
extern "C" {

int gkq_abi_version(void) { return GKQ_ABI_VERSION; }

int gkq_config_default(int dim, gkq_config* cfg) {
    if (!cfg || (dim != 1 && dim != 2)) return GKQ_ERR_INVALID_ARGUMENT;
    memset(cfg, 0, sizeof *cfg);
    cfg->algo = GKQ_ALGO_AUTO;
    cfg->tol.kind = GKQ_TOL_ABS_OR_REL;  // common.rs:44-49
    cfg->tol.abs_tol = 1.49e-8;
    cfg->tol.rel_tol = 1.49e-8;
    cfg->max_evals = dim == 1 ? 2000 : 100000;  // single/common.rs:120, double/common.rs:27
    cfg->workspace_capacity = 0;
    cfg->points = nullptr;
    cfg->npoints = 0;
    return GKQ_OK;
}

int gkq_create(int device, gkq_handle_t* out) { return gkq_create_prio(device, 0, out); }

int gkq_create_prio(int device, int stream_priority, gkq_handle_t* out) {
    if (!out) return fail(nullptr, GKQ_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, GKQ_ERR_NO_DEVICE,
                    std::string("no CUDA device available (there is no CPU fallback): ") +
                        cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, GKQ_ERR_INVALID_ARGUMENT, "bad device index");
    gkq_handle_s* h = new gkq_handle_s();
    h->device = device;
    if (cudaSetDevice(device) != cudaSuccess) {
        delete h;
        return fail(nullptr, GKQ_ERR_CUDA, "cudaSetDevice failed");
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete h;
        return fail(nullptr, GKQ_ERR_CUDA, "cudaGetDeviceProperties failed");
    }
    if (prop.major < 10) {
        delete h;
        return fail(nullptr, GKQ_ERR_UNSUPPORTED,
                    "this library carries sm_100a code only (B200); device is older");
    }
    h->sm_count = prop.multiProcessorCount;
    h->total_mem = prop.totalGlobalMem;
    FillL1<0>::go(h);
    FillL2<0>::go(h);
    bool ok = true;
    auto ev = [&](cudaEvent_t* e_) { ok = ok && cudaEventCreateWithFlags(e_, cudaEventDisableTiming) == cudaSuccess; };
    // the handle's own streams (second chain, tail, host-entry pipeline) take the priority the caller
    // gives its lane; CUDA clamps the value to the device's range (lower number = runs first)
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    const int prio = stream_priority < prio_hi ? prio_hi : (stream_priority > prio_lo ? prio_lo : stream_priority);
    auto strm = [&](cudaStream_t* s_) {
        ok = ok && cudaStreamCreateWithPriority(s_, cudaStreamNonBlocking, prio) == cudaSuccess;
    };
    for (int c = 0; c < 2; ++c) {
        gkq_handle_s::Scratch& S = h->sc[c];
        ok = ok && cudaMalloc(&S.ctl, sizeof(Control)) == cudaSuccess &&
             cudaMallocHost(&S.ctl_host, sizeof(Control)) == cudaSuccess;
        if (ok) memset(S.ctl_host, 0, sizeof(Control));
        strm(&S.side);
        ev(&S.ev_fork);
        ev(&S.ev_join);
        gkq_handle_s::RecBuf& R = h->rb[c];
        ok = ok && cudaMalloc(&R.rctl, sizeof(RecCtl)) == cudaSuccess &&
             cudaMallocHost(&R.rctl_host, sizeof(RecCtl)) == cudaSuccess &&
             cudaMalloc(&R.call_outs, sizeof(Outputs) * gkq_handle_s::MAX_CALLS) == cudaSuccess &&
             cudaMallocHost(&R.call_outs_pin, sizeof(Outputs) * gkq_handle_s::MAX_CALLS) == cudaSuccess;
        if (ok) memset(R.rctl_host, 0, sizeof(RecCtl));
        ev(&R.ev_bulk);
        ev(&R.ev_tail);
    }
    strm(&h->s_tail);
    strm(&h->s_in);
    strm(&h->s_compute);
    strm(&h->s_compute2);
    strm(&h->s_out);
    strm(&h->s_out2);
    ev(&h->ev_c2);
    for (int k = 0; k < gkq_handle_s::MAX_CHUNKS; ++k) {
        ev(&h->ev_in[k]);
        ev(&h->ev_done[k]);
        ev(&h->ev_out[k]);
    }
    ok = ok && cudaMallocHost(&h->fixed_pin, 64 * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMallocHost(&h->flat_ctl_host, 2 * sizeof(Ctl2)) == cudaSuccess;
    ev(&h->ev_flat[0]);
    ev(&h->ev_flat[1]);
    if (!ok) {
        gkq_destroy(h);
        return fail(nullptr, GKQ_ERR_CUDA, "handle allocation failed");
    }
    *out = h;
    return GKQ_OK;
}

int gkq_destroy(gkq_handle_t h) {
    if (!h) return GKQ_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    gkq_comm_destroy(h);
    for (int c = 0; c < 2; ++c) {
        gkq_handle_s::Scratch& s = h->sc[c];
        cudaFree(s.ctl);
        cudaFreeHost(s.ctl_host);
        cudaFree(s.park);
        cudaFree(s.list25);
        cudaFree(s.listS);
        cudaFree(s.listP);
        cudaFree(s.listLoop[0]);
        cudaFree(s.listLoop[1]);
        cudaFree(s.slab_d);
        cudaFree(s.slab_w);
        cudaFree(s.dev_pts);
        cudaFreeHost(s.host_pts);
        if (s.side) cudaStreamDestroy(s.side);
        if (s.ev_fork) cudaEventDestroy(s.ev_fork);
        if (s.ev_join) cudaEventDestroy(s.ev_join);
        gkq_handle_s::RecBuf& R = h->rb[c];
        cudaFree(R.recs);
        cudaFree(R.rctl);
        cudaFreeHost(R.rctl_host);
        cudaFree(R.call_outs);
        cudaFreeHost(R.call_outs_pin);
        cudaFree(R.slab_d);
        cudaFree(R.slab_w);
        cudaFree(R.patch_dev);
        cudaFreeHost(R.patch_pin);
        if (R.ev_bulk) cudaEventDestroy(R.ev_bulk);
        if (R.ev_tail) cudaEventDestroy(R.ev_tail);
    }
    cudaFree(h->stage_dev);
    cudaFree(h->flat_buf);
    cudaFreeHost(h->flat_ctl_host);
    for (int k = 0; k < 2; ++k)
        if (h->ev_flat[k]) cudaEventDestroy(h->ev_flat[k]);
    cudaFreeHost(h->fixed_pin);
    for (int k = 0; k < gkq_handle_s::MAX_CHUNKS; ++k) {
        if (h->ev_in[k]) cudaEventDestroy(h->ev_in[k]);
        if (h->ev_done[k]) cudaEventDestroy(h->ev_done[k]);
        if (h->ev_out[k]) cudaEventDestroy(h->ev_out[k]);
    }
    cudaFreeHost(h->pk_pin);
    if (h->ev_c2) cudaEventDestroy(h->ev_c2);
    prof_clear(h);
    for (auto e : h->prof_pool) cudaEventDestroy(e);
    for (cudaStream_t s : {h->s_tail, h->s_in, h->s_compute, h->s_compute2, h->s_out, h->s_out2})
        if (s) cudaStreamDestroy(s);
    delete h;
    return GKQ_OK;
}

const char* gkq_last_error(gkq_handle_t h) { return h ? h->err.c_str() : g_create_error.c_str(); }

static const FunctorInfo* table_of(int dim, int* count) {
    switch (dim) {
        case 1: *count = F1_COUNT; return F1_TABLE;
        case 2: *count = F2_COUNT; return F2_TABLE;
        case 0: *count = YR_COUNT; return YR_TABLE;
        default: *count = 0; return nullptr;
    }
}
int gkq_functor_count(int dim) {
    int c;
    return table_of(dim, &c) ? c : GKQ_ERR_INVALID_ARGUMENT;
}
int gkq_functor_info(int dim, int id, const char** name, int* nparams, double* flops_per_eval) {
    int c;
    const FunctorInfo* t = table_of(dim, &c);
    if (!t || id < 0 || id >= c) return GKQ_ERR_INVALID_ARGUMENT;
    if (name) *name = t[id].name;
    if (nparams) *nparams = t[id].nparams;
    if (flops_per_eval) *flops_per_eval = t[id].flops;
    return GKQ_OK;
}
int gkq_functor_lookup(int dim, const char* name) {
    int c;
    const FunctorInfo* t = table_of(dim, &c);
    if (!t || !name) return GKQ_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < c; ++i)
        if (strcmp(t[i].name, name) == 0) return i;
    return GKQ_ERR_INVALID_ARGUMENT;
}

int gkq_join(gkq_handle_t h, void* stream) {
    if (!h) return GKQ_ERR_INVALID_ARGUMENT;
    GKQ_CUDA(h, cudaSetDevice(h->device));
    int rc = flush_records(h, (cudaStream_t)stream, false);
    if (rc) return rc;
    // remainders flushed earlier on other streams: order `stream` behind them too (the flags stay
    // set: the buffer's next user still has to order its own stream; waiting twice is harmless)
    for (int c = 0; c < 2; ++c)
        if (h->rb[c].tail_pending) GKQ_CUDA(h, cudaStreamWaitEvent((cudaStream_t)stream, h->rb[c].ev_tail, 0));
    return GKQ_OK;
}

int gkq_set_deferred_join(gkq_handle_t h, int enable) {
    if (!h) return GKQ_ERR_INVALID_ARGUMENT;
    if (!enable && h->deferred) {
        // leaving deferred mode: nothing may stay pending
        int rc = gkq_join(h, nullptr);
        if (rc) return rc;
        GKQ_CUDA(h, cudaStreamSynchronize(nullptr));
        h->rb[0].tail_pending = h->rb[1].tail_pending = false;
    }
    h->deferred = enable != 0;
    return GKQ_OK;
}

int gkq_sync(gkq_handle_t h, void* stream) {
    if (!h) return GKQ_ERR_INVALID_ARGUMENT;
    int rc = gkq_join(h, stream);
    if (rc) return rc;
    GKQ_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream));
    h->rb[0].tail_pending = h->rb[1].tail_pending = false;
    gkq_handle_s::Scratch& S = h->sc[0];
    if (S.ctl_pending) {
        const Control* c = S.ctl_host;
        h->stats.last_survivors_qk25 = c->count25;
        h->stats.last_survivors_loop = c->chain[0].countLoop + c->chain[1].countLoop;
        for (int k = 0; k < 2; ++k) h->stats.last_chain_loop[k] = c->chain[k].countLoop;
        // records of the last profiled batch (profiling flushes after every call)
        h->stats.last_survivors_step2 = h->rb[h->rb_cur ^ 1].rctl_host->count;
        S.ctl_pending = false;
    }
    return GKQ_OK;
}

int gkq_get_stats(gkq_handle_t h, gkq_stats* out) {
    if (!h || !out) return GKQ_ERR_INVALID_ARGUMENT;
    *out = h->stats;
    return GKQ_OK;
}
int gkq_reset_stats(gkq_handle_t h) {
    if (!h) return GKQ_ERR_INVALID_ARGUMENT;
    uint64_t keep = h->stats.scratch_bytes;
    memset(&h->stats, 0, sizeof h->stats);
    h->stats.scratch_bytes = keep;
    return GKQ_OK;
}

int gkq_set_profiling(gkq_handle_t h, int enable) {
    if (!h) return GKQ_ERR_INVALID_ARGUMENT;
    h->profiling = enable != 0;
    if (!h->profiling) prof_clear(h);
    return GKQ_OK;
}

int gkq_get_profile(gkq_handle_t h, int max_entries, const char** names, float* ms) {
    if (!h || max_entries < 0) return GKQ_ERR_INVALID_ARGUMENT;
    int k = 0;
    for (auto& pe : h->prof) {
        if (k >= max_entries) break;
        if (cudaEventSynchronize(pe.e1) != cudaSuccess) return fail(h, GKQ_ERR_CUDA, "profile event sync failed");
        float t = 0.f;
        if (cudaEventElapsedTime(&t, pe.e0, pe.e1) != cudaSuccess) return fail(h, GKQ_ERR_CUDA, "profile event read failed");
        if (names) names[k] = pe.name;
        if (ms) ms[k] = t;
        ++k;
    }
    return k;
}

int gkq_measure_fp64_peak(gkq_handle_t h, int repeats, double* flops_per_s, double* ms_out) {
    if (!h || !flops_per_s) return GKQ_ERR_INVALID_ARGUMENT;
    GKQ_CUDA(h, cudaSetDevice(h->device));
    if (repeats < 1) repeats = 1;
    const int iters = 1 << 16;
    const int grid = h->sm_count * 8;
    cudaEvent_t e0, e1;
    GKQ_CUDA(h, cudaEventCreate(&e0));
    GKQ_CUDA(h, cudaEventCreate(&e1));
    double best = 1e30;
    for (int r = 0; r < repeats + 1; ++r) {  // first trip is a warm-up
        GKQ_CUDA(h, cudaEventRecord(e0, 0));
        k_dfma_peak<<<grid, BLOCK>>>((double*)h->sc[1].ctl, iters, 1.0 + r);
        GKQ_CUDA(h, cudaEventRecord(e1, 0));
        GKQ_CUDA(h, cudaEventSynchronize(e1));
        float ms = 0;
        GKQ_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0 && ms < best) best = ms;
    }
    int rc = launch_check(h, "k_dfma_peak");
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc) return rc;
    const double flops = (double)grid * BLOCK * 8.0 * (double)iters * 2.0;
    *flops_per_s = flops / (best * 1e-3);
    if (ms_out) *ms_out = best;
    return GKQ_OK;
}

// ---- 1-D ----------------------------------------------------------------------------------------
// Bulk kernels of one QAGS batch on `st` (chain 1 forks to the context's side stream).  Appends
// straggler records to the current record buffer; `patch`: host-pipeline mode.
static int run_qags_bulk(gkq_handle_t h, const Launch1D& L, int fid, Batch1D P, cudaStream_t st, bool patch) {
    int rc;
    gkq_handle_s::Scratch& S = h->sc[h->cur];
    if (P.max_evals < 17) {  // qags.rs:86-88 -- only reachable without a worklist split
        const int grid = (int)((P.n + BLOCK - 1) / BLOCK);
        if (P.worklist) {  // the QAGS subset of a CSR AUTO batch (the QAGP subset applies qagp.rs:79-81 itself)
            k_fill_result_list<<<grid < 1184 ? grid : 1184, BLOCK, 0, st>>>(P.worklist, P.worklist_count, P.out, 0.0,
                                                                           F64_MAX, 0u, ST_INSUFFICIENT_ITERATION);
            return launch_check(h, "k_fill_result_list");
        }
        k_fill_result<<<grid, BLOCK, 0, st>>>(P.n, P.out, 0.0, F64_MAX, 0u, ST_INSUFFICIENT_ITERATION);
        return launch_check(h, "k_fill_result");
    }
    unsigned int slot = 0;
    if ((rc = records_begin(h, st, fid, P, (size_t)P.n, patch, slot))) return rc;
    gkq_handle_s::RecBuf& R = h->rb[h->rb_cur];
    P.recs = R.recs;
    P.rctl = R.rctl;
    P.call_slot = slot;
    P.call_outs_w = patch ? nullptr : R.call_outs;

    const long long nblocks = (P.n + IBLOCK - 1) / IBLOCK;
    prof_begin(h, "k_qags_init<qk17>", st);
    L.qags_init17(P, (int)nblocks, st);
    prof_end(h, st);
    if ((rc = launch_check(h, "k_qags_init<qk17>"))) return rc;

    if (h->occ_init25[fid] < 0) h->occ_init25[fid] = L.occupancy_init25();
    long long g25 = (long long)h->sm_count * (h->occ_init25[fid] > 0 ? h->occ_init25[fid] : 1);
    if (g25 > nblocks) g25 = nblocks;
    // the first-step kernel has its own shape: a persistent grid of 128-thread blocks that holds
    // (nearly) all loop entrants of a typical batch at once; the few threads that take a second
    // grid-stride trip overlap the other kernels in flight
    if (h->occ_first[fid] < 0) h->occ_first[fid] = L.occupancy_first();
    long long gfirst = (long long)h->sm_count * (h->occ_first[fid] > 0 ? h->occ_first[fid] : 1);
    const long long nb_first = (P.n + FBLOCK - 1) / FBLOCK;
    if (gfirst > nb_first) gfirst = nb_first;

    // fork: chain 1 (qk25 pass + its first steps) runs on the side stream.  In profiling mode
    // everything stays on one stream so that the per-kernel event brackets are clean.
    const bool serial = h->profiling || patch;  // (host pipeline: chunks already overlap on two streams)
    if (!serial) {
        GKQ_CUDA(h, cudaEventRecord(S.ev_fork, st));
        GKQ_CUDA(h, cudaStreamWaitEvent(S.side, S.ev_fork, 0));
    }
    for (int c = 0; c < 2; ++c) {
        cudaStream_t cs = (c == 0 || serial) ? st : S.side;
        Batch1D Q = P;
        Q.chain = c;
        if (c == 1) {
            prof_begin(h, "k_qags_init<qk25>", cs);
            L.qags_init25(Q, (int)g25, cs);
            prof_end(h, cs);
            if ((rc = launch_check(h, "k_qags_init<qk25>"))) return rc;
        }
        prof_begin(h, c == 0 ? "k_qags_first[0]" : "k_qags_first[1]", cs);
        L.qags_first(Q, (int)gfirst, cs);
        prof_end(h, cs);
        if ((rc = launch_check(h, "k_qags_first"))) return rc;
    }
    if (!serial) {
        GKQ_CUDA(h, cudaEventRecord(S.ev_join, S.side));
        GKQ_CUDA(h, cudaStreamWaitEvent(st, S.ev_join, 0));
    }
    R.ncalls += 1;
    R.n_bound += (size_t)P.n;
    if (!patch) GKQ_CUDA(h, cudaEventRecord(R.ev_bulk, st));  // (the host pipeline orders its own flush)
    return GKQ_OK;
}

static int run_qagp(gkq_handle_t h, const Launch1D& L, int fid, Batch1D P, int npts_max,
                    cudaStream_t st) {
    int rc;
    if (h->occ_qagp[fid] < 0) h->occ_qagp[fid] = L.occupancy_qagp();
    // list length <= nint + (max_evals - 25 nint)/50 <= nint + max_evals/50
    const unsigned long long need = (unsigned long long)npts_max + 1ull + P.max_evals / 50ull + 1ull;
    if (need > (1ull << 26)) return fail(h, GKQ_ERR_OUT_OF_MEMORY, "max_evals too large for the interval slabs");
    const int cap = (int)need;
    const long long nblocks = (P.n + ABLOCK - 1) / ABLOCK;
    int grid = persistent_grid(h, h->occ_qagp[fid], cap, npts_max + 2, ABLOCK);
    if (grid < 1) return fail(h, GKQ_ERR_OUT_OF_MEMORY, "max_evals too large for the interval slabs");
    if ((long long)grid > nblocks) grid = (int)nblocks;
    if ((rc = ensure_slab(h, (unsigned long long)grid * ABLOCK, cap, npts_max + 2, P.slab))) return rc;
    prof_begin(h, "k_adaptive<QAGP>", st);
    L.qagp(P, grid, st);
    prof_end(h, st);
    return launch_check(h, "k_adaptive<QAGP>");
}

// QAG: one thread per integral over a persistent grid (deprecated algorithm: available, not tuned).
static int run_qag(gkq_handle_t h, const Launch1D& L, Batch1D P, cudaStream_t st) {
    int rc;
    const unsigned long long need = P.max_evals / 50ull + 2ull;
    if (need > (1ull << 26)) return fail(h, GKQ_ERR_OUT_OF_MEMORY, "max_evals too large for the interval slabs");
    const int cap = (int)need;
    const long long nblocks = (P.n + ABLOCK - 1) / ABLOCK;
    int grid = persistent_grid(h, 4, cap, 2, ABLOCK);
    if (grid < 1) return fail(h, GKQ_ERR_OUT_OF_MEMORY, "max_evals too large for the interval slabs");
    if ((long long)grid > nblocks) grid = (int)nblocks;
    if ((rc = ensure_slab(h, (unsigned long long)grid * ABLOCK, cap, 2, P.slab))) return rc;
    prof_begin(h, "k_qag", st);
    L.qag(P, grid, st);
    prof_end(h, st);
    return launch_check(h, "k_qag");
}

// The batch call proper, on bulk scratch context h->cur.  patch / idx_offset: host pipeline.
static int integrate_batch_ctx(gkq_handle_t h, const gkq_config* cfg, int functor_id, int64_t n,
                               const double* a, const double* b, int64_t range_stride,
                               const double* params, int64_t param_stride, const int64_t* pts_offsets,
                               const double* pts, double* out_estimate, double* out_delta,
                               uint32_t* out_nevals, uint32_t* out_status, void* stream, bool patch,
                               unsigned long long idx_offset, Control* ctl_zeroed = nullptr,
                               const gkq_overrides* ov = nullptr, long long ov_offset = 0) {
    if (!h) return GKQ_ERR_INVALID_ARGUMENT;
    if (!cfg) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "cfg is NULL");
    if (functor_id < 0 || functor_id >= F1_COUNT) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "unknown functor id");
    if (n < 0 || n > 0xffffffffll) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "n out of range");
    if (cfg->algo < 0 || cfg->algo > 3) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "unknown algorithm");
    if (range_stride != 0 && range_stride != 1) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "range_stride must be 0 or 1");
    if (param_stride != 0 && param_stride < n) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "param_stride must be 0 or >= n");
    if (cfg->max_evals > 0x7fffffffull) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "max_evals must fit in 31 bits");
    if (cfg->npoints < 0 || (cfg->npoints > 0 && !cfg->points)) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "points is NULL");
    int rc;
    if ((rc = check_tol(h, cfg->tol))) return rc;
    for (int64_t k = 0; k < cfg->npoints; ++k)
        if (cfg->points[k] != cfg->points[k])  // integrator.rs:86-90
            return fail(h, GKQ_ERR_NAN_INPUT, "cannot include NAN value in singular points");
    if (n == 0) return GKQ_OK;
    if (!a || !b || !out_estimate || !out_delta || !out_nevals || !out_status)
        return fail(h, GKQ_ERR_INVALID_ARGUMENT, "NULL array argument");
    if (F1_TABLE[functor_id].nparams > 0 && !params) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "params is NULL");
    if (pts_offsets && !pts) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "pts is NULL");
    if (cfg->max_evals > h->max_evals_seen) h->max_evals_seen = cfg->max_evals;

    GKQ_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    gkq_handle_s::Scratch& S = h->sc[h->cur];
    if ((rc = ensure_batch_scratch(h, (size_t)n))) return rc;

    Batch1D P;
    memset(&P, 0, sizeof P);
    P.n = n;
    P.a = a;
    P.b = b;
    P.range_stride = range_stride;
    P.params = params;
    P.param_stride = param_stride;
    P.tol.kind = cfg->tol.kind;
    P.tol.abs_tol = cfg->tol.abs_tol;
    P.tol.rel_tol = cfg->tol.rel_tol;
    P.max_evals = cfg->max_evals;
    P.ws_capacity = cfg->workspace_capacity;
    if (ov) {  // per-integral tolerance / budget (device arrays)
        P.abs_tol_each = ov->abs_tol ? ov->abs_tol + ov_offset : nullptr;
        P.rel_tol_each = ov->rel_tol ? ov->rel_tol + ov_offset : nullptr;
        P.max_evals_each = ov->max_evals ? (const unsigned long long*)ov->max_evals + ov_offset : nullptr;
    }
    P.out.estimate = out_estimate;
    P.out.delta = out_delta;
    P.out.nevals = out_nevals;
    P.out.status = out_status;
    P.ctl = ctl_zeroed ? ctl_zeroed : S.ctl;
    P.park = S.park;
    P.list25 = S.list25;
    P.listLoop[0] = S.listLoop[0];
    P.listLoop[1] = S.listLoop[1];
    P.idx_offset = idx_offset;

    if (h->profiling) prof_clear(h);
    if (!ctl_zeroed) GKQ_CUDA(h, cudaMemsetAsync(S.ctl, 0, sizeof(Control), st));
    const Launch1D& L = h->l1[functor_id];
    h->stats.batches += 1;
    h->stats.integrals += (uint64_t)n;
    if ((cfg->flags & GKQ_FLAG_VALIDATE_INPUTS) && !patch)
        if ((rc = validate_device_inputs(h, st, n, range_stride ? n : 1, a, b, P.abs_tol_each, P.rel_tol_each))) return rc;

    bool did_qags = false;
    if (cfg->algo == GKQ_ALGO_QAG) {
        rc = run_qag(h, L, P, st);  // never looks at points (qag.rs:39-60)
    } else if (!pts_offsets) {
        const bool use_qags = cfg->algo == GKQ_ALGO_QAGS || (cfg->algo == GKQ_ALGO_AUTO && cfg->npoints == 0);
        if (use_qags) {
            rc = run_qags_bulk(h, L, functor_id, P, st, patch);
            did_qags = true;
        } else {
            if ((rc = upload_shared_points(h, cfg->points, (size_t)cfg->npoints, st))) return rc;
            P.pts = S.dev_pts;
            P.npts_shared = (int)cfg->npoints;
            rc = run_qagp(h, L, functor_id, P, (int)cfg->npoints, st);
        }
    } else if (cfg->algo == GKQ_ALGO_QAGS) {
        rc = run_qags_bulk(h, L, functor_id, P, st, patch);  // explicit QAGS ignores points
        did_qags = true;
    } else {
        // CSR points: split for AUTO and find the largest point count (one small sync)
        const int grid = (int)((n + BLOCK - 1) / BLOCK);
        const int mode = cfg->algo == GKQ_ALGO_AUTO ? 0 : 1;
        k_split_auto<<<grid, BLOCK, 0, st>>>(n, (const long long*)pts_offsets, S.listS, S.listP, S.ctl, mode);
        if ((rc = launch_check(h, "k_split_auto"))) return rc;
        GKQ_CUDA(h, cudaMemcpyAsync(S.ctl_host, S.ctl, sizeof(Control), cudaMemcpyDeviceToHost, st));
        GKQ_CUDA(h, cudaStreamSynchronize(st));
        const int npts_max = (int)S.ctl_host->max_pts;
        P.pts_offsets = (const long long*)pts_offsets;
        P.pts = pts;
        if (mode == 0) {
            if (S.ctl_host->countS > 0) {
                Batch1D PS = P;
                PS.worklist = S.listS;
                PS.worklist_count = &S.ctl->countS;
                if ((rc = run_qags_bulk(h, L, functor_id, PS, st, patch))) return rc;
                did_qags = true;
            }
            if (S.ctl_host->countP > 0) {
                Batch1D PP = P;
                PP.worklist = S.listP;
                PP.worklist_count = &S.ctl->countP;
                rc = run_qagp(h, L, functor_id, PP, npts_max, st);
            }
        } else {
            rc = run_qagp(h, L, functor_id, P, npts_max, st);
        }
    }
    if (rc) return rc;
    // ordered mode (the default): finish the stragglers of this batch right away on its stream
    if (did_qags && !patch && (!h->deferred || h->profiling))
        if ((rc = flush_records(h, st, false))) return rc;
    // worklist sizes for gkq_get_stats: read back only in profiling mode (keeps the hot call lean)
    if (h->profiling) {
        GKQ_CUDA(h, cudaMemcpyAsync(S.ctl_host, S.ctl, sizeof(Control), cudaMemcpyDeviceToHost, st));
        S.ctl_pending = true;
    }
    return GKQ_OK;
}

int gkq_integrate_batch_ex(gkq_handle_t h, const gkq_config* cfg, const gkq_overrides* ov,
                           int functor_id, int64_t n, const double* a, const double* b,
                           int64_t range_stride, const double* params, int64_t param_stride,
                           const int64_t* pts_offsets, const double* pts, double* out_estimate,
                           double* out_delta, uint32_t* out_nevals, uint32_t* out_status, void* stream) {
    if (!h) return GKQ_ERR_INVALID_ARGUMENT;
    h->cur = 0;
    return integrate_batch_ctx(h, cfg, functor_id, n, a, b, range_stride, params, param_stride,
                               pts_offsets, pts, out_estimate, out_delta, out_nevals, out_status, stream,
                               false, 0ull, nullptr, ov, 0);
}

int gkq_integrate_batch(gkq_handle_t h, const gkq_config* cfg, int functor_id, int64_t n,
                        const double* a, const double* b, int64_t range_stride,
                        const double* params, int64_t param_stride, const int64_t* pts_offsets,
                        const double* pts, double* out_estimate, double* out_delta,
                        uint32_t* out_nevals, uint32_t* out_status, void* stream) {
    return gkq_integrate_batch_ex(h, cfg, nullptr, functor_id, n, a, b, range_stride, params, param_stride,
                                  pts_offsets, pts, out_estimate, out_delta, out_nevals, out_status, stream);
}

int gkq_qk_batch(gkq_handle_t h, int rule, int functor_id, int64_t n, const double* a,
                 const double* b, int64_t range_stride, const double* params,
                 int64_t param_stride, double* out4, void* stream) {
    if (!h) return GKQ_ERR_INVALID_ARGUMENT;
    if (rule != 17 && rule != 25 && rule != 33 && rule != 41 && rule != 49 && rule != 57)
        return fail(h, GKQ_ERR_INVALID_ARGUMENT, "rule must be 17, 25, 33, 41, 49 or 57");
    if (functor_id < 0 || functor_id >= F1_COUNT) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "unknown functor id");
    if (n < 0 || !a || !b || !out4) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "bad argument");
    if (F1_TABLE[functor_id].nparams > 0 && !params) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "params is NULL");
    if (n == 0) return GKQ_OK;
    GKQ_CUDA(h, cudaSetDevice(h->device));
    Batch1D P;
    memset(&P, 0, sizeof P);
    P.n = n;
    P.a = a;
    P.b = b;
    P.range_stride = range_stride;
    P.params = params;
    P.param_stride = param_stride;
    const int grid = (int)((n + BLOCK - 1) / BLOCK);
    h->l1[functor_id].qk[(rule - 17) / 8](P, out4, grid, (cudaStream_t)stream);
    return launch_check(h, "k_qk_batch");
}

int gkq_functor_eval_batch(gkq_handle_t h, int functor_id, int64_t n, const double* x,
                           const double* params, int64_t param_stride, double* y, void* stream) {
    return gkq_functor_eval_batch_ex(h, functor_id, n, x, params, param_stride, y, 0, stream);
}

int gkq_functor_eval_batch_ex(gkq_handle_t h, int functor_id, int64_t n, const double* x,
                              const double* params, int64_t param_stride, double* y, int mode, void* stream) {
    if (!h) return GKQ_ERR_INVALID_ARGUMENT;
    if (mode != 0 && mode != 1) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "mode must be 0 or 1");
    if (functor_id < 0 || functor_id >= F1_COUNT) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "unknown functor id");
    if (n < 0 || !x || !y) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "bad argument");
    if (F1_TABLE[functor_id].nparams > 0 && !params) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "params is NULL");
    if (n == 0) return GKQ_OK;
    GKQ_CUDA(h, cudaSetDevice(h->device));
    Batch1D P;
    memset(&P, 0, sizeof P);
    P.n = n;
    P.params = params;
    P.param_stride = param_stride;
    const int grid = (int)((n + BLOCK - 1) / BLOCK);
    h->l1[functor_id].eval(P, x, y, mode, grid, (cudaStream_t)stream);
    return launch_check(h, "k_eval");
}

// Host-pointer entry: H2D -> bulk kernels -> D2H pipeline over chunks of the batch.
//   * the whole batch is staged on the device (40 B per integral), so all H2D copies are enqueued
//     first and run back to back; chunk sizes ramp up (32k, 64k, ... 256k) so the first results
//     leave early and later copies are large;
//   * chunks alternate between two compute streams / bulk scratch contexts and two D2H streams;
//     the bulk results of a chunk leave as soon as its bulk kernels end;
//   * the stragglers of ALL chunks are finished by ONE launch of the persistent kernel at the end;
//     their results reach the host as a compact patch that the CPU applies.
// Host-entry calls in flight per device (several host threads, one handle each, overlap one call's
// copies in with another's copies out): the chunk schedule follows it -- see the chunk schedule below.
namespace {
std::atomic<int> g_host_inflight[64];
struct HostInflight {
    int dev;
    int seen;
    explicit HostInflight(int d) : dev(d >= 0 && d < 64 ? d : 0) { seen = g_host_inflight[dev].fetch_add(1) + 1; }
    ~HostInflight() { g_host_inflight[dev].fetch_sub(1); }
};
}  // namespace

int gkq_integrate_batch_host(gkq_handle_t h, const gkq_config* cfg, int functor_id, int64_t n,
                             const double* a, const double* b, int64_t range_stride,
                             const double* params, int64_t param_stride,
                             const int64_t* pts_offsets, const double* pts, double* out_estimate,
                             double* out_delta, uint32_t* out_nevals, uint32_t* out_status) {
    return gkq_integrate_batch_host_ex(h, cfg, nullptr, functor_id, n, a, b, range_stride, params, param_stride,
                                       pts_offsets, pts, out_estimate, out_delta, out_nevals, out_status);
}

int gkq_integrate_batch_host_ex(gkq_handle_t h, const gkq_config* cfg, const gkq_overrides* ov_host,
                                int functor_id, int64_t n, const double* a, const double* b,
                                int64_t range_stride, const double* params, int64_t param_stride,
                                const int64_t* pts_offsets, const double* pts, double* out_estimate,
                                double* out_delta, uint32_t* out_nevals, uint32_t* out_status) {
    if (!h) return GKQ_ERR_INVALID_ARGUMENT;
    if (!cfg) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "cfg is NULL");
    if (functor_id < 0 || functor_id >= F1_COUNT) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "unknown functor id");
    if (n < 0) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "n out of range");
    if (n == 0) return GKQ_OK;
    if (!a || !b || !out_estimate || !out_delta || !out_nevals || !out_status)
        return fail(h, GKQ_ERR_INVALID_ARGUMENT, "NULL array argument");
    GKQ_CUDA(h, cudaSetDevice(h->device));
    const int np = F1_TABLE[functor_id].nparams;
    if (np > 0 && !params) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "params is NULL");
    if (pts_offsets) {
        if (!pts) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "pts is NULL");
        if (host_has_nan(pts, (size_t)pts_offsets[n]))  // integrator.rs:86-90
            return fail(h, GKQ_ERR_NAN_INPUT, "cannot include NAN value in singular points");
    }
    const int np_batch = param_stride ? np : 0;  // per-integral parameter rows
    int rc;
    // the data is visible to the CPU here: refuse what the reference refuses to construct
    // (Range::new, single/common.rs:94; Tolerance, single/integrator.rs:71)
    if (host_has_nan(a, range_stride ? (size_t)n : 1) || host_has_nan(b, range_stride ? (size_t)n : 1))
        return fail(h, GKQ_ERR_NAN_INPUT, "cannot create Range object from Range which contains NaN value.");
    if (ov_host) {
        const bool need_abs = cfg->tol.kind != GKQ_TOL_RELATIVE, need_rel = cfg->tol.kind != GKQ_TOL_ABSOLUTE;
        if ((need_abs && host_has_nan(ov_host->abs_tol, (size_t)n)) || (need_rel && host_has_nan(ov_host->rel_tol, (size_t)n)))
            return fail(h, GKQ_ERR_NAN_INPUT, "Tolerance must not contain NAN.");
    }
    // records pending from deferred device-entry calls are finished first (they own the buffers)
    if (h->rb[0].ncalls || h->rb[1].ncalls || h->rb[0].tail_pending || h->rb[1].tail_pending) {
        if ((rc = gkq_join(h, h->s_compute))) return rc;
        GKQ_CUDA(h, cudaStreamSynchronize(h->s_compute));
        h->rb[0].tail_pending = h->rb[1].tail_pending = false;
    }

    if (pts_offsets) {
        // per-integral break points (CSR): single-shot staging -- such batches are QAGP work (hundreds
        // of samples per integral), the copies are a small share of the call
        const size_t N = (size_t)n, npts = (size_t)pts_offsets[n];
        const size_t nr = range_stride ? N : 1, npar = param_stride ? N : 1;
        const int n_ov = ov_host ? (ov_host->abs_tol != nullptr) + (ov_host->rel_tol != nullptr) + (ov_host->max_evals != nullptr) : 0;
        const size_t in_d = 2 * nr + (size_t)np * npar + (N + 1) + npts + (size_t)n_ov * N;
        const size_t want = in_d * 8 + N * 24 + 256;
        if (want > h->stage_bytes) {
            if (h->stage_dev) GKQ_CUDA(h, cudaFree(h->stage_dev));
            h->stage_dev = nullptr;
            h->stats.scratch_bytes -= h->stage_bytes;
            h->stage_bytes = 0;
            GKQ_CUDA(h, cudaMalloc(&h->stage_dev, want));
            h->stage_bytes = want;
            h->stats.scratch_bytes += want;
        }
        cudaStream_t st = h->s_compute;
        double* d = (double*)h->stage_dev;
        double *d_a = d, *d_b = d + nr, *d_par = d + 2 * nr;
        long long* d_off = (long long*)(d_par + (size_t)np * npar);
        double* d_pts = (double*)(d_off + N + 1);
        double* q = d_pts + npts;
        gkq_overrides ov_dev = {nullptr, nullptr, nullptr};
        if (ov_host && ov_host->abs_tol) {
            ov_dev.abs_tol = q;
            GKQ_CUDA(h, cudaMemcpyAsync(q, ov_host->abs_tol, N * 8, cudaMemcpyHostToDevice, st));
            q += N;
        }
        if (ov_host && ov_host->rel_tol) {
            ov_dev.rel_tol = q;
            GKQ_CUDA(h, cudaMemcpyAsync(q, ov_host->rel_tol, N * 8, cudaMemcpyHostToDevice, st));
            q += N;
        }
        if (ov_host && ov_host->max_evals) {
            ov_dev.max_evals = (const uint64_t*)q;
            GKQ_CUDA(h, cudaMemcpyAsync(q, ov_host->max_evals, N * 8, cudaMemcpyHostToDevice, st));
            q += N;
        }
        double* d_est = q;
        double* d_dlt = d_est + N;
        uint32_t* d_nev = (uint32_t*)(d_dlt + N);
        uint32_t* d_st = d_nev + N;
        GKQ_CUDA(h, cudaMemcpyAsync(d_a, a, nr * 8, cudaMemcpyHostToDevice, st));
        GKQ_CUDA(h, cudaMemcpyAsync(d_b, b, nr * 8, cudaMemcpyHostToDevice, st));
        for (int k = 0; k < np; ++k)
            GKQ_CUDA(h, cudaMemcpyAsync(d_par + (size_t)k * npar, params + (size_t)k * (param_stride ? param_stride : 1),
                                        npar * 8, cudaMemcpyHostToDevice, st));
        GKQ_CUDA(h, cudaMemcpyAsync(d_off, pts_offsets, (N + 1) * 8, cudaMemcpyHostToDevice, st));
        if (npts) GKQ_CUDA(h, cudaMemcpyAsync(d_pts, pts, npts * 8, cudaMemcpyHostToDevice, st));
        const bool was_deferred = h->deferred;
        h->deferred = false;  // ordered: the call's stragglers are finished on `st` before the copies out
        h->cur = 0;
        rc = integrate_batch_ctx(h, cfg, functor_id, n, d_a, d_b, range_stride, d_par, param_stride ? (int64_t)N : 0,
                                 (const int64_t*)d_off, d_pts, d_est, d_dlt, d_nev, d_st, st, false, 0ull, nullptr,
                                 n_ov ? &ov_dev : nullptr, 0);
        h->deferred = was_deferred;
        if (rc) return rc;
        GKQ_CUDA(h, cudaMemcpyAsync(out_estimate, d_est, N * 8, cudaMemcpyDeviceToHost, st));
        GKQ_CUDA(h, cudaMemcpyAsync(out_delta, d_dlt, N * 8, cudaMemcpyDeviceToHost, st));
        GKQ_CUDA(h, cudaMemcpyAsync(out_nevals, d_nev, N * 4, cudaMemcpyDeviceToHost, st));
        GKQ_CUDA(h, cudaMemcpyAsync(out_status, d_st, N * 4, cudaMemcpyDeviceToHost, st));
        GKQ_CUDA(h, cudaStreamSynchronize(st));
        h->rb[0].tail_pending = h->rb[1].tail_pending = false;
        return GKQ_OK;
    }

    // chunk schedule.  A chunk's arrays cross the link as separate copies of 8 B x chunk: 2 MB copies
    // (256 k integrals) reach 75 GB/s with both directions busy, 4 MB copies 89 GB/s
    // (tools/pcie_aggregate_probe.py --sizes), but a bigger first chunk delays the call's first results.
    // One call alone is latency-bound (its copies out trail its copies in): 256 k chunks.  With three or
    // more host calls in flight on the device the link is what counts: 512 k chunks (S1, 1 M integrals,
    // 4 host threads: 0.585 -> 0.541 ms per call; 1 thread: 0.94 -> 1.09; tools/e2e_sweep.py,
    // profiles/e2e_sweep_r02.txt).
    constexpr int MAXC = gkq_handle_s::MAX_CHUNKS;
    const HostInflight inflight(h->device);
    int64_t CH = inflight.seen >= 3 ? (1 << 19) : (1 << 18), C0 = CH;
    if (const char* env = getenv("GKQ_HOST_CHUNK")) {  // tuning knobs (integrals per chunk)
        long long v = atoll(env);
        if (v >= 1024) CH = v;
    }
    if (const char* env = getenv("GKQ_HOST_CHUNK0")) {
        long long v = atoll(env);
        if (v >= 1024) C0 = v;
    }
    if (C0 > CH) C0 = CH;
    while ((n + CH - 1) / CH + 8 > MAXC) CH *= 2;
    int64_t c_off[MAXC], c_len[MAXC];
    int nchunks = 0;
    for (int64_t off = 0, next = C0; off < n; ++nchunks) {
        const int64_t m = (n - off) < next ? (n - off) : next;
        c_off[nchunks] = off;
        c_len[nchunks] = m;
        off += m;
        next = next * 2 > CH ? CH : next * 2;
    }

    // device staging of the whole batch: [64 shared doubles | nchunks Control blocks |
    //   a[n] b[n] (when per-integral) | params[np][n] | est[n] | delta[n] | nevals[n] | status[n]]
    const size_t N = (size_t)n;
    const size_t fixed_bytes = 64 * 8, ctl_bytes = sizeof(Control) * MAXC;
    const size_t in_doubles = (size_t)((range_stride ? 2 : 0) + np_batch) * N;
    const int n_ov = ov_host ? (ov_host->abs_tol != nullptr) + (ov_host->rel_tol != nullptr) + (ov_host->max_evals != nullptr) : 0;
    const size_t want = fixed_bytes + ctl_bytes + in_doubles * 8 + N * 24 + (size_t)n_ov * N * 8 + N * 4 + 256;
    if (want > h->stage_bytes) {
        if (h->stage_dev) GKQ_CUDA(h, cudaFree(h->stage_dev));
        h->stage_dev = nullptr;
        h->stats.scratch_bytes -= h->stage_bytes;
        h->stage_bytes = 0;
        GKQ_CUDA(h, cudaMalloc(&h->stage_dev, want));
        h->stage_bytes = want;
        h->stats.scratch_bytes += want;
    }
    double* d_fixed = (double*)h->stage_dev;
    Control* d_ctl = (Control*)(h->stage_dev + fixed_bytes);
    double* d_in = (double*)(h->stage_dev + fixed_bytes + ctl_bytes);
    double* d_a = range_stride ? d_in : d_fixed;
    double* d_b = range_stride ? d_in + N : d_fixed + 1;
    double* d_par = param_stride ? d_in + (range_stride ? 2 * N : 0) : d_fixed + 2;
    double* d_est = d_in + in_doubles;
    double* d_dlt = d_est + N;
    uint32_t* d_nev = (uint32_t*)(d_dlt + N);
    uint32_t* d_st = d_nev + N;
    // per-integral tolerance / budget arrays (when given) behind the outputs
    gkq_overrides ov_dev = {nullptr, nullptr, nullptr};
    uint32_t* d_pk = nullptr;
    {
        double* q = (double*)(d_st + N);
        if (ov_host && ov_host->abs_tol) { ov_dev.abs_tol = q; q += N; }
        if (ov_host && ov_host->rel_tol) { ov_dev.rel_tol = q; q += N; }
        if (ov_host && ov_host->max_evals) { ov_dev.max_evals = (const uint64_t*)q; q += N; }
        d_pk = (uint32_t*)q;
    }
    // Packed result words (see k_pack_nev_st): when the bulk results of a chunk are final -- every path
    // but the heavy-tail one, which ships the staged arrays whole at the end -- nevals and status cross
    // the link as one word and the calling thread unpacks chunk c while chunk c + 1 is in flight.
    const bool is_qags_call = cfg->algo == GKQ_ALGO_QAGS || (cfg->algo == GKQ_ALGO_AUTO && cfg->npoints == 0);
    const bool heavy_tail_expected = is_qags_call && h->host_tail_fid == functor_id && h->host_tail_share > 1.0 / 16.0;
    // ... when at least two host calls are in flight on the device: then the link is the bound and 4 fewer
    // bytes per integral win (S1, 1 M integrals, 2 / 4 host threads: 0.679 -> 0.612 / 0.631 -> 0.576 ms per
    // call).  A call that is alone is a serial chain with the calling thread's unpack pass at its end
    // (0.2 ms for 1 M integrals at one core's memory bandwidth): four plain copies are faster there
    // (1.02 -> 0.80 ms; profiles/e2e_sweep_r02.txt, second block).
    const char* pk_env = getenv("GKQ_HOST_UNPACKED");
    const bool pk_wanted = pk_env ? pk_env[0] == '0' : inflight.seen >= 2;
    const bool packed = !heavy_tail_expected && cfg->max_evals < (1ull << 29) && pk_wanted;
    if (packed && N > h->pk_cap) {
        if (h->pk_pin) GKQ_CUDA(h, cudaFreeHost(h->pk_pin));
        h->pk_pin = nullptr;
        h->pk_cap = 0;
        GKQ_CUDA(h, cudaMallocHost(&h->pk_pin, (N + N / 8) * sizeof(uint32_t)));
        h->pk_cap = N + N / 8;
    }

    // GKQ_HOST_TRACE=1: per-chunk timeline (timing events) printed to stderr -- tuning aid
    const bool trace = getenv("GKQ_HOST_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    auto mark = [&](cudaStream_t s) {
        if (!trace) return;
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        cudaEventRecord(ev, s);
        tev.push_back(ev);
    };
    mark(h->s_in);

    // ---- phase 1: every H2D copy, back to back on the copy-in stream --------------------------
    // shared (stride-0) inputs go through a pinned bounce buffer (the previous host call has
    // fully synchronised, so it is free)
    memset(h->fixed_pin, 0, 64 * sizeof(double));
    if (!range_stride) {
        h->fixed_pin[0] = a[0];
        h->fixed_pin[1] = b[0];
    }
    if (!param_stride)
        for (int k = 0; k < np; ++k) h->fixed_pin[2 + k] = params[k];
    GKQ_CUDA(h, cudaMemcpyAsync(d_fixed, h->fixed_pin, 64 * sizeof(double), cudaMemcpyHostToDevice, h->s_in));
    GKQ_CUDA(h, cudaMemsetAsync(d_ctl, 0, sizeof(Control) * (size_t)nchunks, h->s_in));
    if (ov_dev.abs_tol)
        GKQ_CUDA(h, cudaMemcpyAsync((void*)ov_dev.abs_tol, ov_host->abs_tol, N * 8, cudaMemcpyHostToDevice, h->s_in));
    if (ov_dev.rel_tol)
        GKQ_CUDA(h, cudaMemcpyAsync((void*)ov_dev.rel_tol, ov_host->rel_tol, N * 8, cudaMemcpyHostToDevice, h->s_in));
    if (ov_dev.max_evals)
        GKQ_CUDA(h, cudaMemcpyAsync((void*)ov_dev.max_evals, ov_host->max_evals, N * 8, cudaMemcpyHostToDevice, h->s_in));
    const bool is_qags = cfg->algo == GKQ_ALGO_QAGS || (cfg->algo == GKQ_ALGO_AUTO && cfg->npoints == 0);
    const int rbi = h->rb_cur;
    if (is_qags) {
        // the whole call appends to ONE record buffer, reset here (before any chunk can append)
        gkq_handle_s::RecBuf& R0 = h->rb[rbi];
        if (N > R0.cap)
            if ((rc = grow(h, R0.recs, R0.cap, N))) return rc;
        GKQ_CUDA(h, cudaMemsetAsync(R0.rctl, 0, sizeof(RecCtl), h->s_in));
    }
    for (int c = 0; c < nchunks; ++c) {
        const int64_t off = c_off[c], m = c_len[c];
        mark(h->s_in);
        if (range_stride) {
            GKQ_CUDA(h, cudaMemcpyAsync(d_a + off, a + off, (size_t)m * 8, cudaMemcpyHostToDevice, h->s_in));
            GKQ_CUDA(h, cudaMemcpyAsync(d_b + off, b + off, (size_t)m * 8, cudaMemcpyHostToDevice, h->s_in));
        }
        for (int k = 0; k < np_batch; ++k)
            GKQ_CUDA(h, cudaMemcpyAsync(d_par + (size_t)k * N + off, params + (size_t)k * param_stride + off,
                                        (size_t)m * 8, cudaMemcpyHostToDevice, h->s_in));
        GKQ_CUDA(h, cudaEventRecord(h->ev_in[c], h->s_in));
        mark(h->s_in);
    }
    // ---- phase 2: bulk kernels + result copies, chunk by chunk ------------------------------------
    rc = GKQ_OK;
    // Whatever way this function is left (every GKQ_CUDA below may return early), the handle must not
    // keep half-updated pipeline state: the "primed" flag, the current scratch context and -- on
    // failure -- the record buffer this call was appending to (drained streams, counters reset).
    struct HostCallGuard {
        gkq_handle_s* h;
        int rbi;
        bool ok = false;
        ~HostCallGuard() {
            h->host_records_primed = false;
            h->cur = 0;
            if (ok) return;
            for (cudaStream_t s : {h->s_in, h->s_compute, h->s_compute2, h->s_out, h->s_out2}) (void)cudaStreamSynchronize(s);
            (void)cudaGetLastError();
            for (int c = 0; c < 2; ++c) {
                gkq_handle_s::RecBuf& B = h->rb[c];
                B.ncalls = 0;
                B.n_bound = 0;
                B.flushed_calls = 0;
                B.tail_pending = false;
            }
            h->rb_cur = rbi;
        }
    } guard{h, rbi};
    h->host_records_primed = is_qags;  // records_begin must not reset the counters again
    for (int c = 0; c < nchunks && rc == GKQ_OK; ++c) {
        const int64_t off = c_off[c], m = c_len[c];
        const int ctx = c & 1;
        cudaStream_t sc = ctx ? h->s_compute2 : h->s_compute;
        cudaStream_t so = ctx ? h->s_out2 : h->s_out;
        GKQ_CUDA(h, cudaStreamWaitEvent(sc, h->ev_in[c], 0));
        mark(sc);
        h->cur = ctx;
        rc = integrate_batch_ctx(h, cfg, functor_id, m, range_stride ? d_a + off : d_a,
                                 range_stride ? d_b + off : d_b, range_stride,
                                 param_stride ? d_par + off : d_par, param_stride ? (int64_t)N : 0, nullptr,
                                 nullptr, d_est + off, d_dlt + off, d_nev + off, d_st + off, sc,
                                 /*patch=*/true, (unsigned long long)off, d_ctl + c, n_ov ? &ov_dev : nullptr,
                                 (long long)off);
        if (rc) break;
        GKQ_CUDA(h, cudaEventRecord(h->ev_done[c], sc));
        mark(sc);
        // the bulk results leave right away
        GKQ_CUDA(h, cudaStreamWaitEvent(so, h->ev_done[c], 0));
        mark(so);
        GKQ_CUDA(h, cudaMemcpyAsync(out_estimate + off, d_est + off, (size_t)m * 8, cudaMemcpyDeviceToHost, so));
        GKQ_CUDA(h, cudaMemcpyAsync(out_delta + off, d_dlt + off, (size_t)m * 8, cudaMemcpyDeviceToHost, so));
        if (packed) {
            k_pack_nev_st<<<(unsigned)((m + BLOCK - 1) / BLOCK), BLOCK, 0, so>>>(m, d_nev + off, d_st + off, d_pk + off);
            if ((rc = launch_check(h, "k_pack_nev_st"))) break;
            GKQ_CUDA(h, cudaMemcpyAsync(h->pk_pin + off, d_pk + off, (size_t)m * 4, cudaMemcpyDeviceToHost, so));
            GKQ_CUDA(h, cudaEventRecord(h->ev_out[c], so));
        } else {
            GKQ_CUDA(h, cudaMemcpyAsync(out_nevals + off, d_nev + off, (size_t)m * 4, cudaMemcpyDeviceToHost, so));
            GKQ_CUDA(h, cudaMemcpyAsync(out_status + off, d_st + off, (size_t)m * 4, cudaMemcpyDeviceToHost, so));
        }
        mark(so);
    }
    h->host_records_primed = false;
    h->cur = 0;

    // Unpack chunk c (host) while the copies of the later chunks and the remainder kernel are in flight.
    // Runs once, after everything has been ENQUEUED and before the first blocking wait of phase 3; the
    // stragglers' patch is applied after it.
    bool unpacked = false;
    auto unpack_chunks = [&]() {
        if (!packed || unpacked || rc != GKQ_OK) return;
        unpacked = true;
        for (int c = 0; c < nchunks; ++c) {
            if (cudaEventSynchronize(h->ev_out[c]) != cudaSuccess) return;  // (reported by the stream checks below)
            const uint32_t* pk = h->pk_pin + c_off[c];
            uint32_t* nv = out_nevals + c_off[c];
            uint32_t* sv = out_status + c_off[c];
            const int64_t m = c_len[c];
            for (int64_t i = 0; i < m; ++i) {
                const uint32_t w = pk[i];
                nv[i] = w >> 3;
                sv[i] = w & 7u;
            }
        }
    };

    // ---- phase 3: the stragglers of every chunk: one launch, results into the patch ---------------
    unsigned long long nrec = 0;
    gkq_handle_s::RecBuf& R = h->rb[rbi];
    const bool have_records = rc == GKQ_OK && is_qags && h->rb_cur == rbi && R.ncalls > 0;
    bool heavy_tail = false;
    if (have_records) {
        GKQ_CUDA(h, cudaEventRecord(h->ev_c2, h->s_compute2));
        GKQ_CUDA(h, cudaStreamWaitEvent(h->s_compute, h->ev_c2, 0));
        // How many stragglers does the bulk leave?  A handful (smooth integrands) travels as a compact
        // patch that the CPU applies; when a large share of the batch needs the persistent kernel
        // (e.g. every integral of an infinite-range batch takes > 1 bisection step) patching a million
        // scattered entries on one host core would dominate the call: the remainder then writes into
        // the staged result arrays and those are copied out once more as a whole.  Both paths are
        // correct for any count; which one is taken follows the share the PREVIOUS host call of this
        // functor saw (asking the device now would stall the enqueue behind the bulk: measured -5 %
        // on the smooth headline workload).
        heavy_tail = heavy_tail_expected;
    }
    if (have_records && heavy_tail) {
        const Outputs whole = {d_est, d_dlt, d_nev, d_st};  // record indices are batch-wide (idx_offset)
        for (int c = 0; c < R.ncalls; ++c) R.call_outs_pin[c] = whole;
        GKQ_CUDA(h, cudaMemcpyAsync(R.call_outs, R.call_outs_pin, sizeof(Outputs) * (size_t)R.ncalls,
                                    cudaMemcpyHostToDevice, h->s_compute));
        GKQ_CUDA(h, cudaEventRecord(R.ev_bulk, h->s_compute));  // (flush_records orders itself behind it)
        if ((rc = flush_records(h, h->s_compute, false)) == GKQ_OK) {
            // the whole-array copies go last on s_out: behind the remainder and behind the chunk copies
            // of both copy-out streams, which target the same host arrays
            GKQ_CUDA(h, cudaEventRecord(h->ev_c2, h->s_out2));
            GKQ_CUDA(h, cudaStreamWaitEvent(h->s_out, h->ev_c2, 0));
            GKQ_CUDA(h, cudaStreamWaitEvent(h->s_out, R.ev_tail, 0));
            GKQ_CUDA(h, cudaMemcpyAsync(out_estimate, d_est, N * 8, cudaMemcpyDeviceToHost, h->s_out));
            GKQ_CUDA(h, cudaMemcpyAsync(out_delta, d_dlt, N * 8, cudaMemcpyDeviceToHost, h->s_out));
            GKQ_CUDA(h, cudaMemcpyAsync(out_nevals, d_nev, N * 4, cudaMemcpyDeviceToHost, h->s_out));
            GKQ_CUDA(h, cudaMemcpyAsync(out_status, d_st, N * 4, cudaMemcpyDeviceToHost, h->s_out));
            GKQ_CUDA(h, cudaMemcpyAsync(R.rctl_host, R.rctl, sizeof(RecCtl), cudaMemcpyDeviceToHost, h->s_compute));
            R.tail_pending = false;  // (the stream synchronisations below cover it)
            R.flushed_calls = 0;
        }
    } else if (have_records) {
        if ((rc = flush_records(h, h->s_compute, true)) == GKQ_OK) {
            GKQ_CUDA(h, cudaMemcpyAsync(R.rctl_host, R.rctl, sizeof(RecCtl), cudaMemcpyDeviceToHost, h->s_compute));
            // most batches leave a handful of stragglers: fetch a first slice together with the count
            const size_t first = R.patch_cap < 4096 ? R.patch_cap : 4096;
            PatchView dv = patch_view(R.patch_dev, R.patch_cap), hv = patch_view(R.patch_pin, R.patch_cap);
            auto fetch = [&](size_t lo, size_t cnt) -> cudaError_t {
                cudaError_t e;
                if ((e = cudaMemcpyAsync(hv.idx + lo, dv.idx + lo, cnt * 8, cudaMemcpyDeviceToHost, h->s_compute))) return e;
                if ((e = cudaMemcpyAsync(hv.est + lo, dv.est + lo, cnt * 8, cudaMemcpyDeviceToHost, h->s_compute))) return e;
                if ((e = cudaMemcpyAsync(hv.dlt + lo, dv.dlt + lo, cnt * 8, cudaMemcpyDeviceToHost, h->s_compute))) return e;
                if ((e = cudaMemcpyAsync(hv.nev + lo, dv.nev + lo, cnt * 4, cudaMemcpyDeviceToHost, h->s_compute))) return e;
                return cudaMemcpyAsync(hv.st + lo, dv.st + lo, cnt * 4, cudaMemcpyDeviceToHost, h->s_compute);
            };
            GKQ_CUDA(h, fetch(0, first));
            unpack_chunks();
            GKQ_CUDA(h, cudaStreamSynchronize(h->s_compute));
            nrec = R.rctl_host->count;
            if (nrec > first) {
                GKQ_CUDA(h, fetch(first, (size_t)nrec - first));
                GKQ_CUDA(h, cudaStreamSynchronize(h->s_compute));
            }
            R.tail_pending = false;
        }
    }
    unpack_chunks();
    cudaError_t e1 = cudaStreamSynchronize(h->s_in);
    cudaError_t e2 = cudaStreamSynchronize(h->s_compute);
    cudaError_t e4 = cudaStreamSynchronize(h->s_compute2);
    cudaError_t e3 = cudaStreamSynchronize(h->s_out);
    if (e3 == cudaSuccess) e3 = cudaStreamSynchronize(h->s_out2);
    if (rc) return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || e4 != cudaSuccess)
        return fail(h, GKQ_ERR_CUDA, "host pipeline failed: a stream reported an error");
    if (have_records) {  // what this call saw steers the next one
        h->host_tail_fid = functor_id;
        h->host_tail_share = (double)R.rctl_host->count / (double)N;
    }
    if (have_records && !heavy_tail) {
        PatchView hv = patch_view(R.patch_pin, R.patch_cap);
        for (unsigned long long k = 0; k < nrec; ++k) {
            const unsigned long long i = hv.idx[k];
            out_estimate[i] = hv.est[k];
            out_delta[i] = hv.dlt[k];
            out_nevals[i] = hv.nev[k];
            out_status[i] = hv.st[k];
        }
    }
    guard.ok = true;
    if (trace && !tev.empty()) {
        // events: [0] start; per chunk h2d_begin, h2d_end (phase 1); then per chunk compute_begin,
        // compute_end, d2h_begin, d2h_end (phase 2)
        auto us = [&](size_t k) {
            float t = 0;
            cudaEventElapsedTime(&t, tev[0], tev[k]);
            return t * 1e3f;
        };
        for (int c = 0; c < nchunks; ++c) {
            const size_t p1 = 1 + 2 * (size_t)c, p2 = 1 + 2 * (size_t)nchunks + 4 * (size_t)c;
            if (p2 + 3 >= tev.size()) break;
            fprintf(stderr, "chunk %d (%lld): h2d %.0f-%.0f  compute %.0f-%.0f  d2h %.0f-%.0f us\n", c,
                    (long long)c_len[c], us(p1), us(p1 + 1), us(p2), us(p2 + 1), us(p2 + 2), us(p2 + 3));
        }
        fprintf(stderr, "stragglers patched on the host: %llu\n", nrec);
        for (auto ev : tev) cudaEventDestroy(ev);
    }
    return GKQ_OK;
}

// ---- 2-D ----------------------------------------------------------------------------------------
}  // extern "C"
namespace gkq {
void launch_flat_plan(bool qagp, const Batch2D& P, int round, int grid, cudaStream_t st) {
    if (qagp)
        k2_plan<true><<<grid, FB, 0, st>>>(P, round);
    else
        k2_plan<false><<<grid, FB, 0, st>>>(P, round);
}
void launch_flat_commit(bool qagp, const Batch2D& P, int round, int grid, cudaStream_t st) {
    if (qagp)
        k2_commit<true><<<grid, FB, 0, st>>>(P, round);
    else
        k2_commit<false><<<grid, FB, 0, st>>>(P, round);
}
}  // namespace gkq

// Bytes of level-synchronous scratch per outer integral (everything but the inner lists).
static size_t flat_bytes_per_outer(int cap_outer, int npts_max) {
    return sizeof(Lane) + sizeof(OuterCtx) + 3 * 4 + 50 * (8 + 4 * 4) + slab_doubles(cap_outer, npts_max + 2) * 8 +
           slab_ints(cap_outer) * 4 + 64;
}

// The nested drivers as rounds of flat kernels (kernels_2d_flat.cuh).  Enqueues rounds on `st` until an
// active count of zero has come back; the host runs one round ahead of the count it waits for.
static int run_nested_flat(gkq_handle_t h, const Launch2D& L, int fid2, int alg, Batch2D Q, int64_t n, cudaStream_t st) {
    int rc;
    const size_t N = (size_t)n;
    // carve the per-outer scratch
    const size_t per = flat_bytes_per_outer(Q.cap_outer, Q.npts_max);
    const size_t want = per * N + 4096;
    if (want > h->flat_bytes) {
        GKQ_CUDA(h, cudaStreamSynchronize(st));
        if (h->flat_buf) GKQ_CUDA(h, cudaFree(h->flat_buf));
        h->flat_buf = nullptr;
        h->stats.scratch_bytes -= h->flat_bytes;
        h->flat_bytes = 0;
        GKQ_CUDA(h, cudaMalloc(&h->flat_buf, want + want / 8));
        h->flat_bytes = want + want / 8;
        h->stats.scratch_bytes += h->flat_bytes;
    }
    char* p = h->flat_buf;
    auto take = [&](size_t bytes) {
        char* q = p;
        p += (bytes + 255) / 256 * 256;
        return q;
    };
    Flat2& F = Q.flat;
    F.ctl = (Ctl2*)take(sizeof(Ctl2));
    F.oslab.dbase = (double*)take(slab_doubles(Q.cap_outer, Q.npts_max + 2) * 8 * N);
    F.oslab.wbase = (int*)take(slab_ints(Q.cap_outer) * 4 * N);
    F.oslab.stride = N;
    F.oslab.cap = Q.cap_outer;
    F.oslab.npts_max = Q.npts_max + 2;
    F.olane = (Lane*)take(sizeof(Lane) * N);
    F.octx = (OuterCtx*)take(sizeof(OuterCtx) * N);
    F.act[0] = (unsigned int*)take(4 * N);
    F.act[1] = (unsigned int*)take(4 * N);
    F.xlist = (unsigned int*)take(4 * N);
    F.t_est = (double*)take(8 * 50 * N);
    F.t_nev = (uint32_t*)take(4 * 50 * N);
    F.t_meta = (uint32_t*)take(4 * 50 * N);
    F.t_owner = (uint32_t*)take(4 * 50 * N);
    F.t_k = (uint32_t*)take(4 * 50 * N);
    if ((size_t)(p - h->flat_buf) > h->flat_bytes) return fail(h, GKQ_ERR_OUT_OF_MEMORY, "level-synchronous scratch carve overflow");

    // persistent grid of the inner kernel: its threads own the inner lists
    if (h->occ_2df[fid2][alg] < 0) h->occ_2df[fid2][alg] = L.occupancy_flat[alg]();
    int grid = persistent_grid(h, h->occ_2df[fid2][alg], Q.cap_inner, Q.npts_max + 2, FB);
    if (grid < 1) return fail(h, GKQ_ERR_OUT_OF_MEMORY, "max_evals too large for the interval slabs");
    // (a round posts up to 50 tasks per outer integral)
    const long long task_blocks = (50ll * n + FB - 1) / FB;
    if ((long long)grid > task_blocks) grid = (int)task_blocks;
    SlabGeom g;
    if ((rc = ensure_slab(h, (unsigned long long)grid * FB, Q.cap_inner, Q.npts_max + 2, g))) return rc;
    Q.slab = g;
    long long flat_grid = (n + FB - 1) / FB;
    const long long flat_cap = (long long)h->sm_count * 16;
    if (flat_grid > flat_cap) flat_grid = flat_cap;

    const char* kname = alg == 0 ? "nested_flat<QAGS2>" : "nested_flat<QAGP2>";
    prof_begin(h, kname, st);
    GKQ_CUDA(h, cudaMemsetAsync(F.ctl, 0, sizeof(Ctl2), st));
    L.flat_start[alg](Q, (int)flat_grid, st);
    if ((rc = launch_check(h, "k2_start"))) return rc;
    const int max_rounds = Q.cap_outer + 8;
    for (int r = 0;; ++r) {
        if (r > max_rounds) return fail(h, GKQ_ERR_CUDA, "level-synchronous nest did not terminate");
        k2_round_reset<<<1, 1, 0, st>>>(F.ctl, r);
        launch_flat_plan(alg == 1, Q, r, (int)flat_grid, st);
        L.flat_inner[alg](Q, grid, st);
        launch_flat_commit(alg == 1, Q, r, (int)flat_grid, st);
        L.flat_exact[alg](Q, grid, st);
        if ((rc = launch_check(h, "nested_flat round"))) return rc;
        h->stats.kernel_launches += 4;
        GKQ_CUDA(h, cudaMemcpyAsync(&h->flat_ctl_host[r & 1], F.ctl, sizeof(Ctl2), cudaMemcpyDeviceToHost, st));
        GKQ_CUDA(h, cudaEventRecord(h->ev_flat[r & 1], st));
        if (r >= 1) {
            GKQ_CUDA(h, cudaEventSynchronize(h->ev_flat[(r - 1) & 1]));
            // after round r - 1: nact[r & 1] outer integrals were left for round r (already enqueued)
            if (h->flat_ctl_host[(r - 1) & 1].nact[r & 1] == 0ull) break;
        }
    }
    prof_end(h, st);
    return GKQ_OK;
}

extern "C" {
int gkq_integrate2_batch_ex(gkq_handle_t h, const gkq_config* cfg, const gkq_overrides* ov, int functor2_id, int yrange_id,
                         int64_t n, const double* x0, const double* x1, int64_t range_stride,
                         const double* fparams, int64_t fparam_stride, const double* yparams,
                         int64_t yparam_stride, const int64_t* pts_offsets, const double* pts_xy,
                         double* out_estimate, double* out_delta, uint32_t* out_nevals,
                         uint32_t* out_status, void* stream) {
    if (!h) return GKQ_ERR_INVALID_ARGUMENT;
    if (!cfg) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "cfg is NULL");
    if (functor2_id < 0 || functor2_id >= F2_COUNT) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "unknown 2-D functor id");
    if (yrange_id < 0 || yrange_id >= YR_COUNT) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "unknown y-range functor id");
    if (n < 0 || n > 0xffffffffll) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "n out of range");
    if (cfg->algo < 0 || cfg->algo > 3) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "unknown algorithm");
    if (range_stride != 0 && range_stride != 1) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "range_stride must be 0 or 1");
    if (cfg->max_evals > 0x7fffffffull) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "max_evals must fit in 31 bits");
    if (cfg->npoints < 0 || (cfg->npoints > 0 && !cfg->points)) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "points is NULL");
    int rc;
    if ((rc = check_tol(h, cfg->tol))) return rc;
    for (int64_t k = 0; k < 2 * cfg->npoints; ++k)
        if (cfg->points[k] != cfg->points[k])  // double/integrator.rs:60-66
            return fail(h, GKQ_ERR_NAN_INPUT, "cannot include NAN value in singular points");
    if (n == 0) return GKQ_OK;
    if (!x0 || !x1 || !out_estimate || !out_delta || !out_nevals || !out_status)
        return fail(h, GKQ_ERR_INVALID_ARGUMENT, "NULL array argument");
    if (F2_TABLE[functor2_id].nparams > 0 && !fparams) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "fparams is NULL");
    if (YR_TABLE[yrange_id].nparams > 0 && !yparams) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "yparams is NULL");
    if (pts_offsets && !pts_xy) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "pts_xy is NULL");
    if (cfg->max_evals > h->max_evals_seen) h->max_evals_seen = cfg->max_evals;

    GKQ_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    h->cur = 0;
    gkq_handle_s::Scratch& S = h->sc[0];
    if ((rc = ensure_batch_scratch(h, (size_t)n))) return rc;

    Batch2D P;
    memset(&P, 0, sizeof P);
    P.n = n;
    P.x0 = x0;
    P.x1 = x1;
    P.range_stride = range_stride;
    P.fparams = fparams;
    P.fparam_stride = fparam_stride;
    P.yparams = yparams;
    P.yparam_stride = yparam_stride;
    P.yrange_id = yrange_id;
    P.nyparams = YR_TABLE[yrange_id].nparams;
    P.swap_xy = (cfg->flags & GKQ_FLAG_SWAP_XY) != 0;
    if (ov) {  // per-integral tolerance values / total budgets (device arrays)
        P.abs_tol_each = ov->abs_tol;
        P.rel_tol_each = ov->rel_tol;
        P.max_evals_each = (const unsigned long long*)ov->max_evals;
    }
    P.tol.kind = cfg->tol.kind;
    P.tol.abs_tol = cfg->tol.abs_tol;
    P.tol.rel_tol = cfg->tol.rel_tol;
    P.max_evals = cfg->max_evals;
    P.out.estimate = out_estimate;
    P.out.delta = out_delta;
    P.out.nevals = out_nevals;
    P.out.status = out_status;
    P.ctl = S.ctl;
    if (h->profiling) prof_clear(h);
    GKQ_CUDA(h, cudaMemsetAsync(S.ctl, 0, sizeof(Control), st));
    h->stats.batches += 1;
    h->stats.integrals += (uint64_t)n;
    if (cfg->flags & GKQ_FLAG_VALIDATE_INPUTS)
        if ((rc = validate_device_inputs(h, st, n, range_stride ? n : 1, x0, x1, P.abs_tol_each, P.rel_tol_each))) return rc;

    if (cfg->algo == GKQ_ALGO_QAG) {
        // QAG2 (deprecated in the reference, double/algorithm/qag.rs): never looks at points; one
        // thread per outer integral (k_nested<.., 2>): available, not tuned
        const Launch2D& L = h->l2[functor2_id];
        const unsigned long long outer_need = 2ull + (cfg->max_evals / 17ull) / 50ull;
        const unsigned long long inner_need = 2ull + cfg->max_evals / 50ull;
        if (inner_need > (1ull << 24)) return fail(h, GKQ_ERR_OUT_OF_MEMORY, "max_evals too large for the interval slabs");
        const int cap_total = (int)(outer_need + inner_need);
        int grid = persistent_grid(h, 1, cap_total, 2 * 2 + 110, BLOCK);
        if (grid < 1) return fail(h, GKQ_ERR_OUT_OF_MEMORY, "max_evals too large for the interval slabs");
        const long long nblocks = (n + BLOCK - 1) / BLOCK;
        if ((long long)grid > nblocks) grid = (int)nblocks;
        Batch2D Q = P;
        SlabGeom g;
        if ((rc = ensure_slab(h, (unsigned long long)grid * BLOCK, cap_total, 2 * 2 + 110, g))) return rc;
        Q.slab = g;
        Q.cap_outer = (int)outer_need;
        Q.cap_inner = (int)inner_need;
        prof_begin(h, "k_nested<QAG2>", st);
        L.run_qag(Q, grid, st);
        prof_end(h, st);
        return launch_check(h, "k_nested<QAG2>");
    }

    // which algorithm(s) run, and over which integrals
    bool run_s = false, run_p = false, split = false;
    int npts_max = 0;
    if (!pts_offsets) {
        run_s = cfg->algo == GKQ_ALGO_QAGS || (cfg->algo == GKQ_ALGO_AUTO && cfg->npoints == 0);
        run_p = !run_s;
        if (run_p) {
            npts_max = (int)cfg->npoints;
            P.npts = npts_max;
            if (npts_max > 0) {
                if ((rc = upload_shared_points(h, cfg->points, (size_t)(2 * npts_max), st))) return rc;
                P.pts_xy = S.dev_pts;
            }
        }
    } else if (cfg->algo == GKQ_ALGO_QAGS) {
        run_s = true;  // explicit QAGS2 never looks at points
    } else {
        const int grid = (int)((n + BLOCK - 1) / BLOCK);
        const int mode = cfg->algo == GKQ_ALGO_AUTO ? 0 : 1;
        k_split_auto<<<grid, BLOCK, 0, st>>>(n, (const long long*)pts_offsets, S.listS, S.listP, S.ctl, mode);
        if ((rc = launch_check(h, "k_split_auto"))) return rc;
        GKQ_CUDA(h, cudaMemcpyAsync(S.ctl_host, S.ctl, sizeof(Control), cudaMemcpyDeviceToHost, st));
        GKQ_CUDA(h, cudaStreamSynchronize(st));
        npts_max = (int)S.ctl_host->max_pts;
        P.pts_offsets = (const long long*)pts_offsets;
        P.pts_xy = pts_xy;
        if (mode == 0) {
            split = true;
            run_s = S.ctl_host->countS > 0;
            run_p = S.ctl_host->countP > 0;
        } else {
            run_p = true;
        }
    }
    P.npts_max = npts_max;

    for (int alg = 0; alg < 2; ++alg) {
        if (alg == 0 ? !run_s : !run_p) continue;
        Batch2D Q = P;
        if (alg == 0) {  // QAGS2 ignores points entirely
            Q.npts = 0;
            Q.pts_offsets = nullptr;
            Q.pts_xy = nullptr;
            Q.npts_max = 0;
        }
        if (split) {
            Q.worklist = alg == 0 ? S.listS : S.listP;
            Q.worklist_count = alg == 0 ? &S.ctl->countS : &S.ctl->countP;
        }
        const Launch2D& L = h->l2[functor2_id];
        if (h->occ_2d[functor2_id][alg] < 0) {
            h->occ_2d[functor2_id][alg] = L.occupancy[alg]();
            h->occ_2dw[functor2_id][alg] = L.occupancy_pool[alg]();
        }
        // list lengths: outer <= npts+1 + (max_evals/17)/50 + 1, inner <= npts+1 + max_evals/50 + 1
        const unsigned long long outer_need = (unsigned long long)Q.npts_max + 2ull + (cfg->max_evals / 17ull) / 50ull;
        const unsigned long long inner_need = (unsigned long long)Q.npts_max + 2ull + cfg->max_evals / 50ull;
        if (inner_need > (1ull << 24)) return fail(h, GKQ_ERR_OUT_OF_MEMORY, "max_evals too large for the interval slabs");
        const int cap_total = (int)(outer_need + inner_need);
        // k_nested_pool: warps own a few outer integrals each and share their inner integrals
        // among the lanes; k_nested (a thread per outer integral) is kept as the reference mapping.
        bool pool_mode = true;
        // level-synchronous rounds pay once a round's inner integrals fill the machine; small batches
        // (and batches whose per-outer scratch would not fit) stay on the pool kernel
        bool flat_mode = n >= 1024 && n < (1ll << 25) &&
                         flat_bytes_per_outer((int)outer_need, Q.npts_max) * (size_t)n < h->total_mem / 3;
        if (const char* env = getenv("GKQ_NESTED_MODE")) {  // test / tuning knob: 1 = thread, 2 = pool, 3 = level-synchronous
            const int m = atoi(env);
            if (m == 1) pool_mode = false;
            if (m == 1 || m == 2) flat_mode = false;
            if (m == 3) flat_mode = n < (1ll << 25);
        }
        if (flat_mode) {
            Q.cap_outer = (int)outer_need;
            Q.cap_inner = (int)inner_need;
            if ((rc = run_nested_flat(h, L, functor2_id, alg, Q, n, st))) return rc;
            continue;
        }
        const int blk = pool_mode ? PB : BLOCK;
        int grid = persistent_grid(h, pool_mode ? h->occ_2dw[functor2_id][alg] : h->occ_2d[functor2_id][alg],
                                   cap_total, 2 * (Q.npts_max + 2) + 110, blk);
        long long nblocks = (n + BLOCK - 1) / BLOCK;
        if (pool_mode && grid >= 1) {
            // outer integrals per warp: as few as still keep every warp of the grid busy
            const long long warps = (long long)grid * (PB / 32);
            long long r = (n + warps - 1) / warps;
            if (const char* env = getenv("GKQ_POOL_R")) r = atoll(env);  // tuning knob
            if (r < 1) r = 1;
            if (r > POOL_R) r = POOL_R;
            Q.pool_r = (int)r;
            nblocks = (n + r * (PB / 32) - 1) / (r * (PB / 32));
        }
        if (grid < 1) return fail(h, GKQ_ERR_OUT_OF_MEMORY, "max_evals too large for the interval slabs");
        if ((long long)grid > nblocks) grid = (int)nblocks;
        SlabGeom g;
        if ((rc = ensure_slab(h, (unsigned long long)grid * blk, cap_total, 2 * (Q.npts_max + 2) + 110, g))) return rc;
        Q.slab = g;
        Q.cap_outer = (int)outer_need;
        Q.cap_inner = (int)inner_need;
        const char* kname = pool_mode ? (alg == 0 ? "k_nested_pool<QAGS2>" : "k_nested_pool<QAGP2>")
                                      : (alg == 0 ? "k_nested<QAGS2>" : "k_nested<QAGP2>");
        prof_begin(h, kname, st);
        if (pool_mode)
            L.run_pool[alg](Q, grid, st);
        else
            L.run[alg](Q, grid, st);
        prof_end(h, st);
        if ((rc = launch_check(h, alg == 0 ? "k_nested<QAGS2>" : "k_nested<QAGP2>"))) return rc;
    }
    return GKQ_OK;
}

int gkq_integrate2_batch(gkq_handle_t h, const gkq_config* cfg, int functor2_id, int yrange_id,
                         int64_t n, const double* x0, const double* x1, int64_t range_stride,
                         const double* fparams, int64_t fparam_stride, const double* yparams,
                         int64_t yparam_stride, const int64_t* pts_offsets, const double* pts_xy,
                         double* out_estimate, double* out_delta, uint32_t* out_nevals,
                         uint32_t* out_status, void* stream) {
    return gkq_integrate2_batch_ex(h, cfg, nullptr, functor2_id, yrange_id, n, x0, x1, range_stride, fparams,
                                   fparam_stride, yparams, yparam_stride, pts_offsets, pts_xy, out_estimate,
                                   out_delta, out_nevals, out_status, stream);
}

int gkq_integrate2_batch_host(gkq_handle_t h, const gkq_config* cfg, int functor2_id,
                              int yrange_id, int64_t n, const double* x0, const double* x1,
                              int64_t range_stride, const double* fparams, int64_t fparam_stride,
                              const double* yparams, int64_t yparam_stride,
                              const int64_t* pts_offsets, const double* pts_xy,
                              double* out_estimate, double* out_delta, uint32_t* out_nevals,
                              uint32_t* out_status) {
    return gkq_integrate2_batch_host_ex(h, cfg, nullptr, functor2_id, yrange_id, n, x0, x1, range_stride, fparams,
                                        fparam_stride, yparams, yparam_stride, pts_offsets, pts_xy,
                                        out_estimate, out_delta, out_nevals, out_status);
}

int gkq_integrate2_batch_host_ex(gkq_handle_t h, const gkq_config* cfg, const gkq_overrides* ov_host,
                                 int functor2_id, int yrange_id, int64_t n, const double* x0,
                                 const double* x1, int64_t range_stride, const double* fparams,
                                 int64_t fparam_stride, const double* yparams, int64_t yparam_stride,
                                 const int64_t* pts_offsets, const double* pts_xy, double* out_estimate,
                                 double* out_delta, uint32_t* out_nevals, uint32_t* out_status) {
    if (!h) return GKQ_ERR_INVALID_ARGUMENT;
    if (n < 0) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "n out of range");
    if (n == 0) return GKQ_OK;
    if (functor2_id < 0 || functor2_id >= F2_COUNT) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "unknown 2-D functor id");
    if (yrange_id < 0 || yrange_id >= YR_COUNT) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "unknown y-range functor id");
    if (!x0 || !x1 || !out_estimate || !out_delta || !out_nevals || !out_status)
        return fail(h, GKQ_ERR_INVALID_ARGUMENT, "NULL array argument");
    if (pts_offsets && !pts_xy) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "pts_xy is NULL");
    GKQ_CUDA(h, cudaSetDevice(h->device));
    const int np = F2_TABLE[functor2_id].nparams, nq = YR_TABLE[yrange_id].nparams;
    const size_t npairs = pts_offsets ? (size_t)pts_offsets[n] : 0;
    if ((np > 0 && !fparams) || (nq > 0 && !yparams)) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "params is NULL");
    if (!cfg) return fail(h, GKQ_ERR_INVALID_ARGUMENT, "cfg is NULL");
    if (host_has_nan(x0, range_stride ? (size_t)n : 1) || host_has_nan(x1, range_stride ? (size_t)n : 1))
        return fail(h, GKQ_ERR_NAN_INPUT, "cannot create Range object from Range which contains NaN value.");
    if (ov_host) {
        const bool need_abs = cfg->tol.kind != GKQ_TOL_RELATIVE, need_rel = cfg->tol.kind != GKQ_TOL_ABSOLUTE;
        if ((need_abs && host_has_nan(ov_host->abs_tol, (size_t)n)) || (need_rel && host_has_nan(ov_host->rel_tol, (size_t)n)))
            return fail(h, GKQ_ERR_NAN_INPUT, "Tolerance must not contain NAN.");
    }
    // single-shot staging (2-D batches are compute-heavy: copies are negligible)
    const size_t N = (size_t)n;
    const size_t nx = range_stride ? N : 1, nfp = fparam_stride ? N : 1, nyp = yparam_stride ? N : 1;
    const size_t in_d = 2 * nx + (size_t)np * nfp + (size_t)nq * nyp;
    const size_t csr_d = pts_offsets ? (N + 1) + 2 * npairs : 0;  // offsets (as 8-byte) + pairs
    const int n_ov = ov_host ? (ov_host->abs_tol != nullptr) + (ov_host->rel_tol != nullptr) + (ov_host->max_evals != nullptr) : 0;
    const size_t bytes = (in_d + csr_d) * 8 + N * 24 + (size_t)n_ov * N * 8 + 256;
    if (bytes > h->stage_bytes) {
        if (h->stage_dev) GKQ_CUDA(h, cudaFree(h->stage_dev));
        h->stage_dev = nullptr;
        h->stats.scratch_bytes -= h->stage_bytes;
        h->stage_bytes = 0;
        GKQ_CUDA(h, cudaMalloc(&h->stage_dev, bytes));
        h->stage_bytes = bytes;
        h->stats.scratch_bytes += bytes;
    }
    cudaStream_t st = h->s_compute;
    double* d = (double*)h->stage_dev;
    double* d_x0 = d;
    double* d_x1 = d + nx;
    double* d_fp = d + 2 * nx;
    double* d_yp = d_fp + (size_t)np * nfp;
    long long* d_off = (long long*)(d + in_d);
    double* d_pairs = (double*)(d_off + (pts_offsets ? N + 1 : 0));
    double* d_est = d + in_d + csr_d;
    double* d_dlt = d_est + N;
    uint32_t* d_nev = (uint32_t*)(d_dlt + N);
    uint32_t* d_st = d_nev + N;
    gkq_overrides ov_dev = {nullptr, nullptr, nullptr};
    {
        double* q = (double*)(d_st + N);
        if (ov_host && ov_host->abs_tol) {
            ov_dev.abs_tol = q;
            GKQ_CUDA(h, cudaMemcpyAsync(q, ov_host->abs_tol, N * 8, cudaMemcpyHostToDevice, st));
            q += N;
        }
        if (ov_host && ov_host->rel_tol) {
            ov_dev.rel_tol = q;
            GKQ_CUDA(h, cudaMemcpyAsync(q, ov_host->rel_tol, N * 8, cudaMemcpyHostToDevice, st));
            q += N;
        }
        if (ov_host && ov_host->max_evals) {
            ov_dev.max_evals = (const uint64_t*)q;
            GKQ_CUDA(h, cudaMemcpyAsync(q, ov_host->max_evals, N * 8, cudaMemcpyHostToDevice, st));
            q += N;
        }
    }
    if (pts_offsets) {
        GKQ_CUDA(h, cudaMemcpyAsync(d_off, pts_offsets, (N + 1) * 8, cudaMemcpyHostToDevice, st));
        if (npairs) GKQ_CUDA(h, cudaMemcpyAsync(d_pairs, pts_xy, npairs * 16, cudaMemcpyHostToDevice, st));
    }
    GKQ_CUDA(h, cudaMemcpyAsync(d_x0, x0, nx * 8, cudaMemcpyHostToDevice, st));
    GKQ_CUDA(h, cudaMemcpyAsync(d_x1, x1, nx * 8, cudaMemcpyHostToDevice, st));
    for (int k = 0; k < np; ++k)
        GKQ_CUDA(h, cudaMemcpyAsync(d_fp + (size_t)k * nfp, fparams + (size_t)k * (fparam_stride ? fparam_stride : 1),
                                    nfp * 8, cudaMemcpyHostToDevice, st));
    for (int k = 0; k < nq; ++k)
        GKQ_CUDA(h, cudaMemcpyAsync(d_yp + (size_t)k * nyp, yparams + (size_t)k * (yparam_stride ? yparam_stride : 1),
                                    nyp * 8, cudaMemcpyHostToDevice, st));
    int rc = gkq_integrate2_batch_ex(h, cfg, n_ov ? &ov_dev : nullptr, functor2_id, yrange_id, n, d_x0, d_x1, range_stride, d_fp,
                                  fparam_stride ? (int64_t)N : 0, d_yp, yparam_stride ? (int64_t)N : 0,
                                  pts_offsets ? (const int64_t*)d_off : nullptr,
                                  pts_offsets ? d_pairs : nullptr, d_est, d_dlt, d_nev, d_st, st);
    if (rc) return rc;
    GKQ_CUDA(h, cudaMemcpyAsync(out_estimate, d_est, N * 8, cudaMemcpyDeviceToHost, st));
    GKQ_CUDA(h, cudaMemcpyAsync(out_delta, d_dlt, N * 8, cudaMemcpyDeviceToHost, st));
    GKQ_CUDA(h, cudaMemcpyAsync(out_nevals, d_nev, N * 4, cudaMemcpyDeviceToHost, st));
    GKQ_CUDA(h, cudaMemcpyAsync(out_status, d_st, N * 4, cudaMemcpyDeviceToHost, st));
    GKQ_CUDA(h, cudaStreamSynchronize(st));
    return GKQ_OK;
}

}  // extern "C"
