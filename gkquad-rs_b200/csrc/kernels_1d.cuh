// kernels_1d.cuh -- the 1-D kernels (templates over the integrand functor).
//
//   k_qags_init<F,8>    qk17 on the whole range of every integral + the exit tests of
//                       initial_integral (single/algorithm/qags.rs:315-376, pass i = 0);
//                       survivors are compacted into two worklists by warp-aggregated
//                       atomics: "needs qk25" and "straight to the bisection loop".
//   k_qags_init<F,12>   qk25 on the first worklist (pass i = 1), survivors join the second.
//   k_adaptive<F,false> persistent self-refilling lanes: QAGS main loop (qags.rs:98-311).
//   k_adaptive<F,true>  the same machine with QAGP's break-point seeding (qagp.rs:68-372).
//   k_qk_batch<F,N>     rule-level entry (single/qk/mod.rs:28-41).
//
// All kernels are thread-per-integral: 100 % lane utilisation in the
// integrand samples, reference summation order for free, node tables broadcast from
// constant memory.
#pragma once
#include "adaptive.cuh"
#include "functors.cuh"

namespace gkq {

constexpr int BLOCK = 256;
// The persistent adaptive kernels use smaller blocks: their per-lane scalars (struct Lane, 200 B)
// live in static shared memory so that the sampling loops run at a low register count.
constexpr int ABLOCK = 128;
// A warp with at most this many live integrals evaluates them cooperatively (latency mode).
#ifndef GKQ_COOP_MAX_LIVE
#define GKQ_COOP_MAX_LIVE 8
#endif
constexpr int COOP_MAX_LIVE = GKQ_COOP_MAX_LIVE;
// Resident 128-thread blocks per SM the adaptive kernels are compiled for (register cap
// 65536 / (128 * N)).  Round 1 ran 6 (80 registers, 24 warps per SM); round 2 measured 5 and 4
// (profiles/ab_adaptive_blocks_r02.txt, 1 M-integral batches): 4 blocks = 128 registers, 16 warps per SM
// wins everywhere -- P0 0.787 -> 0.694 ms, P1 6.03 -> 4.72, S3 0.694 -> 0.621, the S1 straggler launch
// 63 -> 53 us -- because (a) the list / epsilon-table code no longer spills and (b) the 2 x 12 samples
// a transcendental body parks in local memory (192 B per thread) then fit L1 next to the lanes' shared
// memory: 512 threads x 192 B = 98 KB against 137 KB of L1, where 768 threads needed 147 KB of 78 KB.
#ifndef GKQ_ADAPTIVE_MIN_BLOCKS
#define GKQ_ADAPTIVE_MIN_BLOCKS 4
#endif
// result0 of an integral that enters the main loop, parked by the initial pass that found it
// unfinished until k_qags_first picks it up: ONE 32-byte record = one DRAM sector, written whole by its
// thread and read back with two 16-byte loads (round 1 kept four SoA arrays: four partly written sectors
// per entrant, each filled from DRAM first).  Handle-owned scratch: the caller's output arrays are never
// read back -- they may be shared by several deferred calls in flight, or live in a peer GPU's memory.
struct alignas(32) ParkRec {
    double est, dlt, absv;  // estimate, delta, absvalue of the initial rule (qags.rs:315-376 returns them)
    unsigned int nev, pad;
};

// Device-side control block of one batch call (zeroed by a memset node before the kernels).
// The QAGS pipeline runs as two independent chains after the qk17 pass:
//   chain 0: integrals whose qk17 error is far off (delta > 1024 tol) skip qk25 and go straight to
//            bisection:                    qk17 -> first step -> persistent loop
//   chain 1: the others take the qk25 pass first:  qk17 -> qk25 -> first step -> persistent loop
// Each chain has its own worklists, counters and slab, so the two run on concurrent streams.
struct ChainCtl {
    unsigned long long countLoop;   // entries in listLoop[c]  (integrals entering the main loop)
    unsigned long long pad[3];
};
struct Control {
    ChainCtl chain[2];
    unsigned long long count25;  // entries in list25 (integrals that need the qk25 pass)
    unsigned long long headP;    // queue head of the persistent QAGP / nested kernels
    unsigned long long countP;   // QAGP worklist size when built on the device (AUTO split)
    unsigned long long countS;   // QAGS worklist size when built on the device (AUTO split)
    unsigned long long max_pts;  // largest per-integral point count (CSR batches)
    unsigned long long head2;    // second queue head (nested QAGS2 kernel)
    unsigned long long pad[2];
};

// An integral that is still unfinished after its first bisection step: a self-contained record
// (inputs + the state qags.rs leaves behind after iteration 1).  Records of one or many batches
// accumulate in a handle-owned buffer and are finished by ONE launch of the persistent kernel
// -- right away (ordered mode), at gkq_join (deferred mode: the remainders of many batches
// share a launch), or at the end of a host-pipeline call.
struct alignas(16) StragglerRec {
    double a, b;                   // range as given (the kernel re-applies the infinite-range map)
    double p[MAX_PARAMS];
    double est0, delta0, abs0;     // result0 of the initial passes
    double e1, d1, e2, d2;         // the two halves of the first bisection
    double area, errsum;
    unsigned long long idx;        // position in the owning call's output arrays
    unsigned int nevals;           // evaluations so far (result0 + 50)
    unsigned int ro1;              // roundoff_type1 after the first step
    unsigned int call;             // slot of the owning call in the outputs table
    unsigned int pad_;
    double abs_tol, rel_tol;       // the integral's own tolerance and budget (per-integral overrides
    unsigned long long max_evals;  // or the batch configuration)
};
struct RecCtl {
    unsigned long long count;  // records appended
    unsigned long long head;   // queue head of the persistent kernel that finishes them
    unsigned long long pad[2];
};
// Where finished stragglers go: the owning call's output arrays (device entry), or a compact
// patch that the host applies (host pipeline).
struct PatchView {
    unsigned long long* idx;
    double* est;
    double* dlt;
    unsigned int* nev;
    unsigned int* st;
};

// Sink of the QAGS persistent kernel: slot = record index.
struct RecordSink {
    const StragglerRec* recs;
    const Outputs* call_outs;
    PatchView patch;
    GKQ_DEV void put(long long slot, double est, double dlt, uint32_t nev, uint32_t st) const {
        const StragglerRec& r = recs[slot];
        if (patch.idx) {
            patch.idx[slot] = r.idx;
            patch.est[slot] = est;
            patch.dlt[slot] = dlt;
            patch.nev[slot] = nev;
            patch.st[slot] = st;
        } else {
            write_result(call_outs[r.call], (long long)r.idx, est, dlt, nev, st);
        }
    }
};
template <bool QAGP> struct SinkFor { using type = GlobalSink; };
template <> struct SinkFor<false> { using type = RecordSink; };
struct Batch1D {
    long long n;
    const double* a;
    const double* b;
    long long range_stride;
    const double* params;
    long long param_stride;
    const long long* pts_offsets;  // per-integral CSR or nullptr
    const double* pts;             // CSR values, or the shared points (device copy)
    int npts_shared;
    Tol tol;
    unsigned long long max_evals;
    unsigned long long ws_capacity;
    // optional per-integral overrides of tol.abs_tol / tol.rel_tol / max_evals (gkq_overrides);
    // max_evals values are clamped to `max_evals`, which sizes the interval slabs
    const double* abs_tol_each;
    const double* rel_tol_each;
    const unsigned long long* max_evals_each;
    Outputs out;
    // scratch owned by the handle
    Control* ctl;
    ParkRec* park;              // [n] result0 of the loop entrants
    unsigned int* list25;       // [n]
    unsigned int* listLoop[2];  // [n] per chain: integrals entering the main loop
    int chain;                  // which chain this launch serves (k_qags_first)
    // straggler records (written by k_qags_first, finished by k_adaptive<.., false>)
    StragglerRec* recs;
    RecCtl* rctl;
    unsigned int call_slot;         // slot of this call in `call_outs`
    unsigned long long idx_offset;  // added to the integral index stored in a record (host pipeline chunks)
    const Outputs* call_outs;       // [slots] output arrays of the calls that own the records
    // device entry: k_qags_first itself publishes this call's output arrays in its slot of the table
    // (kernel parameters -> device memory, stream-ordered; no host staging buffer that a later call
    // could overwrite while an earlier copy is still queued)
    Outputs* call_outs_w;
    PatchView patch;                // patch.idx != nullptr: finished stragglers go to the patch
    const unsigned int* worklist;        // optional subset of integrals (AUTO split) or nullptr
    const unsigned long long* worklist_count;  // device count for `worklist`
    SlabGeom slab;  // of the chain being launched
};

GKQ_DEV void init_sink(GlobalSink& s, const Batch1D& P) { s.O = P.out; }
GKQ_DEV void init_sink(RecordSink& s, const Batch1D& P) {
    s.recs = P.recs;
    s.call_outs = P.call_outs;
    s.patch = P.patch;
}

// The integral's own tolerance / evaluation budget (per-integral configuration).
GKQ_DEV Tol load_tol(const Batch1D& P, long long i) {
    Tol t = P.tol;
    if (P.abs_tol_each) t.abs_tol = P.abs_tol_each[i];
    if (P.rel_tol_each) t.rel_tol = P.rel_tol_each[i];
    return t;
}
GKQ_DEV unsigned long long load_max_evals(const Batch1D& P, long long i) {
    if (!P.max_evals_each) return P.max_evals;
    const unsigned long long m = P.max_evals_each[i];
    return m < P.max_evals ? m : P.max_evals;
}

template <class F>
GKQ_DEV void load_functor(F& f, const Batch1D& P, long long i, int nparams) {
#pragma unroll
    for (int k = 0; k < MAX_PARAMS; ++k)
        f.p[k] = (k < nparams) ? (P.param_stride ? P.params[(long long)k * P.param_stride + i]
                                                 : P.params[k])
                               : 0.0;
}

// Range of integral i, mapped to [-1, 1] coordinates when either end is infinite
// (qags.rs:48-57, util.rs:162-176).
GKQ_DEV void load_range(const Batch1D& P, long long i, double& a, double& b, bool& transform) {
    a = P.a[i * P.range_stride];
    b = P.b[i * P.range_stride];
    transform = !is_finite(a) || !is_finite(b);
    if (transform) {
        a = transform_point(a);
        b = transform_point(b);
    }
}

// Warp-aggregated append of `i` to a worklist: one atomic per warp per list.
// Must be reached by all 32 lanes of the warp (callers keep whole warps in their loops).
GKQ_DEV void list_append(bool pred, unsigned int i, unsigned int* list, unsigned long long* count) {
    const unsigned mask = __ballot_sync(0xffffffffu, pred);
    if (!pred) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(count, (unsigned long long)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    list[base + __popc(mask & ((1u << lane) - 1u))] = i;
}

template <class F, int N>
struct RuleUnroll {
    // heavy (transcendental) bodies: one node pair per loop trip; light bodies: fully unrolled
    static constexpr int value = F::HEAVY ? 1 : N;
};
// the flat initial-pass kernels (tuning knobs for A/B builds)
#ifndef GKQ_INIT_UNROLL_HEAVY
#define GKQ_INIT_UNROLL_HEAVY 1
#endif
// 128-thread blocks, 6 per SM: a cap of 80 registers (the transcendental bodies use 74) lets ptxas keep
// the node pair's abscissae and arguments instead of recomputing them (9 % fewer instructions per
// pair than at 64 registers), and 24 warps per SM are enough -- 8, 10 and 12 warps per scheduler
// all measured the same or slower (profiles/S1_r01.md, tuning log)
#ifndef GKQ_INIT_BLOCK
#define GKQ_INIT_BLOCK 128
#endif
constexpr int IBLOCK = GKQ_INIT_BLOCK;
#ifndef GKQ_INIT17_MIN_BLOCKS
#define GKQ_INIT17_MIN_BLOCKS 6
#endif
#ifndef GKQ_INIT25_MIN_BLOCKS
#define GKQ_INIT25_MIN_BLOCKS 6
#endif
// bodies without transcendental calls: their rules are fully unrolled and run branch-free (qk_rule_light),
// so registers buy overlap of the samples' dependent chains
#ifndef GKQ_LIGHT_MIN_BLOCKS
#define GKQ_LIGHT_MIN_BLOCKS 6
#endif
#ifndef GKQ_LIGHT_FIRST_MIN_BLOCKS
#define GKQ_LIGHT_FIRST_MIN_BLOCKS 5
#endif
// the persistent kernel: its loop also holds the list / epsilon-table code, so the rule is unrolled
// less than in the flat kernels (instruction cache; profiles/)
#ifndef GKQ_ADAPT_UNROLL_LIGHT
#define GKQ_ADAPT_UNROLL_LIGHT 12
#endif
template <class F, int N>
struct AdaptUnroll {
    static constexpr int value = F::HEAVY ? 1 : (GKQ_ADAPT_UNROLL_LIGHT < N ? GKQ_ADAPT_UNROLL_LIGHT : N);
};
template <class F, int N>
struct InitUnroll {
    static constexpr int value = F::HEAVY ? GKQ_INIT_UNROLL_HEAVY : N;
};

// ---- QAGS initial passes -------------------------------------------------------------------
// PASS 0: N = 8 (qk17) over the whole batch (or `worklist`); PASS 1: N = 12 (qk25) over list25.
template <class F, int N, int NPARAMS>
__global__ void __launch_bounds__(IBLOCK, F::HEAVY ? ((N == 8) ? GKQ_INIT17_MIN_BLOCKS : GKQ_INIT25_MIN_BLOCKS) : GKQ_LIGHT_MIN_BLOCKS)
k_qags_init(const Batch1D P) {
    constexpr int PASS = (N == 8) ? 0 : 1;
#ifndef GKQ_FV_LOCAL
    // transcendental bodies: the 2N samples of the rule live in the thread's shared-memory column
    __shared__ double s_fv[F::HEAVY ? 2 * N * IBLOCK : 1];
#endif
    long long total;
    if (PASS == 0)
        total = P.worklist ? (long long)*P.worklist_count : P.n;
    else
        total = (long long)P.ctl->count25;

    for (long long k = (long long)blockIdx.x * IBLOCK + threadIdx.x;
         k - threadIdx.x % 32 < total;  // keep whole warps in the loop for the ballots below
         k += (long long)gridDim.x * IBLOCK) {
        const bool valid = k < total;
        long long i = 0;
        bool to25 = false, toLoop = false;
        if (valid) {
            i = (PASS == 0) ? (P.worklist ? (long long)P.worklist[k] : k) : (long long)P.list25[k];
            Wrapped<F> g;
            double a, b;
            load_range(P, i, a, b, g.transform);
            load_functor(g.f, P, i, NPARAMS);
            if (PASS == 0 && P.max_evals_each && load_max_evals(P, i) < 17ull) {  // qags.rs:86-88
                write_result(P.out, i, 0.0, F64_MAX, 0u, ST_INSUFFICIENT_ITERATION);
            } else {
#ifndef GKQ_FV_LOCAL
            QK q;
            if constexpr (F::HEAVY)
                q = qk_rule_smem<N>(g, a, b, s_fv + threadIdx.x, IBLOCK);
            else
                q = qk_rule_hoisted<N, InitUnroll<F, N>::value>(g, a, b);
#else
            const QK q = qk_rule_hoisted<N, InitUnroll<F, N>::value>(g, a, b);
#endif
            const uint32_t nevals = (PASS == 0) ? 17u : 42u;

            // initial_integral, pass i = PASS (qags.rs:336-372)
            uint32_t status = ST_NONE;
            bool done = true;
            const double tolerance = load_tol(P, i).to_abs(fabs(q.estimate));
            const double round_off = 100. * F64_EPSILON * q.absvalue;
            if (q.estimate != q.estimate) {
                status = ST_NAN;
            } else if ((q.delta <= tolerance && q.delta != q.asc) || q.delta == 0.0) {
                status = ST_NONE;
            } else if (q.delta <= round_off && q.delta > tolerance) {
                status = ST_ROUNDOFF_ERROR;
            } else if (load_max_evals(P, i) < (unsigned long long)(42 + PASS * 25)) {
                status = ST_INSUFFICIENT_ITERATION;
            } else {
                done = false;
                if (PASS == 0 && !(q.delta > tolerance * 1024.))
                    to25 = true;  // second pass with qk25
                else
                    toLoop = true;  // qags.rs:370-372 (`break`) or end of the 2-pass loop
            }
            // EVERY integral of the pass writes its 24 output bytes -- finished integrals their final
            // answer, the others result0 of this pass as a placeholder that the later kernels overwrite --
            // so whole 32-byte sectors leave the SM and L2 never fills a partly written sector from DRAM
            // (ncu, qk17 pass over 1 M integrals: dram read 37.9 -> 16.8 MB = the parameters, with the
            // 27 % unfinished integrals no longer left as holes; profiles/S1_r02.md).
            write_result(P.out, i, q.estimate, q.delta, nevals, status);
            if (toLoop) {
                ParkRec r;
                r.est = q.estimate;
                r.dlt = q.delta;
                r.absv = q.absvalue;
                r.nev = nevals;
                r.pad = 0u;
                double2* dst = reinterpret_cast<double2*>(P.park + i);
                dst[0] = make_double2(r.est, r.dlt);
                dst[1] = make_double2(r.absv, __hiloint2double((int)r.pad, (int)r.nev));
            }
            }
        }
#ifndef GKQ_EXPERIMENT_NO_APPEND  // (timing experiment only: results of the later passes are wrong)
        if (PASS == 0) list_append(to25, (unsigned int)i, P.list25, &P.ctl->count25);
        list_append(toLoop, (unsigned int)i, P.listLoop[PASS], &P.ctl->chain[PASS].countLoop);
#endif
    }
}

// ---- QAGS: the first bisection step of every loop entrant, as a flat kernel ------------------
// Iteration 1 of the main loop (qags.rs:120-271) works on a one-interval list, so its
// bookkeeping is closed-form; most smooth integrands converge right here.  Running it
// thread-per-integral over the compacted worklist keeps the sampling at full occupancy; only
// the integrals that need a second step go on to the persistent kernel, which resumes them
// from the state this kernel parks (list = the two halves ordered by error, order = [0, 1],
// epsilon table = [result0, area], ktmin = 1: what qags.rs leaves behind after iteration 1).
#ifndef GKQ_FIRST_BLOCK
#define GKQ_FIRST_BLOCK 128
#endif
// 5 blocks per SM (96 registers).  Round 1 went from 7 blocks (72 registers, one wave) to 6 (80): with
// batches pipelined (BatchPipeline) a single wave matters less than the spills (step 0.1185 -> 0.1172 ms,
// profiles/S1_r01.md).  Round 2 measured 6 / 5 / 4 (profiles/ab_adaptive_blocks_r02.txt, last block): the
// kernel alone 37.8 / 35.8 / 34.0 us, the pipelined S1 step 0.1170 / 0.1163 / 0.1177 ms, I1 0.521 / 0.513 /
// 0.507 ms -- all within 3 %; 5 is never worse than 6.
#ifndef GKQ_FIRST_MIN_BLOCKS
#define GKQ_FIRST_MIN_BLOCKS 5
#endif
constexpr int FBLOCK = GKQ_FIRST_BLOCK;
template <class F, int NPARAMS>
__global__ void __launch_bounds__(FBLOCK, F::HEAVY ? GKQ_FIRST_MIN_BLOCKS : GKQ_LIGHT_FIRST_MIN_BLOCKS) k_qags_first(const Batch1D P) {
#ifndef GKQ_FV_LOCAL
    __shared__ double s_fv[F::HEAVY ? 24 * FBLOCK : 1];
#endif
    const long long total = (long long)P.ctl->chain[P.chain].countLoop;
    if (P.call_outs_w && blockIdx.x == 0 && threadIdx.x == 0) P.call_outs_w[P.call_slot] = P.out;
    for (long long k = (long long)blockIdx.x * FBLOCK + threadIdx.x; k - threadIdx.x % 32 < total;
         k += (long long)gridDim.x * FBLOCK) {
        bool survive = false;
        long long i = 0;
        StragglerRec rec;
        if (k < total) {
            i = (long long)P.listLoop[P.chain][k];
            Wrapped<F> g;
            double a, b;
            load_range(P, i, a, b, g.transform);
            load_functor(g.f, P, i, NPARAMS);
            const double2* src = reinterpret_cast<const double2*>(P.park + i);
            const double2 pk0 = src[0], pk1 = src[1];
            const double est0 = pk0.x, delta0 = pk0.y;
            uint32_t nevals = (uint32_t)__double2loint(pk1.y);
            const Tol tol = load_tol(P, i);
            const unsigned long long max_evals = load_max_evals(P, i);
            const unsigned long long max_iters = (max_evals - nevals) / 50;
            if (max_iters >= 1) {  // else: zero trips, result0 already is the answer (qags.rs:274-276)
                const double center = (a + b) * 0.5;
#ifndef GKQ_FV_LOCAL
                QK q1, q2;
                if constexpr (F::HEAVY) {
                    q1 = qk_rule_smem<12>(g, a, center, s_fv + threadIdx.x, FBLOCK);
                    q2 = qk_rule_smem<12>(g, center, b, s_fv + threadIdx.x, FBLOCK);
                } else {
                    q1 = qk_rule<12, RuleUnroll<F, 12>::value>(g, a, center);
                    q2 = qk_rule<12, RuleUnroll<F, 12>::value>(g, center, b);
                }
#else
                const QK q1 = qk_rule<12, RuleUnroll<F, 12>::value>(g, a, center);
                const QK q2 = qk_rule<12, RuleUnroll<F, 12>::value>(g, center, b);
#endif
                nevals += 50;
                if (q1.estimate != q1.estimate || q2.estimate != q2.estimate) {
                    // break before the update: sum_results() over the one-interval list
                    write_result(P.out, i, 0.0 + est0, delta0, nevals, ST_NAN);
                } else {
                    const double area12 = q1.estimate + q2.estimate;
                    const double error12 = q1.delta + q2.delta;
                    double errsum = delta0;
                    errsum += error12 - delta0;
                    double area = est0;
                    area += area12 - est0;
                    const double tolerance = tol.to_abs(fabs(area));
                    int ro1 = 0;
                    if (q1.asc != q1.delta && q2.asc != q2.delta) {
                        if (fabs(est0 - area12) <= 1e-5 * fabs(area12) && error12 >= 0.99 * delta0) ro1 = 1;
                    }
                    const uint32_t error = subrange_too_small(a, center, b) ? ST_SUBRANGE_TOO_SMALL : ST_NONE;
                    // list after ws.update: slot 0 = the half with the larger error (workspace.rs:145-151)
                    const bool swap = q2.delta > q1.delta;
                    const double E0 = swap ? q2.estimate : q1.estimate, E1 = swap ? q1.estimate : q2.estimate;
                    if (errsum <= tolerance) {
                        const double total_est = (0.0 + est0) - est0 + q1.estimate + q2.estimate;
                        write_result(P.out, i, total_est, errsum, nevals, error);
                    } else if (error != ST_NONE || max_iters < 2) {
                        // `break` after the update (error) or loop exhausted: err_ext is still
                        // f64::MAX, so the answer is the plain sum of the list (qags.rs:274-276)
                        write_result(P.out, i, (0.0 + E0) + E1, errsum, nevals, error);
                    } else {
                        survive = true;
                        rec.a = P.a[i * P.range_stride];
                        rec.b = P.b[i * P.range_stride];
#pragma unroll
                        for (int q = 0; q < MAX_PARAMS; ++q) rec.p[q] = g.f.p[q];
                        rec.est0 = est0;
                        rec.delta0 = delta0;
                        rec.abs0 = pk1.x;
                        rec.e1 = q1.estimate;
                        rec.d1 = q1.delta;
                        rec.e2 = q2.estimate;
                        rec.d2 = q2.delta;
                        rec.area = area;
                        rec.errsum = errsum;
                        rec.idx = P.idx_offset + (unsigned long long)i;
                        rec.nevals = nevals;
                        rec.ro1 = (unsigned int)ro1;
                        rec.call = P.call_slot;
                        rec.pad_ = 0;
                        rec.abs_tol = tol.abs_tol;
                        rec.rel_tol = tol.rel_tol;
                        rec.max_evals = max_evals;
                    }
                }
            } else {
                // zero trips: result0 is the answer (qags.rs:274-276, err_ext still f64::MAX)
                write_result(P.out, i, 0.0 + est0, delta0, nevals, ST_NONE);
            }
        }
        // warp-aggregated append of the record
        const unsigned mask = __ballot_sync(0xffffffffu, survive);
        if (survive) {
            const int lane = threadIdx.x & 31;
            const int leader = __ffs(mask) - 1;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(&P.rctl->count, (unsigned long long)__popc(mask));
            base = __shfl_sync(mask, base, leader);
            P.recs[base + __popc(mask & ((1u << lane) - 1u))] = rec;
        }
    }
}

// ---- QAGP break points (qagp.rs:375-400) ------------------------------------------------------
// Builds [begin, sorted user points inside [min,max]..., end] into the lane's PTS slots and
// returns nint = len - 1.  The insertion sort is util.rs:178-211 with a strict comparator in
// range direction.
GKQ_DEV int make_sorted_points(const Slab& S, double begin, double end, const double* pts, int npts,
                               bool transform, int pts_stride = 1, bool pre_transform = false) {
    const bool ascending = begin < end;
    const double mn = ascending ? begin : end;
    const double mx = ascending ? end : begin;
    // pts2[0] = begin, pts2[1..] = user points (transformed when the range is infinite)
    S.PTS(0) = begin;
    for (int k = 0; k < npts; ++k) {
        double v = pts[(long long)k * pts_stride];
        if (pre_transform) v = transform_point(v);  // double/algorithm/qagp.rs:98-106
        S.PTS(1 + k) = transform ? transform_point(v) : v;
    }
    // insert_sort(&mut pts2[1..])
    for (int sp = 2; sp <= npts; ++sp) {
        const double target = S.PTS(sp);
        const double prev = S.PTS(sp - 1);
        if (ascending ? (prev < target) : (prev > target)) continue;
        int insert_pos = 1;
        for (int ip = sp - 2; ip >= 1; --ip) {
            const double another = S.PTS(ip);
            if (ascending ? (another < target) : (another > target)) {
                insert_pos = ip + 1;
                break;
            }
        }
        for (int k = sp; k > insert_pos; --k) S.PTS(k) = S.PTS(k - 1);
        S.PTS(insert_pos) = target;
    }
    // retain(min <= x <= max) then push(end)
    int len = 0;
    for (int k = 0; k <= npts; ++k) {
        const double x = S.PTS(k);
        if (mn <= x && x <= mx) {
            S.PTS(len) = x;
            len += 1;
        }
    }
    S.PTS(len) = end;
    return len;  // = pts.len() - 1 = nint
}

// ---- persistent adaptive kernel ------------------------------------------------------------------
// COOP: compile the cooperative latency mode in.  It pays when the queue is SHORT (the straggler
// records of a smooth QAGS batch: 153 integrals walking 10 sequential steps) and costs when the queue
// is long -- measured on 1M-integral batches with it compiled out: P0 1.08 -> 0.83 ms, P1 7.4 -> 6.3,
// I1 0.61 -> 0.52, I2 2.39 -> 2.12, S3 0.81 -> 0.74 (profiles/ab_coop_r02.txt): a quarter of P0's
// warp instructions were shuffles and redundant reductions of warps that had drawn a thin share of the
// queue's tail.  QAGP batches (the whole batch goes through this kernel) are compiled without it; the
// QAGS instance keeps it but switches it off at run time unless the queue is shorter than the grid.
template <class F, bool QAGP, int NPARAMS, bool COOP = !QAGP>
__global__ void __launch_bounds__(ABLOCK, GKQ_ADAPTIVE_MIN_BLOCKS) k_adaptive(const Batch1D P) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned long long t = (unsigned long long)blockIdx.x * ABLOCK + threadIdx.x;
    // lane-indexed node table for the cooperative path (constant memory would serialise)
    __shared__ double s_xgk[COOP ? 12 : 1];
    if (COOP) {
        if (threadIdx.x < 12) s_xgk[threadIdx.x] = c_xgk25[threadIdx.x];
        __syncthreads();
    }

    const Slab S = make_slab(P.slab.dbase, P.slab.wbase, P.slab.stride, t, P.slab.cap);

    unsigned long long total;
    if (QAGP)
        total = P.worklist ? *P.worklist_count : (unsigned long long)P.n;
    else
        total = P.rctl->count;
    unsigned long long* head = QAGP ? &P.ctl->headP : &P.rctl->head;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * (ABLOCK / 32);
    const unsigned long long warp_id = t >> 5;
    // first fill: queue positions [0, first_span) are owned statically; later refills draw
    // positions first_span + atomicAdd(head, ..)
    const unsigned long long share0 = (total + nwarps - 1ull) / nwarps;
    // cooperative mode only for queues that cannot even fill the grid's lanes once
    const int CML = (COOP && total <= nwarps * 32ull) ? COOP_MAX_LIVE : 0;
    const bool thin_start = share0 <= (unsigned long long)CML;
    const unsigned long long first_span = thin_start ? share0 * nwarps : nwarps * 32ull;
    bool first_fill = true;

    __shared__ Lane s_lanes[ABLOCK];
    Lane& s = s_lanes[threadIdx.x];
    s.active = false;
    s.idx = 0;
    // the lane's own tolerance and budget (batch configuration or per-integral overrides)
    __shared__ double s_abs_tol[ABLOCK], s_rel_tol[ABLOCK];
    __shared__ unsigned long long s_max_evals[ABLOCK];
    double& my_abs_tol = s_abs_tol[threadIdx.x];
    double& my_rel_tol = s_rel_tol[threadIdx.x];
    unsigned long long& my_max_evals = s_max_evals[threadIdx.x];
    my_abs_tol = P.tol.abs_tol;
    my_rel_tol = P.tol.rel_tol;
    my_max_evals = P.max_evals;
    // QAGP lanes write the batch outputs directly; QAGS lanes finish straggler records and write
    // to the owning call's outputs or to the host-pipeline patch (s.idx = record slot)
    typename SinkFor<QAGP>::type sink;
    init_sink(sink, P);
    Wrapped<F> g;
    g.transform = false;
    bool exhausted = false;

    for (;;) {
        // ---- refill idle lanes from the global queue (one atomic per warp per trip) ----
        // While the queue is long every idle lane takes an item (throughput mode).  When what
        // is left would not fill the grid any more, warps only top up to an equal share of the
        // remainder, so the last items are spread thinly over ALL warps and finish in the
        // cooperative latency mode instead of piling up on the few warps that ask first.
        const unsigned want = __ballot_sync(FULL, !s.active && !exhausted);
        if (want) {
            unsigned long long my = ~0ull;
            bool granted = false;
            if (first_fill) {
                // first fill: static assignment, no atomics (every thread owns one slot of the
                // first T queue positions; short queues are dealt round-robin over the warps)
                if (thin_start) {
                    if ((unsigned long long)lane < share0) my = (unsigned long long)lane * nwarps + warp_id;
                } else {
                    my = t;
                }
                granted = my != ~0ull;
                if (!granted) my = total;  // nothing for this lane in the first fill
            } else {
                const int leader = __ffs(want) - 1;
                unsigned long long base = 0;
                int take = __popc(want);
                // (all lanes) number of live integrals this warp already has
                const int live_now = __popc(__ballot_sync(FULL, s.active));
                if (lane == leader) {
                    const unsigned long long cur = first_span + *(volatile unsigned long long*)head;
                    const unsigned long long remaining = total > cur ? total - cur : 0ull;
                    const unsigned long long share = (remaining + nwarps - 1ull) / nwarps;
                    if (remaining > 0ull && share <= (unsigned long long)CML) {
                        int room = (int)share - live_now;
                        if (room < 1) room = live_now == 0 ? 1 : 0;
                        if (room < take) take = room;
                    }
                    if (take > 0) base = first_span + atomicAdd(head, (unsigned long long)take);
                }
                base = __shfl_sync(FULL, base, leader);
                take = __shfl_sync(FULL, take, leader);
                const int rank = __popc(want & ((1u << lane) - 1u));
                granted = rank < take;
                my = base + rank;
            }
            if (!s.active && !exhausted && granted) {
                if (my >= total) {
                    exhausted = true;
                } else {
                    lane_reset_common(s);
                    double a, b;
                    bool continue_refill = false;
                    if (!QAGP) {
                        // resume after iteration 1 from the straggler record
                        const StragglerRec& r = P.recs[my];
                        s.idx = (long long)my;
                        a = r.a;
                        b = r.b;
                        s.transform = !is_finite(a) || !is_finite(b);
                        if (s.transform) {
                            a = transform_point(a);
                            b = transform_point(b);
                        }
#pragma unroll
                        for (int q = 0; q < MAX_PARAMS; ++q) g.f.p[q] = r.p[q];
                        g.transform = s.transform;
                        s.est0 = r.est0;
                        s.delta0 = r.delta0;
                        s.abs0 = r.abs0;
                        s.nevals = r.nevals;  // nevals of result0 + 50
                        const uint32_t nev0 = s.nevals - 50u;
                        my_abs_tol = r.abs_tol;
                        my_rel_tol = r.rel_tol;
                        my_max_evals = r.max_evals;
                        // ws.reserve((max_evals - nevals) / 50 + 1)   qags.rs:98-99
                        s.max_iters = (my_max_evals - nev0) / 50;
                        s.limit = (int)reserve_cap(P.ws_capacity, s.max_iters + 1);
                        const double e1 = r.e1, d1 = r.d1, e2 = r.e2, d2 = r.d2;
                        const double center = (a + b) * 0.5;
                        if (d2 > d1) {  // workspace.rs:145-151
                            ws_push(S, s, center, b, e2, d2, 1);
                            ws_push(S, s, a, center, e1, d1, 1);
                        } else {
                            ws_push(S, s, a, center, e1, d1, 1);
                            ws_push(S, s, center, b, e2, d2, 1);
                        }
                        s.maxlevel = 1;
                        s.area = r.area;
                        s.errsum = r.errsum;
                        s.ro1 = (int)r.ro1;
                        tab_append(S, s, s.est0);
                        tab_append(S, s, s.area);
                        s.ktmin = 1;
                        s.res_ext = s.est0;
                        s.err_ext = F64_MAX;
                        s.ertest = 0.;
                        s.eolr = s.errsum;
                        s.iteration = 2;
                        s.active = true;
                    } else {
                        s.idx = P.worklist ? (long long)P.worklist[my] : (long long)my;
                        load_range(P, s.idx, a, b, s.transform);
                        if (a != a || b != b) {
                            // a NaN bound cannot exist in the reference (Range::new refuses it,
                            // single/common.rs:94); defined here instead of a silent (0, 0, 0, None)
                            write_result(P.out, s.idx, a + b, F64_MAX, 0u, ST_NAN);
                            continue_refill = true;
                        }
                        load_functor(g.f, P, s.idx, NPARAMS);
                        g.transform = s.transform;
                        const double* up;
                        int nup;
                        if (P.pts_offsets) {
                            const long long o0 = P.pts_offsets[s.idx];
                            up = P.pts + o0;
                            nup = (int)(P.pts_offsets[s.idx + 1] - o0);
                        } else {
                            up = P.pts;
                            nup = P.npts_shared;
                        }
                        s.nint = make_sorted_points(S, a, b, up, nup, s.transform);
                        s.nevals = 0;
                        s.est0 = s.delta0 = s.abs0 = s.asc0 = 0.;
                        {
                            const Tol t = load_tol(P, s.idx);
                            my_abs_tol = t.abs_tol;
                            my_rel_tol = t.rel_tol;
                            my_max_evals = load_max_evals(P, s.idx);
                        }
                        if (continue_refill) {
                            // (nothing to do: the lane stays idle and asks again next trip)
                        } else if (my_max_evals < (unsigned long long)s.nint * 25ull) {  // qagp.rs:79-81
                            write_result(P.out, s.idx, 0.0, F64_MAX, 0u, ST_INSUFFICIENT_ITERATION);
                        } else {
                            // ws.reserve(nint + (max_evals - nint*25)/50)   qagp.rs:83-84
                            s.limit = (int)reserve_cap(
                                P.ws_capacity, s.nint + (my_max_evals - (unsigned long long)s.nint * 25ull) / 50ull);
                            s.seeding = true;
                            s.seed_k = 0;
                            s.active = true;
                        }
                    }
                }
            }
        }
        first_fill = false;
        const unsigned live = __ballot_sync(FULL, s.active);
        if (live == 0u) break;

        // ---- plan this trip: up to two ranges to run the 25-point rule on ----
        double ra0 = 0., rb0 = 0., ra1 = 0., rb1 = 0.;
        bool has0 = false, has1 = false;
        if (s.active) {
            if (QAGP && s.seeding) {
                // next two break-point windows wide enough to integrate (qagp.rs:102-106)
                int w0 = -1, w1 = -1;
                int k = s.seed_k;
                for (; k < s.nint && w1 < 0; ++k) {
                    const double lo = S.PTS(k), hi = S.PTS(k + 1);
                    if (fabs(hi - lo) < 100. * F64_MIN_POSITIVE) continue;
                    if (w0 < 0) {
                        w0 = k; ra0 = lo; rb0 = hi; has0 = true;
                    } else {
                        w1 = k; ra1 = lo; rb1 = hi; has1 = true;
                    }
                }
                s.w0 = w0;
                s.w1 = w1;
                s.seed_k = k;
            } else {
                // bisect the interval with the largest error (qags.rs:122-125, util.rs:151-159)
                const int wi = s.wi;
                ra0 = S.A(wi);
                rb1 = S.B(wi);
                rb0 = ra1 = (ra0 + rb1) * 0.5;
                has0 = has1 = true;
            }
        }

        // ---- the hot part: 2 x 25 integrand samples per live integral ----
        QK q0, q1;
        q0.estimate = q0.delta = q0.absvalue = q0.asc = 0.;
        q1 = q0;
        if (!COOP || __popc(live) > CML) {
            // throughput mode: every lane walks its own 50 samples, warp-converged
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {
                const bool on = hh ? has1 : has0;
                if (on) {
                    const QK q = qk_rule<12, AdaptUnroll<F, 12>::value>(g, hh ? ra1 : ra0, hh ? rb1 : rb0);
                    if (hh) q1 = q; else q0 = q;
                }
            }
        } else {
            // latency mode (tail of the batch): the warp spreads the 50 samples of each live
            // integral over its 32 lanes, one live integral after the other
            unsigned todo = live;
#pragma unroll 1
            while (todo) {
                const int owner = __ffs(todo) - 1;
                todo &= todo - 1;
                Wrapped<F> go;
#pragma unroll
                for (int k = 0; k < MAX_PARAMS; ++k)
                    go.f.p[k] = (k < NPARAMS) ? __shfl_sync(FULL, g.f.p[k], owner) : 0.0;
                go.transform = __shfl_sync(FULL, (int)g.transform, owner) != 0;
                const bool oh0 = __shfl_sync(FULL, (int)has0, owner) != 0;
                const bool oh1 = __shfl_sync(FULL, (int)has1, owner) != 0;
                const double oa0 = __shfl_sync(FULL, ra0, owner), ob0 = __shfl_sync(FULL, rb0, owner);
                const double oa1 = __shfl_sync(FULL, ra1, owner), ob1 = __shfl_sync(FULL, rb1, owner);
                QK c0 = q0, c1 = q1;
                qk25_pair_coop(go, s_xgk, oh0, oa0, ob0, oh1, oa1, ob1, c0, c1);
                if (lane == owner) {
                    q0 = c0;
                    q1 = c1;
                }
            }
        }

        // ---- bookkeeping (per lane, divergent, short) ----
        if (s.active) {
            if (QAGP && s.seeding) {
                bool dead = false;
#pragma unroll 1
                for (int hh = 0; hh < 2 && !dead; ++hh) {
                    const int w = hh ? s.w1 : s.w0;
                    if (w < 0) continue;
                    const QK q = hh ? q1 : q0;
                    s.nevals += 25;
                    if (q.estimate != q.estimate) {  // qagp.rs:112-119
                        s.active = false;
                        write_result(P.out, s.idx, s.est0, s.delta0, s.nevals, ST_NAN);
                        dead = true;
                        break;
                    }
                    const int lv = (q.delta == q.asc && q.delta != 0.0) ? 1 : 0;
                    s.est0 += q.estimate;  // add_qkresult, qagp.rs:403-408
                    s.delta0 += q.delta;
                    s.abs0 += q.absvalue;
                    s.asc0 += q.asc;
                    ws_push(S, s, hh ? ra1 : ra0, hh ? rb1 : rb0, q.estimate, q.delta, lv);
                }
                if (!dead && s.seed_k >= s.nint) {
                    // all windows seeded: qagp.rs:132-179
                    s.seeding = false;
                    double deltasum = 0.;
                    for (int k = 0; k < s.size; ++k) {
                        if (S.LV(k) > 0) S.D(k) = s.delta0;
                        deltasum += S.D(k);
                    }
                    for (int k = 0; k < s.size; ++k) S.LV(k) = 0;
                    ws_sort_results(S, s);

                    Tol tol = P.tol;
                    tol.abs_tol = my_abs_tol;
                    tol.rel_tol = my_rel_tol;
                    const double tolerance = tol.to_abs(fabs(s.est0));
                    const double round_off = 100. * F64_EPSILON * s.abs0;
                    if (s.delta0 <= round_off && s.delta0 > tolerance) {
                        s.active = false;
                        write_result(P.out, s.idx, s.est0, s.delta0, s.nevals, ST_ROUNDOFF_ERROR);
                    } else if ((s.delta0 <= tolerance && s.delta0 != s.asc0) || s.delta0 == 0.0) {
                        s.active = false;
                        write_result(P.out, s.idx, s.est0, s.delta0, s.nevals, ST_NONE);
                    } else if ((unsigned long long)s.nevals == my_max_evals) {
                        s.active = false;
                        write_result(P.out, s.idx, s.est0, s.delta0, s.nevals,
                                     ST_INSUFFICIENT_ITERATION);
                    } else {
                        tab_append(S, s, s.est0);
                        s.area = s.est0;
                        s.res_ext = s.est0;
                        s.err_ext = F64_MAX;
                        s.errsum = deltasum;
                        s.eolr = deltasum;
                        s.ertest = tolerance;
                        s.max_iters = (unsigned long long)s.nint + (my_max_evals - s.nevals) / 50ull;
                        s.iteration = (unsigned long long)s.nint;
                    }
                }
            } else {
                const int wi = s.wi;  // the bisected interval (untouched since the plan step)
                Tol tol = P.tol;
                tol.abs_tol = my_abs_tol;
                tol.rel_tol = my_rel_tol;
                adaptive_step<QAGP>(S, s, tol, my_max_evals, sink, ra0, rb1, rb0, S.E(wi), S.D(wi),
                                    S.LV(wi), q0, q1);
            }
        }
    }
}

// ---- QAG (deprecated in the reference, still public): single/algorithm/qag.rs:65-248 -------------
// QAGS' skeleton without the epsilon table.  One thread walks one integral sequentially out of its
// slab (qag.rs has no caller in the adaptive drivers; this keeps the public algorithm available,
// not fast).  Differences from QAGS that matter for parity: the initial pass takes 50 eps (not 100)
// as round-off floor (:217), the loop counts roundoff_type2 for every non-type-1 step (:139-141),
// gives up only while deltasum > tolerance (:147-162), and does report nothing when the budget
// runs out (falls through to `finish`, :176).
template <class F, int NPARAMS>
__global__ void __launch_bounds__(ABLOCK) k_qag(const Batch1D P) {
    const unsigned long long t = (unsigned long long)blockIdx.x * ABLOCK + threadIdx.x;
    const Slab S = make_slab(P.slab.dbase, P.slab.wbase, P.slab.stride, t, P.slab.cap);
    const long long step = (long long)gridDim.x * ABLOCK;
    for (long long i = (long long)t; i < P.n; i += step) {
        Wrapped<F> g;
        double a, b;
        load_range(P, i, a, b, g.transform);
        load_functor(g.f, P, i, NPARAMS);
        const Tol tol = load_tol(P, i);
        const unsigned long long max_evals = load_max_evals(P, i);
        if (max_evals < 17ull) {  // qag.rs:76-78
            write_result(P.out, i, 0.0, F64_MAX, 0u, ST_INSUFFICIENT_ITERATION);
            continue;
        }
        // initial_integral (qag.rs:185-234)
        QK q;
        uint32_t nevals = 0, status = ST_NONE;
        bool finished = false;
#pragma unroll 1
        for (int pass = 0; pass < 2 && !finished; ++pass) {
            if (pass == 0) {
                nevals += 17;
                q = qk_rule<8, 1>(g, a, b);
            } else {
                nevals += 25;
                q = qk_rule<12, 1>(g, a, b);
            }
            const double tolerance = tol.to_abs(fabs(q.estimate));
            const double round_off = 50. * F64_EPSILON * q.absvalue;
            if (q.estimate != q.estimate) {
                status = ST_NAN;
                finished = true;
            } else if ((q.delta <= tolerance && q.delta != q.asc) || q.delta == 0.0) {
                finished = true;
            } else if (q.delta <= round_off && q.delta > tolerance) {
                status = ST_ROUNDOFF_ERROR;
                finished = true;
            } else if (max_evals < (unsigned long long)(42 + pass * 25)) {
                status = ST_INSUFFICIENT_ITERATION;
                finished = true;
            } else if (pass == 0 && q.delta > tolerance * 1024.) {
                break;
            }
        }
        if (finished) {
            write_result(P.out, i, q.estimate, q.delta, nevals, status);
            continue;
        }
        double area = q.estimate, deltasum = q.delta;
        const unsigned long long max_iters = (max_evals - nevals) / 50ull;
        Lane s;
        lane_reset_common(s);
        s.limit = (int)reserve_cap(P.ws_capacity, max_iters + 1ull);  // ws.clear(); ws.reserve(..)  :97-98
        ws_push(S, s, a, b, q.estimate, q.delta, 0);
        int ro1 = 0, ro2 = 0;
        uint32_t error = ST_NONE;
        bool returned = false;
#pragma unroll 1
        for (unsigned long long it = 1; it <= max_iters; ++it) {
            const int wi = s.wi;
            const double ia = S.A(wi), ib = S.B(wi), ie = S.E(wi), id = S.D(wi);
            const int lv = S.LV(wi) + 1;
            const double center = (ia + ib) * 0.5;
            const QK q1 = qk_rule<12, 1>(g, ia, center);
            const QK q2 = qk_rule<12, 1>(g, center, ib);
            nevals += 50;
            if (q1.estimate != q1.estimate || q2.estimate != q2.estimate) {  // :122-125
                error = ST_NAN;
                break;
            }
            const double area12 = q1.estimate + q2.estimate;
            const double delta12 = q1.delta + q2.delta;
            deltasum += delta12 - id;
            area += area12 - ie;
            if (q1.asc != q1.delta && q2.asc != q2.delta) {  // :135-142
                if (fabs(ie - area12) <= 1e-5 * fabs(area12) && delta12 >= 0.99 * id)
                    ro1 += 1;
                else
                    ro2 += 1;
            }
            const double tolerance = tol.to_abs(fabs(area));
            if (deltasum > tolerance) {  // :147-162
                if (ro1 >= 6 || ro2 >= 20)
                    error = ST_ROUNDOFF_ERROR;
                else if (subrange_too_small(ia, center, ib))
                    error = ST_SUBRANGE_TOO_SMALL;
                if (error != ST_NONE) {
                    write_result(P.out, i, ws_sum_results(S, s) - ie + q1.estimate + q2.estimate, deltasum, nevals,
                                 error);
                    returned = true;
                    break;
                }
            }
            ws_update(S, s, ia, center, q1.estimate, q1.delta, center, ib, q2.estimate, q2.delta, lv);
            if (deltasum <= tolerance) break;  // :170-172
        }
        if (!returned) write_result(P.out, i, ws_sum_results(S, s), deltasum, nevals, error);  // :176
    }
}

// ---- rule-level batch (single/qk/mod.rs:28-41) --------------------------------------------------
template <class F, int N, int NPARAMS>
__global__ void __launch_bounds__(BLOCK) k_qk_batch(const Batch1D P, double* out4) {
    const long long i = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= P.n) return;
    Wrapped<F> g;
    g.transform = false;  // the public rules take finite ranges only (naive.rs:53)
    load_functor(g.f, P, i, NPARAMS);
    const double a = P.a[i * P.range_stride], b = P.b[i * P.range_stride];
    const QK q = qk_rule<N, (N <= 12) ? RuleUnroll<F, N>::value : 1>(g, a, b);
    out4[i] = q.estimate;
    out4[P.n + i] = q.delta;
    out4[2 * P.n + i] = q.absvalue;
    out4[3 * P.n + i] = q.asc;
}

// ---- plain integrand evaluation y[i] = f(x[i]; p_i): the Integrand::apply_to_slice of the
// reference (single/common.rs:135-137).  Also the micro-kernel on which the per-functor FP64
// flop counts of the registry are measured with ncu (DADD + DMUL + 2*DFMA per evaluation).
// mode 1: the sample as the rule kernels take it -- branch-free math first (FastMath: own exp / sin / cos /
// log / pow, IEEE division and square root without the slow-path branch), re-evaluated through the
// always-valid path when an operand left the fast range.  Must equal mode 0 bit for bit.
template <class F, int NPARAMS>
__global__ void __launch_bounds__(BLOCK) k_eval(const Batch1D P, const double* x, double* y, int mode) {
    const long long i = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= P.n) return;
    F f;
    load_functor(f, P, i, NPARAMS);
    if (mode == 1) {
        FastMath m;
        double v = f.eval(x[i], m);
        if (!m.ok) v = f(x[i]);
        y[i] = v;
        return;
    }
    y[i] = f(x[i]);
}

// ---- launch table (one entry per registered functor; defined by GKQ_INSTANTIATE_F1) ----------
struct Launch1D {
    void (*qags_init17)(const Batch1D&, int grid, cudaStream_t);
    void (*qags_init25)(const Batch1D&, int grid, cudaStream_t);
    void (*qags_first)(const Batch1D&, int grid, cudaStream_t);
    void (*qags_loop)(const Batch1D&, int grid, cudaStream_t);
    void (*qagp)(const Batch1D&, int grid, cudaStream_t);
    void (*qag)(const Batch1D&, int grid, cudaStream_t);
    void (*qk[6])(const Batch1D&, double* out4, int grid, cudaStream_t);  // 17, 25, 33, 41, 49, 57 points
    void (*eval)(const Batch1D&, const double* x, double* y, int mode, int grid, cudaStream_t);
    int (*occupancy_loop)();   // resident blocks per SM of the persistent kernels
    int (*occupancy_qagp)();
    int (*occupancy_init25)();
    int (*occupancy_first)();
};

template <int ID>
Launch1D make_launch1d();
#define GKQ_DECL_F1(ID, NAME, NP, FL) template <> Launch1D make_launch1d<ID>();
GKQ_F1_LIST(GKQ_DECL_F1)
#undef GKQ_DECL_F1

#define GKQ_INSTANTIATE_F1(ID, NAME, NP, FL)                                                                   \
    template <> Launch1D make_launch1d<ID>() {                                                       \
        Launch1D L;                                                                                  \
        L.qags_init17 = [](const Batch1D& P, int grid, cudaStream_t st) {                            \
            k_qags_init<F1<ID>, 8, NP><<<grid, IBLOCK, 0, st>>>(P);                                   \
        };                                                                                           \
        L.qags_init25 = [](const Batch1D& P, int grid, cudaStream_t st) {                            \
            k_qags_init<F1<ID>, 12, NP><<<grid, IBLOCK, 0, st>>>(P);                                  \
        };                                                                                           \
        L.qags_first = [](const Batch1D& P, int grid, cudaStream_t st) {                             \
            k_qags_first<F1<ID>, NP><<<grid, FBLOCK, 0, st>>>(P);                                    \
        };                                                                                           \
        L.qags_loop = [](const Batch1D& P, int grid, cudaStream_t st) {                              \
            k_adaptive<F1<ID>, false, NP><<<grid, ABLOCK, 0, st>>>(P);                                \
        };                                                                                           \
        L.qagp = [](const Batch1D& P, int grid, cudaStream_t st) {                                   \
            k_adaptive<F1<ID>, true, NP><<<grid, ABLOCK, 0, st>>>(P);                                 \
        };                                                                                           \
        L.qk[0] = [](const Batch1D& P, double* o, int grid, cudaStream_t st) {                       \
            k_qk_batch<F1<ID>, 8, NP><<<grid, BLOCK, 0, st>>>(P, o);                                \
        };                                                                                           \
        L.qk[1] = [](const Batch1D& P, double* o, int grid, cudaStream_t st) {                       \
            k_qk_batch<F1<ID>, 12, NP><<<grid, BLOCK, 0, st>>>(P, o);                                \
        };                                                                                           \
        L.qk[2] = [](const Batch1D& P, double* o, int grid, cudaStream_t st) {                       \
            k_qk_batch<F1<ID>, 16, NP><<<grid, BLOCK, 0, st>>>(P, o);                                \
        };                                                                                           \
        L.qk[3] = [](const Batch1D& P, double* o, int grid, cudaStream_t st) {                       \
            k_qk_batch<F1<ID>, 20, NP><<<grid, BLOCK, 0, st>>>(P, o);                                \
        };                                                                                           \
        L.qk[4] = [](const Batch1D& P, double* o, int grid, cudaStream_t st) {                       \
            k_qk_batch<F1<ID>, 24, NP><<<grid, BLOCK, 0, st>>>(P, o);                                \
        };                                                                                           \
        L.qk[5] = [](const Batch1D& P, double* o, int grid, cudaStream_t st) {                       \
            k_qk_batch<F1<ID>, 28, NP><<<grid, BLOCK, 0, st>>>(P, o);                                \
        };                                                                                           \
        L.qag = [](const Batch1D& P, int grid, cudaStream_t st) {                                    \
            k_qag<F1<ID>, NP><<<grid, ABLOCK, 0, st>>>(P);                                           \
        };                                                                                           \
        L.eval = [](const Batch1D& P, const double* x, double* y, int mode, int grid, cudaStream_t st) { \
            k_eval<F1<ID>, NP><<<grid, BLOCK, 0, st>>>(P, x, y, mode);                                     \
        };                                                                                           \
        L.occupancy_loop = []() {                                                                    \
            int nb = 0;                                                                              \
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_adaptive<F1<ID>, false, NP>, ABLOCK, 0); \
            return nb;                                                                               \
        };                                                                                           \
        L.occupancy_qagp = []() {                                                                    \
            int nb = 0;                                                                              \
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_adaptive<F1<ID>, true, NP>, ABLOCK, 0); \
            return nb;                                                                               \
        };                                                                                           \
        L.occupancy_init25 = []() {                                                                  \
            int nb = 0;                                                                              \
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_qags_init<F1<ID>, 12, NP>, IBLOCK, 0); \
            return nb;                                                                               \
        };                                                                                           \
        L.occupancy_first = []() {                                                                  \
            int nb = 0;                                                                              \
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_qags_first<F1<ID>, NP>, FBLOCK, 0); \
            return nb;                                                                               \
        };                                                                                           \
        return L;                                                                                    \
    }

}  // namespace gkq
