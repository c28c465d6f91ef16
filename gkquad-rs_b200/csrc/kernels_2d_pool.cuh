// kernels_2d_pool.cuh -- the nested 2-D drivers as a two-level lane machine.
//
// (included by kernels_2d.cuh; reference: double/algorithm/qags.rs:46-97, qagp.rs:70-139)
//
// k_nested gives every outer integral one thread that walks the whole nest alone.  That is the
// reference's control flow, but on the GPU it starves: the inner integrals of one outer rule
// differ wildly in length (D1: 23 of 25 take one trip, the two next to the singular point take
// 12-20), so lanes idle, and a batch of a few thousand outer integrals cannot fill the machine
// at all.  Here a WARP owns POOL_R outer integrals at a time and works in rounds:
//
//   1. every outer integral (an "outer lane": its QAGS / QAGP state is a `Lane` in shared
//      memory, exactly the machine k_adaptive runs for 1-D integrals) plans its next trip -- one
//      rule (17 or 25 nodes) or the two rules of a bisection step (50 nodes) -- and posts the
//      inner integrals at those nodes as TASKS on the warp's board;
//   2. the 32 lanes of the warp work the board off as "inner lanes": each owns one inner
//      integral at a time (initial passes / break-point seeding / bisection loop), takes the next
//      task when its own terminates, and all lanes meet once per trip in the same sampling code,
//      so the warp stays converged however different the iteration counts are (a cooperative
//      50-samples-over-32-lanes mode for the last few long inner integrals exists behind
//      GKQ_POOL_COOP_MAX_LIVE; with several outer integrals per warp it measured slower);
//   3. the owners commit their tasks in the reference's call order (below), reduce the values in
//      the reference's summation order (bit-identical outer arithmetic) and advance their outer
//      state machines, POOL_R of them side by side.
//
// What the reference serialises between the inner calls of one outer integral is only the running
// budget (max_evals = total - evaluations so far), the first-error latch and the capacity of the
// re-used inner workspace.  The tasks of a trip therefore run with the budget at the START of the
// trip (an upper bound of their true budgets) and the trip is committed with a running sum when
// that provably did not matter: an inner call that spent `nev` evaluations never looks at its
// budget B or its list capacity as long as B >= 2 nev + 200 (+ 100 per break-point window): its
// iteration count stays below max_iters = (B - nev0) / 50 and its list shorter than limit / 2,
// the only places where qags.rs / qagp.rs / workspace.rs read them.  The first time this cannot
// be shown -- the budget is nearly exhausted -- that outer integral switches for good to the
// exact path: its owner lane walks the inner calls one after the other like k_nested does.
#pragma once

namespace gkq {

constexpr int PB = 128;                 // threads per block of k_nested_pool
constexpr int POOL_R = 6;               // outer integrals per warp (shared memory: 16 warps per SM)
constexpr int POOL_T = POOL_R * 50;     // board slots per warp (a bisection step posts 50 tasks)
#ifndef GKQ_POOL_COOP_MAX_LIVE
#define GKQ_POOL_COOP_MAX_LIVE 0  /* measured: D1 13.6 ms with 0, 15.4 with 2, 17.9 with 8 (profiles/) */
#endif
constexpr int POOL_COOP_MAX_LIVE = GKQ_POOL_COOP_MAX_LIVE;  // <= this many live inner lanes: cooperative sampling

struct TripPlan {
    double ra0, rb0, ra1, rb1;
    bool has17, has0, has1;
};

// ---- one lane of the QAGS / QAGP machine, split into start / plan / finish ----------------
// (the same phases k_adaptive runs; here for both levels of the nest)

// Begin integral `idx` over the (already transformed) range [a, b].  `rsv` (optional) receives how
// the call sized its workspace (RSV_* << 3 | n << 5; the low 3 bits are left for the status).
template <bool QAGP, class Sink>
GKQ_DEV void lane_start(const Slab& S, Lane& s, long long idx, double a, double b, bool transform,
                        const double* pts, int npts, int pts_stride, bool pts_pre_transform,
                        unsigned long long budget, unsigned long long cap0, Sink& sink, uint32_t* rsv) {
    lane_reset_common(s);
    s.idx = idx;
    s.transform = transform;
    s.active = false;
    s.nevals = 0;
    s.est0 = s.delta0 = s.abs0 = s.asc0 = 0.;
    if (rsv) *rsv = 0u;
    if (QAGP) {
        s.nint = make_sorted_points(S, a, b, pts, npts, transform, pts_stride, pts_pre_transform);
        if (budget < (unsigned long long)s.nint * 25ull) {  // qagp.rs:79-81
            sink.put(idx, 0.0, F64_MAX, 0u, ST_INSUFFICIENT_ITERATION);
            return;
        }
        // ws.reserve(nint + (max_evals - nint*25)/50)   qagp.rs:83-84
        s.limit = (int)reserve_cap(cap0, s.nint + (budget - (unsigned long long)s.nint * 25ull) / 50ull);
        if (rsv) *rsv = (RSV_QAGP << 3) | ((uint32_t)s.nint << 5);
        s.seeding = true;
        s.seed_k = 0;
        s.active = true;
    } else {
        if (budget < 17ull) {  // qags.rs:86-88
            sink.put(idx, 0.0, F64_MAX, 0u, ST_INSUFFICIENT_ITERATION);
            return;
        }
        s.seeding = true;  // initial passes: seed_k = 0 -> qk17, 1 -> qk25
        s.seed_k = 0;
        s.active = true;
    }
}

// Which rules the lane needs next.  SEED2: QAGP seeding takes two windows per trip (pure
// integrands only -- a NaN in the first window must leave the second one unevaluated).
template <bool QAGP, bool SEED2>
GKQ_DEV void lane_plan(const Slab& S, Lane& s, double wa, double wb, TripPlan& p) {
    p.ra0 = p.rb0 = p.ra1 = p.rb1 = 0.;
    p.has17 = p.has0 = p.has1 = false;
    if (QAGP && s.seeding) {
        // next break-point window(s) wide enough to integrate (qagp.rs:102-106)
        int w0 = -1, w1 = -1;
        int k = s.seed_k;
        for (; k < s.nint; ++k) {
            const double lo = S.PTS(k), hi = S.PTS(k + 1);
            if (fabs(hi - lo) < 100. * F64_MIN_POSITIVE) continue;
            if (w0 < 0) {
                w0 = k; p.ra0 = lo; p.rb0 = hi; p.has0 = true;
                if (!SEED2) { ++k; break; }
            } else {
                w1 = k; p.ra1 = lo; p.rb1 = hi; p.has1 = true;
                ++k;
                break;
            }
        }
        s.w0 = w0;
        s.w1 = w1;
        s.seed_k = k;
    } else if (!QAGP && s.seeding) {
        p.ra0 = wa;
        p.rb0 = wb;
        if (s.seed_k == 0) p.has17 = true; else p.has0 = true;
    } else {
        // bisect the interval with the largest error (qags.rs:122-125, util.rs:151-159)
        const int wi = s.wi;
        p.ra0 = S.A(wi);
        p.rb1 = S.B(wi);
        p.rb0 = p.ra1 = (p.ra0 + p.rb1) * 0.5;
        p.has0 = p.has1 = true;
    }
}

// Bookkeeping after the rules of the trip (q0 on [ra0, rb0], q1 on [ra1, rb1]).
template <bool QAGP, class Sink>
GKQ_DEV void lane_finish(const Slab& S, Lane& s, const Tol& tol, unsigned long long budget,
                         unsigned long long cap0, Sink& sink, uint32_t* rsv, const TripPlan& p,
                         const QK& q0, const QK& q1, double wa, double wb) {
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
                sink.put(s.idx, s.est0, s.delta0, s.nevals, ST_NAN);
                dead = true;
                break;
            }
            const int lv = (q.delta == q.asc && q.delta != 0.0) ? 1 : 0;
            s.est0 += q.estimate;  // add_qkresult, qagp.rs:403-408
            s.delta0 += q.delta;
            s.abs0 += q.absvalue;
            s.asc0 += q.asc;
            ws_push(S, s, hh ? p.ra1 : p.ra0, hh ? p.rb1 : p.rb0, q.estimate, q.delta, lv);
        }
        if (!dead && s.seed_k >= s.nint) {  // all windows seeded: qagp.rs:132-179
            s.seeding = false;
            double deltasum = 0.;
            for (int k = 0; k < s.size; ++k) {
                if (S.LV(k) > 0) S.D(k) = s.delta0;
                deltasum += S.D(k);
            }
            for (int k = 0; k < s.size; ++k) S.LV(k) = 0;
            ws_sort_results(S, s);
            const double tolerance = tol.to_abs(fabs(s.est0));
            const double round_off = 100. * F64_EPSILON * s.abs0;
            if (s.delta0 <= round_off && s.delta0 > tolerance) {
                s.active = false;
                sink.put(s.idx, s.est0, s.delta0, s.nevals, ST_ROUNDOFF_ERROR);
            } else if ((s.delta0 <= tolerance && s.delta0 != s.asc0) || s.delta0 == 0.0) {
                s.active = false;
                sink.put(s.idx, s.est0, s.delta0, s.nevals, ST_NONE);
            } else if ((unsigned long long)s.nevals == budget) {
                s.active = false;
                sink.put(s.idx, s.est0, s.delta0, s.nevals, ST_INSUFFICIENT_ITERATION);
            } else {
                tab_append(S, s, s.est0);
                s.area = s.est0;
                s.res_ext = s.est0;
                s.err_ext = F64_MAX;
                s.errsum = deltasum;
                s.eolr = deltasum;
                s.ertest = tolerance;
                s.max_iters = (unsigned long long)s.nint + (budget - s.nevals) / 50ull;
                s.iteration = (unsigned long long)s.nint;
            }
        }
    } else if (!QAGP && s.seeding) {
        // initial_integral (qags.rs:315-376), one pass per trip
        const int pass = s.seed_k;
        const QK q = q0;
        s.nevals = pass ? 42u : 17u;
        const double tolerance = tol.to_abs(fabs(q.estimate));
        const double round_off = 100. * F64_EPSILON * q.absvalue;
        if (q.estimate != q.estimate) {
            s.active = false;
            sink.put(s.idx, q.estimate, q.delta, s.nevals, ST_NAN);
        } else if ((q.delta <= tolerance && q.delta != q.asc) || q.delta == 0.0) {
            s.active = false;
            sink.put(s.idx, q.estimate, q.delta, s.nevals, ST_NONE);
        } else if (q.delta <= round_off && q.delta > tolerance) {
            s.active = false;
            sink.put(s.idx, q.estimate, q.delta, s.nevals, ST_ROUNDOFF_ERROR);
        } else if (budget < (unsigned long long)(42 + pass * 25)) {
            s.active = false;
            sink.put(s.idx, q.estimate, q.delta, s.nevals, ST_INSUFFICIENT_ITERATION);
        } else if (pass == 0 && !(q.delta > tolerance * 1024.)) {
            s.seed_k = 1;
        } else {
            // main loop set-up (qags.rs:98-118)
            s.seeding = false;
            s.est0 = q.estimate;
            s.delta0 = q.delta;
            s.abs0 = q.absvalue;
            s.asc0 = 0.;
            s.max_iters = (budget - s.nevals) / 50ull;
            s.limit = (int)reserve_cap(cap0, s.max_iters + 1ull);
            if (rsv) *rsv = (RSV_QAGS << 3) | (s.nevals << 5);
            ws_push(S, s, wa, wb, s.est0, s.delta0, 0);
            tab_append(S, s, s.est0);
            s.area = s.est0;
            s.errsum = s.delta0;
            s.res_ext = s.est0;
            s.err_ext = F64_MAX;
            s.ertest = 0.;
            s.eolr = 0.;
            s.iteration = 1;
            if (s.iteration > s.max_iters) adaptive_epilogue(S, s, sink);
        }
    } else {
        const int wi = s.wi;  // the bisected interval (untouched since the plan step)
        adaptive_step<QAGP>(S, s, tol, budget, sink, p.ra0, p.rb1, p.rb0, S.E(wi), S.D(wi), S.LV(wi),
                            q0, q1);
    }
}

// ---- shared state of one warp --------------------------------------------------------------
struct OuterCtx {
    double fp[MAX_PARAMS];  // integrand parameters of this outer integral
    double yq[2];           // y-range parameters
    const double* pts_xy;   // its singular points (x, y pairs)
    double wa, wb;          // transformed outer range
    unsigned long long nevals, inner_cap, budget0;  // running state; budget handed to this trip's tasks
    uint32_t error;         // first inner error
    Tol tol;                // this outer integral's tolerance and total budget
    unsigned long long max_evals;
    int npts;
    int base, K;            // this trip's tasks: board slots [base, base + K)
    bool xtr, exact;
    NodeSet nodes;
    TripPlan plan;
};
struct PoolShared {
    Lane inner[32];
    Lane outer[POOL_R];
    OuterCtx ctx[POOL_R];
    double est[POOL_T];
    uint32_t nev[POOL_T], meta[POOL_T];  // meta = status | RSV_* << 3 | n << 5
};
// 4 warps per block, 4 blocks per SM must fit the 227 KB of shared memory (1 KB per block is reserved)
static_assert(4 * sizeof(PoolShared) <= 55 * 1024, "PoolShared too large for 16 warps per SM");
struct PoolBoardSink {
    PoolShared* W;
    GKQ_DEV void put(long long t, double est, double, uint32_t nev, uint32_t st) const {
        W->est[t] = est;
        W->nev[t] = nev;
        W->meta[t] = (W->meta[t] & ~7u) | st;
    }
};
struct PoolOuterSink {
    Outputs O;
    const OuterCtx* ctx;
    // result.value.nevals = nevals; if error.is_some() { result.error = error }  (qags.rs:90-93)
    GKQ_DEV void put(long long i, double est, double dlt, uint32_t, uint32_t st) const {
        write_result(O, i, est, dlt, (uint32_t)ctx->nevals, ctx->error != ST_NONE ? ctx->error : st);
    }
};

template <class F2T, bool QAGP, int NPARAMS>
__global__ void __launch_bounds__(PB, 4) k_nested_pool(const Batch2D P) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned long long t = (unsigned long long)blockIdx.x * PB + threadIdx.x;
    // slab regions of this thread: [outer lists | inner lists]
    const Slab Sout = make_slab(P.slab.dbase, P.slab.wbase, P.slab.stride, t, P.cap_outer);
    const Slab Sin = make_slab(P.slab.dbase, P.slab.wbase, P.slab.stride, t, P.cap_inner,
                               slab_doubles(P.cap_outer, P.npts_max + 2), slab_ints(P.cap_outer));

    extern __shared__ __align__(16) unsigned char pool_smem[];
    PoolShared& W = reinterpret_cast<PoolShared*>(pool_smem)[threadIdx.x >> 5];
    __shared__ double s_xgk[12];
    if (threadIdx.x < 12) s_xgk[threadIdx.x] = c_xgk25[threadIdx.x];
    __syncthreads();

    const unsigned long long total = P.worklist ? *P.worklist_count : (unsigned long long)P.n;
    unsigned long long* head = QAGP ? &P.ctl->headP : &P.ctl->head2;
    const bool owner = lane < P.pool_r;
    Lane& so = W.outer[owner ? lane : 0];
    OuterCtx& cx = W.ctx[owner ? lane : 0];
    Lane& s = W.inner[lane];
    s.active = false;
    if (lane < POOL_R) {
        W.outer[lane].active = false;
        W.ctx[lane].K = 0;
        W.ctx[lane].base = 0;
        W.ctx[lane].exact = false;
    }
    PoolOuterSink osink;
    osink.O = P.out;
    osink.ctx = &cx;
    PoolBoardSink bsink;
    bsink.W = &W;
    bool exhausted = false;
    __syncwarp();

    for (;;) {
        // ---- 1. idle outer lanes take the next outer integrals of the batch ----
        while (owner && !so.active && !exhausted) {
            const unsigned long long slot = atomicAdd(head, 1ull);
            if (slot >= total) {
                exhausted = true;
            } else {
                const long long i = P.worklist ? (long long)P.worklist[slot] : (long long)slot;
#pragma unroll
                for (int k = 0; k < MAX_PARAMS; ++k)
                    cx.fp[k] = (k < NPARAMS)
                                   ? (P.fparam_stride ? P.fparams[(long long)k * P.fparam_stride + i] : P.fparams[k])
                                   : 0.0;
                cx.yq[0] = cx.yq[1] = 0.;
                for (int k = 0; k < P.nyparams && k < 2; ++k)
                    cx.yq[k] = P.yparam_stride ? P.yparams[(long long)k * P.yparam_stride + i] : P.yparams[k];
                cx.pts_xy = P.pts_xy;
                cx.npts = P.npts;
                if (P.pts_offsets) {
                    const long long o0 = P.pts_offsets[i];
                    cx.pts_xy = P.pts_xy + 2 * o0;
                    cx.npts = (int)(P.pts_offsets[i + 1] - o0);
                }
                double xa = P.x0[i * P.range_stride];
                double xb = P.x1[i * P.range_stride];
                const bool xtr = !is_finite(xa) || !is_finite(xb);
                if (xtr) {
                    xa = transform_point(xa);
                    xb = transform_point(xb);
                }
                cx.wa = xa;
                cx.wb = xb;
                cx.xtr = xtr;
                cx.exact = false;
                cx.nevals = 0;
                cx.error = ST_NONE;
                cx.inner_cap = 0;  // WorkSpace::new()
                cx.tol = load_tol2(P, i);
                cx.max_evals = load_max_evals2(P, i);
                // outer points = x of the 2-D points, transformed HERE when the x-range is infinite
                // (double/algorithm/qagp.rs:98-106) and once more inside make_sorted_points
                lane_start<QAGP>(Sout, so, i, xa, xb, xtr, cx.pts_xy, cx.npts, 2, xtr, cx.max_evals / 17ull, 0ull,
                                 osink, (uint32_t*)nullptr);
            }
        }
        // ---- 2. outer lanes plan their trip and post its nodes ----
        const bool oact = owner && so.active;
        int K = 0;
        if (oact) {
            lane_plan<QAGP, false>(Sout, so, cx.wa, cx.wb, cx.plan);
            NodeSet& ns = cx.nodes;
            ns.N = cx.plan.has17 ? 8 : 12;
            ns.lo[0] = cx.plan.ra0;
            ns.hi[0] = cx.plan.rb0;
            ns.lo[1] = cx.plan.ra1;
            ns.hi[1] = cx.plan.rb1;
            K = cx.plan.has17 ? 17 : (cx.plan.has0 ? (cx.plan.has1 ? 50 : 25) : 0);
            ns.K = K;
            // config1.max_evals = if error.is_some() { 0 } else { config.max_evals - nevals }
            cx.budget0 = (cx.error != ST_NONE || cx.nevals > cx.max_evals) ? 0ull : (cx.max_evals - cx.nevals);
        }
        if (__ballot_sync(FULL, oact) == 0u) break;  // nothing in flight and the batch is handed out
        int incl = K;
#pragma unroll
        for (int d = 1; d < POOL_R; d <<= 1) {
            const int up = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += up;
        }
        const int T = __shfl_sync(FULL, incl, POOL_R - 1);
        if (owner) {
            cx.base = incl - K;
            cx.K = K;
        }
        __syncwarp();

        // ---- 3. the inner lanes work the board off ----
        {
            Wrapped<InnerF<F2T>> g;
#pragma unroll
            for (int k = 0; k < MAX_PARAMS; ++k) g.f.f.p[k] = 0.;
            g.f.x = 0.;
            g.f.swap = P.swap_xy;
            g.transform = false;
            double ya = 0., yb = 0.;  // the lane's whole inner range (initial passes of QAGS)
            unsigned long long budget0 = 0, cap0 = 0;
            Tol tol = P.tol;  // the owning outer integral's tolerance
            int next = 0;
            for (;;) {
                const unsigned idle = __ballot_sync(FULL, !s.active);
                if (idle && next < T) {
                    const int tk = next + __popc(idle & ((1u << lane) - 1u));
                    next += __popc(idle);
                    if (!s.active && tk < T) {
                        int r = 0;
#pragma unroll
                        for (int q = 1; q < POOL_R; ++q)
                            if (tk >= W.ctx[q].base && W.ctx[q].K > 0) r = q;
                        // (bases ascend; owners without tasks have K = 0)
                        const OuterCtx& c = W.ctx[r];
                        if (!c.exact) {
                            double x = c.nodes(tk - c.base);
                            if (c.xtr) x = x * (1. / (1. - fabs(x)));
                            YRange yr;
                            yr.id = P.yrange_id;
                            yr.q[0] = c.yq[0];
                            yr.q[1] = c.yq[1];
                            double y0, y1;
                            yr(x, y0, y1);
                            const bool ytr = !is_finite(y0) || !is_finite(y1);
                            if (ytr) {
                                y0 = transform_point(y0);
                                y1 = transform_point(y1);
                            }
#pragma unroll
                            for (int k = 0; k < MAX_PARAMS; ++k) g.f.f.p[k] = c.fp[k];
                            g.f.x = x;
                            g.transform = ytr;
                            ya = y0;
                            yb = y1;
                            budget0 = c.budget0;
                            cap0 = c.inner_cap;
                            tol = c.tol;
                            lane_start<QAGP>(Sin, s, tk, y0, y1, ytr, c.pts_xy + 1, c.npts, 2, false, budget0,
                                             cap0, bsink, &W.meta[tk]);
                        }
                    }
                }
                const unsigned live = __ballot_sync(FULL, s.active);
                if (live == 0u) {
                    // (nothing started in this pass -- the lanes drew tasks of an outer integral that is on
                    // the exact path, or tasks that ended inside lane_start -- is not "board drained":
                    // round 1 left the loop here and committed stale values for the tasks behind them)
                    if (next >= T) break;
                    continue;
                }

                TripPlan p;
                p.ra0 = p.rb0 = p.ra1 = p.rb1 = 0.;
                p.has17 = p.has0 = p.has1 = false;
                if (s.active) lane_plan<QAGP, true>(Sin, s, ya, yb, p);

                // the hot part: integrand samples, warp-converged
                QK q0, q1;
                q0.estimate = q0.delta = q0.absvalue = q0.asc = 0.;
                q1 = q0;
                if (!QAGP && __any_sync(FULL, p.has17)) {
                    if (p.has17) q0 = qk_rule<8, NESTED_UNROLL>(g, p.ra0, p.rb0);
                }
                if (POOL_COOP_MAX_LIVE == 0 || __popc(live) > POOL_COOP_MAX_LIVE) {
#pragma unroll 1
                    for (int hh = 0; hh < 2; ++hh) {
                        const bool on = hh ? p.has1 : p.has0;
                        if (on) {
                            const QK q = qk_rule<12, NESTED_UNROLL>(g, hh ? p.ra1 : p.ra0, hh ? p.rb1 : p.rb0);
                            if (hh) q1 = q; else q0 = q;
                        }
                    }
                } else {
                    // latency mode: the warp spreads the 50 samples of each live lane over its lanes
                    unsigned todo = __ballot_sync(FULL, p.has0 || p.has1);
#pragma unroll 1
                    while (todo) {
                        const int own = __ffs(todo) - 1;
                        todo &= todo - 1;
                        Wrapped<InnerF<F2T>> go;
#pragma unroll
                        for (int k = 0; k < MAX_PARAMS; ++k)
                            go.f.f.p[k] = (k < NPARAMS) ? __shfl_sync(FULL, g.f.f.p[k], own) : 0.0;
                        go.f.x = __shfl_sync(FULL, g.f.x, own);
                        go.f.swap = P.swap_xy;
                        go.transform = __shfl_sync(FULL, (int)g.transform, own) != 0;
                        const bool oh0 = __shfl_sync(FULL, (int)p.has0, own) != 0;
                        const bool oh1 = __shfl_sync(FULL, (int)p.has1, own) != 0;
                        const double oa0 = __shfl_sync(FULL, p.ra0, own), ob0 = __shfl_sync(FULL, p.rb0, own);
                        const double oa1 = __shfl_sync(FULL, p.ra1, own), ob1 = __shfl_sync(FULL, p.rb1, own);
                        QK c0 = q0, c1 = q1;
                        qk25_pair_coop(go, s_xgk, oh0, oa0, ob0, oh1, oa1, ob1, c0, c1);
                        if (lane == own) {
                            q0 = c0;
                            q1 = c1;
                        }
                    }
                }
                if (s.active)
                    lane_finish<QAGP>(Sin, s, tol, budget0, cap0, bsink, &W.meta[s.idx], p, q0, q1, ya, yb);
            }
        }
        __syncwarp();

        // ---- 4. owners: commit in call order, reduce in summation order, advance the outer machine ----
        if (oact) {
            const double nan = __longlong_as_double(0x7ff8000000000000ll);
            const int base = cx.base;
            bool fine = !cx.exact;
            if (fine) {
                const unsigned long long budget0 = cx.budget0;
                unsigned long long spent = 0, cap = cx.inner_cap;
                uint32_t err = ST_NONE;
                bool failed = false;
#pragma unroll 1
                for (int k = 0; k < K; ++k) {
                    if (failed) {
                        // calls after the first failure get a zero budget: no evaluations, an
                        // InsufficientIteration of their own, NaN into the outer rule
                        W.est[base + k] = nan;
                        continue;
                    }
                    const uint32_t nev = W.nev[base + k], meta = W.meta[base + k];
                    const uint32_t st = meta & 7u, kind = (meta >> 3) & 3u;
                    const unsigned long long rn = meta >> 5;
                    if (spent != 0ull) {
                        const unsigned long long margin = 2ull * nev + 200ull + (kind == RSV_QAGP ? 100ull * rn : 0ull);
                        if (budget0 < spent || budget0 - spent < margin) {
                            fine = false;
                            break;
                        }
                    }
                    const unsigned long long mine = budget0 - spent;
                    if (kind == RSV_QAGS) cap = reserve_cap(cap, (mine - rn) / 50ull + 1ull);
                    if (kind == RSV_QAGP) cap = reserve_cap(cap, rn + (mine - 25ull * rn) / 50ull);
                    spent += nev;
                    if (st != ST_NONE) {
                        failed = true;
                        err = st;
                        W.est[base + k] = nan;
                    }
                }
                if (fine) {
                    cx.nevals += spent;
                    if (failed && cx.error == ST_NONE) cx.error = err;
                    cx.inner_cap = cap;
                }
            }
            if (!fine) {
                // exact path: this lane walks the inner calls of the trip in order
                cx.exact = true;
                OuterIntegrand<F2T, QAGP> G;
#pragma unroll
                for (int k = 0; k < MAX_PARAMS; ++k) G.f.p[k] = cx.fp[k];
                G.yr.id = P.yrange_id;
                G.yr.q[0] = cx.yq[0];
                G.yr.q[1] = cx.yq[1];
                G.tol = cx.tol;
                G.total_evals = cx.max_evals;
                G.nevals = cx.nevals;
                G.error = cx.error;
                G.inner_cap = cx.inner_cap;
                G.Sin = Sin;
                G.pts_xy = cx.pts_xy;
                G.npts = cx.npts;
                G.swap_xy = P.swap_xy;
                G.transform = false;  // applied below, like the speculative path
#pragma unroll 1
                for (int k = 0; k < K; ++k) {
                    double x = cx.nodes(k);
                    if (cx.xtr) x = x * (1. / (1. - fabs(x)));
                    W.est[base + k] = G.inner_integral(x);
                }
                cx.nevals = G.nevals;
                cx.error = G.error;
                cx.inner_cap = G.inner_cap;
            }
            if (cx.xtr) {
                // single/util.rs:74-103: f(t / (1 - |t|)) / (1 - |t|)^2
#pragma unroll 1
                for (int k = 0; k < K; ++k) {
                    const double coef = 1. / (1. - fabs(cx.nodes(k)));
                    W.est[base + k] = W.est[base + k] * coef * coef;
                }
            }
            QK q0, q1;
            q0.estimate = q0.delta = q0.absvalue = q0.asc = 0.;
            q1 = q0;
            if (cx.plan.has17) {
                double fv1[8], fv2[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    fv1[j] = W.est[base + j];
                    fv2[j] = W.est[base + 8 + j];
                }
                q0 = qk_reduce<8>(fv1, fv2, W.est[base + 16], cx.plan.ra0, cx.plan.rb0);
            } else {
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    if (half ? !cx.plan.has1 : !cx.plan.has0) continue;
                    const int b0 = base + 25 * half;
                    double fv1[12], fv2[12];
#pragma unroll
                    for (int j = 0; j < 12; ++j) {
                        fv1[j] = W.est[b0 + j];
                        fv2[j] = W.est[b0 + 12 + j];
                    }
                    const QK q = qk_reduce<12>(fv1, fv2, W.est[b0 + 24], half ? cx.plan.ra1 : cx.plan.ra0,
                                               half ? cx.plan.rb1 : cx.plan.rb0);
                    if (half) q1 = q; else q0 = q;
                }
            }
            lane_finish<QAGP>(Sout, so, cx.tol, cx.max_evals / 17ull, 0ull, osink, (uint32_t*)nullptr, cx.plan, q0, q1,
                              cx.wa, cx.wb);
        }
        __syncwarp();
    }
}

}  // namespace gkq
