// kernels_2d_flat.cuh -- the nested 2-D drivers as a LEVEL-SYNCHRONOUS pipeline (large batches).
//
// (included by kernels_2d.cuh; reference: double/algorithm/qags.rs:46-97, qagp.rs:70-139)
//
// k_nested_pool keeps a few outer integrals per warp and lets the warp's 32 lanes work their inner
// integrals off a shared-memory board.  Its lanes idle whenever the board holds fewer long inner
// integrals than lanes (D1: 2 of the 25 inner integrals of an outer rule take 12-20 trips, the others
// one), the outer bookkeeping runs with 1-6 of 32 threads active, and the kernel needs 128 registers
// and ~55 KB of shared memory per block.  For batches that fill the machine many times over the nest
// is unrolled in TIME instead: all outer integrals advance in lock-step rounds, and the inner
// integrals of a round -- tens of tasks per outer integral, millions per round -- are ONE flat 1-D
// batch for the same self-refilling lane machine that k_adaptive runs:
//
//   k2_start   thread per outer integral: parameters -> OuterCtx, outer lane_start, active list
//   per round:
//     k2_plan    thread per active outer integral: plan the trip (17 / 25 / 50 nodes), allocate its
//                slice of the round's task arrays, budget handed to the tasks (budget at trip start)
//     k2_inner   persistent grid, one lane per inner integral, refilled from ONE global queue the
//                moment its integral terminates: full warps whatever the mix of long and short tasks
//     k2_commit  thread per active outer integral: commit the tasks in the reference's call order
//                (running budget, first-error latch, workspace capacity -- the margin argument of
//                kernels_2d_pool.cuh), reduce in the reference's summation order, advance the outer
//                machine; survivors form the next round's active list
//     k2_exact   the (rare) outer integrals whose budget is nearly exhausted leave the rounds: a
//                thread walks the rest of the nest alone, inner call after inner call, like k_nested
// The host enqueues rounds until an active count of zero comes back (one round of look-ahead, so the
// GPU never waits for the host).  Results are bit-identical to the oracle and to both other mappings.
#pragma once

namespace gkq {

constexpr int FB = 128;  // threads per block of the flat kernels
#ifndef GKQ_FLAT_UNROLL
#define GKQ_FLAT_UNROLL NESTED_UNROLL
#endif
#ifndef GKQ_FLAT_MIN_BLOCKS
#define GKQ_FLAT_MIN_BLOCKS 4
#endif

struct FlatBoardSink {
    double* est;
    uint32_t* nev;
    uint32_t* meta;
    GKQ_DEV void put(long long t, double e, double, uint32_t n, uint32_t st) const {
        est[t] = e;
        nev[t] = n;
        meta[t] = (meta[t] & ~7u) | st;
    }
};

GKQ_DEV Slab flat_outer_slab(const Batch2D& P, unsigned int o) {
    return make_slab(P.flat.oslab.dbase, P.flat.oslab.wbase, P.flat.oslab.stride, o, P.cap_outer);
}
GKQ_DEV Slab flat_inner_slab(const Batch2D& P, unsigned long long t) {
    return make_slab(P.slab.dbase, P.slab.wbase, P.slab.stride, t, P.cap_inner);
}

static __global__ void k2_round_reset(Ctl2* c, int round) {
    c->nact[(round + 1) & 1] = 0ull;
    c->ntasks = c->qhead = c->nexact = c->xhead = 0ull;
}

// ---- start: one thread per outer integral ---------------------------------------------------------
template <class F2T, bool QAGP, int NPARAMS>
__global__ void __launch_bounds__(FB) k2_start(const Batch2D P) {
    const unsigned long long total = P.worklist ? *P.worklist_count : (unsigned long long)P.n;
    for (unsigned long long k = (unsigned long long)blockIdx.x * FB + threadIdx.x; k - (threadIdx.x & 31) < total;
         k += (unsigned long long)gridDim.x * FB) {
        bool alive = false;
        if (k < total) {
            const unsigned int o = (unsigned int)k;
            const long long i = P.worklist ? (long long)P.worklist[k] : (long long)k;
            OuterCtx cx;
#pragma unroll
            for (int q = 0; q < MAX_PARAMS; ++q)
                cx.fp[q] = (q < NPARAMS) ? (P.fparam_stride ? P.fparams[(long long)q * P.fparam_stride + i] : P.fparams[q]) : 0.0;
            cx.yq[0] = cx.yq[1] = 0.;
            for (int q = 0; q < P.nyparams && q < 2; ++q)
                cx.yq[q] = P.yparam_stride ? P.yparams[(long long)q * P.yparam_stride + i] : P.yparams[q];
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
            cx.budget0 = 0;
            cx.base = 0;
            cx.K = 0;
            cx.tol = load_tol2(P, i);
            cx.max_evals = load_max_evals2(P, i);
            cx.nodes.N = 12;
            cx.nodes.K = 0;
            cx.nodes.lo[0] = cx.nodes.lo[1] = cx.nodes.hi[0] = cx.nodes.hi[1] = 0.;
            cx.plan.ra0 = cx.plan.rb0 = cx.plan.ra1 = cx.plan.rb1 = 0.;
            cx.plan.has17 = cx.plan.has0 = cx.plan.has1 = false;
            Lane so;
            PoolOuterSink osink;
            osink.O = P.out;
            osink.ctx = &cx;
            const Slab Sout = flat_outer_slab(P, o);
            // outer points = x of the 2-D points, transformed HERE when the x-range is infinite
            // (double/algorithm/qagp.rs:98-106) and once more inside make_sorted_points
            lane_start<QAGP>(Sout, so, i, xa, xb, xtr, cx.pts_xy, cx.npts, 2, xtr, cx.max_evals / 17ull, 0ull, osink,
                             (uint32_t*)nullptr);
            alive = so.active;
            P.flat.olane[o] = so;
            P.flat.octx[o] = cx;
        }
        list_append(alive, (unsigned int)k, P.flat.act[0], &P.flat.ctl->nact[0]);
    }
}

// ---- plan: one thread per active outer integral ---------------------------------------------------
template <bool QAGP>
__global__ void __launch_bounds__(FB) k2_plan(const Batch2D P, int round) {
    const unsigned long long total = P.flat.ctl->nact[round & 1];
    const unsigned int* act = P.flat.act[round & 1];
    for (unsigned long long k = (unsigned long long)blockIdx.x * FB + threadIdx.x; k < total;
         k += (unsigned long long)gridDim.x * FB) {
        const unsigned int o = act[k];
        Lane so = P.flat.olane[o];
        OuterCtx& cx = P.flat.octx[o];
        const Slab Sout = flat_outer_slab(P, o);
        TripPlan plan;
        lane_plan<QAGP, false>(Sout, so, cx.wa, cx.wb, plan);
        NodeSet ns;
        ns.N = plan.has17 ? 8 : 12;
        ns.lo[0] = plan.ra0;
        ns.hi[0] = plan.rb0;
        ns.lo[1] = plan.ra1;
        ns.hi[1] = plan.rb1;
        const int K = plan.has17 ? 17 : (plan.has0 ? (plan.has1 ? 50 : 25) : 0);
        ns.K = K;
        cx.plan = plan;
        cx.nodes = ns;
        cx.K = K;
        // config1.max_evals = if error.is_some() { 0 } else { config.max_evals - nevals }
        cx.budget0 = (cx.error != ST_NONE || cx.nevals > cx.max_evals) ? 0ull : (cx.max_evals - cx.nevals);
        const unsigned long long base = K > 0 ? atomicAdd(&P.flat.ctl->ntasks, (unsigned long long)K) : 0ull;
        cx.base = (int)base;
        for (int j = 0; j < K; ++j) {
            P.flat.t_owner[base + j] = o;
            P.flat.t_k[base + j] = (uint32_t)j;
        }
        // (lane_plan only moves the seeding cursor of the outer machine)
        P.flat.olane[o].seed_k = so.seed_k;
        P.flat.olane[o].w0 = so.w0;
        P.flat.olane[o].w1 = so.w1;
    }
}

// ---- inner: persistent self-refilling lanes over the round's tasks -------------------------------
template <class F2T, bool QAGP, int NPARAMS>
__global__ void __launch_bounds__(FB, GKQ_FLAT_MIN_BLOCKS) k2_inner(const Batch2D P) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned long long t = (unsigned long long)blockIdx.x * FB + threadIdx.x;
    const Slab Sin = flat_inner_slab(P, t);
    __shared__ Lane s_lanes[FB];
    Lane& s = s_lanes[threadIdx.x];
    s.active = false;
    s.idx = 0;
    FlatBoardSink bsink;
    bsink.est = P.flat.t_est;
    bsink.nev = P.flat.t_nev;
    bsink.meta = P.flat.t_meta;
    const unsigned long long total = P.flat.ctl->ntasks;
    unsigned long long* head = &P.flat.ctl->qhead;

    Wrapped<InnerF<F2T>> g;
#pragma unroll
    for (int k = 0; k < MAX_PARAMS; ++k) g.f.f.p[k] = 0.;
    g.f.x = 0.;
    g.f.swap = P.swap_xy;
    g.transform = false;
    double ya = 0., yb = 0.;  // the lane's whole inner range (initial passes of QAGS)
    unsigned long long budget0 = 0, cap0 = 0;
    Tol tol = P.tol;
    bool exhausted = false;

    for (;;) {
        const unsigned want = __ballot_sync(FULL, !s.active && !exhausted);
        if (want) {
            const int leader = __ffs(want) - 1;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(head, (unsigned long long)__popc(want));
            base = __shfl_sync(FULL, base, leader);
            if (!s.active && !exhausted) {
                const unsigned long long tk = base + (unsigned long long)__popc(want & ((1u << lane) - 1u));
                if (tk >= total) {
                    exhausted = true;
                } else {
                    const OuterCtx& c = P.flat.octx[P.flat.t_owner[tk]];
                    double x = c.nodes((int)P.flat.t_k[tk]);
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
                    lane_start<QAGP>(Sin, s, (long long)tk, y0, y1, ytr, c.pts_xy + 1, c.npts, 2, false, budget0, cap0,
                                     bsink, &P.flat.t_meta[tk]);
                }
            }
        }
        const unsigned live = __ballot_sync(FULL, s.active);
        if (live == 0u) {
            if (__ballot_sync(FULL, !exhausted) == 0u) break;
            continue;  // (a lane's task ended inside lane_start: ask again)
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
            if (p.has17) q0 = qk_rule<8, GKQ_FLAT_UNROLL>(g, p.ra0, p.rb0);
        }
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
            const bool on = hh ? p.has1 : p.has0;
            if (on) {
                const QK q = qk_rule<12, GKQ_FLAT_UNROLL>(g, hh ? p.ra1 : p.ra0, hh ? p.rb1 : p.rb0);
                if (hh) q1 = q; else q0 = q;
            }
        }
        if (s.active) lane_finish<QAGP>(Sin, s, tol, budget0, cap0, bsink, &P.flat.t_meta[s.idx], p, q0, q1, ya, yb);
    }
}

// Reduce the committed values of one outer trip in the reference's summation order and advance the
// outer machine.  `val(k)` = value of node k of the trip (NaN injected, infinite-range factor applied).
template <bool QAGP, class Val>
GKQ_DEV void flat_outer_advance(const Slab& Sout, Lane& so, OuterCtx& cx, PoolOuterSink& osink, Val val) {
    QK q0, q1;
    q0.estimate = q0.delta = q0.absvalue = q0.asc = 0.;
    q1 = q0;
    if (cx.plan.has17) {
        double fv1[8], fv2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            fv1[j] = val(j);
            fv2[j] = val(8 + j);
        }
        q0 = qk_reduce<8>(fv1, fv2, val(16), cx.plan.ra0, cx.plan.rb0);
    } else {
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            if (half ? !cx.plan.has1 : !cx.plan.has0) continue;
            const int b0 = 25 * half;
            double fv1[12], fv2[12];
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                fv1[j] = val(b0 + j);
                fv2[j] = val(b0 + 12 + j);
            }
            const QK q = qk_reduce<12>(fv1, fv2, val(b0 + 24), half ? cx.plan.ra1 : cx.plan.ra0,
                                       half ? cx.plan.rb1 : cx.plan.rb0);
            if (half) q1 = q; else q0 = q;
        }
    }
    lane_finish<QAGP>(Sout, so, cx.tol, cx.max_evals / 17ull, 0ull, osink, (uint32_t*)nullptr, cx.plan, q0, q1, cx.wa,
                      cx.wb);
}

// ---- commit: one thread per active outer integral --------------------------------------------------
template <bool QAGP>
__global__ void __launch_bounds__(FB) k2_commit(const Batch2D P, int round) {
    const unsigned long long total = P.flat.ctl->nact[round & 1];
    const unsigned int* act = P.flat.act[round & 1];
    for (unsigned long long k = (unsigned long long)blockIdx.x * FB + threadIdx.x; k - (threadIdx.x & 31) < total;
         k += (unsigned long long)gridDim.x * FB) {
        bool alive = false, to_exact = false;
        unsigned int o = 0;
        if (k < total) {
            o = act[k];
            OuterCtx cx = P.flat.octx[o];
            const int K = cx.K;
            const unsigned long long base = (unsigned long long)(unsigned int)cx.base;
            const double nan = __longlong_as_double(0x7ff8000000000000ll);
            // commit in the reference's call order with a running budget (kernels_2d_pool.cuh: a call that
            // spent nev evaluations never read its budget B or its list capacity while B >= 2 nev + 200
            // (+ 100 per break-point window))
            bool fine = true, failed = false;
            const unsigned long long budget0 = cx.budget0;
            unsigned long long spent = 0, cap = cx.inner_cap;
            uint32_t err = ST_NONE;
            int first_failed = K;
#pragma unroll 1
            for (int j = 0; j < K; ++j) {
                if (failed) break;  // later calls get a zero budget: no evaluations, NaN into the outer rule
                const uint32_t nev = P.flat.t_nev[base + j], meta = P.flat.t_meta[base + j];
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
                    first_failed = j;
                }
            }
            if (!fine) {
                // budget nearly exhausted: this outer integral leaves the rounds (k2_exact)
                to_exact = true;
            } else {
                cx.nevals += spent;
                if (failed && cx.error == ST_NONE) cx.error = err;
                cx.inner_cap = cap;
                Lane so = P.flat.olane[o];
                PoolOuterSink osink;
                osink.O = P.out;
                osink.ctx = &cx;
                const Slab Sout = flat_outer_slab(P, o);
                const double* te = P.flat.t_est + base;
                const bool xtr = cx.xtr;
                const NodeSet ns = cx.nodes;
                auto val = [&](int j) -> double {
                    double v = j >= first_failed ? nan : te[j];
                    if (xtr) {  // single/util.rs:74-103: f(t / (1 - |t|)) / (1 - |t|)^2
                        const double coef = 1. / (1. - fabs(ns(j)));
                        v = v * coef * coef;
                    }
                    return v;
                };
                flat_outer_advance<QAGP>(Sout, so, cx, osink, val);
                alive = so.active;
                P.flat.olane[o] = so;
                P.flat.octx[o] = cx;
            }
        }
        list_append(alive, o, P.flat.act[(round + 1) & 1], &P.flat.ctl->nact[(round + 1) & 1]);
        list_append(to_exact, o, P.flat.xlist, &P.flat.ctl->nexact);
    }
}

// ---- exact: a thread walks the rest of the nest of one outer integral alone ----------------------
template <class F2T, bool QAGP, int NPARAMS>
__global__ void __launch_bounds__(FB) k2_exact(const Batch2D P) {
    const unsigned long long t = (unsigned long long)blockIdx.x * FB + threadIdx.x;
    const Slab Sin = flat_inner_slab(P, t);
    const unsigned long long total = P.flat.ctl->nexact;
    for (;;) {
        const unsigned long long slot = atomicAdd(&P.flat.ctl->xhead, 1ull);
        if (slot >= total) break;
        const unsigned int o = P.flat.xlist[slot];
        Lane so = P.flat.olane[o];
        OuterCtx cx = P.flat.octx[o];
        const Slab Sout = flat_outer_slab(P, o);
        PoolOuterSink osink;
        osink.O = P.out;
        osink.ctx = &cx;
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
        double est[50];
        // the current trip is already planned (cx.plan / cx.nodes / cx.K); later ones are planned here
        while (so.active) {
            const int K = cx.K;
#pragma unroll 1
            for (int k = 0; k < K; ++k) {
                double x = cx.nodes(k);
                if (cx.xtr) x = x * (1. / (1. - fabs(x)));
                double v = G.inner_integral(x);
                if (cx.xtr) {
                    const double coef = 1. / (1. - fabs(cx.nodes(k)));
                    v = v * coef * coef;
                }
                est[k] = v;
            }
            cx.nevals = G.nevals;
            cx.error = G.error;
            cx.inner_cap = G.inner_cap;
            flat_outer_advance<QAGP>(Sout, so, cx, osink, [&](int j) -> double { return est[j]; });
            if (!so.active) break;
            lane_plan<QAGP, false>(Sout, so, cx.wa, cx.wb, cx.plan);
            cx.nodes.N = cx.plan.has17 ? 8 : 12;
            cx.nodes.lo[0] = cx.plan.ra0;
            cx.nodes.hi[0] = cx.plan.rb0;
            cx.nodes.lo[1] = cx.plan.ra1;
            cx.nodes.hi[1] = cx.plan.rb1;
            cx.K = cx.plan.has17 ? 17 : (cx.plan.has0 ? (cx.plan.has1 ? 50 : 25) : 0);
            cx.nodes.K = cx.K;
        }
    }
}

}  // namespace gkq
