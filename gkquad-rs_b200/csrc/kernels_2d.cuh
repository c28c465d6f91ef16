// kernels_2d.cuh -- the nested double-integral drivers QAGS2 / QAGP2.
//
// Reference: double/algorithm/qags.rs:46-97 and double/algorithm/qagp.rs:70-139 -- an outer
// 1-D algorithm over x whose integrand runs an inner 1-D algorithm over y(x), with
//   * one running evaluation budget: before EACH inner call max_evals = total - nevals so far
//     (0 once an inner call has failed), in the outer rule's evaluation order
//     (single/qk/naive.rs:60-68: lower nodes j = 0..n-1, upper nodes, centre);
//   * an inner workspace that is re-used, so its Vec capacity (= `limit`) is the running max
//     of the reserve requests (qags.rs:60-61);
//   * an inner failure injected as NaN into the outer integrand and the FIRST inner error
//     overriding the outer status (qags.rs:75-80, 90-93); nevals = sum of inner evaluations.
//
// Mapping: one thread owns one outer integral and runs the whole nest sequentially out of
// two slab regions (outer list + inner list); threads pull outer integrals from a global
// queue until it is empty.  Inner rule evaluations (the hot part) are ordinary pure functor
// samples; the outer rule uses the ORDERED variant because its "integrand" has state.
#pragma once
#include "adaptive.cuh"
#include "functors.cuh"
#include "kernels_1d.cuh"

namespace gkq {

// Node pairs inlined per trip of the rule loops inside the nested kernels.  Lanes / warps of these
// kernels sit in different phases of different inner integrals, so the instruction stream jumps
// between the rule and the bookkeeping code all the time: what counts is that the whole trip fits
// the 32 KB instruction cache (ncu: `no_instruction` was the top stall with fully unrolled rules).
constexpr int NESTED_UNROLL = 1;

struct Res {
    double est, dlt;
    uint32_t nev, st;
    // how the call sized its workspace (Vec::reserve request as a function of the budget B):
    //   RSV_QAGS: (B - rsv_n) / 50 + 1 with rsv_n = evaluations of the initial passes (qags.rs:98-99)
    //   RSV_QAGP: rsv_n + (B - 25 rsv_n) / 50 with rsv_n = number of windows (qagp.rs:84-86)
    // (k_nested_warp replays the persistent inner workspace's capacity from these)
    uint32_t rsv_kind, rsv_n;
};
enum : uint32_t { RSV_NONE = 0, RSV_QAGS = 1, RSV_QAGP = 2 };

// How a sequential driver obtains its rule values:
//   RM_PURE        pure integrand, samples in any order (inner integrals)
//   RM_ORDERED     integrand with state, evaluated node by node in the reference's order
//   RM_COLLECTIVE  the integrand evaluates a whole rule (or the two rules of a bisection step)
//                  itself: g.rule<N>(a, b), g.pair(a, c, b, q1, q2) (warp-cooperative outer)
enum : int { RM_PURE = 0, RM_ORDERED = 1, RM_COLLECTIVE = 2 };

// qk rule in the reference's evaluation order, for integrands with side effects.
template <int N, class G>
__device__ __noinline__ QK qk_rule_ordered(G& g, double a, double b) {
    using R = Rule<N>;
    double fv1[N], fv2[N];
    const double center = 0.5 * (a + b);
    const double half_length = 0.5 * (b - a);
    const double abs_half_length = fabs(half_length);
#pragma unroll 1
    for (int j = 0; j < N; ++j) fv1[j] = g(center - half_length * R::xgk(j));
#pragma unroll 1
    for (int j = 0; j < N; ++j) fv2[j] = g(center + half_length * R::xgk(j));
    const double f_center = g(center);

    double result_gauss = 0.;
    double result_kronrod = f_center * R::WCK;
    double result_abs = fabs(result_kronrod);
#pragma unroll 1
    for (int j = 0; j < N; ++j) {
        double fsum = fv1[j] + fv2[j];
        result_kronrod += R::wgk(j) * fsum;
        result_abs += R::wgk(j) * (fabs(fv1[j]) + fabs(fv2[j]));
        if ((j & 1) == 0) result_gauss += R::wg(j >> 1) * fsum;
    }
    const double mean = result_kronrod * 0.5;
    double result_asc = R::WCK * fabs(f_center - mean);
#pragma unroll 1
    for (int j = 0; j < N; ++j)
        result_asc += R::wgk(j) * (fabs(fv1[j] - mean) + fabs(fv2[j] - mean));
    double err = (result_kronrod - result_gauss) * half_length;
    result_kronrod *= half_length;
    result_abs *= abs_half_length;
    result_asc *= abs_half_length;
    QK q;
    q.estimate = result_kronrod;
    q.delta = rescale_error(err, result_abs, result_asc);
    q.absvalue = result_abs;
    q.asc = result_asc;
    return q;
}

template <int MODE, int N, int U, class G>
GKQ_DEV QK rule_any(G& g, double a, double b) {
    if constexpr (MODE == RM_COLLECTIVE)
        return g.template rule<N>(a, b);
    else if constexpr (MODE == RM_ORDERED)
        return qk_rule_ordered<N>(g, a, b);
    else
        return qk_rule<N, U>(g, a, b);
}
// the two 25-point rules of one bisection step, left half first
template <int MODE, int U, class G>
GKQ_DEV void rule_pair_any(G& g, double a, double c, double b, QK& q1, QK& q2) {
    if constexpr (MODE == RM_COLLECTIVE) {
        g.pair(a, c, b, q1, q2);
    } else {
        q1 = rule_any<MODE, 12, U>(g, a, c);
        q2 = rule_any<MODE, 12, U>(g, c, b);
    }
}

// Sequential QAGS for one thread: single/algorithm/qags.rs:67-311 on an already
// transformed range.  `ws_cap` is the workspace capacity before the call and is updated
// like Vec::reserve would (persistent inner workspace).
template <int MODE, int U, class G>
__device__ __noinline__ Res seq_qags(G& g, double a, double b, const Tol& tol,
                                     unsigned long long max_evals, unsigned long long& ws_cap,
                                     const Slab& S) {
    Res r;
    r.est = 0.;
    r.dlt = F64_MAX;
    r.nev = 0;
    r.st = ST_INSUFFICIENT_ITERATION;
    r.rsv_kind = RSV_NONE;
    r.rsv_n = 0;
    if (max_evals < 17) return r;  // qags.rs:86-88

    // initial_integral (qags.rs:315-376)
    QK q;
    uint32_t nevals = 0;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 0) {
            nevals += 17;
            q = rule_any<MODE, 8, U>(g, a, b);
        } else {
            nevals += 25;
            q = rule_any<MODE, 12, U>(g, a, b);
        }
        r.est = q.estimate;
        r.dlt = q.delta;
        r.nev = nevals;
        if (q.estimate != q.estimate) {
            r.st = ST_NAN;
            return r;
        }
        const double tolerance = tol.to_abs(fabs(q.estimate));
        if ((q.delta <= tolerance && q.delta != q.asc) || q.delta == 0.0) {
            r.st = ST_NONE;
            return r;
        }
        const double round_off = 100. * F64_EPSILON * q.absvalue;
        if (q.delta <= round_off && q.delta > tolerance) {
            r.st = ST_ROUNDOFF_ERROR;
            return r;
        }
        if (max_evals < (unsigned long long)(42 + pass * 25)) {
            r.st = ST_INSUFFICIENT_ITERATION;
            return r;
        }
        if (pass == 0 && q.delta > tolerance * 1024.) break;
    }

    Lane s;
    lane_reset_common(s);
    s.idx = 0;
    s.transform = false;
    s.est0 = q.estimate;
    s.delta0 = q.delta;
    s.abs0 = q.absvalue;
    s.asc0 = 0.;
    s.nevals = nevals;
    s.max_iters = (max_evals - nevals) / 50;
    ws_cap = reserve_cap(ws_cap, s.max_iters + 1);  // ws.clear(); ws.reserve(..)  qags.rs:98-99
    r.rsv_kind = RSV_QAGS;
    r.rsv_n = nevals;
    s.limit = (int)ws_cap;
    ws_push(S, s, a, b, s.est0, s.delta0, 0);
    tab_append(S, s, s.est0);
    s.area = s.est0;
    s.errsum = s.delta0;
    s.res_ext = s.est0;
    s.err_ext = F64_MAX;
    s.ertest = 0.;
    s.eolr = 0.;
    s.iteration = 1;
    s.active = true;
    LocalSink sink;
    sink.est = 0.;
    sink.dlt = 0.;
    sink.nev = 0;
    sink.st = 0;
    if (s.iteration > s.max_iters) adaptive_epilogue(S, s, sink);
    while (s.active) {
        const int wi = s.wi;
        const double ia = S.A(wi), ib = S.B(wi), ie = S.E(wi), id = S.D(wi);
        const int ilv = S.LV(wi);
        const double center = (ia + ib) * 0.5;
        QK q1, q2;
        rule_pair_any<MODE, U>(g, ia, center, ib, q1, q2);
        adaptive_step<false>(S, s, tol, max_evals, sink, ia, ib, center, ie, id, ilv, q1, q2);
    }
    r.est = sink.est;
    r.dlt = sink.dlt;
    r.nev = sink.nev;
    r.st = sink.st;
    return r;
}

// Sequential QAG for one thread (deprecated in the reference): single/algorithm/qag.rs:65-248.
template <int MODE, int U, class G>
__device__ __noinline__ Res seq_qag(G& g, double a, double b, const Tol& tol, unsigned long long max_evals,
                                    unsigned long long& ws_cap, const Slab& S) {
    Res r;
    r.est = 0.;
    r.dlt = F64_MAX;
    r.nev = 0;
    r.st = ST_INSUFFICIENT_ITERATION;
    r.rsv_kind = RSV_NONE;
    r.rsv_n = 0;
    if (max_evals < 17) return r;  // qag.rs:76-78
    QK q;
    uint32_t nevals = 0;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {  // initial_integral, qag.rs:185-234
        if (pass == 0) {
            nevals += 17;
            q = rule_any<MODE, 8, U>(g, a, b);
        } else {
            nevals += 25;
            q = rule_any<MODE, 12, U>(g, a, b);
        }
        r.est = q.estimate;
        r.dlt = q.delta;
        r.nev = nevals;
        if (q.estimate != q.estimate) {
            r.st = ST_NAN;
            return r;
        }
        const double tolerance = tol.to_abs(fabs(q.estimate));
        if ((q.delta <= tolerance && q.delta != q.asc) || q.delta == 0.0) {
            r.st = ST_NONE;
            return r;
        }
        const double round_off = 50. * F64_EPSILON * q.absvalue;  // :217
        if (q.delta <= round_off && q.delta > tolerance) {
            r.st = ST_ROUNDOFF_ERROR;
            return r;
        }
        if (max_evals < (unsigned long long)(42 + pass * 25)) {
            r.st = ST_INSUFFICIENT_ITERATION;
            return r;
        }
        if (pass == 0 && q.delta > tolerance * 1024.) break;
    }
    double area = q.estimate, deltasum = q.delta;
    const unsigned long long max_iters = (max_evals - nevals) / 50ull;
    Lane s;
    lane_reset_common(s);
    ws_cap = reserve_cap(ws_cap, max_iters + 1ull);  // :97-98
    s.limit = (int)ws_cap;
    ws_push(S, s, a, b, q.estimate, q.delta, 0);
    int ro1 = 0, ro2 = 0;
    uint32_t error = ST_NONE;
#pragma unroll 1
    for (unsigned long long it = 1; it <= max_iters; ++it) {
        const int wi = s.wi;
        const double ia = S.A(wi), ib = S.B(wi), ie = S.E(wi), id = S.D(wi);
        const int lv = S.LV(wi) + 1;
        const double center = (ia + ib) * 0.5;
        QK q1, q2;
        rule_pair_any<MODE, U>(g, ia, center, ib, q1, q2);
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
                r.est = ws_sum_results(S, s) - ie + q1.estimate + q2.estimate;
                r.dlt = deltasum;
                r.nev = nevals;
                r.st = error;
                return r;
            }
        }
        ws_update(S, s, ia, center, q1.estimate, q1.delta, center, ib, q2.estimate, q2.delta, lv);
        if (deltasum <= tolerance) break;  // :170-172
    }
    r.est = ws_sum_results(S, s);  // :176
    r.dlt = deltasum;
    r.nev = nevals;
    r.st = error;
    return r;
}

// Sequential QAGP for one thread: single/algorithm/qagp.rs:68-372.  `pts` are the user
// break points (read through `PT(k)`), the range is already transformed.
template <int MODE, int U, class G, class PointAt>
__device__ __noinline__ Res seq_qagp(G& g, double a, double b, bool transform, PointAt user_pt,
                                     int npts, const Tol& tol, unsigned long long max_evals,
                                     unsigned long long& ws_cap, const Slab& S) {
    Res r;
    r.rsv_kind = RSV_NONE;
    r.rsv_n = 0;
    // make_sorted_points (qagp.rs:375-400) into the slab's PTS slots
    const bool ascending = a < b;
    const double mn = ascending ? a : b, mx = ascending ? b : a;
    S.PTS(0) = a;
    for (int k = 0; k < npts; ++k) {
        const double v = user_pt(k);
        S.PTS(1 + k) = transform ? transform_point(v) : v;
    }
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
    int len = 0;
    for (int k = 0; k <= npts; ++k) {
        const double x = S.PTS(k);
        if (mn <= x && x <= mx) {
            S.PTS(len) = x;
            len += 1;
        }
    }
    S.PTS(len) = b;
    const int nint = len;

    if (max_evals < (unsigned long long)nint * 25ull) {  // qagp.rs:79-81
        r.est = 0.;
        r.dlt = F64_MAX;
        r.nev = 0;
        r.st = ST_INSUFFICIENT_ITERATION;
        return r;
    }

    Lane s;
    lane_reset_common(s);
    s.idx = 0;
    s.transform = transform;
    ws_cap = reserve_cap(ws_cap, nint + (max_evals - (unsigned long long)nint * 25ull) / 50ull);
    r.rsv_kind = RSV_QAGP;
    r.rsv_n = (uint32_t)nint;
    s.limit = (int)ws_cap;
    s.nint = nint;
    s.nevals = 0;
    s.est0 = s.delta0 = s.abs0 = s.asc0 = 0.;

    // first pass over the break-point windows (qagp.rs:102-130)
#pragma unroll 1
    for (int w = 0; w < nint; ++w) {
        const double lo = S.PTS(w), hi = S.PTS(w + 1);
        if (fabs(hi - lo) < 100. * F64_MIN_POSITIVE) continue;
        const QK q = rule_any<MODE, 12, U>(g, lo, hi);
        s.nevals += 25;
        if (q.estimate != q.estimate) {
            r.est = s.est0;
            r.dlt = s.delta0;
            r.nev = s.nevals;
            r.st = ST_NAN;
            return r;
        }
        const int lv = (q.delta == q.asc && q.delta != 0.0) ? 1 : 0;
        s.est0 += q.estimate;
        s.delta0 += q.delta;
        s.abs0 += q.absvalue;
        s.asc0 += q.asc;
        ws_push(S, s, lo, hi, q.estimate, q.delta, lv);
    }
    double deltasum = 0.;
    for (int k = 0; k < s.size; ++k) {
        if (S.LV(k) > 0) S.D(k) = s.delta0;
        deltasum += S.D(k);
    }
    for (int k = 0; k < s.size; ++k) S.LV(k) = 0;
    ws_sort_results(S, s);

    const double tolerance = tol.to_abs(fabs(s.est0));
    const double round_off = 100. * F64_EPSILON * s.abs0;
    r.est = s.est0;
    r.dlt = s.delta0;
    r.nev = s.nevals;
    if (s.delta0 <= round_off && s.delta0 > tolerance) {
        r.st = ST_ROUNDOFF_ERROR;
        return r;
    } else if ((s.delta0 <= tolerance && s.delta0 != s.asc0) || s.delta0 == 0.0) {
        r.st = ST_NONE;
        return r;
    } else if ((unsigned long long)s.nevals == max_evals) {
        r.st = ST_INSUFFICIENT_ITERATION;
        return r;
    }

    tab_append(S, s, s.est0);
    s.area = s.est0;
    s.res_ext = s.est0;
    s.err_ext = F64_MAX;
    s.errsum = deltasum;
    s.eolr = deltasum;
    s.ertest = tolerance;
    s.max_iters = (unsigned long long)nint + (max_evals - s.nevals) / 50ull;
    s.iteration = (unsigned long long)nint;
    s.active = true;
    LocalSink sink;
    sink.est = 0.;
    sink.dlt = 0.;
    sink.nev = 0;
    sink.st = 0;
    while (s.active) {
        const int wi = s.wi;
        const double ia = S.A(wi), ib = S.B(wi), ie = S.E(wi), id = S.D(wi);
        const int ilv = S.LV(wi);
        const double center = (ia + ib) * 0.5;
        QK q1, q2;
        rule_pair_any<MODE, U>(g, ia, center, ib, q1, q2);
        adaptive_step<true>(S, s, tol, max_evals, sink, ia, ib, center, ie, id, ilv, q1, q2);
    }
    r.est = sink.est;
    r.dlt = sink.dlt;
    r.nev = sink.nev;
    r.st = sink.st;
    return r;
}

struct OuterCtx;  // kernels_2d_pool.cuh
struct Ctl2 {
    unsigned long long nact[2];  // active outer integrals of this round / the next
    unsigned long long ntasks;   // inner tasks posted this round
    unsigned long long qhead;    // queue head of k2_inner
    unsigned long long nexact;   // outer integrals handed to k2_exact this round
    unsigned long long xhead;    // queue head of k2_exact
    unsigned long long pad[2];
};

// Scratch of the level-synchronous path (handle-owned, sized by the batch).
struct Flat2 {
    Lane* olane;            // [n] outer QAGS / QAGP machines
    OuterCtx* octx;         // [n]
    unsigned int* act[2];   // [n] active outer slots, ping-pong
    unsigned int* xlist;    // [n] outer slots for k2_exact
    double* t_est;          // [50 n] task results ...
    uint32_t* t_nev;
    uint32_t* t_meta;       // status | RSV_* << 3 | n << 5
    uint32_t* t_owner;      // ... and what they are: outer slot,
    uint32_t* t_k;          // node index inside the trip
    Ctl2* ctl;
    SlabGeom oslab;         // outer interval lists: one column per outer slot (stride n)
};

struct Batch2D {
    long long n;
    const double* x0;
    const double* x1;
    long long range_stride;
    const double* fparams;
    long long fparam_stride;
    const double* yparams;
    long long yparam_stride;
    int yrange_id;
    int nyparams;
    const double* pts_xy;  // (x, y) pairs: shared by the batch (device copy) or CSR values
    int npts;              // shared count (when pts_offsets == nullptr)
    const long long* pts_offsets;  // per-integral CSR offsets (in pairs) or nullptr
    int npts_max;                  // largest per-integral count (slab geometry)
    const unsigned int* worklist;  // optional subset (AUTO split) or nullptr
    const unsigned long long* worklist_count;
    Tol tol;
    unsigned long long max_evals;
    Outputs out;
    Control* ctl;
    SlabGeom slab;
    int cap_outer, cap_inner;
    bool swap_xy;  // GKQ_FLAG_SWAP_XY
    // optional per-integral tolerance values / total budget (gkq_overrides); budgets are clamped
    // to max_evals, which sizes the interval slabs
    const double* abs_tol_each;
    const double* rel_tol_each;
    const unsigned long long* max_evals_each;
    int pool_r;  // k_nested_pool: outer integrals per warp (<= POOL_R)
    Flat2 flat;  // level-synchronous path (kernels_2d_flat.cuh)
};

GKQ_DEV Tol load_tol2(const Batch2D& P, long long i) {
    Tol t = P.tol;
    if (P.abs_tol_each) t.abs_tol = P.abs_tol_each[i];
    if (P.rel_tol_each) t.rel_tol = P.rel_tol_each[i];
    return t;
}
GKQ_DEV unsigned long long load_max_evals2(const Batch2D& P, long long i) {
    if (!P.max_evals_each) return P.max_evals;
    const unsigned long long m = P.max_evals_each[i];
    return m < P.max_evals ? m : P.max_evals;
}

// f(x, .) as a 1-D integrand of y
template <class F2T>
struct InnerF {
    F2T f;
    double x;
    bool swap;  // GKQ_FLAG_SWAP_XY: the integrand takes (inner, outer) -- DynamicX, qags.rs:37
    static constexpr bool PAIRWISE = false;
    static constexpr bool HEAVY = F2T::HEAVY;
    GKQ_DEV double operator()(double y) const { return swap ? f(y, x) : f(x, y); }
    template <class M> GKQ_DEV double eval(double y, M& m) const {
        return swap ? f.eval(y, x, m) : f.eval(x, y, m);
    }
};

// The outer integrand of double/algorithm/qags.rs:66-87 / qagp.rs:108-131 with the infinite
// range wrapper of the outer algorithm folded in (single/util.rs:74-103).
// ALG: 0 = QAGS2, 1 = QAGP2, 2 = QAG2 (the inner algorithm; `bool QAGP` arguments convert)
template <class F2T, int ALG>
struct OuterIntegrand {
    F2T f;
    YRange yr;
    Tol tol;
    unsigned long long total_evals;
    unsigned long long nevals;  // running sum of inner evaluations (usize in the reference)
    uint32_t error;             // first inner error
    unsigned long long inner_cap;
    Slab Sin;
    const double* pts_xy;
    int npts;
    bool transform;  // outer x-range infinite
    bool swap_xy;

    // One inner integral over y(x) with an explicit budget and workspace capacity; touches no
    // state of the outer integrand.
    __device__ __noinline__ Res run_inner(double x, unsigned long long budget,
                                          unsigned long long& cap) const {
        double y0, y1;
        yr(x, y0, y1);
        const bool ytr = !is_finite(y0) || !is_finite(y1);
        if (ytr) {
            y0 = transform_point(y0);
            y1 = transform_point(y1);
        }
        Wrapped<InnerF<F2T>> g;
        g.f.f = f;
        g.f.x = x;
        g.f.swap = swap_xy;
        g.transform = ytr;
        if (ALG == 1) {
            const double* p = pts_xy;
            auto user_pt = [p](int k) { return p[2 * k + 1]; };  // every y of the 2-D points
            return seq_qagp<RM_PURE, NESTED_UNROLL>(g, y0, y1, ytr, user_pt, npts, tol,
                                                                 budget, cap, Sin);
        }
        if (ALG == 2) return seq_qag<RM_PURE, NESTED_UNROLL>(g, y0, y1, tol, budget, cap, Sin);
        return seq_qags<RM_PURE, NESTED_UNROLL>(g, y0, y1, tol, budget, cap, Sin);
    }
    // config1.max_evals = if error.is_some() { 0 } else { config.max_evals - nevals }
    // (if an inner QAGP call overshot the budget by its last 50 samples the reference's
    // usize subtraction wraps and Vec::reserve panics; we hand out a zero budget instead)
    GKQ_DEV unsigned long long budget_now() const {
        return (error != ST_NONE || nevals > total_evals) ? 0ull : (total_evals - nevals);
    }

    GKQ_DEV double inner_integral(double x) {
        const Res r = run_inner(x, budget_now(), inner_cap);
        nevals += r.nev;
        if (r.st != ST_NONE) {
            if (error == ST_NONE) error = r.st;
            return __longlong_as_double(0x7ff8000000000000ll);  // f64::NAN
        }
        return r.est;
    }

    GKQ_DEV double operator()(double t) {
        if (transform) {
            const double coef = 1. / (1. - fabs(t));
            const double x2 = t * coef;
            return inner_integral(x2) * coef * coef;
        }
        return inner_integral(t);
    }
};

template <class F2T, int ALG, int NPARAMS>
__global__ void __launch_bounds__(BLOCK) k_nested(const Batch2D P) {
    constexpr bool QAGP = ALG == 1;
    const unsigned long long t = (unsigned long long)blockIdx.x * BLOCK + threadIdx.x;
    // slab regions of this thread: [outer lists | inner lists]
    // slab regions of this thread: [outer lists | inner lists]
    const Slab Sout = make_slab(P.slab.dbase, P.slab.wbase, P.slab.stride, t, P.cap_outer);
    const Slab Sin = make_slab(P.slab.dbase, P.slab.wbase, P.slab.stride, t, P.cap_inner,
                               slab_doubles(P.cap_outer, P.npts_max + 2), slab_ints(P.cap_outer));

    const unsigned long long total = P.worklist ? *P.worklist_count : (unsigned long long)P.n;
    unsigned long long* head = QAGP ? &P.ctl->headP : &P.ctl->head2;
    for (;;) {
        const unsigned long long slot = atomicAdd(head, 1ull);
        if (slot >= total) break;
        const unsigned long long i = P.worklist ? (unsigned long long)P.worklist[slot] : slot;
        const double* my_pts = P.pts_xy;
        int my_npts = P.npts;
        if (P.pts_offsets) {
            const long long o0 = P.pts_offsets[i];
            my_pts = P.pts_xy + 2 * o0;
            my_npts = (int)(P.pts_offsets[i + 1] - o0);
        }

        OuterIntegrand<F2T, ALG> G;
#pragma unroll
        for (int k = 0; k < MAX_PARAMS; ++k)
            G.f.p[k] = (k < NPARAMS)
                           ? (P.fparam_stride ? P.fparams[(long long)k * P.fparam_stride + (long long)i]
                                              : P.fparams[k])
                           : 0.0;
        G.yr.id = P.yrange_id;
        G.yr.q[0] = G.yr.q[1] = 0.;
        for (int k = 0; k < P.nyparams && k < 2; ++k)
            G.yr.q[k] = P.yparam_stride ? P.yparams[(long long)k * P.yparam_stride + (long long)i]
                                        : P.yparams[k];
        const Tol my_tol = load_tol2(P, (long long)i);
        const unsigned long long my_max_evals = load_max_evals2(P, (long long)i);
        G.tol = my_tol;
        G.total_evals = my_max_evals;
        G.nevals = 0;
        G.error = ST_NONE;
        G.inner_cap = 0;  // WorkSpace::new()
        G.Sin = Sin;
        G.pts_xy = my_pts;
        G.npts = my_npts;
        G.swap_xy = P.swap_xy;

        double xa = P.x0[(long long)i * P.range_stride];
        double xb = P.x1[(long long)i * P.range_stride];
        const bool xtr = !is_finite(xa) || !is_finite(xb);
        if (xtr) {
            xa = transform_point(xa);
            xb = transform_point(xb);
        }
        G.transform = xtr;

        unsigned long long outer_cap = 0;  // QAGS::new() / QAGP::new(): fresh outer workspace
        const unsigned long long outer_evals = my_max_evals / 17ull;
        Res r;
        if (QAGP) {
            // outer points = x of the 2-D points, transformed HERE when the x-range is infinite
            // (double/algorithm/qagp.rs:98-106) and once more inside make_sorted_points.
            const double* p = my_pts;
            auto user_pt = [p, xtr](int k) {
                const double x = p[2 * k];
                return xtr ? transform_point(x) : x;
            };
            r = seq_qagp<RM_ORDERED, 1>(G, xa, xb, xtr, user_pt, my_npts, my_tol, outer_evals, outer_cap, Sout);
        } else if (ALG == 2) {
            r = seq_qag<RM_ORDERED, 1>(G, xa, xb, my_tol, outer_evals, outer_cap, Sout);
        } else {
            r = seq_qags<RM_ORDERED, 1>(G, xa, xb, my_tol, outer_evals, outer_cap, Sout);
        }
        // result.value.nevals = nevals; if error.is_some() { result.error = error }
        write_result(P.out, (long long)i, r.est, r.dlt, (uint32_t)G.nevals,
                     G.error != ST_NONE ? G.error : r.st);
    }
}

// The nodes of one trip of an outer integral in the reference's evaluation order
// (single/qk/naive.rs:60-68: lower nodes j = 0..N-1, upper nodes, centre), for one rule or for
// the two rules of a bisection step.
struct NodeSet {
    int N, K;  // node pairs per rule, nodes in the trip (rules * (2N + 1))
    double lo[2], hi[2];
    GKQ_DEV double operator()(int k) const {
        const int per = 2 * N + 1;
        const int half = k >= per ? 1 : 0;
        const int t = k - per * half;
        const double center = 0.5 * (lo[half] + hi[half]);
        const double half_length = 0.5 * (hi[half] - lo[half]);
        if (t >= 2 * N) return center;
        const int j = t < N ? t : t - N;
        const double abscissa = half_length * (N == 8 ? c_xgk17[j] : c_xgk25[j]);
        return t < N ? center - abscissa : center + abscissa;
    }
};

}  // namespace gkq
#include "kernels_2d_pool.cuh"
#include "kernels_2d_flat.cuh"
namespace gkq {

struct Launch2D {
    void (*run[2])(const Batch2D&, int grid, cudaStream_t);  // [0] QAGS2, [1] QAGP2 (thread per outer integral)
    int (*occupancy[2])();
    void (*run_qag)(const Batch2D&, int grid, cudaStream_t);  // QAG2: k_nested<.., 2> (BLOCK threads)
    void (*run_pool[2])(const Batch2D&, int grid, cudaStream_t);  // k_nested_pool (PB threads per block)
    int (*occupancy_pool[2])();
    // level-synchronous path (kernels_2d_flat.cuh): [0] QAGS2, [1] QAGP2
    void (*flat_start[2])(const Batch2D&, int grid, cudaStream_t);
    void (*flat_inner[2])(const Batch2D&, int grid, cudaStream_t);
    void (*flat_exact[2])(const Batch2D&, int grid, cudaStream_t);
    int (*occupancy_flat[2])();
};
// the kernels of the level-synchronous path that do not depend on the integrand
void launch_flat_plan(bool qagp, const Batch2D& P, int round, int grid, cudaStream_t st);
void launch_flat_commit(bool qagp, const Batch2D& P, int round, int grid, cudaStream_t st);
template <int ID>
Launch2D make_launch2d();
#define GKQ_DECL_F2(ID, NAME, NP, FL) template <> Launch2D make_launch2d<ID>();
GKQ_F2_LIST(GKQ_DECL_F2)
#undef GKQ_DECL_F2

#define GKQ_INSTANTIATE_F2(ID, NAME, NP, FL)                                                      \
    template <> Launch2D make_launch2d<ID>() {                                                    \
        Launch2D L;                                                                               \
        L.run[0] = [](const Batch2D& P, int grid, cudaStream_t st) {                              \
            k_nested<F2<ID>, false, NP><<<grid, BLOCK, 0, st>>>(P);                               \
        };                                                                                        \
        L.run[1] = [](const Batch2D& P, int grid, cudaStream_t st) {                              \
            k_nested<F2<ID>, true, NP><<<grid, BLOCK, 0, st>>>(P);                                \
        };                                                                                        \
        L.run_qag = [](const Batch2D& P, int grid, cudaStream_t st) {                             \
            k_nested<F2<ID>, 2, NP><<<grid, BLOCK, 0, st>>>(P);                                   \
        };                                                                                        \
        L.occupancy[0] = []() {                                                                   \
            int nb = 0;                                                                           \
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_nested<F2<ID>, false, NP>, BLOCK, 0); \
            return nb;                                                                            \
        };                                                                                        \
        L.occupancy[1] = []() {                                                                   \
            int nb = 0;                                                                           \
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_nested<F2<ID>, true, NP>, BLOCK, 0); \
            return nb;                                                                            \
        };                                                                                        \
        L.run_pool[0] = [](const Batch2D& P, int grid, cudaStream_t st) {                         \
            const int smem = (PB / 32) * (int)sizeof(PoolShared);                                 \
            /* per device: set on every launch (an engine per GPU may live in one process) */     \
            cudaFuncSetAttribute(k_nested_pool<F2<ID>, false, NP>,                            \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, smem);          \
            k_nested_pool<F2<ID>, false, NP><<<grid, PB, smem, st>>>(P);                          \
        };                                                                                        \
        L.run_pool[1] = [](const Batch2D& P, int grid, cudaStream_t st) {                         \
            const int smem = (PB / 32) * (int)sizeof(PoolShared);                                 \
            /* per device: set on every launch (an engine per GPU may live in one process) */     \
            cudaFuncSetAttribute(k_nested_pool<F2<ID>, true, NP>,                             \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, smem);          \
            k_nested_pool<F2<ID>, true, NP><<<grid, PB, smem, st>>>(P);                           \
        };                                                                                        \
        L.occupancy_pool[0] = []() {                                                              \
            int nb = 0;                                                                           \
            const int smem = (PB / 32) * (int)sizeof(PoolShared);                                 \
            cudaFuncSetAttribute(k_nested_pool<F2<ID>, false, NP>,                                \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, smem);              \
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_nested_pool<F2<ID>, false, NP>, PB, smem); \
            return nb;                                                                            \
        };                                                                                        \
        L.occupancy_pool[1] = []() {                                                              \
            int nb = 0;                                                                           \
            const int smem = (PB / 32) * (int)sizeof(PoolShared);                                 \
            cudaFuncSetAttribute(k_nested_pool<F2<ID>, true, NP>,                                 \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, smem);              \
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_nested_pool<F2<ID>, true, NP>, PB, smem); \
            return nb;                                                                            \
        };                                                                                        \
        L.flat_start[0] = [](const Batch2D& P, int grid, cudaStream_t st) {                       \
            k2_start<F2<ID>, false, NP><<<grid, FB, 0, st>>>(P);                                  \
        };                                                                                        \
        L.flat_start[1] = [](const Batch2D& P, int grid, cudaStream_t st) {                       \
            k2_start<F2<ID>, true, NP><<<grid, FB, 0, st>>>(P);                                   \
        };                                                                                        \
        L.flat_inner[0] = [](const Batch2D& P, int grid, cudaStream_t st) {                       \
            k2_inner<F2<ID>, false, NP><<<grid, FB, 0, st>>>(P);                                  \
        };                                                                                        \
        L.flat_inner[1] = [](const Batch2D& P, int grid, cudaStream_t st) {                       \
            k2_inner<F2<ID>, true, NP><<<grid, FB, 0, st>>>(P);                                   \
        };                                                                                        \
        L.flat_exact[0] = [](const Batch2D& P, int grid, cudaStream_t st) {                       \
            k2_exact<F2<ID>, false, NP><<<grid, FB, 0, st>>>(P);                                  \
        };                                                                                        \
        L.flat_exact[1] = [](const Batch2D& P, int grid, cudaStream_t st) {                       \
            k2_exact<F2<ID>, true, NP><<<grid, FB, 0, st>>>(P);                                   \
        };                                                                                        \
        L.occupancy_flat[0] = []() {                                                              \
            int nb = 0;                                                                           \
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k2_inner<F2<ID>, false, NP>, FB, 0); \
            return nb;                                                                            \
        };                                                                                        \
        L.occupancy_flat[1] = []() {                                                              \
            int nb = 0;                                                                           \
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k2_inner<F2<ID>, true, NP>, FB, 0); \
            return nb;                                                                            \
        };                                                                                        \
        return L;                                                                                 \
    }

}  // namespace gkq
