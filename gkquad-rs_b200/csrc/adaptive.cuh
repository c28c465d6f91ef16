// adaptive.cuh -- the per-lane QAGS / QAGP state machine.
//
// B200 mapping: ONE THREAD OWNS ONE INTEGRAL and replays the
// reference's strictly sequential bisection order for it (max-error-first through the
// sorted `order` list, the nrmax side channel, the Wynn epsilon table), because status /
// nevals parity with gkquad requires that order.  Parallelism
// comes from the batch: a persistent grid of warps in which every lane runs its own state
// machine and *refills itself* from a global queue the moment its integral terminates, so
// the expensive part of each trip -- two 25-point rule evaluations = 50 integrand
// samples -- is always executed by (nearly) full warps no matter how different the
// integrals' iteration counts are.  This replaces the per-integral heap of the reference
// (single/workspace.rs) with structure-of-arrays slabs in HBM, interleaved by thread so
// that lanes touching the same list slot coalesce.
//
// Reference files restated here (paths relative to /root/reference/gkquad/src/):
//   single/workspace.rs:102-277   push / sort_results / update / qpsrt / increase_nrmax /
//                                 reset_nrmax / sum_results
//   single/qelg.rs:33-323         ExtrapolationTable::append / qelg
//   single/algorithm/qags.rs:98-311   main loop + epilogue
//   single/algorithm/qagp.rs:68-372   break-point seeding, main loop + epilogue
#pragma once
#include "gk_common.cuh"

namespace gkq {

// Geometry of the interleaved per-thread slabs: all arrays of one thread share one `double` slab and
// one `int` slab (layout: struct Slab below).
struct SlabGeom {
    double* dbase;
    int* wbase;
    unsigned long long stride;  // number of persistent threads
    int cap;                    // interval slots per thread
    int npts_max;               // break-point slots per thread (QAGP) incl. the two endpoints
};

// View of one thread's column of the slabs.  The four doubles of an interval (a, b, estimate, delta)
// form ONE 32-byte record -- a single address computation and one DRAM sector per interval instead
// of four of each; the records of a column are `stride` records apart (records of the same slot k
// of consecutive threads are adjacent, so lanes walking in step still coalesce).  Behind the cap
// records sit the thread's epsilon table, result history and break points as plain doubles, and in
// the int slab its level and order arrays.  Addresses are base + k * (byte stride): one 32 x 32 -> 64
// bit multiply-add each (round 1 spent a quarter of the persistent kernel's code on 64-bit index
// arithmetic: profiles/qagp_r02.md).
struct Slab {
    char* iv;  // interval records: record k at iv + k * sb32
    char* ex;  // extras: double j at ex + j * sb8   (EPS[52] | R3[3] | PTS[..])
    char* lw;  // ints: LV[k] at lw + k * sb4, ORD[k] at lw + (cap + k) * sb4
    unsigned int sb32, sb8, sb4;  // byte strides = stride * 32 / 8 / 4
    int cap;
    // -DGKQ_BOUNDS_CHECK: a debug build that traps on any list / table index outside its slot
    // range (compute-sanitizer is not available on every pool); run the GPU tests against it.
#ifdef GKQ_BOUNDS_CHECK
    GKQ_DEV void chk(int k, int n) const { if (k < 0 || k >= n) __trap(); }
#else
    GKQ_DEV void chk(int, int) const {}
#endif
    GKQ_DEV double* rec(int k) const { chk(k, cap); return reinterpret_cast<double*>(iv + (size_t)(unsigned)k * sb32); }
    GKQ_DEV double& A(int k) const { return rec(k)[0]; }
    GKQ_DEV double& B(int k) const { return rec(k)[1]; }
    GKQ_DEV double& E(int k) const { return rec(k)[2]; }
    GKQ_DEV double& D(int k) const { return rec(k)[3]; }
    GKQ_DEV double& X(int j) const { return *reinterpret_cast<double*>(ex + (size_t)(unsigned)j * sb8); }
    GKQ_DEV double& EPS(int k) const { chk(k, 52); return X(k); }
    GKQ_DEV double& R3(int k) const { chk(k, 3); return X(52 + k); }
    GKQ_DEV double& PTS(int k) const { chk(k, 1 << 20); return X(55 + k); }
    GKQ_DEV int& LV(int k) const { chk(k, cap); return *reinterpret_cast<int*>(lw + (size_t)(unsigned)k * sb4); }
    GKQ_DEV int& ORD(int k) const { chk(k, cap); return *reinterpret_cast<int*>(lw + (size_t)(unsigned)(cap + k) * sb4); }
};
// Column `t` of a slab pair whose first double / int column group starts at dcol / wcol (in units of
// `stride` elements: the nested kernels keep an outer and an inner list per thread, one behind the
// other).  stride < 2^27 (checked by the host: byte strides are 32-bit).
GKQ_DEV Slab make_slab(double* dbase, int* wbase, unsigned long long stride, unsigned long long t, int cap,
                       unsigned long long dcol = 0, unsigned long long wcol = 0) {
    Slab S;
    double* d0 = dbase + dcol * stride;
    S.iv = reinterpret_cast<char*>(d0 + 4ull * t);
    S.ex = reinterpret_cast<char*>(d0 + 4ull * (unsigned long long)cap * stride + t);
    S.lw = reinterpret_cast<char*>(wbase + wcol * stride + t);
    S.sb32 = (unsigned int)(stride * 32ull);
    S.sb8 = (unsigned int)(stride * 8ull);
    S.sb4 = (unsigned int)(stride * 4ull);
    S.cap = cap;
    return S;
}
__host__ __device__ inline unsigned long long slab_doubles(int cap, int npts_max) {
    return 4ull * cap + 55ull + (unsigned long long)npts_max;
}
__host__ __device__ inline unsigned long long slab_ints(int cap) { return 2ull * cap; }

// Vec::reserve on a cleared vector of capacity c0 (alloc::raw_vec grow_amortized):
// the capacity afterwards is the `limit` of qpsrt / increase_nrmax (workspace.rs:165,236).
__host__ __device__ inline unsigned long long reserve_cap(unsigned long long c0,
                                                          unsigned long long need) {
    if (c0 >= need) return c0;
    unsigned long long nc = c0 * 2 > need ? c0 * 2 : need;
    return nc < 4 ? 4 : nc;
}

// Scalars of one in-flight integral (registers / local memory of the owning lane).
struct Lane {
    long long idx;
    bool active;
    bool transform;
    // WorkSpace scalars (workspace.rs:32-48)
    int size, nrmax, wi, maxlevel, limit;
    // ExtrapolationTable scalars (qelg.rs:12-25)
    int tab_n, tab_nres;
    // integrate_impl locals (qags.rs:74-116 / qagp.rs:75-179)
    double area, errsum, res_ext, err_ext, ertest, eolr, correc;
    double est0, delta0, abs0, asc0;  // result0
    int ktmin, ro1, ro2, ro3;
    uint32_t error;
    bool error2, extrapolate, disallow;
    unsigned long long iteration, max_iters;
    uint32_t nevals;
    // QAGP seeding phase
    bool seeding;
    int nint, seed_k, w0, w1;
    int pad_[2];
};
// k_adaptive keeps one Lane per thread in shared memory: a size of 8 x (odd) bytes makes the
// 64-bit fields of consecutive lanes fall into distinct bank pairs.
static_assert(sizeof(Lane) % 8 == 0 && (sizeof(Lane) / 8) % 2 == 1, "Lane must be 8 x odd bytes");

// ---- WorkSpace on a slab --------------------------------------------------------------
GKQ_DEV void ws_push(const Slab& S, Lane& s, double a, double b, double e, double d, int lv) {
    int k = s.size;  // workspace.rs:102-105
    S.A(k) = a;
    S.B(k) = b;
    S.E(k) = e;
    S.D(k) = d;
    S.LV(k) = lv;
    S.ORD(k) = k;
    s.size = k + 1;
}

// workspace.rs:107-138 (order[] indexed by value in the swap, restated literally)
static __device__ __noinline__ void ws_sort_results(const Slab& S, Lane& s) {
    int nint = s.size;
    if (nint == 0) return;
    for (int i = 0; i < nint; ++i) {
        int i1 = S.ORD(i);
        double e1 = S.D(i1);
        int i_max = i1;
        for (int k = i + 1; k < nint; ++k) {
            int i2 = S.ORD(k);
            double e2 = S.D(i2);
            if (e2 >= e1) {
                i_max = i2;
                e1 = e2;
            }
        }
        if (i_max != i1) {
            S.ORD(i) = S.ORD(i_max);
            S.ORD(i_max) = i1;
        }
    }
    s.wi = S.ORD(0);
}

// workspace.rs:163-223
static __device__ __noinline__ void ws_qpsrt(const Slab& S, Lane& s) {
    const int last = s.size - 1;
    const int limit = s.limit;
    int i_nrmax = s.nrmax;
    int i_maxdelta = S.ORD(i_nrmax);

    if (last < 2) {
        S.ORD(0) = 0;
        S.ORD(1) = 1;
        s.wi = i_maxdelta;
        return;
    }

    const double deltamax = S.D(i_maxdelta);
    while (i_nrmax > 0) {
        int prev = S.ORD(i_nrmax - 1);
        if (!(deltamax > S.D(prev))) break;
        S.ORD(i_nrmax) = prev;
        i_nrmax -= 1;
    }

    const int top = (last < (limit / 2 + 2)) ? last : limit - last + 1;

    int i = i_nrmax + 1;
    while (i < top) {
        int cur = S.ORD(i);
        if (!(deltamax < S.D(cur))) break;
        S.ORD(i - 1) = cur;
        i += 1;
    }
    S.ORD(i - 1) = i_maxdelta;

    const double errmin = S.D(last);
    int k = top - 1;
    while (k >= i - 1) {
        int cur = S.ORD(k);
        if (!(errmin >= S.D(cur))) break;
        S.ORD(k + 1) = cur;
        k -= 1;
    }
    S.ORD(k + 1) = last;

    i_maxdelta = S.ORD(i_nrmax);
    s.wi = i_maxdelta;
    s.nrmax = i_nrmax;
}

// workspace.rs:141-160.  Both halves carry level `lv`.
GKQ_DEV void ws_update(const Slab& S, Lane& s, double a1, double b1, double e1, double d1,
                       double a2, double b2, double e2, double d2, int lv) {
    const int wi = s.wi;
    const int k = s.size;
    if (d2 > d1) {
        S.A(wi) = a2; S.B(wi) = b2; S.E(wi) = e2; S.D(wi) = d2;
        S.A(k) = a1;  S.B(k) = b1;  S.E(k) = e1;  S.D(k) = d1;
    } else {
        S.A(wi) = a1; S.B(wi) = b1; S.E(wi) = e1; S.D(wi) = d1;
        S.A(k) = a2;  S.B(k) = b2;  S.E(k) = e2;  S.D(k) = d2;
    }
    S.LV(wi) = lv;
    S.LV(k) = lv;
    S.ORD(k) = k;
    s.size = k + 1;
    if (lv > s.maxlevel) s.maxlevel = lv;
    ws_qpsrt(S, s);
}

// workspace.rs:233-259
static __device__ __noinline__ bool ws_increase_nrmax(const Slab& S, Lane& s) {
    const int id = s.nrmax;
    const int limit = s.limit;
    const int last = s.size - 1;
    const int jupbnd = (last > (1 + limit / 2)) ? limit + 1 - last : last;
    for (int k = id; k <= jupbnd; ++k) {
        int i_max = S.ORD(s.nrmax);
        s.wi = i_max;
        if (S.LV(i_max) < s.maxlevel) return true;
        s.nrmax += 1;
    }
    return false;
}
// workspace.rs:262-265
GKQ_DEV void ws_reset_nrmax(const Slab& S, Lane& s) {
    s.nrmax = 0;
    s.wi = S.ORD(0);
}
// workspace.rs:275-277
GKQ_DEV double ws_sum_results(const Slab& S, const Lane& s) {
    double acc = 0.0;
    for (int k = 0; k < s.size; ++k) acc += S.E(k);
    return acc;
}

// ---- Wynn epsilon table on a slab (qelg.rs:33-37, 50-323) --------------------------------
GKQ_DEV void tab_append(const Slab& S, Lane& s, double y) {
    S.EPS(s.tab_n) = y;
    s.tab_n += 1;
}

static __device__ __noinline__ void tab_qelg(const Slab& S, Lane& s, double& result, double& abserr) {
    const int n = s.tab_n - 1;
    const double current = S.EPS(n);
    const int newelm = n / 2;
    const int n_orig = n;
    int n_final = n;
    const int nres_orig = s.tab_nres;

    result = current;
    abserr = F64_MAX;
    if (n < 2) return;

    S.EPS(n + 2) = S.EPS(n);
    S.EPS(n) = F64_MAX;

    for (int i = 0; i < newelm; ++i) {
        const int ep = n - 2 * i;
        double res = S.EPS(ep + 2);
        const double e0 = S.EPS(ep - 2);
        const double e1 = S.EPS(ep - 1);
        const double e2 = res;

        const double elabs = fabs(e1);
        const double delta2 = e2 - e1;
        const double err2 = fabs(delta2);
        const double tol2 = fmax(fabs(e2), elabs) * F64_EPSILON;
        const double delta3 = e1 - e0;
        const double err3 = fabs(delta3);
        const double tol3 = fmax(elabs, fabs(e0)) * F64_EPSILON;

        if (err2 <= tol2 && err3 <= tol3) {
            // e0, e1 and e2 agree to machine accuracy: convergence (table left untouched)
            result = res;
            const double absolute = err2 + err3;
            const double relative = 5. * F64_EPSILON * fabs(res);
            abserr = fmax(absolute, relative);
            return;
        }

        const double e3 = S.EPS(ep);
        S.EPS(ep) = e1;
        const double delta1 = e1 - e3;
        const double err1 = fabs(delta1);
        const double tol1 = fmax(elabs, fabs(e3)) * F64_EPSILON;

        if (err1 <= tol1 || err2 <= tol2 || err3 <= tol3) {
            n_final = 2 * i;
            break;
        }

        const double ss = (1. / delta1 + 1. / delta2) - 1. / delta3;
        if (fabs(ss * e1) <= 0.0001) {
            n_final = 2 * i;
            break;
        }

        res = e1 + 1. / ss;
        S.EPS(ep) = res;

        const double error = err2 + fabs(res - e2) + err3;
        if (error <= abserr) {
            abserr = error;
            result = res;
        }
    }

    const int LIMEXP = 50 - 1;
    if (n_final == LIMEXP) n_final = 2 * (LIMEXP / 2);

    {
        const int start = (n_orig & 1) ? 1 : 0;
        for (int k = 0; k <= newelm; ++k) S.EPS(start + 2 * k) = S.EPS(start + 2 * k + 2);
    }
    if (n_orig != n_final) {
        const int shift = n_orig - n_final;
        for (int k = 0; k <= n_final; ++k) S.EPS(k) = S.EPS(shift + k);
    }
    s.tab_n = n_final + 1;

    if (nres_orig < 3) {
        S.R3(nres_orig) = result;
        abserr = F64_MAX;
    } else {
        const double r0 = S.R3(0), r1 = S.R3(1), r2 = S.R3(2);
        abserr = fabs(result - r2) + fabs(result - r1) + fabs(result - r0);
        S.R3(0) = r1;
        S.R3(1) = r2;
        S.R3(2) = result;
    }
    s.tab_nres = nres_orig + 1;
    abserr = fmax(abserr, 5. * F64_EPSILON * fabs(result));
}

// ---- result sink -------------------------------------------------------------------------
struct Outputs {
    double* estimate;
    double* delta;
    uint32_t* nevals;
    uint32_t* status;
};
GKQ_DEV void write_result(const Outputs& O, long long i, double est, double dlt, uint32_t nev,
                          uint32_t st) {
    O.estimate[i] = est;
    O.delta[i] = dlt;
    O.nevals[i] = nev;
    O.status[i] = st;
}
// Sink concept: put(i, estimate, delta, nevals, status).  GlobalSink writes the batch
// outputs; LocalSink keeps the result in registers (inner integrals of the nested 2-D driver).
struct GlobalSink {
    Outputs O;
    GKQ_DEV void put(long long i, double est, double dlt, uint32_t nev, uint32_t st) const {
        write_result(O, i, est, dlt, nev, st);
    }
};
struct LocalSink {
    double est, dlt;
    uint32_t nev, st;
    GKQ_DEV void put(long long, double e, double d, uint32_t n, uint32_t s) {
        est = e;
        dlt = d;
        nev = n;
        st = s;
    }
};

// Epilogue shared by QAGS (qags.rs:274-310) and QAGP (qagp.rs:338-371): choose between the
// summed and the extrapolated result, then the divergence test.
template <class Sink>
__device__ __noinline__ void adaptive_epilogue(const Slab& S, Lane& s, Sink& O) {
    s.active = false;
    uint32_t error = s.error;
    double err_ext = s.err_ext;
    const double res_ext = s.res_ext, area = s.area, errsum = s.errsum;

    if (err_ext == F64_MAX) {
        O.put(s.idx, ws_sum_results(S, s), errsum, s.nevals, error);
        return;
    }
    if (error != ST_NONE || s.error2) {
        if (s.error2) err_ext += s.correc;
        if (error == ST_NONE) error = ST_ROUNDOFF_ERROR;
        if (res_ext != 0.0 && area != 0.0) {
            if (err_ext / fabs(res_ext) > errsum / fabs(area)) {
                O.put(s.idx, ws_sum_results(S, s), errsum, s.nevals, error);
                return;
            }
        } else if (err_ext > errsum) {
            O.put(s.idx, ws_sum_results(S, s), errsum, s.nevals, error);
            return;
        } else if (area == 0.0) {
            O.put(s.idx, res_ext, err_ext, s.nevals, error);
            return;
        }
    }
    const bool positive_integrand = test_positivity(s.est0, s.abs0);
    if (!positive_integrand && fmax(fabs(res_ext), fabs(area)) < 0.01 * s.abs0) {
        O.put(s.idx, res_ext, err_ext, s.nevals, error);
        return;
    }
    const double ratio = res_ext / area;
    if ((ratio < 0.01 || ratio > 100.0 || errsum > fabs(area)) && error == ST_NONE)
        error = ST_DIVERGENT;
    O.put(s.idx, res_ext, err_ext, s.nevals, error);
}

// One trip of the main loop AFTER the two rule evaluations (qags.rs:132-271 /
// qagp.rs:191-335).  (a0,b0,e0,d0,lv0) is the interval that was bisected at `center`;
// q1 / q2 are the rule results on its halves.  Leaves s.active == false when the integral
// has terminated (result written).
template <bool QAGP, class Sink>
GKQ_DEV void adaptive_step(const Slab& S, Lane& s, const Tol& tol, unsigned long long max_evals,
                           Sink& O, double a0, double b0, double center, double e0,
                           double d0, int lv0, const QK& q1, const QK& q2) {
    const int current_level = lv0 + 1;
    s.nevals += 50;

    if (q1.estimate != q1.estimate || q2.estimate != q2.estimate) {
        s.error = ST_NAN;
        adaptive_epilogue(S, s, O);
        return;
    }

    const double area12 = q1.estimate + q2.estimate;
    const double error12 = q1.delta + q2.delta;
    const double last_e_i = d0;

    s.errsum += error12 - d0;
    s.area += area12 - e0;
    const double tolerance = tol.to_abs(fabs(s.area));

    if (q1.asc != q1.delta && q2.asc != q2.delta) {
        if (fabs(e0 - area12) <= 1e-5 * fabs(area12) && error12 >= 0.99 * d0) {
            if (!s.extrapolate)
                s.ro1 += 1;
            else
                s.ro2 += 1;
        }
        if (s.iteration > 10 && error12 > d0) s.ro3 += 1;
    }
    if (s.ro1 + s.ro2 >= 10 || s.ro3 >= 20) s.error = ST_ROUNDOFF_ERROR;
    if (s.ro2 >= 5) s.error2 = true;
    if (subrange_too_small(a0, center, b0)) s.error = ST_SUBRANGE_TOO_SMALL;

    if (s.errsum <= tolerance) {
        const double total = ws_sum_results(S, s) - e0 + q1.estimate + q2.estimate;
        s.active = false;
        O.put(s.idx, total, s.errsum, s.nevals, s.error);
        return;
    }

    ws_update(S, s, a0, center, q1.estimate, q1.delta, center, b0, q2.estimate, q2.delta,
              current_level);

    if (s.error != ST_NONE) {
        adaptive_epilogue(S, s, O);
        return;
    }

    if (QAGP) {
        if (s.iteration >= s.max_iters - 1) {  // qagp.rs:259-262
            s.error = ST_INSUFFICIENT_ITERATION;
            adaptive_epilogue(S, s, O);
            return;
        }
    } else {
        if (s.iteration >= max_evals - 1) {  // qags.rs:200-203 (compares with evaluations: sic)
            s.error = ST_INSUFFICIENT_ITERATION;
            adaptive_epilogue(S, s, O);
            return;
        }
    }

    // `continue` in the reference == fall out of this do/while(0) to the loop increment
    do {
        if (!QAGP && s.iteration == 2) {  // qags.rs:206-211
            s.eolr = s.errsum;
            s.ertest = tolerance;
            tab_append(S, s, s.area);
            break;
        }
        if (s.disallow) break;

        s.eolr -= last_e_i;
        if (current_level < s.maxlevel) s.eolr += error12;

        if (!s.extrapolate) {
            if (S.LV(s.wi) < s.maxlevel) break;
            s.extrapolate = true;
            s.nrmax = 1;
        }

        if (QAGP) {
            if (!s.error2 && s.eolr > s.ertest && ws_increase_nrmax(S, s)) break;
        } else {
            if (s.eolr > s.ertest && ws_increase_nrmax(S, s)) break;
        }

        tab_append(S, s, s.area);
        if (QAGP && s.tab_n < 3) {  // qagp.rs:295-300
            ws_reset_nrmax(S, s);
            s.extrapolate = false;
            s.eolr = s.errsum;
            break;
        }

        double reseps = 0., abseps = 0.;
        tab_qelg(S, s, reseps, abseps);

        s.ktmin += 1;
        if (s.ktmin > 5 && s.err_ext < 0.001 * s.errsum) s.error = ST_ROUNDOFF_ERROR;

        if (abseps < s.err_ext) {
            s.ktmin = 0;
            s.err_ext = abseps;
            s.res_ext = reseps;
            s.correc = s.eolr;
            s.ertest = tol.to_abs(fabs(reseps));
            if (s.err_ext <= s.ertest) {
                adaptive_epilogue(S, s, O);
                return;
            }
        }

        if (s.tab_n == 1) s.disallow = true;

        if (s.error != ST_NONE) {
            adaptive_epilogue(S, s, O);
            return;
        }

        ws_reset_nrmax(S, s);
        s.extrapolate = false;
        s.eolr = s.errsum;
    } while (0);

    s.iteration += 1;
    if (s.iteration > s.max_iters) adaptive_epilogue(S, s, O);
}

GKQ_DEV void lane_reset_common(Lane& s) {
    s.size = 0;
    s.nrmax = 0;
    s.wi = 0;
    s.maxlevel = 0;
    s.tab_n = 0;
    s.tab_nres = 0;
    s.correc = 0.;
    s.ktmin = 0;
    s.ro1 = s.ro2 = s.ro3 = 0;
    s.error = ST_NONE;
    s.error2 = false;
    s.extrapolate = false;
    s.disallow = false;
    s.seeding = false;
}

}  // namespace gkq
