// gk_common.cuh -- constants, Gauss-Kronrod tables and small device helpers.
//
// Arithmetic contract: this library is compiled with -fmad=false so that every a*b+c in
// the rule bookkeeping, the interval list, the epsilon table and the polynomial/rational
// functors rounds exactly like the reference (Rust never contracts to FMA).  Each thread
// walks the sums in the reference's own order, so for IEEE-exact integrands (+,-,*,/,sqrt)
// the engine is bit-identical to gkquad on the CPU.  libdevice's transcendental functions
// keep their internal FMAs (they are explicit fma intrinsics, untouched by -fmad).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace gkq {

#define GKQ_DEV __device__ __forceinline__

}  // namespace gkq
#include "gk_math.cuh"
namespace gkq {

constexpr double F64_EPSILON = 2.220446049250313e-16;        // core::f64::EPSILON
constexpr double F64_MIN_POSITIVE = 2.2250738585072014e-308; // core::f64::MIN_POSITIVE
constexpr double F64_MAX = 1.7976931348623157e308;           // core::f64::MAX

enum : uint32_t {
    ST_NONE = 0,
    ST_INSUFFICIENT_ITERATION = 1,
    ST_ROUNDOFF_ERROR = 2,
    ST_SUBRANGE_TOO_SMALL = 3,
    ST_DIVERGENT = 4,
    ST_NAN = 5,
};
enum : int { ALGO_AUTO = 0, ALGO_QAGS = 1, ALGO_QAGP = 2 };

// Node tables of the Kronrod rules (values: single/qk/mod.rs:78-412; gk_tables.inc is generated
// from first principles by tools/gen_gk_tables.py and checked bit-identical to those literals).
// All lanes of a warp read the same entry at the same time (thread-per-integral mapping),
// so constant memory broadcasts them at register speed.  The adaptive drivers use the 17- and
// 25-point rules only (qags.rs:315-376); 33 ... 57 are the public rule-level entries
// (single/qk/mod.rs:44-73) behind gkq_qk_batch.
#include "gk_tables.inc"
template <int N> struct Rule;
static __constant__ double c_xgk17[8] = {GK_XGK17};
static __constant__ double c_wg17[4] = {GK_WG17};
static __constant__ double c_wgk17[8] = {GK_WGK17};
static __constant__ double c_xgk25[12] = {GK_XGK25};
static __constant__ double c_wg25[6] = {GK_WG25};
static __constant__ double c_wgk25[12] = {GK_WGK25};
static __constant__ double c_xgk33[16] = {GK_XGK33};
static __constant__ double c_wg33[8] = {GK_WG33};
static __constant__ double c_wgk33[16] = {GK_WGK33};
static __constant__ double c_xgk41[20] = {GK_XGK41};
static __constant__ double c_wg41[10] = {GK_WG41};
static __constant__ double c_wgk41[20] = {GK_WGK41};
static __constant__ double c_xgk49[24] = {GK_XGK49};
static __constant__ double c_wg49[12] = {GK_WG49};
static __constant__ double c_wgk49[24] = {GK_WGK49};
static __constant__ double c_xgk57[28] = {GK_XGK57};
static __constant__ double c_wg57[14] = {GK_WG57};
static __constant__ double c_wgk57[28] = {GK_WGK57};

#define GKQ_RULE(N_, NPTS_)                                                     \
    template <> struct Rule<N_> {                                               \
        static constexpr int NPTS = NPTS_;                                      \
        static constexpr double WCK = GK_WCK##NPTS_;                            \
        GKQ_DEV static double xgk(int j) { return c_xgk##NPTS_[j]; }            \
        GKQ_DEV static double wgk(int j) { return c_wgk##NPTS_[j]; }            \
        GKQ_DEV static double wg(int j) { return c_wg##NPTS_[j]; }              \
    };
GKQ_RULE(8, 17)
GKQ_RULE(12, 25)
GKQ_RULE(16, 33)
GKQ_RULE(20, 41)
GKQ_RULE(24, 49)
GKQ_RULE(28, 57)
#undef GKQ_RULE

// Tolerance::to_abs (common.rs:34-41).  fmax/fmin return the non-NaN operand, which is
// what Rust's f64::max / f64::min do.
struct Tol {
    int kind;
    double abs_tol;
    double rel_tol;
    GKQ_DEV double to_abs(double value) const {
        switch (kind) {
            case 0: return abs_tol;
            case 1: return value * rel_tol;
            case 2: return fmax(abs_tol, value * rel_tol);
            default: return fmin(abs_tol, value * rel_tol);
        }
    }
};

struct QK {
    double estimate, delta, absvalue, asc;
};

// single/util.rs:118-141
GKQ_DEV double rescale_error(double err, double result_abs, double result_asc) {
    err = fabs(err);
    if (result_asc != 0.0 && err != 0.0) {
        double scale = 200.0 * err;
        if (scale < result_asc) {
            scale /= result_asc;
            err = result_asc * scale * sqrt(scale);
        } else {
            err = result_asc;
        }
    }
    if (result_abs > F64_MIN_POSITIVE / (50.0 * F64_EPSILON)) {
        double min_err = 50.0 * F64_EPSILON * result_abs;
        if (min_err > err) err = min_err;
    }
    return err;
}

// single/util.rs:162-170
GKQ_DEV double transform_point(double x) {
    if (x == -INFINITY) return -1.0;
    if (x == INFINITY) return 1.0;
    return x / (1.0 + fabs(x));
}
GKQ_DEV bool is_finite(double x) { return fabs(x) <= F64_MAX; }

// single/util.rs:110-115
GKQ_DEV bool subrange_too_small(double a1, double a2, double b2) {
    double tmp = (1. + 100. * F64_EPSILON) * (fabs(a2) + 1000. * F64_MIN_POSITIVE);
    return fabs(a1) <= tmp && fabs(b2) <= tmp;
}
// single/util.rs:146-148
GKQ_DEV bool test_positivity(double result, double resabs) {
    return fabs(result) >= (1.0 - 50.0 * F64_EPSILON) * resabs;
}

// Integrand adaptor = IntegrandWrapper (single/util.rs:74-103): point-wise infinite-range
// map t -> f(t/(1-|t|)) / (1-|t|)^2, selected per integral at run time (a batch may mix
// finite and infinite ranges; a warp-uniform flag costs nothing).
template <class F>
struct Wrapped {
    F f;
    bool transform;
    static constexpr bool HEAVY = F::HEAVY;
    // Branch-free forms for a caller that keeps ONE FastMath for a whole rule (qk_rule_light): no
    // re-evaluation here, the caller redoes the rule when m.ok went false.
    template <bool TR>
    GKQ_DEV void pair_fast(double x1, double x2, double& f1, double& f2, FastMath& m) const {
        double c1 = 1., c2 = 1.;
        if (TR) {
            c1 = m.div(1., 1. - fabs(x1));
            c2 = m.div(1., 1. - fabs(x2));
            x1 = x1 * c1;
            x2 = x2 * c2;
        }
        f1 = f.eval(x1, m);
        f2 = f.eval(x2, m);
        if (TR) {
            f1 = f1 * c1 * c1;
            f2 = f2 * c2 * c2;
        }
    }
    template <bool TR>
    GKQ_DEV double at_fast(double x, FastMath& m) const {
        if (TR) {
            const double coef = m.div(1., 1. - fabs(x));
            const double x2 = x * coef;
            return f.eval(x2, m) * coef * coef;
        }
        return f.eval(x, m);
    }
    // TR = whether the infinite-range map is on, as a compile-time constant (kernels that hoist
    // the test out of the node loop); the run-time forms below dispatch on `transform`.
    template <bool TR>
    GKQ_DEV double at(double x) const {
        if (TR) {
            double coef = 1. / (1. - fabs(x));
            double x2 = x * coef;
            return f(x2) * coef * coef;
        }
        return f(x);
    }
    GKQ_DEV double operator()(double x) const { return transform ? at<true>(x) : at<false>(x); }
    // Two samples at once.  The integrand bodies run branch-free (FastMath) in ONE basic block so
    // that ptxas interleaves their dependent chains (DFMA latency is ~9 cycles, profiles/); the
    // rare out-of-range argument re-evaluates both samples through the always-valid path.
    template <bool TR>
    GKQ_DEV void pair_t(double x1, double x2, double& f1, double& f2) const {
        double c1 = 1., c2 = 1.;
        if (TR) {
            c1 = 1. / (1. - fabs(x1));
            c2 = 1. / (1. - fabs(x2));
            x1 = x1 * c1;
            x2 = x2 * c2;
        }
        PairMath m;
        if constexpr (F::PAIRWISE) {
            const D2 r = f.eval(D2{x1, x2}, m);  // both samples through the same statements (ILP 2)
            f1 = r.a;
            f2 = r.b;
        } else {
            f1 = f.eval(x1, m);
            f2 = f.eval(x2, m);
        }
        if (!m.ok) {
            f1 = f(x1);
            f2 = f(x2);
        }
        if (TR) {
            f1 = f1 * c1 * c1;
            f2 = f2 * c2 * c2;
        }
    }
    GKQ_DEV void pair(double x1, double x2, double& f1, double& f2) const {
        if (transform)
            pair_t<true>(x1, x2, f1, f2);
        else
            pair_t<false>(x1, x2, f1, f2);
    }
};

// Second half of the rule (single/qk/naive.rs:70-105): Kronrod / Gauss / |f| sums and the asc
// pass over already evaluated samples, in the reference's order.
template <int N>
GKQ_DEV QK qk_reduce(const double (&fv1)[N], const double (&fv2)[N], double f_center, double a,
                     double b) {
    using R = Rule<N>;
    const double half_length = 0.5 * (b - a);
    const double abs_half_length = fabs(half_length);
    double result_gauss = 0.;
    double result_kronrod = f_center * R::WCK;
    double result_abs = fabs(result_kronrod);
#pragma unroll
    for (int j = 0; j < N; ++j) {
        double fsum = fv1[j] + fv2[j];
        result_kronrod += R::wgk(j) * fsum;
        result_abs += R::wgk(j) * (fabs(fv1[j]) + fabs(fv2[j]));
        if ((j & 1) == 0) result_gauss += R::wg(j >> 1) * fsum;
    }
    const double mean = result_kronrod * 0.5;
    double result_asc = R::WCK * fabs(f_center - mean);
#pragma unroll
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

// qk_reduce over samples kept in SHARED memory, one column per thread: sample j of the lower / upper
// half at fv[j * stride] / fv[(N + j) * stride] (stride = threads per block, conflict-free).  Same
// operations in the same order as qk_reduce.
template <int N>
GKQ_DEV QK qk_reduce_smem(const double* fv, int stride, double f_center, double a, double b) {
    using R = Rule<N>;
    const double half_length = 0.5 * (b - a);
    const double abs_half_length = fabs(half_length);
    double result_gauss = 0.;
    double result_kronrod = f_center * R::WCK;
    double result_abs = fabs(result_kronrod);
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const double f1 = fv[j * stride], f2 = fv[(N + j) * stride];
        double fsum = f1 + f2;
        result_kronrod += R::wgk(j) * fsum;
        result_abs += R::wgk(j) * (fabs(f1) + fabs(f2));
        if ((j & 1) == 0) result_gauss += R::wg(j >> 1) * fsum;
    }
    const double mean = result_kronrod * 0.5;
    double result_asc = R::WCK * fabs(f_center - mean);
#pragma unroll
    for (int j = 0; j < N; ++j)
        result_asc += R::wgk(j) * (fabs(fv[j * stride] - mean) + fabs(fv[(N + j) * stride] - mean));

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
// The rule for a transcendental body in the flat kernels: one node pair per loop trip, the 2N
// samples parked in the thread's shared-memory column instead of local memory (local memory
// competes for L1 with itself -- 176 B x 768 threads per SM -- and its misses and write-backs go
// to L2: ncu showed 10 M L2 sectors of local traffic per launch of the qk17 pass).
template <int N, bool TR, class G>
GKQ_DEV QK qk_rule_smem_fixed(const G& g, double a, double b, double* fv, int stride) {
    using R = Rule<N>;
    const double center = 0.5 * (a + b);
    const double half_length = 0.5 * (b - a);
#pragma unroll 1
    for (int j = 0; j < N; ++j) {
        double abscissa = half_length * R::xgk(j);
        double f1, f2;
        g.template pair_t<TR>(center - abscissa, center + abscissa, f1, f2);
        fv[j * stride] = f1;
        fv[(N + j) * stride] = f2;
    }
    const double f_center = g.template at<TR>(center);
    return qk_reduce_smem<N>(fv, stride, f_center, a, b);
}
template <int N, class G>
GKQ_DEV QK qk_rule_smem(const G& g, double a, double b, double* fv, int stride) {
    if (g.transform) return qk_rule_smem_fixed<N, true>(g, a, b, fv, stride);
    return qk_rule_smem_fixed<N, false>(g, a, b, fv, stride);
}

// The rule for a body WITHOUT transcendental calls, fully unrolled: all 2N+1 samples run through the
// branch-free math (m.div / m.sqrt, gk_math.cuh) in ONE basic block, so ptxas overlaps the dependent
// chains of as many samples as the registers hold; if any operand left the fast range (denormal-range
// quotients, huge divisors, negative or tiny radicands) the whole rule is redone by the out-of-line
// always-valid version.  Same operations, same order, same results.  (In between -- 2 or 4 node pairs per
// loop trip with one flag per group, for the 2-D nest whose instruction footprint rules the full unrolling
// out -- measured SLOWER than the pair-at-a-time loop with the compiler's division: D0 4.33 -> 5.19 / 5.48 ms,
// D1 73.0 -> 81.2 / 90.5 ms, profiles/ab_branch_free_div_r02.txt; the fully unrolled rule in the nest's inner
// kernel: D1 96 ms.  The nest keeps one pair per trip.)
template <int N, bool TR, class G>
__device__ __noinline__ QK qk_rule_slow(const G& g, double a, double b) {
    using R = Rule<N>;
    double fv1[N], fv2[N];
    const double center = 0.5 * (a + b);
    const double half_length = 0.5 * (b - a);
#pragma unroll 1
    for (int j = 0; j < N; ++j) {
        double abscissa = half_length * R::xgk(j);
        fv1[j] = g.template at<TR>(center - abscissa);
        fv2[j] = g.template at<TR>(center + abscissa);
    }
    const double f_center = g.template at<TR>(center);
    return qk_reduce<N>(fv1, fv2, f_center, a, b);
}
template <int N, bool TR, class G>
GKQ_DEV QK qk_rule_light(const G& g, double a, double b) {
    using R = Rule<N>;
    double fv1[N], fv2[N];
    const double center = 0.5 * (a + b);
    const double half_length = 0.5 * (b - a);
    FastMath m;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        double abscissa = half_length * R::xgk(j);
        g.template pair_fast<TR>(center - abscissa, center + abscissa, fv1[j], fv2[j], m);
    }
    const double f_center = g.template at_fast<TR>(center, m);
    if (!m.ok) return qk_rule_slow<N, TR>(g, a, b);
    return qk_reduce<N>(fv1, fv2, f_center, a, b);
}

// The Gauss-Kronrod rule of single/qk/naive.rs:32-106 for ONE thread: the thread evaluates
// all 2N+1 nodes itself and accumulates the Kronrod / Gauss / |f| sums and the asc pass in
// the reference's order (j = 0..N-1 from the centre outwards), so results are bit-equal
// to the CPU for IEEE-exact integrands.  UNROLL = how many node pairs are inlined per loop
// trip (N = fully unrolled, values stay in registers; smaller = the 2N values live in
// local memory, which is lane-interleaved and L1-resident).
template <int N, int UNROLL, class G>
GKQ_DEV QK qk_rule(const G& g, double a, double b) {
    using R = Rule<N>;
#ifndef GKQ_NO_RULE_LIGHT
    if constexpr (UNROLL >= N && !G::HEAVY) {
        if (g.transform) return qk_rule_light<N, true>(g, a, b);
        return qk_rule_light<N, false>(g, a, b);
    }
#endif
    double fv1[N], fv2[N];
    const double center = 0.5 * (a + b);
    const double half_length = 0.5 * (b - a);
#pragma unroll UNROLL
    for (int j = 0; j < N; ++j) {
        double abscissa = half_length * R::xgk(j);
        g.pair(center - abscissa, center + abscissa, fv1[j], fv2[j]);
    }
    const double f_center = g(center);
    return qk_reduce<N>(fv1, fv2, f_center, a, b);
}
// The same rule with the infinite-range test hoisted out of the node loop (two copies of the
// loop: worth it in the flat bulk kernels, where the loop is all there is).
template <int N, int UNROLL, bool TR, class G>
GKQ_DEV QK qk_rule_fixed(const G& g, double a, double b) {
    using R = Rule<N>;
#ifndef GKQ_NO_RULE_LIGHT
    if constexpr (UNROLL >= N && !G::HEAVY) return qk_rule_light<N, TR>(g, a, b);
#endif
    double fv1[N], fv2[N];
    const double center = 0.5 * (a + b);
    const double half_length = 0.5 * (b - a);
#pragma unroll UNROLL
    for (int j = 0; j < N; ++j) {
        double abscissa = half_length * R::xgk(j);
        g.template pair_t<TR>(center - abscissa, center + abscissa, fv1[j], fv2[j]);
    }
    const double f_center = g.template at<TR>(center);
    return qk_reduce<N>(fv1, fv2, f_center, a, b);
}
template <int N, int UNROLL, class G>
GKQ_DEV QK qk_rule_hoisted(const G& g, double a, double b) {
    if (g.transform) return qk_rule_fixed<N, UNROLL, true>(g, a, b);
    return qk_rule_fixed<N, UNROLL, false>(g, a, b);
}

// Warp-cooperative 25-point rule on up to two ranges (one bisection step = 50 samples) of ONE
// integral: lane L evaluates sample slots L and L+32; slot s < 25 belongs to range 0, 25 <= s
// < 50 to range 1; inside a range t = s % 25: t < 12 -> centre - h*x_t, t < 24 -> centre +
// h*x_(t-12), t = 24 -> centre.  Every lane then pulls the samples with shuffles in the
// reference's summation order and reduces redundantly (uniform control flow), so the result is
// bit-identical to qk_rule.  Used when a warp has only a few live integrals left (the tail),
// where one thread walking 50 samples alone would be the critical path of the whole batch.
template <class G>
GKQ_DEV void qk25_pair_coop(const G& g, const double* xgk_sh, bool h0, double a0, double b0, bool h1,
                            double a1, double b1, QK& q0, QK& q1) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    double v[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int slot = lane + 32 * r;
        const int half = slot >= 25 ? 1 : 0;
        const int t = slot - 25 * half;
        const bool on = slot < 50 && (half ? h1 : h0);
        v[r] = 0.;
        if (on) {
            const double a = half ? a1 : a0, b = half ? b1 : b0;
            const double center = 0.5 * (a + b);
            const double half_length = 0.5 * (b - a);
            double x = center;
            if (t < 24) {
                const double abscissa = half_length * xgk_sh[t < 12 ? t : t - 12];
                x = t < 12 ? center - abscissa : center + abscissa;
            }
            v[r] = g(x);
        }
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        if (half ? !h1 : !h0) continue;  // warp-uniform
        double fv1[12], fv2[12];
        auto pull = [&](int slot) {
            return slot < 32 ? __shfl_sync(FULL, v[0], slot) : __shfl_sync(FULL, v[1], slot - 32);
        };
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            fv1[j] = pull(25 * half + j);
            fv2[j] = pull(25 * half + 12 + j);
        }
        const double fc = pull(25 * half + 24);
        const QK q = qk_reduce<12>(fv1, fv2, fc, half ? a1 : a0, half ? b1 : b0);
        if (half) q1 = q; else q0 = q;
    }
}

}  // namespace gkq
