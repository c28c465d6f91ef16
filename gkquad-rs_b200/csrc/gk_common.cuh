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

// Node tables of the 17- and 25-point Kronrod rules (values: single/qk/mod.rs:78-148).
// All lanes of a warp read the same entry at the same time (thread-per-integral mapping),
// so constant memory broadcasts them at register speed.
template <int N> struct Rule;
static __constant__ double c_xgk17[8] = {
    0.183434642495649804939476142360184, 0.360701097928131957192548622296891,
    0.525532409916328985817739049189246, 0.672354070945158677156310738092831,
    0.796666477413626739591553936475830, 0.894120906847456421948361017538251,
    0.960289856497536231683560868569473, 0.993379875881716155935888069019671};
static __constant__ double c_wg17[4] = {
    0.362683783378361982965150449277196, 0.313706645877887287337962201986601,
    0.222381034453374470544355994426241, 0.101228536290376259152531354309962};
static __constant__ double c_wgk17[8] = {
    0.181400025068034643061748525172550, 0.172070608555211311857294880203857,
    0.156652606168188400490248088486969, 0.136263109255172215262338745254506,
    0.111646370826839613222108158933941, 0.082482298931358330688625193445608,
    0.049439395002139308500363969446998, 0.017822383320710355152786961202750};
static __constant__ double c_xgk25[12] = {
    0.125233408511468915472441369463853, 0.248505748320469276267790960362718,
    0.367831498998180193752691536643718, 0.481339450478157092935943615018832,
    0.587317954286617447296702418940534, 0.684059895470055893944929100341155,
    0.769902674194304687036893833212818, 0.843558124161153244792141885059839,
    0.904117256370474856678465866119096, 0.950537795943121296549060195131619,
    0.981560634246719250690549090149281, 0.996933922529595426912350237258385};
static __constant__ double c_wg25[6] = {
    0.249147045813402785000562436042951, 0.233492536538354808760849898924878,
    0.203167426723065921749064455809798, 0.160078328543346226334652529543359,
    0.106939325995318430960254718193996, 0.047175336386511827194615961485017};
static __constant__ double c_wgk25[12] = {
    0.124584164536156073437312473209229, 0.121626303523948383246099758091310,
    0.116712053501756826293580745305730, 0.110022604977644072635907398742250,
    0.101649732279060277715688770491228, 0.091549468295049210528171939739614,
    0.079920275333601701493392609529783, 0.067250907050839930304940940047316,
    0.053697017607756251228889163320458, 0.038915230469299477115089632285863,
    0.023036084038982232591084580367969, 0.008257711433168395757693922439212};

template <> struct Rule<8> {
    static constexpr int NPTS = 17;
    static constexpr double WCK = 0.184446405744691643528970955705643;
    GKQ_DEV static double xgk(int j) { return c_xgk17[j]; }
    GKQ_DEV static double wgk(int j) { return c_wgk17[j]; }
    GKQ_DEV static double wg(int j) { return c_wg17[j]; }
};
template <> struct Rule<12> {
    static constexpr int NPTS = 25;
    static constexpr double WCK = 0.125556893905474335304296132860078;
    GKQ_DEV static double xgk(int j) { return c_xgk25[j]; }
    GKQ_DEV static double wgk(int j) { return c_wgk25[j]; }
    GKQ_DEV static double wg(int j) { return c_wg25[j]; }
};

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
        FastMath m;
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

// The Gauss-Kronrod rule of single/qk/naive.rs:32-106 for ONE thread: the thread evaluates
// all 2N+1 nodes itself and accumulates the Kronrod / Gauss / |f| sums and the asc pass in
// the reference's order (j = 0..N-1 from the centre outwards), so results are bit-equal
// to the CPU for IEEE-exact integrands.  UNROLL = how many node pairs are inlined per loop
// trip (N = fully unrolled, values stay in registers; smaller = the 2N values live in
// local memory, which is lane-interleaved and L1-resident).
template <int N, int UNROLL, class G>
GKQ_DEV QK qk_rule(const G& g, double a, double b) {
    using R = Rule<N>;
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
