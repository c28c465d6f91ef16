// gkq_oracle.hpp -- CPU oracle for the gkquad hot path.  TEST INFRASTRUCTURE ONLY.
//
// This header is a literal, sequential C++ restatement of the reference crate
// `gkquad` 0.0.4 (Kogia-sima/gkquad-rs).  It exists so that the CUDA engine in
// `gkquad-rs_b200/csrc/` can be checked against the reference's behaviour: only
// `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may
// build, link, import or execute anything under `oracle/`.  The product library
// (`libgkquad_b200.so`) never includes this file.
//
// Parity status: PINNED.  The reference cannot be compiled here (no Rust
// toolchain); the oracle is pinned instead by every golden vector the reference
// keeps in gkquad/tests/{qk_single,algorithms_single,algorithms_double}.rs (see
// tests/test_oracle_golden.py) at the reference's own tolerances.
//
// Rules followed (SURVEY.md section 4 gotchas):
//   * compile with -ffp-contract=off and without -ffast-math (Rust never fuses);
//   * every floating-point expression keeps the reference's operand order;
//   * `limit` seen by qpsrt/increase_nrmax is the Vec *capacity* (modelled below);
//   * f64::max / f64::min ignore NaN the way Rust's do.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference/gkquad/src/).
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <vector>

namespace gkqo {

constexpr double F64_EPSILON = 2.220446049250313e-16;      // core::f64::EPSILON
constexpr double F64_MIN_POSITIVE = 2.2250738585072014e-308;  // core::f64::MIN_POSITIVE
constexpr double F64_MAX = 1.7976931348623157e308;         // core::f64::MAX

// Rust's f64::max / f64::min: if one operand is NaN the other is returned.
inline double rmax(double a, double b) {
    if (a != a) return b;
    if (b != b) return a;
    return a > b ? a : b;
}
inline double rmin(double a, double b) {
    if (a != a) return b;
    if (b != b) return a;
    return a < b ? a : b;
}

// error.rs:140-178 -- numbering follows declaration order, 0 = no error.
enum Status : uint32_t {
    ST_NONE = 0,
    ST_INSUFFICIENT_ITERATION = 1,
    ST_ROUNDOFF_ERROR = 2,
    ST_SUBRANGE_TOO_SMALL = 3,
    ST_DIVERGENT = 4,
    ST_NAN_VALUE_ENCOUNTERED = 5,
};

// common.rs:9-42
enum TolKind : int { TOL_ABSOLUTE = 0, TOL_RELATIVE = 1, TOL_ABS_OR_REL = 2, TOL_ABS_AND_REL = 3 };

struct Tolerance {
    int kind;
    double abs_;
    double rel_;
    // common.rs:34-41
    double to_abs(double value) const {
        switch (kind) {
            case TOL_ABSOLUTE: return abs_;
            case TOL_RELATIVE: return value * rel_;
            case TOL_ABS_OR_REL: return rmax(abs_, value * rel_);
            default: return rmin(abs_, value * rel_);
        }
    }
};

struct Range {
    double begin, end;
};

// common.rs:213-231 + the Option<RuntimeError> of ValueWithError (common.rs:60-63)
struct Result {
    double estimate = 0.0;
    double delta = F64_MAX;
    uint64_t nevals = 0;
    uint32_t status = ST_NONE;
};

struct Config {
    Tolerance tol;
    uint64_t max_evals;
    const double* points;
    size_t npoints;
};

// single/qk/mod.rs:16-25
struct QKResult {
    double estimate, delta, absvalue, asc;
};

// ---------------------------------------------------------------------------
// Node tables: single/qk/mod.rs:78-412 (data).  gk_tables.inc is generated from first principles
// by tools/gen_gk_tables.py (roots of the Legendre / Stieltjes polynomials at 120 digits) and was
// checked there to be bit-identical to the reference's literals.
// ---------------------------------------------------------------------------
#include "gk_tables.inc"
static const double XGK17[8] = {GK_XGK17};
static const double WG17[4] = {GK_WG17};
static const double WCK17 = GK_WCK17;
static const double WGK17[8] = {GK_WGK17};
static const double XGK25[12] = {GK_XGK25};
static const double WG25[6] = {GK_WG25};
static const double WCK25 = GK_WCK25;
static const double WGK25[12] = {GK_WGK25};
static const double XGK33[16] = {GK_XGK33};
static const double WG33[8] = {GK_WG33};
static const double WCK33 = GK_WCK33;
static const double WGK33[16] = {GK_WGK33};
static const double XGK41[20] = {GK_XGK41};
static const double WG41[10] = {GK_WG41};
static const double WCK41 = GK_WCK41;
static const double WGK41[20] = {GK_WGK41};
static const double XGK49[24] = {GK_XGK49};
static const double WG49[12] = {GK_WG49};
static const double WCK49 = GK_WCK49;
static const double WGK49[24] = {GK_WGK49};
static const double XGK57[28] = {GK_XGK57};
static const double WG57[14] = {GK_WG57};
static const double WCK57 = GK_WCK57;
static const double WGK57[28] = {GK_WGK57};

// single/util.rs:118-141
inline double rescale_error(double err, double result_abs, double result_asc) {
    err = std::fabs(err);
    if (result_asc != 0.0 && err != 0.0) {
        double scale = 200.0 * err;
        if (scale < result_asc) {
            scale /= result_asc;
            err = result_asc * scale * std::sqrt(scale);
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

// single/util.rs:74-103 -- the infinite-range wrapper, applied point by point
// exactly like apply_to_slice does.
template <class F>
struct IntegrandWrapper {
    F& inner;
    bool transform;
    inline double apply(double x) {
        if (transform) {
            double coef = 1. / (1. - std::fabs(x));
            double x2 = x * coef;
            return inner(x2) * coef * coef;
        }
        return inner(x);
    }
};

// single/qk/naive.rs:32-106 -- generic rule.  N = number of node pairs.
// buf layout (naive.rs:60-66): buf[j] = c - h x_j, buf[j+n] = c + h x_j, buf[2n] = c,
// evaluated in slice order (single/common.rs:135-137).
template <int N, class W>
inline QKResult qk(W& f, const Range& range, const double* xgk, const double* wg,
                   const double* wgk, double wck) {
    double buf[2 * N + 1];
    const int n = N;
    double center = 0.5 * (range.begin + range.end);
    double half_length = 0.5 * (range.end - range.begin);
    double abs_half_length = std::fabs(half_length);

    buf[n << 1] = center;
    for (int j = 0; j < n; ++j) {
        double abscissa = half_length * xgk[j];
        buf[j] = center - abscissa;
        buf[j + n] = center + abscissa;
    }
    for (int j = 0; j < 2 * n + 1; ++j) buf[j] = f.apply(buf[j]);

    double f_center = buf[n << 1];
    double result_gauss = 0.;
    double result_kronrod = f_center * wck;
    double result_abs = std::fabs(result_kronrod);

    for (int j = 0; j < n; ++j) {
        double fval1 = buf[j];
        double fval2 = buf[j + n];
        double fsum = fval1 + fval2;
        result_kronrod += wgk[j] * fsum;
        result_abs += wgk[j] * (std::fabs(fval1) + std::fabs(fval2));
        if (j % 2 == 0) result_gauss += wg[j / 2] * fsum;
    }

    double mean = result_kronrod * 0.5;
    double result_asc = wck * std::fabs(f_center - mean);
    for (int j = 0; j < n; ++j) {
        result_asc += wgk[j] * (std::fabs(buf[j] - mean) + std::fabs(buf[j + n] - mean));
    }

    double err = (result_kronrod - result_gauss) * half_length;
    result_kronrod *= half_length;
    result_abs *= abs_half_length;
    result_asc *= abs_half_length;

    return QKResult{result_kronrod, rescale_error(err, result_abs, result_asc), result_abs,
                    result_asc};
}

// single/qk/mod.rs:28-41
template <class W>
inline QKResult qk17(W& f, const Range& r) { return qk<8>(f, r, XGK17, WG17, WGK17, WCK17); }
template <class W>
inline QKResult qk25(W& f, const Range& r) { return qk<12>(f, r, XGK25, WG25, WGK25, WCK25); }
// single/qk/mod.rs:44-73
template <class W>
inline QKResult qk33(W& f, const Range& r) { return qk<16>(f, r, XGK33, WG33, WGK33, WCK33); }
template <class W>
inline QKResult qk41(W& f, const Range& r) { return qk<20>(f, r, XGK41, WG41, WGK41, WCK41); }
template <class W>
inline QKResult qk49(W& f, const Range& r) { return qk<24>(f, r, XGK49, WG49, WGK49, WCK49); }
template <class W>
inline QKResult qk57(W& f, const Range& r) { return qk<28>(f, r, XGK57, WG57, WGK57, WCK57); }

// single/util.rs:110-115
inline bool subrange_too_small(double a1, double a2, double b2) {
    double tmp = (1. + 100. * F64_EPSILON) * (std::fabs(a2) + 1000. * F64_MIN_POSITIVE);
    return std::fabs(a1) <= tmp && std::fabs(b2) <= tmp;
}
// single/util.rs:146-148
inline bool test_positivity(double result, double resabs) {
    return std::fabs(result) >= (1.0 - 50.0 * F64_EPSILON) * resabs;
}
// single/util.rs:162-170
inline double transform_point(double x) {
    if (x == -std::numeric_limits<double>::infinity()) return -1.0;
    if (x == std::numeric_limits<double>::infinity()) return 1.0;
    return x / (1.0 + std::fabs(x));
}
// single/util.rs:174-176
inline Range transform_range(const Range& r) {
    return Range{transform_point(r.begin), transform_point(r.end)};
}

// single/workspace.rs:7-28
struct SubRangeInfo {
    Range range;
    double estimate;
    double delta;
    size_t level;
};

// single/workspace.rs:32-278.  `cap` models Vec::capacity() of `subranges`
// (== that of `order`, workspace.rs:81-84): with_capacity(n) gives exactly n;
// reserve()/push() grow like alloc::raw_vec grow_amortized:
// new_cap = max(2*cap, required, 4).
struct WorkSpace {
    size_t nrmax = 0;
    size_t i = 0;
    size_t maximum_level = 0;
    std::vector<SubRangeInfo> subranges;
    std::vector<size_t> order;
    size_t cap = 0;

    WorkSpace() = default;
    explicit WorkSpace(size_t n) : cap(n) {}  // workspace.rs:62-70
    // Become a fresh WorkSpace::with_capacity(n) without giving the std::vector storage
    // back to the allocator (keeps the batch drivers allocation-free like the reference,
    // README.md:45 "0 allocs").
    void reset_fresh(size_t n) {
        clear();
        cap = n;
    }

    size_t size() const { return subranges.size(); }   // :74-78
    size_t capacity() const { return cap; }            // :81-84
    void grow_to(size_t required) {
        size_t nc = cap * 2 > required ? cap * 2 : required;
        if (nc < 4) nc = 4;
        cap = nc;
    }
    void reserve(size_t n) {                           // :87-90
        if (cap - subranges.size() >= n) return;
        grow_to(subranges.size() + n);
    }
    void clear() {                                     // :93-99
        subranges.clear();
        order.clear();
        i = 0;
        nrmax = 0;
        maximum_level = 0;
    }
    void push_raw(const SubRangeInfo& s) {
        if (subranges.size() == cap) grow_to(subranges.size() + 1);
        subranges.push_back(s);
    }
    void push(const SubRangeInfo& s) {                 // :102-105
        push_raw(s);
        order.push_back(order.size());
    }

    // :107-138 -- note order[] is indexed by *value* in the swap (restated literally).
    void sort_results() {
        size_t nint = size();
        if (nint == 0) return;
        for (size_t i = 0; i < nint; ++i) {
            size_t i1 = order[i];
            double e1 = subranges[i1].delta;
            size_t i_max = i1;
            for (size_t k = i + 1; k < nint; ++k) {
                size_t i2 = order[k];
                double e2 = subranges[i2].delta;
                if (e2 >= e1) {
                    i_max = i2;
                    e1 = e2;
                }
            }
            if (i_max != i1) {
                order[i] = order[i_max];
                order[i_max] = i1;
            }
        }
        this->i = order[0];
    }

    // :141-160
    void update(const SubRangeInfo& s1, const SubRangeInfo& s2) {
        size_t new_level = subranges[i].level + 1;
        if (s2.delta > s1.delta) {
            subranges[i] = s2;
            push_raw(s1);
        } else {
            subranges[i] = s1;
            push_raw(s2);
        }
        order.push_back(order.size());
        if (new_level > maximum_level) maximum_level = new_level;
        qpsrt();
    }

    // :163-223
    void qpsrt() {
        size_t last = size() - 1;
        size_t limit = capacity();
        size_t i_nrmax = nrmax;
        size_t i_maxdelta = order[i_nrmax];

        if (last < 2) {
            order[0] = 0;
            order[1] = 1;
            this->i = i_maxdelta;
            return;
        }

        double deltamax = subranges[i_maxdelta].delta;
        while (i_nrmax > 0 && deltamax > subranges[order[i_nrmax - 1]].delta) {
            order[i_nrmax] = order[i_nrmax - 1];
            i_nrmax -= 1;
        }

        size_t top = (last < (limit / 2 + 2)) ? last : limit - last + 1;

        size_t i = i_nrmax + 1;
        while (i < top && deltamax < subranges[order[i]].delta) {
            order[i - 1] = order[i];
            i += 1;
        }
        order[i - 1] = i_maxdelta;

        double errmin = subranges[last].delta;
        int64_t k = (int64_t)top - 1;
        while (k >= (int64_t)i - 1 && errmin >= subranges[order[(size_t)k]].delta) {
            order[(size_t)k + 1] = order[(size_t)k];
            k -= 1;
        }
        order[(size_t)(k + 1)] = last;

        i_maxdelta = order[i_nrmax];
        this->i = i_maxdelta;
        this->nrmax = i_nrmax;
    }

    // :233-259
    bool increase_nrmax() {
        size_t id = nrmax;
        size_t limit = capacity();
        size_t last = size() - 1;
        size_t jupbnd = (last > (1 + limit / 2)) ? limit + 1 - last : last;
        for (size_t k = id; k <= jupbnd; ++k) {
            size_t i_max = order[nrmax];
            this->i = i_max;
            if (subranges[i_max].level < maximum_level) return true;
            nrmax += 1;
        }
        return false;
    }
    // :262-265
    void reset_nrmax() {
        nrmax = 0;
        i = order[0];
    }
    const SubRangeInfo& get() const { return subranges[i]; }  // :269-271
    // :275-277 -- left-to-right sum starting from 0.0
    double sum_results() const {
        double s = 0.0;
        for (const auto& r : subranges) s += r.estimate;
        return s;
    }
};

// single/qelg.rs:12-336
struct ExtrapolationTable {
    size_t n = 0;
    double rlist2[52] = {0};
    size_t nres = 0;
    double res3la[3] = {0, 0, 0};

    void append(double y) {  // :33-37
        rlist2[n] = y;
        n += 1;
    }

    void qelg(double& result, double& abserr) {  // :50-323
        double* epstab = rlist2;
        const size_t n = this->n - 1;
        const double current = epstab[n];
        const size_t newelm = n / 2;
        const size_t n_orig = n;
        size_t n_final = n;
        const size_t nres_orig = nres;

        result = current;
        abserr = F64_MAX;
        if (n < 2) {
            result = current;
            abserr = F64_MAX;
            return;
        }

        epstab[n + 2] = epstab[n];
        epstab[n] = F64_MAX;

        for (size_t i = 0; i < newelm; ++i) {
            double* ep = epstab + (n - 2 * i);
            double res = ep[2];
            double e0 = *(ep - 2);
            double e1 = *(ep - 1);
            double e2 = res;

            double elabs = std::fabs(e1);
            double delta2 = e2 - e1;
            double err2 = std::fabs(delta2);
            double tol2 = rmax(std::fabs(e2), elabs) * F64_EPSILON;
            double delta3 = e1 - e0;
            double err3 = std::fabs(delta3);
            double tol3 = rmax(elabs, std::fabs(e0)) * F64_EPSILON;

            if (err2 <= tol2 && err3 <= tol3) {
                result = res;
                double absolute = err2 + err3;
                double relative = 5. * F64_EPSILON * std::fabs(res);
                abserr = rmax(absolute, relative);
                return;
            }

            double e3 = *ep;
            *ep = e1;
            double delta1 = e1 - e3;
            double err1 = std::fabs(delta1);
            double tol1 = rmax(elabs, std::fabs(e3)) * F64_EPSILON;

            if (err1 <= tol1 || err2 <= tol2 || err3 <= tol3) {
                n_final = 2 * i;
                break;
            }

            double ss = (1. / delta1 + 1. / delta2) - 1. / delta3;

            if (std::fabs(ss * e1) <= 0.0001) {
                n_final = 2 * i;
                break;
            }

            res = e1 + 1. / ss;
            *ep = res;

            double error = err2 + std::fabs(res - e2) + err3;
            if (error <= abserr) {
                abserr = error;
                result = res;
            }
        }

        const size_t LIMEXP = 50 - 1;
        if (n_final == LIMEXP) n_final = 2 * (LIMEXP / 2);

        {
            size_t start = (n_orig & 1) ? 1 : 0;
            for (size_t k = 0; k <= newelm; ++k) epstab[start + 2 * k] = epstab[start + 2 * k + 2];
        }

        if (n_orig != n_final) {
            // core::ptr::copy == memmove; source is above destination so a forward loop is exact
            for (size_t k = 0; k <= n_final; ++k) epstab[k] = epstab[n_orig - n_final + k];
        }
        this->n = n_final + 1;

        if (nres_orig < 3) {
            res3la[nres_orig] = result;
            abserr = F64_MAX;
        } else {
            abserr = std::fabs(result - res3la[2]) + std::fabs(result - res3la[1]) +
                     std::fabs(result - res3la[0]);
            res3la[0] = res3la[1];
            res3la[1] = res3la[2];
            res3la[2] = result;
        }

        nres = nres_orig + 1;
        abserr = rmax(abserr, 5. * F64_EPSILON * std::fabs(result));
    }
};

inline Result finish(double estimate, double delta, uint64_t nevals, uint32_t error) {
    Result r;
    r.estimate = estimate;
    r.delta = delta;
    r.nevals = nevals;
    r.status = error;
    return r;
}

// single/algorithm/qags.rs:315-376
template <class W>
inline Result initial_integral(W& f, const Range& range, const Config& config, double& absvalue,
                               bool& finished) {
    Result solution;  // Solution::default(): estimate 0, delta f64::MAX, nevals 0
    absvalue = 0.0;
    finished = true;
    for (int i = 0; i < 2; ++i) {
        QKResult result0;
        if (i == 0) {
            solution.nevals += 17;
            result0 = qk17(f, range);
        } else {
            solution.nevals += 25;
            result0 = qk25(f, range);
        }
        solution.estimate = result0.estimate;
        solution.delta = result0.delta;
        absvalue = result0.absvalue;

        if (result0.estimate != result0.estimate) {
            solution.status = ST_NAN_VALUE_ENCOUNTERED;
            absvalue = 0.0;
            return solution;
        }

        double tolerance = config.tol.to_abs(std::fabs(result0.estimate));
        if ((result0.delta <= tolerance && result0.delta != result0.asc) || result0.delta == 0.0) {
            absvalue = 0.0;
            return solution;
        }

        double round_off = 100. * F64_EPSILON * result0.absvalue;
        if (result0.delta <= round_off && result0.delta > tolerance) {
            solution.status = ST_ROUNDOFF_ERROR;
            absvalue = 0.0;
            return solution;
        }

        if (config.max_evals < (uint64_t)(42 + i * 25)) {
            solution.status = ST_INSUFFICIENT_ITERATION;
            absvalue = 0.0;
            return solution;
        }

        if (i == 0 && result0.delta > tolerance * 1024.) break;
    }
    finished = false;
    return solution;
}

// single/algorithm/qags.rs:67-311
template <class W>
Result qags_impl(W& f, const Range& range, const Config& config, WorkSpace& ws) {
    double ertest = 0;
    double error_over_large_ranges = 0;
    double correc = 0.;
    size_t ktmin = 0;
    int roundoff_type1 = 0, roundoff_type2 = 0, roundoff_type3 = 0;
    uint32_t error = ST_NONE;
    int error2 = 0;
    bool extrapolate = false;
    bool disallow_extrapolation = false;

    if (config.max_evals < 17) {
        Result r;
        r.status = ST_INSUFFICIENT_ITERATION;
        return r;
    }

    double absvalue;
    bool finished;
    Result result0 = initial_integral(f, range, config, absvalue, finished);
    if (finished) return result0;

    uint64_t nevals = result0.nevals;

    ws.clear();
    ws.reserve((config.max_evals - nevals) / 50 + 1);
    ws.push(SubRangeInfo{range, result0.estimate, result0.delta, 0});

    ExtrapolationTable table;
    table.append(result0.estimate);

    double area = result0.estimate;
    double errsum = result0.delta;
    double res_ext = result0.estimate;
    double err_ext = F64_MAX;
    const uint64_t max_iters = (config.max_evals - nevals) / 50;

    for (uint64_t iteration = 1; iteration <= max_iters; ++iteration) {
        const SubRangeInfo info = ws.get();
        size_t current_level = info.level + 1;

        // util.rs:151-159
        double center = (info.range.begin + info.range.end) * 0.5;
        Range r1{info.range.begin, center};
        Range r2{center, info.range.end};

        QKResult result1 = qk25(f, r1);
        QKResult result2 = qk25(f, r2);
        nevals += 50;

        if (result1.estimate != result1.estimate || result2.estimate != result2.estimate) {
            error = ST_NAN_VALUE_ENCOUNTERED;
            break;
        }

        double area12 = result1.estimate + result2.estimate;
        double error12 = result1.delta + result2.delta;
        double last_e_i = info.delta;

        errsum += error12 - info.delta;
        area += area12 - info.estimate;

        double tolerance = config.tol.to_abs(std::fabs(area));

        if (result1.asc != result1.delta && result2.asc != result2.delta) {
            if (std::fabs(info.estimate - area12) <= 1e-5 * std::fabs(area12) &&
                error12 >= 0.99 * info.delta) {
                if (!extrapolate)
                    roundoff_type1 += 1;
                else
                    roundoff_type2 += 1;
            }
            if (iteration > 10 && error12 > info.delta) roundoff_type3 += 1;
        }

        if (roundoff_type1 + roundoff_type2 >= 10 || roundoff_type3 >= 20) error = ST_ROUNDOFF_ERROR;
        if (roundoff_type2 >= 5) error2 = 1;

        if (subrange_too_small(r1.begin, r1.end, r2.end)) error = ST_SUBRANGE_TOO_SMALL;

        if (errsum <= tolerance) {
            return finish(ws.sum_results() - info.estimate + result1.estimate + result2.estimate,
                          errsum, nevals, error);
        }

        ws.update(SubRangeInfo{r1, result1.estimate, result1.delta, current_level},
                  SubRangeInfo{r2, result2.estimate, result2.delta, current_level});

        if (error != ST_NONE) break;

        // qags.rs:200 -- compares the iteration counter with max_evals - 1 (sic)
        if (iteration >= config.max_evals - 1) {
            error = ST_INSUFFICIENT_ITERATION;
            break;
        }

        if (iteration == 2) {
            error_over_large_ranges = errsum;
            ertest = tolerance;
            table.append(area);
            continue;
        }

        if (disallow_extrapolation) continue;

        error_over_large_ranges -= last_e_i;
        if (current_level < ws.maximum_level) error_over_large_ranges += error12;

        if (!extrapolate) {
            if (ws.get().level < ws.maximum_level) continue;
            extrapolate = true;
            ws.nrmax = 1;
        }

        if (error_over_large_ranges > ertest && ws.increase_nrmax()) continue;

        double reseps = 0., abseps = 0.;
        table.append(area);
        table.qelg(reseps, abseps);

        ktmin += 1;
        if (ktmin > 5 && err_ext < 0.001 * errsum) error = ST_ROUNDOFF_ERROR;

        if (abseps < err_ext) {
            ktmin = 0;
            err_ext = abseps;
            res_ext = reseps;
            correc = error_over_large_ranges;
            ertest = config.tol.to_abs(std::fabs(reseps));
            if (err_ext <= ertest) break;
        }

        if (table.n == 1) disallow_extrapolation = true;

        if (error != ST_NONE) break;

        ws.reset_nrmax();
        extrapolate = false;
        error_over_large_ranges = errsum;
    }

    if (err_ext == F64_MAX) return finish(ws.sum_results(), errsum, nevals, error);

    if (error != ST_NONE || error2 > 0) {
        if (error2 > 0) err_ext += correc;
        if (error == ST_NONE) error = ST_ROUNDOFF_ERROR;

        if (res_ext != 0.0 && area != 0.0) {
            if (err_ext / std::fabs(res_ext) > errsum / std::fabs(area))
                return finish(ws.sum_results(), errsum, nevals, error);
        } else if (err_ext > errsum) {
            return finish(ws.sum_results(), errsum, nevals, error);
        } else if (area == 0.0) {
            return finish(res_ext, err_ext, nevals, error);
        }
    }

    bool positive_integrand = test_positivity(result0.estimate, absvalue);
    if (!positive_integrand && rmax(std::fabs(res_ext), std::fabs(area)) < 0.01 * absvalue)
        return finish(res_ext, err_ext, nevals, error);

    double ratio = res_ext / area;
    if ((ratio < 0.01 || ratio > 100.0 || errsum > std::fabs(area)) && error == ST_NONE)
        error = ST_DIVERGENT;

    return finish(res_ext, err_ext, nevals, error);
}

// single/algorithm/qags.rs:41-63
template <class F>
Result qags(F& f, const Range& range, const Config& config, WorkSpace& ws) {
    bool transform = !std::isfinite(range.begin) || !std::isfinite(range.end);
    IntegrandWrapper<F> w{f, transform};
    Range r = transform ? transform_range(range) : range;
    return qags_impl(w, r, config, ws);
}

// ---------------------------------------------------------------------------
// QAG (deprecated in the reference, still public): single/algorithm/qag.rs:65-248.
// The same skeleton as QAGS without the epsilon table; its initial pass uses 50 eps (not 100)
// for the round-off floor (qag.rs:217) and keeps no absvalue.
// ---------------------------------------------------------------------------
template <class W>
inline Result qag_initial_integral(W& f, const Range& range, const Config& config, bool& finished) {
    Result solution;  // Solution::default()
    finished = true;
    for (int i = 0; i < 2; ++i) {
        QKResult result0;
        if (i == 0) {
            solution.nevals += 17;
            result0 = qk17(f, range);
        } else {
            solution.nevals += 25;
            result0 = qk25(f, range);
        }
        solution.estimate = result0.estimate;
        solution.delta = result0.delta;
        if (result0.estimate != result0.estimate) {  // qag.rs:204-209
            solution.status = ST_NAN_VALUE_ENCOUNTERED;
            return solution;
        }
        double tolerance = config.tol.to_abs(std::fabs(result0.estimate));
        if ((result0.delta <= tolerance && result0.delta != result0.asc) || result0.delta == 0.0)
            return solution;  // :212-214
        double round_off = 50. * F64_EPSILON * result0.absvalue;  // :217
        if (result0.delta <= round_off && result0.delta > tolerance) {
            solution.status = ST_ROUNDOFF_ERROR;
            return solution;
        }
        if (config.max_evals < (uint64_t)(42 + i * 25)) {  // :222-227
            solution.status = ST_INSUFFICIENT_ITERATION;
            return solution;
        }
        if (i == 0 && result0.delta > tolerance * 1024.) break;  // :229-231
    }
    finished = false;
    return solution;
}

// qag.rs:65-183
template <class W>
Result qag_impl(W& f, const Range& range, const Config& config, WorkSpace& ws) {
    int roundoff_type1 = 0, roundoff_type2 = 0;
    uint32_t error = ST_NONE;
    if (config.max_evals < 17) {
        Result r;
        r.status = ST_INSUFFICIENT_ITERATION;
        return r;
    }
    bool finished;
    Result result0 = qag_initial_integral(f, range, config, finished);
    if (finished) return result0;

    double area = result0.estimate;
    double deltasum = result0.delta;
    uint64_t nevals = result0.nevals;

    ws.clear();
    ws.reserve((config.max_evals - nevals) / 50 + 1);
    ws.push(SubRangeInfo{range, result0.estimate, result0.delta, 0});

    const uint64_t max_iters = (config.max_evals - nevals) / 50;
    for (uint64_t it = 1; it <= max_iters; ++it) {
        const SubRangeInfo info = ws.get();
        size_t current_level = info.level + 1;
        double center = (info.range.begin + info.range.end) * 0.5;  // util.rs:151-159
        Range r1{info.range.begin, center};
        Range r2{center, info.range.end};
        QKResult result1 = qk25(f, r1);
        QKResult result2 = qk25(f, r2);
        nevals += 50;
        if (result1.estimate != result1.estimate || result2.estimate != result2.estimate) {
            error = ST_NAN_VALUE_ENCOUNTERED;  // :122-125
            break;
        }
        double area12 = result1.estimate + result2.estimate;
        double delta12 = result1.delta + result2.delta;
        deltasum += delta12 - info.delta;
        area += area12 - info.estimate;
        if (result1.asc != result1.delta && result2.asc != result2.delta) {  // :135-142
            if (std::fabs(info.estimate - area12) <= 1e-5 * std::fabs(area12) && delta12 >= 0.99 * info.delta)
                roundoff_type1 += 1;
            else
                roundoff_type2 += 1;
        }
        double tolerance = config.tol.to_abs(std::fabs(area));
        if (deltasum > tolerance) {  // :147-162
            if (roundoff_type1 >= 6 || roundoff_type2 >= 20)
                error = ST_ROUNDOFF_ERROR;
            else if (subrange_too_small(r1.begin, r1.end, r2.end))
                error = ST_SUBRANGE_TOO_SMALL;
            if (error != ST_NONE)
                return finish(ws.sum_results() - info.estimate + result1.estimate + result2.estimate, deltasum,
                              nevals, error);
        }
        ws.update(SubRangeInfo{r1, result1.estimate, result1.delta, current_level},
                  SubRangeInfo{r2, result2.estimate, result2.delta, current_level});
        if (deltasum <= tolerance) break;  // :170-172
    }
    return finish(ws.sum_results(), deltasum, nevals, error);  // :176
}

// qag.rs:39-60
template <class F>
Result qag(F& f, const Range& range, const Config& config, WorkSpace& ws) {
    bool transform = !std::isfinite(range.begin) || !std::isfinite(range.end);
    IntegrandWrapper<F> w{f, transform};
    Range r = transform ? transform_range(range) : range;
    return qag_impl(w, r, config, ws);
}

// single/util.rs:178-211 -- insertion sort with a strict comparator.  An element is
// left in place when is_less(prev, target); otherwise it moves down to just above the
// first lower neighbour for which is_less(another, target) holds.
template <class Less>
inline void insert_sort(double* s, size_t len, Less is_less) {
    for (size_t sp = 1; sp < len; ++sp) {
        double target = s[sp];
        if (is_less(s[sp - 1], target)) continue;
        size_t insert_pos = 0;
        for (int64_t ip = (int64_t)sp - 2; ip >= 0; --ip) {
            if (is_less(s[ip], target)) {
                insert_pos = (size_t)ip + 1;
                break;
            }
        }
        for (size_t k = sp; k > insert_pos; --k) s[k] = s[k - 1];
        s[insert_pos] = target;
    }
}

// single/algorithm/qagp.rs:375-400
inline std::vector<double> make_sorted_points(const Range& range, const double* pts, size_t npts,
                                              bool transform) {
    double mn, mx;
    if (range.begin < range.end) {
        mn = range.begin;
        mx = range.end;
    } else {
        mn = range.end;
        mx = range.begin;
    }
    std::vector<double> pts2;
    pts2.reserve(npts + 2);
    pts2.push_back(range.begin);
    for (size_t k = 0; k < npts; ++k) pts2.push_back(transform ? transform_point(pts[k]) : pts[k]);
    if (range.begin < range.end)
        insert_sort(pts2.data() + 1, pts2.size() - 1, [](double a, double b) { return a < b; });
    else
        insert_sort(pts2.data() + 1, pts2.size() - 1, [](double a, double b) { return a > b; });
    std::vector<double> kept;
    kept.reserve(pts2.size() + 1);
    for (double x : pts2)
        if (mn <= x && x <= mx) kept.push_back(x);
    kept.push_back(range.end);
    return kept;
}

// single/algorithm/qagp.rs:68-372
template <class W>
Result qagp_impl(W& f, const Range& range, const Config& config, bool transform, WorkSpace& ws) {
    std::vector<double> pts = make_sorted_points(range, config.points, config.npoints, transform);
    const uint64_t nint = pts.size() - 1;
    uint64_t nevals = 0;

    if (config.max_evals < nint * 25) {
        Result r;
        r.status = ST_INSUFFICIENT_ITERATION;
        return r;
    }

    ws.clear();
    ws.reserve(nint + (config.max_evals - nint * 25) / 50);

    double correc = 0.0;
    int ktmin = 0;
    int roundoff_type1 = 0, roundoff_type2 = 0, roundoff_type3 = 0;
    uint32_t error = ST_NONE;
    bool error2 = false;
    bool extrapolate = false;
    bool disallow_extrapolation = false;

    QKResult result0{0., 0., 0., 0.};

    for (size_t w = 0; w + 1 < pts.size(); ++w) {
        if (std::fabs(pts[w + 1] - pts[w]) < 100. * F64_MIN_POSITIVE) continue;
        Range r{pts[w], pts[w + 1]};
        QKResult result1 = qk25(f, r);
        nevals += 25;

        if (result1.estimate != result1.estimate)
            return finish(result0.estimate, result0.delta, nevals, ST_NAN_VALUE_ENCOUNTERED);

        size_t current_level = (result1.delta == result1.asc && result1.delta != 0.0) ? 1 : 0;
        // qagp.rs:403-408
        result0.estimate += result1.estimate;
        result0.delta += result1.delta;
        result0.absvalue += result1.absvalue;
        result0.asc += result1.asc;

        ws.push(SubRangeInfo{r, result1.estimate, result1.delta, current_level});
    }

    double deltasum = 0.;
    for (auto& si : ws.subranges) {
        if (si.level > 0) si.delta = result0.delta;
        deltasum += si.delta;
    }
    for (auto& si : ws.subranges) si.level = 0;

    ws.sort_results();

    double tolerance = config.tol.to_abs(std::fabs(result0.estimate));
    double round_off = 100. * F64_EPSILON * result0.absvalue;

    if (result0.delta <= round_off && result0.delta > tolerance) {
        return finish(result0.estimate, result0.delta, nevals, ST_ROUNDOFF_ERROR);
    } else if ((result0.delta <= tolerance && result0.delta != result0.asc) ||
               result0.delta == 0.0) {
        return finish(result0.estimate, result0.delta, nevals, ST_NONE);
    } else if (nevals == config.max_evals) {
        return finish(result0.estimate, result0.delta, nevals, ST_INSUFFICIENT_ITERATION);
    }

    ExtrapolationTable table;
    table.append(result0.estimate);

    double area = result0.estimate;
    double res_ext = result0.estimate;
    double err_ext = F64_MAX;
    double error_over_large_ranges = deltasum;
    double ertest = tolerance;
    const uint64_t max_iters = nint + (config.max_evals - nevals) / 50;

    for (uint64_t iteration = nint; iteration <= max_iters; ++iteration) {
        const SubRangeInfo info = ws.get();
        size_t current_level = info.level + 1;
        double center = (info.range.begin + info.range.end) * 0.5;
        Range r1{info.range.begin, center};
        Range r2{center, info.range.end};

        QKResult result1 = qk25(f, r1);
        QKResult result2 = qk25(f, r2);
        nevals += 50;

        if (result1.estimate != result1.estimate || result2.estimate != result2.estimate) {
            error = ST_NAN_VALUE_ENCOUNTERED;
            break;
        }

        double area12 = result1.estimate + result2.estimate;
        double error12 = result1.delta + result2.delta;
        double last_e_i = info.delta;

        deltasum += error12 - info.delta;
        area += area12 - info.estimate;
        double tolerance = config.tol.to_abs(std::fabs(area));

        if (result1.asc != result1.delta && result2.asc != result2.delta) {
            if (std::fabs(info.estimate - area12) <= 1e-5 * std::fabs(area12) &&
                error12 >= 0.99 * info.delta) {
                if (!extrapolate)
                    roundoff_type1 += 1;
                else
                    roundoff_type2 += 1;
            }
            if (iteration > 10 && error12 > info.delta) roundoff_type3 += 1;
        }

        if (roundoff_type1 + roundoff_type2 >= 10 || roundoff_type3 >= 20) error = ST_ROUNDOFF_ERROR;
        if (roundoff_type2 >= 5) error2 = true;

        if (subrange_too_small(r1.begin, r1.end, r2.end)) error = ST_SUBRANGE_TOO_SMALL;

        if (deltasum <= tolerance) {
            return finish(ws.sum_results() - info.estimate + result1.estimate + result2.estimate,
                          deltasum, nevals, error);
        }

        ws.update(SubRangeInfo{r1, result1.estimate, result1.delta, current_level},
                  SubRangeInfo{r2, result2.estimate, result2.delta, current_level});

        if (error != ST_NONE) break;

        // qagp.rs:259-262 (usize arithmetic: max_iters >= nint >= 1 so no wrap)
        if (iteration >= max_iters - 1) {
            error = ST_INSUFFICIENT_ITERATION;
            break;
        }

        if (disallow_extrapolation) continue;

        error_over_large_ranges += -last_e_i;
        if (current_level < ws.maximum_level) error_over_large_ranges += error12;

        if (!extrapolate) {
            if (ws.get().level < ws.maximum_level) continue;
            extrapolate = true;
            ws.nrmax = 1;
        }

        if (!error2 && error_over_large_ranges > ertest && ws.increase_nrmax()) continue;

        table.append(area);
        if (table.n < 3) {
            ws.reset_nrmax();
            extrapolate = false;
            error_over_large_ranges = deltasum;
            continue;
        }

        double reseps = 0.0, abseps = 0.0;
        table.qelg(reseps, abseps);
        ktmin += 1;

        if (ktmin > 5 && err_ext < 0.001 * deltasum) error = ST_ROUNDOFF_ERROR;

        if (abseps < err_ext) {
            ktmin = 0;
            err_ext = abseps;
            res_ext = reseps;
            correc = error_over_large_ranges;
            ertest = config.tol.to_abs(std::fabs(reseps));
            if (err_ext <= ertest) break;
        }

        if (table.n == 1) disallow_extrapolation = true;

        if (error != ST_NONE) break;

        ws.reset_nrmax();
        extrapolate = false;
        error_over_large_ranges = deltasum;
    }

    if (err_ext == F64_MAX) return finish(ws.sum_results(), deltasum, nevals, error);

    if (error != ST_NONE || error2) {
        if (error2) err_ext += correc;
        if (error == ST_NONE) error = ST_ROUNDOFF_ERROR;

        if (res_ext != 0. && area != 0.) {
            if (err_ext / std::fabs(res_ext) > deltasum / std::fabs(area))
                return finish(ws.sum_results(), deltasum, nevals, error);
        } else if (err_ext > deltasum) {
            return finish(ws.sum_results(), deltasum, nevals, error);
        } else if (area == 0.0) {
            return finish(res_ext, err_ext, nevals, error);
        }
    }

    bool positive_integrand = test_positivity(result0.estimate, result0.absvalue);
    if (!positive_integrand && rmax(std::fabs(res_ext), std::fabs(area)) < 0.01 * result0.absvalue)
        return finish(res_ext, err_ext, nevals, error);

    double ratio = res_ext / area;
    if ((ratio < 0.01 || ratio > 100. || deltasum > std::fabs(area)) && error == ST_NONE)
        error = ST_DIVERGENT;

    return finish(res_ext, err_ext, nevals, error);
}

// single/algorithm/qagp.rs:42-64
template <class F>
Result qagp(F& f, const Range& range, const Config& config, WorkSpace& ws) {
    bool transform = !std::isfinite(range.begin) || !std::isfinite(range.end);
    IntegrandWrapper<F> w{f, transform};
    Range r = transform ? transform_range(range) : range;
    return qagp_impl(w, r, config, transform, ws);
}

enum Algo : int { ALGO_AUTO = 0, ALGO_QAGS = 1, ALGO_QAGP = 2, ALGO_QAG = 3 };

// single/algorithm/auto.rs:7-35 -- AUTO owns one workspace per algorithm.
struct Workspaces {
    WorkSpace qags_ws;
    WorkSpace qagp_ws;
};

template <class F>
Result integrate(int algo, F& f, const Range& range, const Config& config, Workspaces& w) {
    if (algo == ALGO_QAGS) return qags(f, range, config, w.qags_ws);
    if (algo == ALGO_QAGP) return qagp(f, range, config, w.qagp_ws);
    if (algo == ALGO_QAG) return qag(f, range, config, w.qags_ws);  // (shares the QAGS slot: one list)
    if (config.npoints == 0) return qags(f, range, config, w.qags_ws);
    return qagp(f, range, config, w.qagp_ws);
}

// ---------------------------------------------------------------------------
// Nested 2-D drivers.
// ---------------------------------------------------------------------------
struct Config2 {
    Tolerance tol;
    uint64_t max_evals;
    const double* points_xy;  // interleaved (x0,y0,x1,y1,...)
    size_t npoints;
};

// double/algorithm/qags.rs:46-97.  F2(x,y) -> f64, YR(x) -> Range.
template <class F2, class YR>
Result qags2(F2& f, const Range& xrange, YR& yrange, const Config2& config) {
    Config config1{config.tol, config.max_evals / 17, nullptr, 0};
    const Config config2 = config1;

    WorkSpace inner_ws;  // re-used by every inner call: capacity persists (qags.rs:60-61)
    uint32_t error = ST_NONE;
    uint64_t nevals = 0;

    auto integrand = [&](double x) -> double {
        auto integrand2 = [&](double y) -> double { return f(x, y); };
        // usize arithmetic with overflow-checks=false (root Cargo.toml profile) wraps
        config1.max_evals = (error != ST_NONE) ? 0 : (uint64_t)(config.max_evals - nevals);
        Range yr = yrange(x);
        Result result = qags(integrand2, yr, config1, inner_ws);
        if (result.status != ST_NONE) {
            if (error == ST_NONE) error = result.status;
            nevals += result.nevals;
            return std::numeric_limits<double>::quiet_NaN();
        }
        nevals += result.nevals;
        return result.estimate;
    };

    WorkSpace outer_ws;
    Result result = qags(integrand, xrange, config2, outer_ws);
    result.nevals = nevals;
    if (error != ST_NONE) result.status = error;
    return result;
}

// double/algorithm/qagp.rs:70-139 (incl. the double transform of outer points when
// the x-range is infinite: here :100-101 and again in make_sorted_points).
template <class F2, class YR>
Result qagp2(F2& f, const Range& xrange, YR& yrange, const Config2& config) {
    std::vector<double> inner_pts, outer_pts;
    bool xtransform = !std::isfinite(xrange.begin) || !std::isfinite(xrange.end);
    for (size_t k = 0; k < config.npoints; ++k) {
        double x = config.points_xy[2 * k], y = config.points_xy[2 * k + 1];
        outer_pts.push_back(xtransform ? transform_point(x) : x);
        inner_pts.push_back(y);
    }
    Config inner_config{config.tol, 0, inner_pts.data(), inner_pts.size()};
    Config outer_config{config.tol, config.max_evals / 17, outer_pts.data(), outer_pts.size()};

    WorkSpace inner_ws;
    uint32_t error = ST_NONE;
    uint64_t nevals = 0;

    auto integrand = [&](double x) -> double {
        auto integrand2 = [&](double y) -> double { return f(x, y); };
        inner_config.max_evals = (error != ST_NONE) ? 0 : (uint64_t)(config.max_evals - nevals);
        Range yr = yrange(x);
        Result result = qagp(integrand2, yr, inner_config, inner_ws);
        if (result.status != ST_NONE) {
            if (error == ST_NONE) error = result.status;
            nevals += result.nevals;
            return std::numeric_limits<double>::quiet_NaN();
        }
        nevals += result.nevals;
        return result.estimate;
    };

    WorkSpace outer_ws;
    Result result = qagp(integrand, xrange, outer_config, outer_ws);
    result.nevals = nevals;
    if (error != ST_NONE) result.status = error;
    return result;
}

// double/algorithm/qag.rs:46-100 (deprecated QAG2): the QAGS2 nest with QAG inside and outside.
template <class F2, class YR>
Result qag2(F2& f, const Range& xrange, YR& yrange, const Config2& config) {
    Config config1{config.tol, config.max_evals / 17, nullptr, 0};
    const Config config2 = config1;

    WorkSpace inner_ws;
    uint32_t error = ST_NONE;
    uint64_t nevals = 0;

    auto integrand = [&](double x) -> double {
        auto integrand2 = [&](double y) -> double { return f(x, y); };
        config1.max_evals = (error != ST_NONE) ? 0 : (uint64_t)(config.max_evals - nevals);  // :63-67
        Range yr = yrange(x);
        Result result = qag(integrand2, yr, config1, inner_ws);
        if (result.status != ST_NONE) {  // :71-76
            if (error == ST_NONE) error = result.status;
            nevals += result.nevals;
            return std::numeric_limits<double>::quiet_NaN();
        }
        nevals += result.nevals;
        return result.estimate;
    };

    WorkSpace outer_ws;
    Result result = qag(integrand, xrange, config2, outer_ws);  // :85
    result.nevals = nevals;
    if (error != ST_NONE) result.status = error;
    return result;
}

// double/algorithm/auto.rs:15-36
template <class F2, class YR>
Result integrate2(int algo, F2& f, const Range& xrange, YR& yrange, const Config2& config) {
    if (algo == ALGO_QAGS) return qags2(f, xrange, yrange, config);
    if (algo == ALGO_QAGP) return qagp2(f, xrange, yrange, config);
    if (algo == ALGO_QAG) return qag2(f, xrange, yrange, config);
    if (config.npoints == 0) return qags2(f, xrange, yrange, config);
    return qagp2(f, xrange, yrange, config);
}

}  // namespace gkqo
