// gkquad_b200.hpp -- C++ host-side mirror of the gkquad crate's public surface for the hot path,
// on top of the C ABI (include/gkquad_b200.h).  Header-only; link with -lgkquad_b200.
//
// The reference is a Rust crate and this image has no Rust toolchain, so the compiled host side
// above the C ABI is C++ (the Rust crate source that binds the same ABI is in ../rust/, uncompiled).
// Names, argument meaning and error behaviour follow the reference:
//
//   gkquad::Tolerance, RuntimeError, Solution, IntegrationResult        (common.rs, error.rs)
//   gkquad::single::{Range, Points, IntegrationConfig, Integrator, integral,
//                    integral_with_config, QAGS, QAGP, AUTO}             (single/*.rs)
//   gkquad::double_::{Rectangle, DynamicY, IntegrationConfig2, Integrator2, integral2,
//                     QAGS2, QAGP2, AUTO2}                                (double/*.rs)
//
// The one difference: an integrand is a DeviceIntegrand (a functor of the device registry plus a
// batch of parameters) instead of a closure, and run() integrates the whole batch on the GPU.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/gkquad_b200.h"

namespace gkquad {

// error.rs:140-178
enum class RuntimeError : uint32_t {
    InsufficientIteration = 1,
    RoundoffError = 2,
    SubrangeTooSmall = 3,
    Divergent = 4,
    NanValueEncountered = 5,
};

// common.rs:9-49
struct Tolerance {
    gkq_tolerance raw;
    static Tolerance Absolute(double x) { return {{GKQ_TOL_ABSOLUTE, 0, x, 0.0}}; }
    static Tolerance Relative(double y) { return {{GKQ_TOL_RELATIVE, 0, 0.0, y}}; }
    static Tolerance AbsOrRel(double x, double y) { return {{GKQ_TOL_ABS_OR_REL, 0, x, y}}; }
    static Tolerance AbsAndRel(double x, double y) { return {{GKQ_TOL_ABS_AND_REL, 0, x, y}}; }
    static Tolerance Default() { return AbsOrRel(1.49e-8, 1.49e-8); }  // common.rs:44-49
    bool contains_nan() const {  // common.rs:22-31
        switch (raw.kind) {
            case GKQ_TOL_ABSOLUTE: return std::isnan(raw.abs_tol);
            case GKQ_TOL_RELATIVE: return std::isnan(raw.rel_tol);
            default: return std::isnan(raw.abs_tol) || std::isnan(raw.rel_tol);
        }
    }
    double to_abs(double value) const {  // common.rs:34-41
        switch (raw.kind) {
            case GKQ_TOL_ABSOLUTE: return raw.abs_tol;
            case GKQ_TOL_RELATIVE: return value * raw.rel_tol;
            case GKQ_TOL_ABS_OR_REL: return std::fmax(raw.abs_tol, value * raw.rel_tol);
            default: return std::fmin(raw.abs_tol, value * raw.rel_tol);
        }
    }
};

// common.rs:213-231
struct Solution {
    double estimate = 0.0;
    double delta = std::numeric_limits<double>::max();
    std::size_t nevals = 0;
};

// ValueWithError<Solution, RuntimeError> (common.rs:60-209): the partial value is always present.
struct IntegrationResult {
    Solution value;
    std::optional<RuntimeError> error;
    bool has_err() const { return error.has_value(); }
    const Solution& unwrap() const {  // panics in the reference (common.rs:193-208)
        if (error) throw std::runtime_error("called `unwrap()` on a result with an error");
        return value;
    }
    const Solution& unwrap_unchecked() const { return value; }
};

// Structure-of-arrays results of a batch.
struct BatchResult {
    std::vector<double> estimate, delta;
    std::vector<uint32_t> nevals, status;
    std::size_t size() const { return estimate.size(); }
    IntegrationResult operator[](std::size_t i) const {
        IntegrationResult r;
        r.value = {estimate[i], delta[i], nevals[i]};
        if (status[i]) r.error = static_cast<RuntimeError>(status[i]);
        return r;
    }
};

class GkqError : public std::runtime_error {
public:
    GkqError(int code, const std::string& msg) : std::runtime_error(msg), code(code) {}
    int code;
};

// One engine (gkq_handle_t) per device; the counterpart of the state an Integrator owns.
class Engine {
public:
    // stream_priority: CUDA priority of the streams the handle owns (0 = default, negative = scheduled
    // first; gkq_create_prio) -- for the lane of a pipeline whose batches have to leave the GPU first
    explicit Engine(int device = 0, int stream_priority = 0) {
        int rc = gkq_create_prio(device, stream_priority, &h_);
        if (rc != GKQ_OK) throw GkqError(rc, gkq_last_error(nullptr));
    }
    ~Engine() { gkq_destroy(h_); }
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;
    gkq_handle_t handle() const { return h_; }
    void check(int rc) const {
        if (rc != GKQ_OK) throw GkqError(rc, gkq_last_error(h_));
    }
    static std::shared_ptr<Engine> shared(int device = 0) {
        static std::shared_ptr<Engine> e = std::make_shared<Engine>(device);
        return e;
    }

private:
    gkq_handle_t h_ = nullptr;
};

namespace single {

// single/common.rs:18-96
struct Range {
    double begin, end;
    Range(double b, double e) : begin(b), end(e) {
        if (std::isnan(b) || std::isnan(e))
            throw std::invalid_argument("cannot create Range object from Range which contains NaN value.");
    }
    static Range unbounded() { return {-INFINITY, INFINITY}; }  // `..`
    static Range from(double b) { return {b, INFINITY}; }       // `b..`
    static Range to(double e) { return {-INFINITY, e}; }        // `..e`
};
using Points = std::vector<double>;  // SmallVec<[f64; 8]> (single/common.rs:12)

// single/common.rs:108-126
struct IntegrationConfig {
    Tolerance tolerance = Tolerance::Default();
    std::size_t max_evals = 2000;
    Points points;
};

// Stands where the reference takes `F: Integrand` (single/common.rs:129-145).  `params` is
// structure-of-arrays [nparams][n]; n == 0 means one integral with params[k][0] (or none).
struct DeviceIntegrand {
    std::string functor;
    std::vector<std::vector<double>> params;
    std::size_t batch() const { return params.empty() ? 1 : params[0].size(); }
};

// trait Algorithm<F> (single/algorithm/mod.rs:16-23)
class Algorithm {
public:
    explicit Algorithm(gkq_algo a, std::size_t workspace_capacity = 0) : algo_(a), ws_cap_(workspace_capacity) {}
    BatchResult integrate(const DeviceIntegrand& f, const Range& range, const IntegrationConfig& cfg,
                          std::shared_ptr<Engine> eng = nullptr) const {
        if (!eng) eng = Engine::shared();
        const int fid = gkq_functor_lookup(1, f.functor.c_str());
        if (fid < 0) throw GkqError(fid, "unknown functor " + f.functor);
        const std::size_t n = f.batch();
        std::vector<double> flat;
        for (const auto& row : f.params) {
            if (row.size() != n) throw std::invalid_argument("ragged parameter rows");
            flat.insert(flat.end(), row.begin(), row.end());
        }
        gkq_config c;
        gkq_config_default(1, &c);
        c.algo = algo_;
        c.tol = cfg.tolerance.raw;
        c.max_evals = cfg.max_evals;
        c.workspace_capacity = ws_cap_;
        c.points = cfg.points.empty() ? nullptr : cfg.points.data();
        c.npoints = (int64_t)cfg.points.size();
        BatchResult r;
        r.estimate.resize(n);
        r.delta.resize(n);
        r.nevals.resize(n);
        r.status.resize(n);
        eng->check(gkq_integrate_batch_host(eng->handle(), &c, fid, (int64_t)n, &range.begin, &range.end, 0,
                                            flat.empty() ? nullptr : flat.data(), flat.empty() ? 0 : (int64_t)n,
                                            nullptr, nullptr, r.estimate.data(), r.delta.data(),
                                            r.nevals.data(), r.status.data()));
        return r;
    }

private:
    gkq_algo algo_;
    std::size_t ws_cap_;
};
struct QAGS : Algorithm { explicit QAGS(std::size_t ws_cap = 0) : Algorithm(GKQ_ALGO_QAGS, ws_cap) {} };  // qags.rs
struct QAGP : Algorithm { explicit QAGP(std::size_t ws_cap = 0) : Algorithm(GKQ_ALGO_QAGP, ws_cap) {} };  // qagp.rs
struct AUTO : Algorithm { AUTO() : Algorithm(GKQ_ALGO_AUTO) {} };                                         // auto.rs
struct QAG : Algorithm { explicit QAG(std::size_t ws_cap = 0) : Algorithm(GKQ_ALGO_QAG, ws_cap) {} };     // qag.rs (deprecated)

// single/integrator.rs:31-106
class Integrator {
public:
    explicit Integrator(DeviceIntegrand f) : f_(std::move(f)), alg_(AUTO()) {}
    Integrator& algorithm(const Algorithm& a) { alg_ = a; return *this; }
    Integrator& tolerance(const Tolerance& t) {
        if (t.contains_nan()) throw std::invalid_argument("Tolerance must not contain NAN.");  // :71
        cfg_.tolerance = t;
        return *this;
    }
    Integrator& max_evals(std::size_t m) { cfg_.max_evals = m; return *this; }
    Integrator& points(const Points& p) {
        for (double x : p)
            if (std::isnan(x)) throw std::invalid_argument("cannot include NAN value in singular points");  // :86-90
        cfg_.points = p;
        return *this;
    }
    BatchResult run(const Range& r) const { return alg_.integrate(f_, r, cfg_); }

private:
    DeviceIntegrand f_;
    Algorithm alg_;
    IntegrationConfig cfg_;
};

// single/integral.rs:17-19: fresh QAGS + default config
inline BatchResult integral(const DeviceIntegrand& f, const Range& r) { return QAGS().integrate(f, r, IntegrationConfig()); }
// single/integral.rs:25-31: AUTO
inline BatchResult integral_with_config(const DeviceIntegrand& f, const Range& r, const IntegrationConfig& c) {
    return AUTO().integrate(f, r, c);
}

}  // namespace single

namespace double_ {
using single::Range;

// double/range.rs:77-112 (a Rectangle is DynamicY with the "const" y-range functor, qags.rs:18-27)
struct DynamicY {
    Range xrange;
    std::string yrange;            // registry name: "const" | "zero_to_x" | "circle"
    std::vector<double> yparams;   // its parameters
    bool swap_xy = false;          // built by DynamicX(): the roles of x and y are exchanged
};
// double/range.rs:38-75: y in [y1, y2], x in xrange(y).  The algorithms swap the integrand's
// arguments and the point coordinates (double/algorithm/qags.rs:29-44, qagp.rs:34-53).
inline DynamicY DynamicX(const std::string& xrange, double y1, double y2, std::vector<double> xparams = {}) {
    DynamicY r{Range(y1, y2), xrange, std::move(xparams)};
    r.swap_xy = true;
    return r;
}
inline DynamicY Rectangle(double x1, double x2, double y1, double y2) {
    if (!(std::isfinite(x1) && std::isfinite(x2) && std::isfinite(y1) && std::isfinite(y2)))
        throw std::invalid_argument("Infinite interval in 2-dimension is not already supported.");  // range.rs:20-22
    return {Range(x1, x2), "const", {y1, y2}};
}
// From<(R1, R2)> for Rectangle (range.rs:27-34) allows infinite sides
inline DynamicY RectangleFrom(const Range& x, const Range& y) { return {x, "const", {y.begin, y.end}}; }

// double/common.rs:14-32
struct IntegrationConfig2 {
    Tolerance tolerance = Tolerance::Default();
    std::size_t max_evals = 100000;
    std::vector<std::pair<double, double>> points;
};
struct DeviceIntegrand2 {
    std::string functor;
    std::vector<std::vector<double>> params;
    std::size_t batch() const { return params.empty() ? 1 : params[0].size(); }
};

// trait Algorithm2<F, R> (double/algorithm/mod.rs:7-10)
class Algorithm2 {
public:
    explicit Algorithm2(gkq_algo a) : algo_(a) {}
    BatchResult integrate(const DeviceIntegrand2& f, const DynamicY& range, const IntegrationConfig2& cfg,
                          std::shared_ptr<Engine> eng = nullptr) const {
        if (!eng) eng = Engine::shared();
        const int fid = gkq_functor_lookup(2, f.functor.c_str());
        const int yid = gkq_functor_lookup(0, range.yrange.c_str());
        if (fid < 0 || yid < 0) throw GkqError(GKQ_ERR_INVALID_ARGUMENT, "unknown functor");
        const std::size_t n = f.batch();
        std::vector<double> flat, pts;
        for (const auto& row : f.params) flat.insert(flat.end(), row.begin(), row.end());
        for (const auto& p : cfg.points) {
            pts.push_back(range.swap_xy ? p.second : p.first);  // qagp.rs:49
            pts.push_back(range.swap_xy ? p.first : p.second);
        }
        gkq_config c;
        gkq_config_default(2, &c);
        c.flags = range.swap_xy ? GKQ_FLAG_SWAP_XY : 0;
        c.algo = algo_;
        c.tol = cfg.tolerance.raw;
        c.max_evals = cfg.max_evals;
        c.points = pts.empty() ? nullptr : pts.data();
        c.npoints = (int64_t)cfg.points.size();
        BatchResult r;
        r.estimate.resize(n);
        r.delta.resize(n);
        r.nevals.resize(n);
        r.status.resize(n);
        eng->check(gkq_integrate2_batch_host(
            eng->handle(), &c, fid, yid, (int64_t)n, &range.xrange.begin, &range.xrange.end, 0,
            flat.empty() ? nullptr : flat.data(), flat.empty() ? 0 : (int64_t)n,
            range.yparams.empty() ? nullptr : range.yparams.data(), 0, nullptr, nullptr, r.estimate.data(),
            r.delta.data(), r.nevals.data(), r.status.data()));
        return r;
    }

private:
    gkq_algo algo_;
};
struct QAGS2 : Algorithm2 { QAGS2() : Algorithm2(GKQ_ALGO_QAGS) {} };
struct QAGP2 : Algorithm2 { QAGP2() : Algorithm2(GKQ_ALGO_QAGP) {} };
struct AUTO2 : Algorithm2 { AUTO2() : Algorithm2(GKQ_ALGO_AUTO) {} };
struct QAG2 : Algorithm2 { QAG2() : Algorithm2(GKQ_ALGO_QAG) {} };  // double/algorithm/qag.rs (deprecated)

// double/integral.rs:9-16
inline BatchResult integral2(const DeviceIntegrand2& f, const DynamicY& r) { return QAGS2().integrate(f, r, IntegrationConfig2()); }
inline BatchResult integral2_with_config(const DeviceIntegrand2& f, const DynamicY& r, const IntegrationConfig2& c) {
    return AUTO2().integrate(f, r, c);
}

}  // namespace double_
}  // namespace gkquad
