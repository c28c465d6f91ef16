// oracle_capi.cpp -- C entry points of the CPU oracle (ctypes-loadable).  TEST
// INFRASTRUCTURE ONLY: used by tests/, __graft_entry__.smoke() and bench.py's CPU
// baseline legs; never linked into libgkquad_b200.so.  See gkq_oracle.hpp.
//
// Batch semantics mirror the engine's C ABI (include/gkquad_b200.h): every integral gets
// a fresh workspace (reference: QAGS::new()/QAGP::new() per call, single/integral.rs:17-31)
// unless ws_capacity > 0, which models WorkSpace::with_capacity(ws_capacity) as used by the
// reference's own tests (tests/algorithms_single.rs:53).
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "functors.hpp"
#include "gkq_oracle.hpp"

using namespace gkqo;

namespace {

// Dynamic chunked parallel-for over [0, n) on `nt` host threads (std::thread; the
// equivalent of rayon's par_iter over cloned Integrators named by BASELINE.json).
template <class Body>
void parallel_chunks(int64_t n, int64_t chunk, int nt, Body body) {
    if (nt <= 1 || n <= chunk) {
        body((int64_t)0, n);
        return;
    }
    std::atomic<int64_t> next{0};
    auto worker = [&]() {
        for (;;) {
            int64_t lo = next.fetch_add(chunk);
            if (lo >= n) break;
            int64_t hi = lo + chunk < n ? lo + chunk : n;
            body(lo, hi);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
}

int default_threads() {
    unsigned h = std::thread::hardware_concurrency();
    return h ? (int)h : 1;
}

struct BatchArgs {
    int algo;
    int64_t n;
    const double* a;
    const double* b;
    int64_t range_stride;   // 1 = per-integral ranges, 0 = one shared range
    const double* params;   // SoA: params[k*param_stride + i]; param_stride 0 = shared
    int64_t param_stride;
    const int64_t* pts_offsets;  // n+1 entries or NULL (then all integrals share pts[0..npts))
    const double* pts;
    int64_t npts;
    Tolerance tol;
    uint64_t max_evals;
    uint64_t ws_capacity;
    double* out_estimate;
    double* out_delta;
    uint32_t* out_nevals;
    uint32_t* out_status;
    int nthreads;
    // per-integral tolerance / budget (an Integrator configured anew for every integral of a sweep,
    // single/integrator.rs:69-84); NULL = the batch value
    const double* abs_tol_each = nullptr;
    const double* rel_tol_each = nullptr;
    const uint64_t* max_evals_each = nullptr;
};

template <int ID>
void run_batch(const BatchArgs& A) {
    const int np = F1_TABLE[ID].nparams;
    int nt = A.nthreads > 0 ? A.nthreads : default_threads();
    parallel_chunks(A.n, 1024, nt, [&](int64_t lo, int64_t hi) {
        Workspaces w;
        for (int64_t i = lo; i < hi; ++i) {
            double p[4] = {0, 0, 0, 0};
            for (int k = 0; k < np; ++k)
                p[k] = A.param_stride ? A.params[(int64_t)k * A.param_stride + i] : A.params[k];
            F1<ID> f{p};
            Range r{A.a[i * A.range_stride], A.b[i * A.range_stride]};
            Config cfg;
            cfg.tol = A.tol;
            cfg.max_evals = A.max_evals;
            if (A.abs_tol_each) cfg.tol.abs_ = A.abs_tol_each[i];
            if (A.rel_tol_each) cfg.tol.rel_ = A.rel_tol_each[i];
            if (A.max_evals_each) cfg.max_evals = A.max_evals_each[i] < A.max_evals ? A.max_evals_each[i] : A.max_evals;
            if (A.pts_offsets) {
                cfg.points = A.pts + A.pts_offsets[i];
                cfg.npoints = (size_t)(A.pts_offsets[i + 1] - A.pts_offsets[i]);
            } else {
                cfg.points = A.pts;
                cfg.npoints = (size_t)A.npts;
            }
            w.qags_ws.reset_fresh((size_t)A.ws_capacity);
            w.qagp_ws.reset_fresh((size_t)A.ws_capacity);
            Result res = integrate(A.algo, f, r, cfg, w);
            A.out_estimate[i] = res.estimate;
            A.out_delta[i] = res.delta;
            A.out_nevals[i] = (uint32_t)res.nevals;
            A.out_status[i] = res.status;
        }
    });
}

template <int ID>
struct Dispatch1 {
    static int go(int id, const BatchArgs& A) {
        if (id == ID) {
            run_batch<ID>(A);
            return 0;
        }
        return Dispatch1<ID + 1>::go(id, A);
    }
};
template <>
struct Dispatch1<F1_COUNT> {
    static int go(int, const BatchArgs&) { return -1; }
};

struct YRange {
    int id;
    const double* q;
    inline Range operator()(double x) const {
        switch (id) {
            case YR_CONST: return Range{q[0], q[1]};
            case YR_ZERO_TO_X: return Range{0.0, x};
            default: {
                double ymax = std::sqrt(q[0] - x * x);
                return Range{-ymax, ymax};
            }
        }
    }
};

struct Batch2Args {
    int algo;
    int yrange_id;
    int64_t n;
    const double* x0;
    const double* x1;
    int64_t range_stride;
    const double* fparams;
    int64_t fparam_stride;
    const double* yparams;
    int64_t yparam_stride;
    const int64_t* pts_offsets;  // in points (pairs), n+1 entries, or NULL
    const double* pts_xy;        // interleaved pairs
    int64_t npts;
    Tolerance tol;
    uint64_t max_evals;
    double* out_estimate;
    double* out_delta;
    uint32_t* out_nevals;
    uint32_t* out_status;
    int nthreads;
    int swap_xy = 0;  // DynamicX: the algorithms see g(x, y) = f(y, x) (double/algorithm/qags.rs:36-37)
    // per-integral tolerance values / total budget (an Integrator2 configured anew per integral)
    const double* abs_tol_each = nullptr;
    const double* rel_tol_each = nullptr;
    const uint64_t* max_evals_each = nullptr;
};

// double/algorithm/qags.rs:36-37, qagp.rs:41-42: `let mut g = |x, y| f.apply((y, x));`
template <class F>
struct SwappedF2 {
    F& f;
    double operator()(double x, double y) { return f(y, x); }
};

template <int ID>
void run_batch2(const Batch2Args& A) {
    const int np = F2_TABLE[ID].nparams;
    const int nq = YR_TABLE[A.yrange_id].nparams;
    int nt = A.nthreads > 0 ? A.nthreads : default_threads();
    parallel_chunks(A.n, 16, nt, [&](int64_t lo, int64_t hi) {
      for (int64_t i = lo; i < hi; ++i) {
        double p[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
        for (int k = 0; k < np; ++k)
            p[k] = A.fparam_stride ? A.fparams[(int64_t)k * A.fparam_stride + i] : A.fparams[k];
        for (int k = 0; k < nq; ++k)
            q[k] = A.yparam_stride ? A.yparams[(int64_t)k * A.yparam_stride + i] : A.yparams[k];
        F2<ID> f{p};
        YRange yr{A.yrange_id, q};
        Range xr{A.x0[i * A.range_stride], A.x1[i * A.range_stride]};
        Config2 cfg;
        cfg.tol = A.tol;
        cfg.max_evals = A.max_evals;
        if (A.abs_tol_each) cfg.tol.abs_ = A.abs_tol_each[i];
        if (A.rel_tol_each) cfg.tol.rel_ = A.rel_tol_each[i];
        if (A.max_evals_each) cfg.max_evals = A.max_evals_each[i] < A.max_evals ? A.max_evals_each[i] : A.max_evals;
        if (A.pts_offsets) {
            cfg.points_xy = A.pts_xy + 2 * A.pts_offsets[i];
            cfg.npoints = (size_t)(A.pts_offsets[i + 1] - A.pts_offsets[i]);
        } else {
            cfg.points_xy = A.pts_xy;
            cfg.npoints = (size_t)A.npts;
        }
        Result res;
        if (A.swap_xy) {
            SwappedF2<F2<ID>> g{f};
            res = integrate2(A.algo, g, xr, yr, cfg);
        } else {
            res = integrate2(A.algo, f, xr, yr, cfg);
        }
        A.out_estimate[i] = res.estimate;
        A.out_delta[i] = res.delta;
        A.out_nevals[i] = (uint32_t)res.nevals;
        A.out_status[i] = res.status;
      }
    });
}

template <int ID>
struct Dispatch2 {
    static int go(int id, const Batch2Args& A) {
        if (id == ID) {
            run_batch2<ID>(A);
            return 0;
        }
        return Dispatch2<ID + 1>::go(id, A);
    }
};
template <>
struct Dispatch2<F2_COUNT> {
    static int go(int, const Batch2Args&) { return -1; }
};

// rule-level: qk17 ... qk57 on one range
template <int ID>
struct DispatchQK {
    static int go(int id, int rule, double a, double b, const double* p, double* out4) {
        if (id == ID) {
            F1<ID> f{p};
            IntegrandWrapper<F1<ID>> w{f, false};
            Range r{a, b};
            QKResult q;
            switch (rule) {
                case 17: q = qk17(w, r); break;
                case 25: q = qk25(w, r); break;
                case 33: q = qk33(w, r); break;
                case 41: q = qk41(w, r); break;
                case 49: q = qk49(w, r); break;
                default: q = qk57(w, r); break;
            }
            out4[0] = q.estimate;
            out4[1] = q.delta;
            out4[2] = q.absvalue;
            out4[3] = q.asc;
            return 0;
        }
        return DispatchQK<ID + 1>::go(id, rule, a, b, p, out4);
    }
};
template <>
struct DispatchQK<F1_COUNT> {
    static int go(int, int, double, double, const double*, double*) { return -1; }
};

// single integral with the workspace exposed (for the `order` golden vectors)
template <int ID>
struct DispatchOne {
    static int go(int id, int algo, double a, double b, const double* p, const Config& cfg,
                  uint64_t ws_capacity, double* out2, uint32_t* out_nevals, uint32_t* out_status,
                  int64_t* order_out, int64_t order_cap, int64_t* order_len) {
        if (id == ID) {
            F1<ID> f{p};
            Workspaces w;
            w.qags_ws = WorkSpace((size_t)ws_capacity);
            w.qagp_ws = WorkSpace((size_t)ws_capacity);
            Range r{a, b};
            Result res = integrate(algo, f, r, cfg, w);
            out2[0] = res.estimate;
            out2[1] = res.delta;
            *out_nevals = (uint32_t)res.nevals;
            *out_status = res.status;
            bool use_qags = (algo == ALGO_QAGS) || (algo == ALGO_QAG) || (algo == ALGO_AUTO && cfg.npoints == 0);
            const WorkSpace& ws = use_qags ? w.qags_ws : w.qagp_ws;
            *order_len = (int64_t)ws.order.size();
            for (int64_t k = 0; k < (int64_t)ws.order.size() && k < order_cap; ++k)
                order_out[k] = (int64_t)ws.order[(size_t)k];
            return 0;
        }
        return DispatchOne<ID + 1>::go(id, algo, a, b, p, cfg, ws_capacity, out2, out_nevals,
                                       out_status, order_out, order_cap, order_len);
    }
};
template <>
struct DispatchOne<F1_COUNT> {
    static int go(int, int, double, double, const double*, const Config&, uint64_t, double*,
                  uint32_t*, uint32_t*, int64_t*, int64_t, int64_t*) {
        return -1;
    }
};

}  // namespace

extern "C" {

int gkqo_num_threads() { return default_threads(); }

int gkqo_functor_info(int dim, int id, const char** name, int* nparams) {
    const FunctorInfo* t = nullptr;
    int count = 0;
    if (dim == 1) {
        t = F1_TABLE;
        count = F1_COUNT;
    } else if (dim == 2) {
        t = F2_TABLE;
        count = F2_COUNT;
    } else if (dim == 0) {
        t = YR_TABLE;
        count = YR_COUNT;
    }
    if (!t || id < 0 || id >= count) return -1;
    *name = t[id].name;
    *nparams = t[id].nparams;
    return 0;
}

int gkqo_integrate_batch(int algo, int functor_id, int64_t n, const double* a, const double* b,
                         int64_t range_stride, const double* params, int64_t param_stride,
                         const int64_t* pts_offsets, const double* pts, int64_t npts, int tol_kind,
                         double tol_abs, double tol_rel, uint64_t max_evals, uint64_t ws_capacity,
                         double* out_estimate, double* out_delta, uint32_t* out_nevals,
                         uint32_t* out_status, int nthreads) {
    BatchArgs A{algo,        n,          a,         b,          range_stride,
                params,      param_stride, pts_offsets, pts,      npts,
                Tolerance{tol_kind, tol_abs, tol_rel}, max_evals, ws_capacity,
                out_estimate, out_delta,  out_nevals, out_status, nthreads};
    return Dispatch1<0>::go(functor_id, A);
}

int gkqo_integrate_batch_ex(int algo, int functor_id, int64_t n, const double* a, const double* b,
                            int64_t range_stride, const double* params, int64_t param_stride,
                            const int64_t* pts_offsets, const double* pts, int64_t npts, int tol_kind,
                            double tol_abs, double tol_rel, uint64_t max_evals, uint64_t ws_capacity,
                            const double* abs_tol_each, const double* rel_tol_each,
                            const uint64_t* max_evals_each, double* out_estimate, double* out_delta,
                            uint32_t* out_nevals, uint32_t* out_status, int nthreads) {
    BatchArgs A{algo,        n,          a,         b,          range_stride,
                params,      param_stride, pts_offsets, pts,      npts,
                Tolerance{tol_kind, tol_abs, tol_rel}, max_evals, ws_capacity,
                out_estimate, out_delta,  out_nevals, out_status, nthreads};
    A.abs_tol_each = abs_tol_each;
    A.rel_tol_each = rel_tol_each;
    A.max_evals_each = max_evals_each;
    return Dispatch1<0>::go(functor_id, A);
}

int gkqo_integrate2_batch(int algo, int functor2_id, int yrange_id, int64_t n, const double* x0,
                          const double* x1, int64_t range_stride, const double* fparams,
                          int64_t fparam_stride, const double* yparams, int64_t yparam_stride,
                          const int64_t* pts_offsets, const double* pts_xy, int64_t npts,
                          int tol_kind, double tol_abs, double tol_rel, uint64_t max_evals,
                          double* out_estimate, double* out_delta, uint32_t* out_nevals,
                          uint32_t* out_status, int nthreads, int swap_xy, const double* abs_tol_each,
                          const double* rel_tol_each, const uint64_t* max_evals_each) {
    if (yrange_id < 0 || yrange_id >= YR_COUNT) return -1;
    Batch2Args A{algo,       yrange_id,     n,          x0,         x1,         range_stride,
                 fparams,    fparam_stride, yparams,    yparam_stride, pts_offsets, pts_xy,
                 npts,       Tolerance{tol_kind, tol_abs, tol_rel}, max_evals,
                 out_estimate, out_delta,   out_nevals, out_status, nthreads};
    A.swap_xy = swap_xy;
    A.abs_tol_each = abs_tol_each;
    A.rel_tol_each = rel_tol_each;
    A.max_evals_each = max_evals_each;
    return Dispatch2<0>::go(functor2_id, A);
}

int gkqo_qk(int rule, int functor_id, double a, double b, const double* params, double* out4) {
    if (rule != 17 && rule != 25 && rule != 33 && rule != 41 && rule != 49 && rule != 57) return -2;
    return DispatchQK<0>::go(functor_id, rule, a, b, params, out4);
}

int gkqo_integrate_one(int algo, int functor_id, double a, double b, const double* params,
                       const double* pts, int64_t npts, int tol_kind, double tol_abs,
                       double tol_rel, uint64_t max_evals, uint64_t ws_capacity, double* out2,
                       uint32_t* out_nevals, uint32_t* out_status, int64_t* order_out,
                       int64_t order_cap, int64_t* order_len) {
    Config cfg;
    cfg.tol = Tolerance{tol_kind, tol_abs, tol_rel};
    cfg.max_evals = max_evals;
    cfg.points = pts;
    cfg.npoints = (size_t)npts;
    return DispatchOne<0>::go(functor_id, algo, a, b, params, cfg, ws_capacity, out2, out_nevals,
                              out_status, order_out, order_cap, order_len);
}

}  // extern "C"
