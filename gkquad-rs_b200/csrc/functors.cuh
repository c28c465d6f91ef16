// functors.cuh -- the device integrand registry.
//
// The reference takes closures (trait Integrand, single/common.rs:129-145; Integrand2,
// double/common.rs:35-45; DynamicY's range closure, double/range.rs:77-112).  Closures cannot
// cross an FFI, so integrands are device functors compiled into the kernels as template
// arguments (a function-pointer table would block inlining) and selected by ID through
// gkq_functor_lookup / the `functor_id` argument of the batch calls.
//
// To add an integrand: add an id to the enum, a GKQ_F1 / GKQ_F2 body, a row in the table
// (name, nparams, flops) and rebuild.  Bodies use plain C arithmetic; with -fmad=false the
// polynomial/rational ones round exactly like the Rust closures they restate
// (/root/reference/gkquad/benches/*.rs, tests/common/functions.rs).
#pragma once
#include "gk_common.cuh"
#include "gk_math.cuh"

namespace gkq {

constexpr int MAX_PARAMS = 4;

struct FunctorInfo {
    const char* name;
    int nparams;
    // Algorithmic FP64 flops of one evaluation: counted from source
    // for polynomial/rational bodies; for transcendental bodies the DADD+DMUL+2*DFMA SASS
    // count of one evaluation (measured, see profiles/).
    double flops;
};

// ---- 1-D ------------------------------------------------------------------------------
enum F1Id : int {
    F1_SQUARE = 0, F1_RECIP_P, F1_LORENTZ_P, F1_EXPCOS_P, F1_GAUSS_P, F1_ABSPOW_P, F1_LOGABS_P,
    F1_F1, F1_F2, F1_F3, F1_F4, F1_F5, F1_F6, F1_EXP_P, F1_SQRT_SHIFT_P, F1_INV_ABS_SHIFT_P,
    F1_COS_P, F1_SIN_P, F1_SEMICIRCLE_P, F1_SINPOW_P, F1_COUNT
};

// HEAVY = body calls transcendental libdevice routines: the rule loop is then only partly
// unrolled to keep the instruction footprint inside the instruction cache.
template <int ID> struct F1;
// BODY computes f(x) with the math policy `m` (m.exp, m.cos, m.sin, m.log, m.pow):
//   eval(x, FastMath&)  branch-free sample; clears m.ok when an argument leaves the fast range
//                       (m.div / m.sqrt: IEEE division / square root without nvcc's slow-path branch)
//   operator()(x)       always-valid sample (SafeMath)
//   GKQ_F1P: the body only uses + - * / and m.exp / m.cos / m.sin, so it also instantiates on the
//   two-lane type D2 and a node pair is evaluated with interleaved chains (PAIRWISE).
#define GKQ_F1_IMPL(ID, HEAVY_, PAIR_, BODY)                                     \
    template <> struct F1<ID> {                                                  \
        static constexpr bool HEAVY = HEAVY_;                                    \
        static constexpr bool PAIRWISE = PAIR_;                                  \
        double p[MAX_PARAMS];                                                    \
        template <class T, class M> GKQ_DEV T eval(T x, M& m) const {            \
            (void)m;                                                             \
            BODY                                                                 \
        }                                                                        \
        GKQ_DEV double operator()(double x) const {                              \
            SafeMath m;                                                          \
            return eval(x, m);                                                   \
        }                                                                        \
    };
#define GKQ_F1(ID, HEAVY_, BODY) GKQ_F1_IMPL(ID, HEAVY_, false, BODY)
#ifdef GKQ_NO_PAIRWISE
#define GKQ_F1P(ID, HEAVY_, BODY) GKQ_F1_IMPL(ID, HEAVY_, false, BODY)
#else
#define GKQ_F1P(ID, HEAVY_, BODY) GKQ_F1_IMPL(ID, HEAVY_, true, BODY)
#endif
GKQ_F1(F1_SQUARE, false, return x * x;)                                   // benches/single.rs:11
GKQ_F1(F1_RECIP_P, false, return m.div(p[0], x);)                               // benches/single.rs:20
GKQ_F1(F1_LORENTZ_P, false, return m.div(1.0, 1.0 + p[0] * (x * x));)         // benches/single.rs:32
GKQ_F1P(F1_EXPCOS_P, true, return m.exp(-p[0] * x) * m.cos(p[1] * x);)         // S1 / S3
GKQ_F1P(F1_GAUSS_P, true, return m.exp(-p[0] * (x * x));)                    // I2
GKQ_F1(F1_ABSPOW_P, true, return m.pow(fabs(x - p[0]), -p[1]);)             // P1
GKQ_F1(F1_LOGABS_P, true, return m.log(fabs(x - p[0]));)                    // P2
GKQ_F1(F1_F1, true, return m.pow(x, 2.6) * m.log(m.div(1., x));)                    // functions.rs:3-5
GKQ_F1(F1_F2, true, return m.pow(x, -0.9) * m.log(m.div(1., x));)                   // functions.rs:8-10
GKQ_F1P(F1_F3, true, return m.cos(2.4622888266898326 * m.sin(x));)             // functions.rs:13-15 (2^1.3)
GKQ_F1(F1_F4, true, return m.log(m.div(1., x));)                                  // functions.rs:17-19
GKQ_F1(F1_F5, true, double x2 = x * x;
       return x2 * x * m.log(fabs((x2 - 1.) * (x2 - 2.)));)                 // functions.rs:22-25
GKQ_F1(F1_F6, false, double x2 = x * x; double x3 = x2 * x;
       return m.div(1.0, x3 - x2 + x - 1.);)                                  // functions.rs:27-31
GKQ_F1P(F1_EXP_P, true, return m.exp(p[0] * x);)
GKQ_F1(F1_SQRT_SHIFT_P, false, return m.sqrt(x - p[0]);)
GKQ_F1(F1_INV_ABS_SHIFT_P, false, return m.div(1.0, fabs(x - p[0]));)
GKQ_F1P(F1_COS_P, true, return m.cos(p[0] * x);)
GKQ_F1P(F1_SIN_P, true, return m.sin(p[0] * x);)
GKQ_F1(F1_SEMICIRCLE_P, false, return m.sqrt(p[0] - x * x);)                  // examples/sphere.rs:11 (p0 = 1)
// examples/wallis_integral.rs:13: (x * PI).sin().powi(m), m = p0; powi = compiler-rt __powidf2
GKQ_F1(F1_SINPOW_P, true, double s = m.sin(x * 3.141592653589793); int n = (int)p[0];
       double r = 1.0; for (int k = n < 0 ? -n : n;;) { if (k & 1) r *= s; k >>= 1; if (k == 0) break; s *= s; }
       return n < 0 ? m.div(1.0, r) : r;)
#undef GKQ_F1
#undef GKQ_F1P
#undef GKQ_F1_IMPL

// ---- 2-D ------------------------------------------------------------------------------
enum F2Id : int {
    F2_XY = 0, F2_RECIP_XY_P, F2_RSQRT3, F2_G1, F2_G2, F2_G3, F2_G4, F2_GP1, F2_GP2, F2_GP3,
    F2_D1_P, F2_HEMISPHERE_P, F2_COUNT
};
template <int ID> struct F2;
#define GKQ_F2(ID, HEAVY_, BODY)                                                      \
    template <> struct F2<ID> {                                                       \
        static constexpr bool HEAVY = HEAVY_;                                         \
        static constexpr bool PAIRWISE = false;                                       \
        double p[MAX_PARAMS];                                                         \
        template <class M> GKQ_DEV double eval(double x, double y, M& m) const {      \
            (void)m;                                                                  \
            BODY                                                                      \
        }                                                                             \
        GKQ_DEV double operator()(double x, double y) const {                         \
            SafeMath m;                                                               \
            return eval(x, y, m);                                                     \
        }                                                                             \
    };
GKQ_F2(F2_XY, false, return x * y;)                                       // benches/double.rs:10
GKQ_F2(F2_RECIP_XY_P, false, return m.div(p[0], x * y);)                      // benches/double.rs:19
GKQ_F2(F2_RSQRT3, false, double denom = 1.0 + x * x + y * y;
       return m.div(1.0, denom * m.sqrt(denom));)                               // benches/double.rs:31-33
GKQ_F2(F2_G1, false, return 1.0 - x * y;)                                 // functions.rs:37-39
GKQ_F2(F2_G2, true, return m.exp(m.div(y, x));)                                   // functions.rs:41-43
GKQ_F2(F2_G3, false, return m.sqrt(0.25 - x * x - y * y);)                  // functions.rs:45-47
GKQ_F2(F2_G4, false, double a = 1.0 + x + y; return m.div(1.0, a * (a * a));) // functions.rs:49-51 powi(3)
GKQ_F2(F2_GP1, false, return m.div(1.0, m.sqrt(fabs(x) + fabs(y)));)              // functions.rs:54-56
GKQ_F2(F2_GP2, false, return m.div(x, x * x + y * y);)                        // functions.rs:58-60
GKQ_F2(F2_GP3, false, return m.div(0.5, m.sqrt(x * x + y * y));)                  // functions.rs:62-64
GKQ_F2(F2_D1_P, false, return m.div(1.0, m.sqrt(fabs(x - p[0]) + fabs(y - p[1])));)  // D1
GKQ_F2(F2_HEMISPHERE_P, false, return m.sqrt(p[0] - x * x - y * y);)        // examples/sphere.rs:23 (p0 = 1)
#undef GKQ_F2

// ---- y-range functors: run-time switch (evaluated once per inner integral) -------------
enum YRId : int { YR_CONST = 0, YR_ZERO_TO_X = 1, YR_CIRCLE = 2, YR_COUNT = 3 };
struct YRange {
    int id;
    double q[2];
    GKQ_DEV void operator()(double x, double& y0, double& y1) const {
        if (id == YR_CONST) {
            y0 = q[0];
            y1 = q[1];
        } else if (id == YR_ZERO_TO_X) {
            y0 = 0.0;
            y1 = x;
        } else {
            double ymax = sqrt(q[0] - x * x);
            y0 = -ymax;
            y1 = ymax;
        }
    }
};

// Registry rows: X(id, name, nparams, flops_per_eval).  Single source of truth for the host
// tables (engine.cu), the launch tables and the explicit instantiations (inst_*.cu).
// flops: counted from source for polynomial/rational bodies (add, sub,
// mul, div, sqrt = 1); for transcendental bodies the DADD+DMUL+2*DFMA SASS count of one
// evaluation (see profiles/functor_flops_r01.md for how they were obtained; the pow / log functors were
// re-measured with the engine's own pow / log: profiles/functor_flops_r02.csv).
#define GKQ_F1_LIST(X)                       \
    X(F1_SQUARE, "square", 0, 1.0)           \
    X(F1_RECIP_P, "recip_p", 1, 1.0)         \
    X(F1_LORENTZ_P, "lorentz_p", 1, 4.0)     \
    X(F1_EXPCOS_P, "expcos_p", 2, 61.0)     \
    X(F1_GAUSS_P, "gauss_p", 1, 31.0)        \
    X(F1_ABSPOW_P, "abspow_p", 2, 68.0)     \
    X(F1_LOGABS_P, "logabs_p", 1, 30.0)      \
    X(F1_F1, "f1", 0, 104.0)                 \
    X(F1_F2, "f2", 0, 104.0)                 \
    X(F1_F3, "f3", 0, 59.0)                 \
    X(F1_F4, "f4", 0, 38.0)                  \
    X(F1_F5, "f5", 0, 35.0)                  \
    X(F1_F6, "f6", 0, 6.0)                   \
    X(F1_EXP_P, "exp_p", 1, 30.0)            \
    X(F1_SQRT_SHIFT_P, "sqrt_shift_p", 1, 2.0)       \
    X(F1_INV_ABS_SHIFT_P, "inv_abs_shift_p", 1, 2.0) \
    X(F1_COS_P, "cos_p", 1, 30.0)            \
    X(F1_SIN_P, "sin_p", 1, 30.0)            \
    X(F1_SEMICIRCLE_P, "semicircle_p", 1, 3.0)       \
    X(F1_SINPOW_P, "sinpow_p", 1, 40.0)

#define GKQ_F2_LIST(X)                       \
    X(F2_XY, "xy", 0, 1.0)                   \
    X(F2_RECIP_XY_P, "recip_xy_p", 1, 2.0)   \
    X(F2_RSQRT3, "rsqrt3", 0, 8.0)           \
    X(F2_G1, "g1", 0, 2.0)                   \
    X(F2_G2, "g2", 0, 45.0)                  \
    X(F2_G3, "g3", 0, 5.0)                   \
    X(F2_G4, "g4", 0, 5.0)                   \
    X(F2_GP1, "gp1", 0, 3.0)                 \
    X(F2_GP2, "gp2", 0, 4.0)                 \
    X(F2_GP3, "gp3", 0, 5.0)                 \
    X(F2_D1_P, "d1_p", 2, 5.0)               \
    X(F2_HEMISPHERE_P, "hemisphere_p", 1, 5.0)

#define GKQ_YR_LIST(X)                       \
    X(YR_CONST, "const", 2, 0.0)             \
    X(YR_ZERO_TO_X, "zero_to_x", 0, 0.0)     \
    X(YR_CIRCLE, "circle", 1, 3.0)

}  // namespace gkq
