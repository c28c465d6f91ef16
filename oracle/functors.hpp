// functors.hpp -- CPU integrand registry of the oracle.  TEST INFRASTRUCTURE ONLY.
//
// The reference takes Rust closures (single/common.rs:129-145, double/common.rs:35-45);
// closures cannot cross an FFI, so the engine and this oracle both select integrands by
// ID from a registry.  IDs, names and parameter counts here MUST match
// gkquad-rs_b200/csrc/functors.cuh (tests/test_abi.py checks the table through both
// libraries).  Bodies restate, operation by operation, the closures in
// /root/reference/gkquad/{benches,tests/common/functions.rs} and the parameterised
// families of SURVEY.md section 8(d); they call the platform libm exactly where the
// Rust code calls f64::{powf,ln,exp,sin,cos,sqrt}.
#pragma once
#include <cmath>

namespace gkqo {

struct FunctorInfo {
    int id;
    const char* name;
    int nparams;
};

// ---- 1-D ----------------------------------------------------------------------------
enum F1Id : int {
    F1_SQUARE = 0,      // x*x                         benches/single.rs:11
    F1_RECIP_P = 1,     // p0 / x                      benches/single.rs:20 (p0=1), functions.rs:33-35
    F1_LORENTZ_P = 2,   // 1/(1 + p0*(x*x))            benches/single.rs:32 (p0=1); S2, I1
    F1_EXPCOS_P = 3,    // exp(-p0 x) cos(p1 x)        S1, S3
    F1_GAUSS_P = 4,     // exp(-p0 x^2)                I2
    F1_ABSPOW_P = 5,    // |x-p0|^(-p1)                P1
    F1_LOGABS_P = 6,    // ln|x-p0|                    P2
    F1_F1 = 7,          // x^2.6 ln(1/x)               functions.rs:3-5
    F1_F2 = 8,          // x^-0.9 ln(1/x)              functions.rs:8-10
    F1_F3 = 9,          // cos(2^1.3 sin x)            functions.rs:13-15
    F1_F4 = 10,         // ln(1/x)                     functions.rs:17-19
    F1_F5 = 11,         // x^3 ln|(x^2-1)(x^2-2)|      functions.rs:22-25
    F1_F6 = 12,         // 1/(x^3-x^2+x-1)             functions.rs:27-31
    F1_EXP_P = 13,      // exp(p0*x)                   SURVEY appendix A (half-infinite ranges)
    F1_SQRT_SHIFT_P = 14,   // sqrt(x-p0): NaN left of p0  (NanValueEncountered paths)
    F1_INV_ABS_SHIFT_P = 15,  // 1/|x-p0|: +inf at a node  (appendix A, inf is not NaN)
    F1_COS_P = 16,      // cos(p0*x)                   budget-exhaustion case (section 0.4)
    F1_SIN_P = 17,      // sin(p0*x)
    F1_SEMICIRCLE_P = 18,   // sqrt(p0 - x*x)           examples/sphere.rs:11
    F1_SINPOW_P = 19,   // sin(pi x)^p0 (powi)         examples/wallis_integral.rs:13
    // oracle-only probe (not in the product registry): S1's integrand with cos() moved one ulp up,
    // i.e. "the reference linked against a libm that is 1 ulp off" -- tools/libm_sensitivity.py uses
    // it to measure how often gkquad's own status / nevals flip under ulp-level libm differences
    F1_EXPCOS_P_ULP = 20,
    F1_COUNT = 21,
};

static const FunctorInfo F1_TABLE[F1_COUNT] = {
    {F1_SQUARE, "square", 0},   {F1_RECIP_P, "recip_p", 1},   {F1_LORENTZ_P, "lorentz_p", 1},
    {F1_EXPCOS_P, "expcos_p", 2}, {F1_GAUSS_P, "gauss_p", 1}, {F1_ABSPOW_P, "abspow_p", 2},
    {F1_LOGABS_P, "logabs_p", 1}, {F1_F1, "f1", 0},           {F1_F2, "f2", 0},
    {F1_F3, "f3", 0},           {F1_F4, "f4", 0},             {F1_F5, "f5", 0},
    {F1_F6, "f6", 0},           {F1_EXP_P, "exp_p", 1},       {F1_SQRT_SHIFT_P, "sqrt_shift_p", 1},
    {F1_INV_ABS_SHIFT_P, "inv_abs_shift_p", 1},               {F1_COS_P, "cos_p", 1},
    {F1_SIN_P, "sin_p", 1},     {F1_SEMICIRCLE_P, "semicircle_p", 1}, {F1_SINPOW_P, "sinpow_p", 1},
    {F1_EXPCOS_P_ULP, "expcos_p_ulp", 2},
};

template <int ID>
struct F1;
#define GKQO_F1(ID, BODY)                                             \
    template <>                                                       \
    struct F1<ID> {                                                   \
        const double* p;                                              \
        inline double operator()(double x) const { BODY }             \
    };
GKQO_F1(F1_SQUARE, return x * x;)
GKQO_F1(F1_RECIP_P, return p[0] / x;)
GKQO_F1(F1_LORENTZ_P, return 1.0 / (1.0 + p[0] * (x * x));)
GKQO_F1(F1_EXPCOS_P, return std::exp(-p[0] * x) * std::cos(p[1] * x);)
GKQO_F1(F1_GAUSS_P, return std::exp(-p[0] * (x * x));)
GKQO_F1(F1_ABSPOW_P, return std::pow(std::fabs(x - p[0]), -p[1]);)
GKQO_F1(F1_LOGABS_P, return std::log(std::fabs(x - p[0]));)
GKQO_F1(F1_F1, return std::pow(x, 2.6) * std::log(1. / x);)
GKQO_F1(F1_F2, return std::pow(x, -0.9) * std::log(1. / x);)
GKQO_F1(F1_F3, return std::cos(std::pow(2.0, 1.3) * std::sin(x));)
GKQO_F1(F1_F4, return std::log(1. / x);)
GKQO_F1(F1_F5, double x2 = x * x; return x2 * x * std::log(std::fabs((x2 - 1.) * (x2 - 2.)));)
GKQO_F1(F1_F6, double x2 = x * x; double x3 = x2 * x; return 1.0 / (x3 - x2 + x - 1.);)
GKQO_F1(F1_EXP_P, return std::exp(p[0] * x);)
GKQO_F1(F1_SQRT_SHIFT_P, return std::sqrt(x - p[0]);)
GKQO_F1(F1_INV_ABS_SHIFT_P, return 1.0 / std::fabs(x - p[0]);)
GKQO_F1(F1_COS_P, return std::cos(p[0] * x);)
GKQO_F1(F1_SIN_P, return std::sin(p[0] * x);)
GKQO_F1(F1_SEMICIRCLE_P, return std::sqrt(p[0] - x * x);)
GKQO_F1(F1_SINPOW_P, double s = std::sin(x * 3.141592653589793); int n = (int)p[0];
        double r = 1.0; for (int k = n < 0 ? -n : n;;) { if (k & 1) r *= s; k >>= 1; if (k == 0) break; s *= s; }
        return n < 0 ? 1.0 / r : r;)
GKQO_F1(F1_EXPCOS_P_ULP, return std::exp(-p[0] * x) * std::nextafter(std::cos(p[1] * x), 2.0);)
#undef GKQO_F1

// ---- 2-D ----------------------------------------------------------------------------
enum F2Id : int {
    F2_XY = 0,          // x*y                         benches/double.rs:10
    F2_RECIP_XY_P = 1,  // p0/(x*y)                    benches/double.rs:19 (p0=1)
    F2_RSQRT3 = 2,      // 1/(d*sqrt(d)), d=1+x*x+y*y  benches/double.rs:31-33
    F2_G1 = 3,          // 1 - x*y                     functions.rs:37-39
    F2_G2 = 4,          // exp(y/x)                    functions.rs:41-43
    F2_G3 = 5,          // sqrt(0.25 - x*x - y*y)      functions.rs:45-47
    F2_G4 = 6,          // 1/(1+x+y)^3 (powi(3))       functions.rs:49-51
    F2_GP1 = 7,         // 1/sqrt(|x|+|y|)             functions.rs:54-56
    F2_GP2 = 8,         // x/(x*x+y*y)                 functions.rs:58-60
    F2_GP3 = 9,         // 0.5/sqrt(x*x+y*y)           functions.rs:62-64
    F2_D1_P = 10,       // 1/sqrt(|x-p0|+|y-p1|)       D1 (section 8d cfg-5)
    F2_HEMISPHERE_P = 11,   // sqrt(p0 - x*x - y*y)    examples/sphere.rs:23
    F2_COUNT = 12,
};

static const FunctorInfo F2_TABLE[F2_COUNT] = {
    {F2_XY, "xy", 0},   {F2_RECIP_XY_P, "recip_xy_p", 1}, {F2_RSQRT3, "rsqrt3", 0},
    {F2_G1, "g1", 0},   {F2_G2, "g2", 0},                 {F2_G3, "g3", 0},
    {F2_G4, "g4", 0},   {F2_GP1, "gp1", 0},               {F2_GP2, "gp2", 0},
    {F2_GP3, "gp3", 0}, {F2_D1_P, "d1_p", 2},             {F2_HEMISPHERE_P, "hemisphere_p", 1},
};

template <int ID>
struct F2;
#define GKQO_F2(ID, BODY)                                                  \
    template <>                                                            \
    struct F2<ID> {                                                        \
        const double* p;                                                   \
        inline double operator()(double x, double y) const { BODY }        \
    };
GKQO_F2(F2_XY, return x * y;)
GKQO_F2(F2_RECIP_XY_P, return p[0] / (x * y);)
GKQO_F2(F2_RSQRT3, double denom = 1.0 + x * x + y * y; return 1.0 / (denom * std::sqrt(denom));)
GKQO_F2(F2_G1, return 1.0 - x * y;)
GKQO_F2(F2_G2, return std::exp(y / x);)
GKQO_F2(F2_G3, return std::sqrt(0.25 - x * x - y * y);)
// powi(3) == compiler-rt __powidf2: r = a; a = a*a; r = r*a  (SURVEY section 4 gotcha b)
GKQO_F2(F2_G4, double a = 1.0 + x + y; return 1.0 / (a * (a * a));)
GKQO_F2(F2_GP1, return 1.0 / std::sqrt(std::fabs(x) + std::fabs(y));)
GKQO_F2(F2_GP2, return x / (x * x + y * y);)
GKQO_F2(F2_GP3, return 0.5 / std::sqrt(x * x + y * y);)
GKQO_F2(F2_D1_P, return 1.0 / std::sqrt(std::fabs(x - p[0]) + std::fabs(y - p[1]));)
GKQO_F2(F2_HEMISPHERE_P, return std::sqrt(p[0] - x * x - y * y);)
#undef GKQO_F2

// ---- y-range functors (double/range.rs:77-112: DynamicY's closure) -------------------
enum YRId : int {
    YR_CONST = 0,      // (q0, q1)                       Rectangle -> DynamicY, qags.rs:18-27
    YR_ZERO_TO_X = 1,  // (0, x)                         tests/algorithms_double.rs:76
    YR_CIRCLE = 2,     // (-sqrt(q0-x*x), +sqrt(q0-x*x)) tests/algorithms_double.rs:100-103,159-162
    YR_COUNT = 3,
};
static const FunctorInfo YR_TABLE[YR_COUNT] = {
    {YR_CONST, "const", 2}, {YR_ZERO_TO_X, "zero_to_x", 0}, {YR_CIRCLE, "circle", 1}};

}  // namespace gkqo
