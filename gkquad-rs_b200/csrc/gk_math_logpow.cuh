// gk_math_logpow.cuh -- the engine's own FP64 log and pow for integrand functors.
//
// Why: the QAGP workload families (|x - c|^-p, ln|x - c|, x^a ln(1/x): SURVEY.md section 8d, tests
// f1 / f2 / f4 / f5) spend most of their instructions inside libdevice's pow (246 warp instructions
// per sample, 120 of them FP64) and log.  The rule kernels are issue-bound, so what pays is fewer
// instructions per sample at the same accuracy.
//
// Method (the classic table-driven scheme of modern libms, with this library's own tables from
// tools/gen_math_coeffs.py):
//   x = 2^k z, z in [0.706, 1.412) so that x near 1 has k = 0; the top 7 mantissa bits of
//   (bits(x) - OFF) select one of 128 subintervals with a centre c whose reciprocal has <= 9
//   significant bits, hence r = z / c - 1 is EXACT in one fma and |r| < 2^-7;
//   log(x) = k ln2 + log(c) + log1p(r), log1p(r) = r - r^2 / 2 + r^3 Q(r), Q of degree 6;
//   log(c) is stored as a head rounded to 2^-43 (k ln2hi + head is exact) plus a tail.
//   log: evaluated in plain double with the small terms summed first: <= 0.53 ulp.
//   pow: log(x) as a double-double (relative error ~2^-68), e = y log(x) as head + tail,
//   exp(e) = 2^m T[j] (1 + (e^r - 1)) with T = 2^(j/64) as head + tail, |r| <= ln2/128: <= 0.53 ulp.
// Both are branch-free; arguments outside the fast range (x not a positive normal number, |y log x|
// >= 700, NaN) are reported through `ok` and re-evaluated by the caller through libdevice.
//
// The same source compiles for the host (-DGKQ_MATH_HOST, tools/check_logpow.cpp) so that accuracy
// can be measured against the correctly rounded result without a GPU.
#pragma once
#include "gk_math_coeffs.inc"

#ifdef GKQ_MATH_HOST
#include <cmath>
#include <cstdint>
#include <cstring>
#define GKQ_LP_DEV inline
#define GKQ_LP_TABLE static const
#define GKQ_LP_CONST static const
namespace gkq {
inline double lp_fma(double a, double b, double c) { return std::fma(a, b, c); }
inline long long lp_bits(double x) { long long v; std::memcpy(&v, &x, 8); return v; }
inline double lp_from_bits(long long v) { double x; std::memcpy(&x, &v, 8); return x; }
inline unsigned long long lp_ld(const unsigned long long* p) { return *p; }
}  // namespace gkq
#else
#define GKQ_LP_DEV __device__ __forceinline__
#define GKQ_LP_TABLE static __device__ const
#define GKQ_LP_CONST static __constant__
namespace gkq {
GKQ_LP_DEV double lp_fma(double a, double b, double c) { return fma(a, b, c); }
GKQ_LP_DEV long long lp_bits(double x) { return __double_as_longlong(x); }
GKQ_LP_DEV double lp_from_bits(long long v) { return __longlong_as_double(v); }
GKQ_LP_DEV unsigned long long lp_ld(const unsigned long long* p) { return __ldg(p); }
}  // namespace gkq
#endif

namespace gkq {

struct LogPowConsts {
    unsigned long long q[8];        // Q(r) ascending, degree 6 (+pad)
    unsigned long long ln2hi, ln2lo;  // ln2 split so that k * ln2hi is exact for |k| < 2^11
    unsigned long long e2[8];       // (e^r - 1 - r) / r^2, degree 4 (+pad), |r| <= ln2 / 128
    unsigned long long log2e_64, ln2_64_hi, ln2_64_lo, pad;
};
// lane-indexed tables: read-only global memory (L1-resident)
GKQ_LP_TABLE unsigned long long g_log_tab[128][3] = GKQ_LOG_TAB;
GKQ_LP_TABLE unsigned long long g_exp2_hl[64][2] = GKQ_EXP2_TAB_HL;
// warp-uniform operands
GKQ_LP_CONST LogPowConsts c_logpow = {GKQ_LOG_Q, GKQ_LN2_HI43, GKQ_LN2_LO43, GKQ_EXP_E2B,
                                      GKQ_64_LOG2E, GKQ_LN2_64_HI, GKQ_LN2_64_LO, 0ull};

GKQ_LP_DEV double lp_c(unsigned long long bits) { return lp_from_bits((long long)bits); }

// x is a positive, normal, finite double?  (integer pipe; false for NaN)
GKQ_LP_DEV bool gk_log_in_range(double x) {
    const unsigned hi = (unsigned)((unsigned long long)lp_bits(x) >> 32);
    return hi - 0x00100000u < 0x7fe00000u;
}

struct LogReduced {
    double kd, r, invc, logc, tail;
};
GKQ_LP_DEV LogReduced gk_log_reduce(double x) {
    const long long OFF = 0x3fe6955500000000ll;
    const long long ix = lp_bits(x);
    const long long tmp = ix - OFF;
    const int i = (int)(tmp >> 45) & 127;
    const long long k = tmp >> 52;  // arithmetic shift
    const double z = lp_from_bits(ix - (long long)((unsigned long long)tmp & 0xfff0000000000000ull));
    LogReduced L;
    L.kd = (double)(int)k;
    L.invc = lp_c(lp_ld(&g_log_tab[i][0]));
    L.logc = lp_c(lp_ld(&g_log_tab[i][1]));
    L.tail = lp_c(lp_ld(&g_log_tab[i][2]));
    L.r = lp_fma(z, L.invc, -1.0);  // exact: 1/c has <= 9 significant bits
    return L;
}
GKQ_LP_DEV double gk_log_q(double r) {
    double q = lp_c(c_logpow.q[6]);
    q = lp_fma(q, r, lp_c(c_logpow.q[5]));
    q = lp_fma(q, r, lp_c(c_logpow.q[4]));
    q = lp_fma(q, r, lp_c(c_logpow.q[3]));
    q = lp_fma(q, r, lp_c(c_logpow.q[2]));
    q = lp_fma(q, r, lp_c(c_logpow.q[1]));
    q = lp_fma(q, r, lp_c(c_logpow.q[0]));
    return q;
}

// log(x) for x a positive normal number (gk_log_in_range); <= 0.53 ulp.
GKQ_LP_DEV double gk_log_fast(double x) {
    const LogReduced L = gk_log_reduce(x);
    const double r = L.r;
    const double t1 = lp_fma(L.kd, lp_c(c_logpow.ln2hi), L.logc);  // exact
    const double hi = t1 + r;
    const double lo = (t1 - hi) + r;  // exact: |t1| >= |r| or t1 == 0
    const double q = gk_log_q(r);
    const double p = (r * r) * lp_fma(r, q, -0.5);  // log1p(r) - r
    const double small = lp_fma(L.kd, lp_c(c_logpow.ln2lo), L.tail);
    return hi + ((small + lo) + p);
}

// |e| < 700: exp(e) is a normal number with room to spare (false for NaN / inf)
GKQ_LP_DEV bool gk_pow_exp_in_range(double e) {
    const unsigned hi = (unsigned)((unsigned long long)lp_bits(e) >> 32) & 0x7fffffffu;
    return hi < 0x4085e000u;  // 700.0 = 0x4085e00000000000
}

// pow(x, y) for x a positive normal number and |y log x| < 700 (else ok = false and the value
// returned is unspecified); <= 0.53 ulp.
GKQ_LP_DEV double gk_pow_fast(double x, double y, bool& ok) {
    const LogReduced L = gk_log_reduce(x);
    const double r = L.r;
    // log(x) = lhi + llo, relative error ~2^-68
    const double t1 = lp_fma(L.kd, lp_c(c_logpow.ln2hi), L.logc);  // exact
    const double t2 = t1 + r;
    const double lo1 = lp_fma(L.kd, lp_c(c_logpow.ln2lo), L.tail);
    const double lo2 = (t1 - t2) + r;
    const double ar = -0.5 * r;
    const double ar2 = r * ar;  // -r^2 / 2, head
    const double hi = t2 + ar2;
    const double lo3 = lp_fma(ar, r, -ar2);  // ... its rounding error
    const double lo4 = (t2 - hi) + ar2;
    const double q = gk_log_q(r);
    const double p = (r * (r * r)) * q;
    const double lo = (((lo1 + lo2) + lo3) + lo4) + p;
    const double lhi = hi + lo;
    const double llo = (hi - lhi) + lo;
    // e = y log(x) as head + tail
    const double ehi = y * lhi;
    const double elo = lp_fma(y, llo, lp_fma(y, lhi, -ehi));
    ok = ok && gk_log_in_range(x) && gk_pow_exp_in_range(ehi);
    // exp(ehi + elo) = 2^m * T[j] * e^rr,  ehi = (64 m + j) ln2 / 64 + rr
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52
    const double t = lp_fma(ehi, lp_c(c_logpow.log2e_64), MAGIC);
    const double md = t - MAGIC;
    const int m = (int)(unsigned)(unsigned long long)lp_bits(t);  // low word: rint(ehi * 64 / ln2)
    const double Thi = lp_c(lp_ld(&g_exp2_hl[m & 63][0]));
    const double Tlo = lp_c(lp_ld(&g_exp2_hl[m & 63][1]));
    double rr = lp_fma(md, -lp_c(c_logpow.ln2_64_hi), ehi);
    rr = lp_fma(md, -lp_c(c_logpow.ln2_64_lo), rr);
    rr += elo;
    double e = lp_c(c_logpow.e2[4]);
    e = lp_fma(e, rr, lp_c(c_logpow.e2[3]));
    e = lp_fma(e, rr, lp_c(c_logpow.e2[2]));
    e = lp_fma(e, rr, lp_c(c_logpow.e2[1]));
    e = lp_fma(e, rr, lp_c(c_logpow.e2[0]));
    const double em1 = lp_fma(rr * rr, e, rr);  // e^rr - 1
    const double res = Thi + lp_fma(Thi, em1, Tlo);
    // scale by 2^(m >> 6): add to the exponent field (the result stays normal for |e| < 700)
    return lp_from_bits(lp_bits(res) + ((long long)(m >> 6) << 52));
}

}  // namespace gkq
