// Accuracy of gk_log_fast / gk_pow_fast (gkquad-rs_b200/csrc/gk_math_logpow.cuh) measured on the host
// against x87 long double (64-bit significand): the same source compiles with -DGKQ_MATH_HOST.
//   g++ -O2 -ffp-contract=off -DGKQ_MATH_HOST -I gkquad-rs_b200/csrc tools/hostcheck/check_logpow.cpp -o /tmp/check_logpow
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "gk_math_logpow.cuh"

static double ulp_err(double got, long double want) {
    const double w = (double)want;
    const long double u = (long double)std::fabs(std::nextafter(w, INFINITY) - w);
    return (double)(fabsl((long double)got - want) / u);
}

int main() {
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    double worst_log = 0, worst_pow = 0, sum_log = 0, sum_pow = 0;
    long long nlog = 0, npow = 0, differ_log = 0, differ_pow = 0, badrange = 0;
    double wl_x = 0, wp_x = 0, wp_y = 0;
    for (int it = 0; it < 4000000; ++it) {
        double x;
        switch (it & 3) {
            case 0: x = U(rng); break;                                  // (0, 1)
            case 1: x = 0.5 + U(rng); break;                            // around 1
            case 2: x = std::exp((U(rng) - 0.5) * 1400.0); break;       // whole range
            default: x = 1.0 + (U(rng) - 0.5) * 1e-3; break;            // very near 1
        }
        if (!gkq::gk_log_in_range(x)) continue;
        const double g = gkq::gk_log_fast(x);
        const long double w = logl((long double)x);
        const double e = ulp_err(g, w);
        if (e > worst_log) { worst_log = e; wl_x = x; }
        sum_log += e;
        ++nlog;
        if (g != std::log(x)) ++differ_log;
        // pow
        double y;
        switch ((it >> 2) & 3) {
            case 0: y = -(0.1 + 0.5 * U(rng)); break;   // P1
            case 1: y = 2.6; break;                     // f1
            case 2: y = -0.9; break;                    // f2
            default: y = (U(rng) - 0.5) * 40.0; break;
        }
        bool ok = true;
        const double gp = gkq::gk_pow_fast(x, y, ok);
        if (!ok) { ++badrange; continue; }
        const long double wp = powl((long double)x, (long double)y);
        const double ep = ulp_err(gp, wp);
        if (ep > worst_pow) { worst_pow = ep; wp_x = x; wp_y = y; }
        sum_pow += ep;
        ++npow;
        if (gp != std::pow(x, y)) ++differ_pow;
    }
    std::printf("log: n %lld max ulp %.4f (x = %.17g) mean %.4f, differs from glibc on %lld (%.4f %%)\n", nlog, worst_log,
                wl_x, sum_log / nlog, differ_log, 100.0 * differ_log / nlog);
    std::printf("pow: n %lld max ulp %.4f (x = %.17g, y = %.17g) mean %.4f, differs from glibc on %lld (%.4f %%), out of fast range %lld\n",
                npow, worst_pow, wp_x, wp_y, sum_pow / npow, differ_pow, 100.0 * differ_pow / npow, badrange);
    // edge values
    bool ok = true;
    std::printf("pow(1, 5) = %.17g, pow(2, 10) = %.17g, pow(0.5, -3) = %.17g, pow(7, 0) = %.17g, log(1) = %.17g, log(2) - M_LN2 = %g\n",
                gkq::gk_pow_fast(1.0, 5.0, ok), gkq::gk_pow_fast(2.0, 10.0, ok), gkq::gk_pow_fast(0.5, -3.0, ok),
                gkq::gk_pow_fast(7.0, 0.0, ok), gkq::gk_log_fast(1.0), gkq::gk_log_fast(2.0) - M_LN2);
    return (worst_log <= 0.6 && worst_pow <= 0.6) ? 0 : 1;
}
