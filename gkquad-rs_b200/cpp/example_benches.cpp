// The reference's six micro-benchmarks (gkquad/benches/single.rs, double.rs) through the C++
// mirror: same integrands, ranges, configurations and assertions, one integral per call.
// Build: g++ -std=c++17 example_benches.cpp -L../lib -lgkquad_b200 -Wl,-rpath,../lib
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "gkquad_b200.hpp"

using namespace gkquad;

#define CHECK(cond)                                                   \
    do {                                                              \
        if (!(cond)) {                                                \
            std::fprintf(stderr, "assertion failed: %s\n", #cond);    \
            return 1;                                                 \
        }                                                             \
    } while (0)

int main() {
    const double PI = 3.141592653589793;
    {  // benches/single.rs:10-17
        auto r = single::Integrator(single::DeviceIntegrand{"square", {}}).run({0.0, 1.0})[0].unwrap();
        CHECK(std::fabs(r.estimate - 1.0 / 3.0) <= 1.49e-8 && r.nevals == 17);
    }
    {  // benches/single.rs:19-29
        auto it = single::Integrator(single::DeviceIntegrand{"recip_p", {{1.0}}});
        it.points({0.0}).tolerance(Tolerance::Absolute(1.49e-8));
        auto r = it.run({-1.0, 1.0})[0].unwrap();
        CHECK(std::fabs(r.estimate) <= 1.49e-8 && r.nevals == 250);
    }
    {  // benches/single.rs:31-39
        auto r = single::Integrator(single::DeviceIntegrand{"lorentz_p", {{1.0}}}).run(single::Range::unbounded())[0].unwrap();
        CHECK(std::fabs(r.estimate - PI) <= 1.49e-8 && r.nevals == 67);
    }
    {  // benches/double.rs:9-16
        auto r = double_::integral2(double_::DeviceIntegrand2{"xy", {}}, double_::Rectangle(0, 1, 0, 1))[0].unwrap();
        CHECK(std::fabs(r.estimate - 0.25) <= 1.49e-8 && r.nevals == 289);
    }
    {  // benches/double.rs:18-28
        double_::IntegrationConfig2 c;
        c.tolerance = Tolerance::Absolute(1.49e-8);
        c.points = {{0.0, 0.0}};
        auto r = double_::integral2_with_config(double_::DeviceIntegrand2{"recip_xy_p", {{1.0}}},
                                                double_::Rectangle(-1, 1, -1, 1), c)[0].unwrap();
        CHECK(std::fabs(r.estimate) <= 1.49e-8 && r.nevals == 12500);
    }
    {  // benches/double.rs:30-41
        auto r = double_::integral2(double_::DeviceIntegrand2{"rsqrt3", {}},
                                    double_::RectangleFrom(single::Range::unbounded(), single::Range::unbounded()))[0].unwrap();
        CHECK(std::fabs(r.estimate - 2.0 * PI) <= 1.49e-8 && r.nevals == 11239);
    }
    {  // numerical failure is a status, not an exception (ValueWithError)
        auto it = single::Integrator(single::DeviceIntegrand{"f6", {}});
        it.points({1.0}).tolerance(Tolerance::Absolute(1e-10)).algorithm(single::QAGP());
        auto r = it.run({0.0, 2.0})[0];
        CHECK(r.has_err() && *r.error == RuntimeError::Divergent && r.unwrap_unchecked().nevals == 550);
    }
    {  // a parameter sweep in one call (the examples/wallis_integral.rs use-case)
        single::DeviceIntegrand f{"expcos_p", {{0.5, 1.0, 1.5, 2.0}, {1.0, 4.0, 7.0, 10.0}}};
        auto it = single::Integrator(f);
        it.tolerance(Tolerance::Relative(1e-10));
        auto r = it.run({0.0, 1.0});
        for (std::size_t i = 0; i < r.size(); ++i) {
            const double p0 = f.params[0][i], p1 = f.params[1][i];
            const double truth = (p0 * (1 - std::exp(-p0) * std::cos(p1)) + p1 * std::exp(-p0) * std::sin(p1)) / (p0 * p0 + p1 * p1);
            CHECK(std::fabs(r.estimate[i] - truth) <= 1e-10 * std::fabs(truth));
        }
    }
    {  // the deprecated algorithms: tests/algorithms_double.rs:44-54 (qag_g1) and a QAG run
        double_::IntegrationConfig2 c2;
        c2.tolerance = Tolerance::Relative(1e-10);
        auto r = double_::QAG2().integrate(double_::DeviceIntegrand2{"g1", {}}, double_::Rectangle(0., 1., 0., 1.), c2)[0].unwrap();
        CHECK(std::fabs(r.estimate - 0.75) <= 1e-15 && r.nevals == 289);
        auto it = single::Integrator(single::DeviceIntegrand{"lorentz_p", {{1.0}}});
        it.algorithm(single::QAG()).tolerance(Tolerance::Absolute(1e-10));
        auto q = it.run({-1.0, 1.0})[0].unwrap();
        CHECK(std::fabs(q.estimate - PI / 2) <= 1e-10);
    }
    {  // DynamicX (double/range.rs:38-75): x in [1, 2] inside, y in [0, 1] outside, f(x, y) = exp(y / x)
        double_::IntegrationConfig2 c2;
        c2.tolerance = Tolerance::Relative(1e-9);
        auto r = double_::QAGS2().integrate(double_::DeviceIntegrand2{"g2", {}},
                                            double_::DynamicX("const", 0.0, 1.0, {1.0, 2.0}), c2)[0].unwrap();
        const double truth = 1.4483309393877208;  // scipy.integrate.dblquad, 1e-13
        CHECK(std::fabs(r.estimate - truth) <= 1e-8 * truth);
    }
    std::printf("all reference bench assertions hold through the C++ mirror\n");
    return 0;
}
