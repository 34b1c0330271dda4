// tests/emul/test_integrator.cpp -- TEST PROGRAM: emme::RungeKutta3 (emme_b200/host/integrator.hpp)
// on the harmonic oscillator x'' = -x of the reference's test/test_integrator.cpp, fixed step and
// step_adaptive, checked (i) against sin t with the reference test's own 1e-5 bound and (ii) bit
// for bit against what the reference's Integrator template produced for the same state
// (tests/golden/oscillator_rk3.bin, written by oracle/pic_driver.cpp `oscillator`).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <limits>
#include <vector>

#include "../../emme_b200/host/integrator.hpp"

struct double_state {
    using value_type = double;
    struct velocity_type {
        value_type v0, v1;
    };
    friend velocity_type operator*(value_type a, velocity_type v) { return {a * v.v0, a * v.v1}; }
    friend velocity_type operator+(velocity_type l, velocity_type r) { return {l.v0 + r.v0, l.v1 + r.v1}; }
    velocity_type initial_velocity_storage() const { return {}; }
    void put_velocity(velocity_type& v) {   // d^2x/dt^2 = -x
        v.v0 = x1;
        v.v1 = -x0;
    }
    void update(const velocity_type& v, value_type dt) {
        x0 += v.v0 * dt;
        x1 += v.v1 * dt;
        t += dt;
    }
    value_type get_update_err(const velocity_type& v, value_type dt) {
        auto l2 = [](value_type a, value_type b) { return std::sqrt(.5 * (a * a + b * b)); };
        return l2(x0, x1) < std::numeric_limits<value_type>::epsilon() ? l2(v.v0 * dt, v.v1 * dt)
                                                                        : l2(v.v0 * dt, v.v1 * dt) / l2(x0, x1);
    }
    double t, x0, x1;
};

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    std::ifstream f(argv[1], std::ios::binary);
    std::uint64_t n = 0;
    int bad = 0;
    double worst = 0;
    f.read(reinterpret_cast<char*>(&n), 8);
    {
        double_state s{0, 0, 1};
        emme::RungeKutta3<double_state> rk(s, 1.e-5, 1.e-7);
        for (std::uint64_t i = 0; i < n; ++i) {
            rk.step(0.01);
            double rec[3];
            f.read(reinterpret_cast<char*>(rec), sizeof(rec));
            if (rec[0] != s.t || rec[1] != s.x0 || rec[2] != s.x1) ++bad;
            worst = std::fmax(worst, std::fabs(s.x0 - std::sin(s.t)));
        }
        std::printf("fixed: %llu steps, mismatches %d, max |x - sin t| %.3e\n", (unsigned long long)n, bad, worst);
        if (n != 1000 || worst > 1e-5) ++bad;
    }
    f.read(reinterpret_cast<char*>(&n), 8);
    {
        double_state s{0, 0, 1};
        emme::RungeKutta3<double_state> rk(s, 1.e-5, 1.e-7);
        std::uint64_t c = 0;
        double worst2 = 0;
        while (s.t < 10) {
            const double dt = rk.step_adaptive();
            double rec[4] = {0, 0, 0, 0};
            if (c < n) f.read(reinterpret_cast<char*>(rec), sizeof(rec));
            if (c >= n || rec[0] != dt || rec[1] != s.t || rec[2] != s.x0 || rec[3] != s.x1) ++bad;
            worst2 = std::fmax(worst2, std::fabs(s.x0 - std::sin(s.t)));
            ++c;
        }
        std::printf("adaptive: %llu steps (reference %llu), mismatches %d, max |x - sin t| %.3e\n",
                    (unsigned long long)c, (unsigned long long)n, bad, worst2);
        if (c != n || worst2 > 1e-5) ++bad;
    }
    return bad ? 1 : 0;
}
