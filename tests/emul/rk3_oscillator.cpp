// tests/emul/rk3_oscillator.cpp -- TEST PROGRAM: emme::RungeKutta3 (tests/emul/integrator.hpp)
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

#include "integrator.hpp"

// x'' = -x as a state for emme::RungeKutta3 (interface: tests/emul/integrator.hpp)
struct Oscillator {
    using value_type = double;
    struct velocity_type {
        double dpos, dvel;
        friend velocity_type operator*(double a, const velocity_type& v) { return {a * v.dpos, a * v.dvel}; }
        friend velocity_type operator+(const velocity_type& l, const velocity_type& r) {
            return {l.dpos + r.dpos, l.dvel + r.dvel};
        }
    };
    double time = 0, pos = 0, vel = 1;

    velocity_type initial_velocity_storage() const { return {0, 0}; }
    void put_velocity(velocity_type& k) const { k = {vel, -pos}; }
    void update(const velocity_type& k, double h) {
        pos += k.dpos * h;
        vel += k.dvel * h;
        time += h;
    }
    double get_update_err(const velocity_type& k, double h) const {
        const double inc = std::sqrt(.5 * (k.dpos * h * (k.dpos * h) + k.dvel * h * (k.dvel * h)));
        const double mag = std::sqrt(.5 * (pos * pos + vel * vel));
        return mag < std::numeric_limits<double>::epsilon() ? inc : inc / mag;
    }
};

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    std::ifstream f(argv[1], std::ios::binary);
    std::uint64_t n = 0;
    int bad = 0;
    double worst = 0;
    f.read(reinterpret_cast<char*>(&n), 8);
    {
        Oscillator s;
        emme::RungeKutta3<Oscillator> rk(s, 1.e-5, 1.e-7);
        for (std::uint64_t i = 0; i < n; ++i) {
            rk.step(0.01);
            double rec[3];
            f.read(reinterpret_cast<char*>(rec), sizeof(rec));
            if (rec[0] != s.time || rec[1] != s.pos || rec[2] != s.vel) ++bad;
            worst = std::fmax(worst, std::fabs(s.pos - std::sin(s.time)));
        }
        std::printf("fixed: %llu steps, mismatches %d, max |x - sin t| %.3e\n", (unsigned long long)n, bad, worst);
        if (n != 1000 || worst > 1e-5) ++bad;
    }
    f.read(reinterpret_cast<char*>(&n), 8);
    {
        Oscillator s;
        emme::RungeKutta3<Oscillator> rk(s, 1.e-5, 1.e-7);
        std::uint64_t c = 0;
        double worst2 = 0;
        while (s.time < 10) {
            const double dt = rk.step_adaptive();
            double rec[4] = {0, 0, 0, 0};
            if (c < n) f.read(reinterpret_cast<char*>(rec), sizeof(rec));
            if (c >= n || rec[0] != dt || rec[1] != s.time || rec[2] != s.pos || rec[3] != s.vel) ++bad;
            worst2 = std::fmax(worst2, std::fabs(s.pos - std::sin(s.time)));
            ++c;
        }
        std::printf("adaptive: %llu steps (reference %llu), mismatches %d, max |x - sin t| %.3e\n",
                    (unsigned long long)c, (unsigned long long)n, bad, worst2);
        if (c != n || worst2 > 1e-5) ++bad;
    }
    return bad ? 1 : 0;
}
