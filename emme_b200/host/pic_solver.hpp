// pic_solver.hpp -- C++ host mirror of the reference's PIC method over the C ABI (row N4).
//
// Same names as include/solver_pic.h of the reference so that solve_once_pic
// (src/main.cpp:82-137) reads the same: PIC_State(para, marker_per_cell), Integrator(state),
// integrator.step(dt), state.current_field(), util::calculate_omega(stats, dt).  Markers and
// field live on the GPU; the field of every step is recorded there and downloaded on demand.
#pragma once
#include <array>
#include <complex>
#include <vector>

#include "../../include/emme_b200.h"
#include "parameters.hpp"

namespace emme {

emme_pic_params pic_pod(const Parameters& para);

class PIC_State {
   public:
    using value_type = double;
    using complex_type = std::complex<double>;
    using field_type = std::vector<complex_type>;

    // seed < 0: std::random_device like the reference (include/solver_pic.h:356-359)
    PIC_State(const Parameters& para, std::size_t marker_num_per_cell, long long seed = -1, int device = 0);
    ~PIC_State();
    PIC_State(const PIC_State&) = delete;
    PIC_State& operator=(const PIC_State&) = delete;

    std::size_t marker_num() const noexcept;
    const field_type& current_field();              // downloads the field
    field_type field_history(long first, long count);
    std::vector<std::array<double, 3>> field_stats(long first, long count);
    void step(double dt, int nsteps);               // nsteps x Integrator::step
    long steps_done() const;
    double last_step_call_ms() const;

   private:
    void check(int rc) const;
    emme_pic* h_ = nullptr;
    field_type field_;
};

class Integrator {
   public:
    static constexpr std::size_t order = 3;
    explicit Integrator(PIC_State& initial_state) : state(initial_state) {}
    void step(double dt) { state.step(dt, 1); }

   private:
    PIC_State& state;
};

namespace util {
std::complex<double> calculate_omega(const std::vector<std::array<double, 3>>& stats, double dt);
}

}  // namespace emme
