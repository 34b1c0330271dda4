// integrator.hpp -- the reference's three-stage Runge-Kutta `Integrator<T>`
// (include/solver_pic.h:406-471) for ANY state type, as a host-side template.
//
// The PIC method itself runs the same scheme on the device (emme_b200/csrc/pic.cu uses the same
// coefficient table, EMME_PIC_RK_COEF); this header carries the generic control logic of the
// reference class -- fixed step and the error-controlled `step_adaptive` -- for small host states
// such as the harmonic oscillator of the reference's test/test_integrator.cpp.
//
// Requirements on State (those of the reference, include/solver_pic.h:420-456):
//   using value_type, velocity_type;
//   velocity_type initial_velocity_storage() const;
//   void put_velocity(velocity_type&);
//   void update(const velocity_type&, value_type dt);
//   value_type get_update_err(const velocity_type&, value_type dt);
//   value_type * velocity_type and velocity_type + velocity_type.
#pragma once
#include <array>
#include <cstddef>

namespace emme {

template <typename State>
struct RungeKutta3 {
    using state_type = State;
    using value_type = typename State::value_type;
    using velocity_type = typename State::velocity_type;
    static constexpr std::size_t order = 3;

    explicit RungeKutta3(state_type& initial_state, value_type upper_err_bound = 1.e-7,
                         value_type lower_err_bound = 1.e-10)
        : current_dt(0.1),
          state(initial_state),
          upper_err_bound_(upper_err_bound),
          lower_err_bound_(lower_err_bound),
          intermediates{initial_state.initial_velocity_storage(), initial_state.initial_velocity_storage(),
                        initial_state.initial_velocity_storage()} {}

    // three stages: k_p = f(state); state += (sum_{k<=p} coef[p][k] k_k) * (coef[p][p+1] dt)
    void step(value_type dt) {
        for (std::size_t p = 0; p < order; ++p) {
            state.put_velocity(intermediates[p]);
            state.update(combine(p, p + 1), coef[p][p + 1] * dt);
        }
    }

    // Error-controlled step (include/solver_pic.h:436-456): a trial step with the current dt is
    // accepted when the embedded estimate (row 3 of the table) is below the upper bound, otherwise
    // the state is rolled back and dt halved; an accepted step whose estimate is also below the
    // lower bound doubles dt for the NEXT call.  Returns the dt of the accepted step.
    value_type step_adaptive() {
        const state_type saved = state;
        for (;;) {
            step(current_dt);
            const value_type err = state.get_update_err(combine(order, 3), current_dt);
            if (err < upper_err_bound_) {
                const value_type used = current_dt;
                if (err < lower_err_bound_) current_dt *= 2.;
                return used;
            }
            current_dt *= .5;
            state = saved;
        }
    }

    static constexpr std::array<std::array<double, 4>, 4> coef{{{1, 0.62653829327080},
                                                                {0, 1, -0.55111240553326},
                                                                {0, 1.5220585509963, -0.52205855099628, 0.92457411226246},
                                                                {1., 0.13686116839369, -1.1368611683937}}};

   private:
    // left fold (... + coef[row][k] * intermediates[k]) over k < count, like the reference
    velocity_type combine(std::size_t row, std::size_t count) const {
        velocity_type v = coef[row][0] * intermediates[0];
        for (std::size_t k = 1; k < count; ++k) v = v + coef[row][k] * intermediates[k];
        return v;
    }

    value_type current_dt;
    state_type& state;
    value_type upper_err_bound_;
    value_type lower_err_bound_;
    std::array<velocity_type, 3> intermediates;
};

}  // namespace emme
