// pic_solver.cpp -- see pic_solver.hpp.
#include "pic_solver.hpp"

#include <stdexcept>

namespace emme {

emme_pic_params pic_pod(const Parameters& para) {
    emme_pic_params p{};
    p.q = para.q; p.R = para.R; p.vt = para.vt; p.tau = para.tau;
    p.shat = para.shat; p.b_theta = para.b_theta; p.length = para.length;
    p.eta_i = para.eta_i; p.omega_s_i = para.omega_s_i; p.omega_d_bar = para.omega_d_bar;
    p.water_bag_weight_vpara = para.water_bag_weight_vpara;
    p.water_bag_weight_vperp = para.water_bag_weight_vperp;
    p.npoints = para.npoints;
    p.drift_center_transformation_switch = para.drift_center_transformation_switch ? 1 : 0;
    return p;
}

void PIC_State::check(int rc) const {
    if (rc != 0) throw std::runtime_error(emme_last_error());
}

PIC_State::PIC_State(const Parameters& para, std::size_t marker_num_per_cell, long long seed, int device) {
    const emme_pic_params p = pic_pod(para);
    const long n = (long)(marker_num_per_cell * (std::size_t)para.npoints);
    std::vector<double> eta(n), v_para(n), v_perp(n), weight(2 * (std::size_t)n);
    check(emme_pic_load_markers(&p, n, seed, eta.data(), v_para.data(), v_perp.data(), weight.data()));
    check(emme_pic_create(&p, n, eta.data(), v_para.data(), v_perp.data(), weight.data(), device, &h_));
    field_.resize(para.npoints);
}

PIC_State::~PIC_State() { emme_pic_destroy(h_); }

std::size_t PIC_State::marker_num() const noexcept { return (std::size_t)emme_pic_marker_num(h_); }
long PIC_State::steps_done() const { return emme_pic_steps_done(h_); }

const PIC_State::field_type& PIC_State::current_field() {
    check(emme_pic_current_field(h_, field_.data()));
    return field_;
}

PIC_State::field_type PIC_State::field_history(long first, long count) {
    field_type out((std::size_t)count * field_.size());
    check(emme_pic_field_history(h_, first, count, out.data()));
    return out;
}

std::vector<std::array<double, 3>> PIC_State::field_stats(long first, long count) {
    std::vector<std::array<double, 3>> out(count);
    check(emme_pic_field_stats(h_, first, count, count ? out[0].data() : nullptr));
    return out;
}

void PIC_State::step(double dt, int nsteps) { check(emme_pic_step(h_, dt, nsteps)); }

double PIC_State::last_step_call_ms() const {
    double ms = 0;
    emme_pic_get_timing(h_, &ms, nullptr);
    return ms;
}

std::complex<double> util::calculate_omega(const std::vector<std::array<double, 3>>& stats, double dt) {
    double re = 0, im = 0;
    if (emme_pic_calculate_omega(stats.empty() ? nullptr : stats[0].data(), (long)stats.size(), dt, &re, &im) != 0)
        throw std::runtime_error(emme_last_error());
    return {re, im};
}

}  // namespace emme
