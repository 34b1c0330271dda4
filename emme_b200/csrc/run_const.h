// run_const.h -- emme_params (C ABI) -> RunConst (kernel constants), host side.
//
// Every derived scalar is formed with the reference's own association order so that the
// kernel multiplies by bit-identical doubles: e.g. beta_1 = (q*R)/vt*omega_d_bar*(g-g')
// (src/Parameters.cpp:87-90) becomes c_beta*(g-g') with c_beta = ((q*R)/vt)*omega_d_bar.
#pragma once
#include <cmath>

#include "../../include/emme_b200.h"
#include "emme_eval.cuh"

namespace emme {

inline RunConst make_run_const(const emme_params& p, int npoints, double wr, double wi) {
    RunConst rc{};
    rc.qR = p.q * p.R;
    rc.vt = p.vt;
    rc.arc = p.arc_coeff;
    rc.inv_arc = 1.0 / p.arc_coeff;
    rc.omega_s_i = p.omega_s_i;
    rc.eta_i = p.eta_i;
    rc.wsi_etai = p.omega_s_i * p.eta_i;
    rc.c_beta = (p.q * p.R) / p.vt * (p.omega_d_bar);
    rc.c_beta_e = (p.q * p.R) / p.vt * (p.omega_d_bar * p.omega_s_e / p.omega_s_i);
    rc.kappa_pref = (p.q * p.R) / (p.vt * std::sqrt((2.0 * M_PI)));
    rc.ke1_pref = (p.q * p.R) / (2.0 * p.vt * p.tau);
    rc.ke2_pref = (p.q * p.q * p.R * p.R) / (2.0 * p.vt * p.vt * p.tau);
    rc.omega_s_e = p.omega_s_e;
    rc.eta_e = p.eta_e;
    rc.diag_es = 1.0 + 1.0 / p.tau;
    rc.em = std::fpclassify(p.beta_e) != FP_ZERO;
    rc.diag_em = rc.em ? (2.0 * p.tau) / p.beta_e : 0.0;
    rc.tol = p.integration_precision;
    rc.prec = p.integration_accuracy;
    const double a = 0.0, b = M_PI / 2.0;  // std::numbers::pi/2.0, include/functions.h:327
    rc.thr_len = 0.99 * (b - a);
    rc.inv_scale = 2. / (b - a);
    rc.half_pi = b;
    rc.dx = p.dx;
    rc.wr = wr;
    rc.wi = wi;
    rc.a0 = wr - p.omega_s_i * (1.0 - 1.5 * p.eta_i);
    rc.omi = -std::copysign(1.0, wr);
    rc.maxdepth = p.integration_iteration_limit;
    rc.order = p.integration_start_points;
    rc.N = npoints;
    return rc;
}

}  // namespace emme
