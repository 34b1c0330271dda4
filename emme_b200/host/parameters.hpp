// parameters.hpp -- host mirror of the reference's Parameters hierarchy (include/Parameters.h).
//
// Same member names, same derived constants, same five geometries selected by "conf"
// (reference src/Parameters.cpp:10-34).  Only what the eigen hot path needs is kept as
// behaviour: the per-node geometry functions g_integration_f(eta) and bi(eta), evaluated on
// the host with the reference's own operation order so that the tables handed to the GPU are
// bit-identical to what the reference would compute (tests/test_host.py checks this against
// fixtures dumped from the compiled reference).  The integrand itself lives in the CUDA kernel.
#pragma once
#include <memory>
#include <vector>

#include "../../include/emme_b200.h"
#include "json.hpp"

namespace emme {

struct Parameters {
    static std::unique_ptr<Parameters> generate(const json::Value& input);
    virtual ~Parameters() = default;

    double q, shat, tau, epsilon_n, epsilon_r, eta_i, eta_e, b_theta, beta_e, R, vt,
        omega_d_coeff, length, theta;
    int npoints, iteration_step_limit;
    double integration_precision, integration_accuracy;
    int integration_iteration_limit, integration_start_points;
    double arc_coeff, alpha, water_bag_weight_vpara, water_bag_weight_vperp;
    double omega_s_i, omega_s_e, omega_d_bar;
    bool drift_center_transformation_switch;

    virtual double g_integration_f(double eta) const;
    virtual double bi(double eta) const;
    double beta_1(double eta, double eta_p) const;
    double beta_1_e(double eta, double eta_p) const;

    bool electromagnetic() const;  // beta_e is not FP_ZERO (include/solver.h:407,439)
    int dim() const { return electromagnetic() ? 2 * npoints : npoints; }

    // POD for the C ABI (dx from Grid) and the per-node tables eta/g/bi of length npoints
    emme_params to_pod() const;
    void tables(std::vector<double>& eta, std::vector<double>& g, std::vector<double>& b) const;

   protected:
    explicit Parameters(const json::Value& input);
};

struct Stellarator : Parameters {
    double eta_k;
    int lh, mh;
    double epsilon_h_t, alpha_0, r_over_R, deltap, beta_e_p, rdeltapp, curvature_aver;
    explicit Stellarator(const json::Value& input);
    double sigma_f(double eta) const;
    double bi(double eta) const override;
    double g_integration_f(double eta) const override;
};

struct Cylinder : Parameters {
    double shat_coeff;
    explicit Cylinder(const json::Value& input);
    double g_integration_f(double eta) const override;
};

struct CylinderOld : Parameters {
    explicit CylinderOld(const json::Value& input) : Parameters(input) {}
    double g_integration_f(double eta) const override;
};

struct TaylorMagneticDrift : Parameters {
    explicit TaylorMagneticDrift(const json::Value& input) : Parameters(input) {}
    double g_integration_f(double eta) const override;
};

// Grid<double> (include/Grid.h:8-14)
struct Grid {
    Grid(double len, unsigned npoints);
    double len;
    unsigned npoints;
    double dx;
    std::vector<double> grid;
};

// first zero of cos x + a x sin x on [0, pi] by bisection and the mean of that function
// up to it (reference src/functions.cpp:32-83), used by the "cylinder" geometry only
double find_zero_point(double a, double tolerance = 1e-9, int max_iterations = 100);
double calculate_average_value(double a);

}  // namespace emme
