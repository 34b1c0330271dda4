// parameters.cpp -- see parameters.hpp.
//
// The geometry functions are node-only (they depend on one eta), so the GPU path evaluates
// them once per grid node on the host and ships three tables.  Bit-identity with the
// reference matters here: beta_1 is a DIFFERENCE of neighbouring g values, so one ulp in g
// is amplified by g/(g-g') before it reaches the integrand.  Every expression below therefore
// keeps the reference's association order, including its quirks:
//   * std::pow(x, 3 / 2) has the INTEGER exponent 1          (src/Parameters.cpp:81)
//   * mh / lh is an integer division                           (src/Parameters.cpp:222)
#include "parameters.hpp"

#include <cmath>
#include <stdexcept>
#include <string>

namespace emme {

std::unique_ptr<Parameters> Parameters::generate(const json::Value& input) {
    const std::string conf = input.at("conf").as_string();
    if (conf == "tokamak") return std::unique_ptr<Parameters>(new Parameters(input));
    if (conf == "stellarator") return std::make_unique<Stellarator>(input);
    if (conf == "cylinder") return std::make_unique<Cylinder>(input);
    if (conf == "taloyMagneticDrift") return std::make_unique<TaylorMagneticDrift>(input);
    if (conf == "cylinder old") return std::make_unique<CylinderOld>(input);
    throw std::runtime_error("Input configuration not supported yet.");
}

// key order = the reference constructor's member-initialiser order (src/Parameters.cpp:36-66),
// which decides WHICH missing key is reported first
Parameters::Parameters(const json::Value& input)
    : q(input.at("q")),
      shat(input.at("shat")),
      tau(input.at("tau")),
      epsilon_n(input.at("epsilon_n")),
      epsilon_r(input.at("epsilon_r")),
      eta_i(input.at("eta_i")),
      eta_e(input.at("eta_e")),
      b_theta(input.at("k_rho").number() * input.at("k_rho").number()),
      beta_e(input.at("beta_e")),
      R(input.at("R")),
      vt(input.at("vt")),
      omega_d_coeff(input.at("omega_d_coeff")),
      length(input.at("length")),
      theta(input.at("theta")),
      npoints((int)input.at("npoints").number()),
      iteration_step_limit((int)input.at("iteration_step_limit").number()),
      integration_precision(input.at("integration_precision")),
      integration_accuracy(input.at("integration_accuracy")),
      integration_iteration_limit((int)input.at("integration_iteration_limit").number()),
      integration_start_points((int)input.at("integration_start_points").number()),
      arc_coeff(input.at("arc_coeff")),
      alpha(q * q * R * beta_e / (epsilon_n * R) * ((1 + eta_e) + 1 / tau * (1 + eta_i))),
      water_bag_weight_vpara(input.at("water_bag_weight_vpara")),
      water_bag_weight_vperp(input.at("water_bag_weight_vperp")),
      omega_s_i(-(std::sqrt(b_theta) * vt) / (epsilon_n * R)),
      omega_s_e(-tau * omega_s_i),
      omega_d_bar(2.0 * epsilon_n * omega_s_i * omega_d_coeff),
      drift_center_transformation_switch(
          input.at("drift_center_transformation_switch").as_boolean()) {}

bool Parameters::electromagnetic() const { return std::fpclassify(beta_e) != FP_ZERO; }

// tokamak s-alpha geometry (src/Parameters.cpp:76-85)
double Parameters::g_integration_f(double eta) const {
    const double ce = std::cos(eta), se = std::sin(eta);
    const double shear_term =
        (1 - shat) * q * epsilon_r / std::pow((std::pow(epsilon_r, 2) + std::pow(q, 2)), 3 / 2) * eta;
    return -((alpha * eta) / 2.0) + shat * theta * ce - shat * eta * ce + se + shat * se +
           0.25 * alpha * std::sin(2.0 * eta) - shear_term;
}

double Parameters::bi(double eta) const {
    return b_theta * (1.0 + std::pow(shat * (eta - theta) - alpha * std::sin(eta), 2));
}

double Parameters::beta_1(double eta, double eta_p) const {
    return (q * R) / vt * (omega_d_bar) * (g_integration_f(eta) - g_integration_f(eta_p));
}

double Parameters::beta_1_e(double eta, double eta_p) const {
    return (q * R) / vt * (omega_d_bar * omega_s_e / omega_s_i) *
           (g_integration_f(eta) - g_integration_f(eta_p));
}

emme_params Parameters::to_pod() const {
    emme_params p{};
    p.q = q;
    p.R = R;
    p.vt = vt;
    p.tau = tau;
    p.beta_e = beta_e;
    p.eta_i = eta_i;
    p.eta_e = eta_e;
    p.omega_s_i = omega_s_i;
    p.omega_s_e = omega_s_e;
    p.omega_d_bar = omega_d_bar;
    p.arc_coeff = arc_coeff;
    p.integration_precision = integration_precision;
    p.integration_accuracy = integration_accuracy;
    p.integration_iteration_limit = integration_iteration_limit;
    p.integration_start_points = integration_start_points;
    p.dx = Grid(length, (unsigned)npoints).dx;
    return p;
}

void Parameters::tables(std::vector<double>& eta, std::vector<double>& g,
                        std::vector<double>& b) const {
    Grid grid(length, (unsigned)npoints);
    eta = grid.grid;
    g.resize(npoints);
    b.resize(npoints);
    for (int i = 0; i < npoints; ++i) {
        g[i] = g_integration_f(eta[i]);
        b[i] = bi(eta[i]);
    }
}

Grid::Grid(double leni, unsigned npointsi)
    : len(leni), npoints(npointsi), dx((2 * leni) / (npoints - 1)), grid(npoints) {
    for (unsigned i = 0; i < npoints; i++) grid[i] = -len + i * dx;
}

// ------------------------------------------------------------------------- stellarator
Stellarator::Stellarator(const json::Value& input)
    : Parameters(input),
      eta_k(input.at("eta_k")),
      lh((int)input.at("lh").number()),
      mh((int)input.at("mh").number()),
      epsilon_h_t(input.at("epsilon_h_t")),
      alpha_0(input.at("alpha_0")),
      r_over_R(input.at("r_over_R")),
      deltap(-0.25 * alpha),
      beta_e_p(beta_e * (1.0 + eta_e) / (epsilon_n * R)),
      rdeltapp((-alpha + (2.0 * shat - 3) * deltap)),
      curvature_aver(mh / lh * r_over_R / (q * R) * (4.0 - shat) +
                     (-alpha + 2 * shat * deltap + 0) / R) {}

double Stellarator::sigma_f(double eta) const {
    return shat * (eta - eta_k) + (deltap * (1 + shat) + rdeltapp) * std::sin(eta);
}

double Stellarator::bi(double eta) const { return b_theta * (1.0 + std::pow(sigma_f(eta), 2)); }

namespace {
// One product term of the stellarator drift integral, multiplied strictly left to right in
// the canonical factor order the reference uses for every one of its 70 monomials:
//   [coef] [deltap] [eps_h] [lh^a] [mh^b] [q^c] [rdeltapp] [shat] [trig]
struct Mono {
    signed char sign;  // +1 / -1 : how the term enters the running sum
    signed char coef;  // 0 = absent
    bool dp, eh;       // deltap, epsilon_h_t prefix
    signed char a, b, c;  // powers of lh, mh, q (0 = absent)
    bool rd, sh;       // rdeltapp, shat suffix
    signed char trig;  // 0 sin(eta) 1 sin(2eta) 2 sin(A) 3 sin(B)
};

// powers pattern shared by the "cubic" blocks: (coef, a, b, c, sign for sin(A), sign for sin(B))
struct Pat {
    signed char coef, a, b, c, sA, sB;
};
constexpr Pat kHel[7] = {{0, 3, 0, 0, +1, -1}, {0, 4, 0, 0, -1, -1}, {2, 2, 1, 1, -1, +1},
                         {3, 3, 1, 1, +1, +1}, {0, 1, 2, 2, +1, -1}, {3, 2, 2, 2, -1, -1},
                         {0, 1, 3, 3, +1, +1}};
// sin(eta) block: +-coef * lh^a mh^b q^c
constexpr Pat kS1[8] = {{2, 2, 0, 0, -1, 0}, {2, 4, 0, 0, +1, 0}, {4, 1, 1, 1, +1, 0},
                        {8, 3, 1, 1, -1, 0}, {2, 0, 2, 2, -1, 0}, {12, 2, 2, 2, +1, 0},
                        {8, 1, 3, 3, -1, 0}, {2, 0, 4, 4, +1, 0}};
// sin(2 eta) block
constexpr Pat kS2[4] = {{0, 1, 1, 1, -1, 0}, {2, 3, 1, 1, +1, 0}, {3, 2, 2, 2, -1, 0},
                        {2, 1, 3, 3, +1, 0}};
}  // namespace

// Drift integral of the helical geometry (src/Parameters.cpp:248-393): a sum of 74 terms over
// a common denominator.  Restated as data (the monomial tables above) plus the four
// structured terms; evaluation order -- left-to-right products, left-to-right sum -- is the
// reference's, so the value is bit-identical (tests/test_host.py::test_tables_*).
double Stellarator::g_integration_f(double eta) const {
    const double hq = lh - mh * q;        // lh - mh*q
    const double hm = -1 + lh - mh * q;   // -1 + lh - mh*q
    const double hp = 1 + lh - mh * q;    //  1 + lh - mh*q
    const double s1 = std::sin(eta), s2 = std::sin(2 * eta);
    const double sA = std::sin(eta + eta * lh - alpha_0 * mh - eta * mh * q);
    const double sB = std::sin(eta - eta * lh + alpha_0 * mh + eta * mh * q);
    const double trig[4] = {s1, s2, sA, sB};

    auto mono = [&](const Mono& t) {
        double v = 0.0;
        bool have = false;
        auto mul = [&](double f) {
            v = have ? v * f : f;
            have = true;
        };
        if (t.coef) mul((double)t.coef);
        if (t.dp) mul(deltap);
        if (t.eh) mul(epsilon_h_t);
        // literal exponents, as in the reference, so the compiler treats pow(x, 2) identically
        auto ipow = [](double x, int k) {
            switch (k) {
                case 1: return x;
                case 2: return std::pow(x, 2);
                case 3: return std::pow(x, 3);
                default: return std::pow(x, 4);
            }
        };
        if (t.a) mul(ipow((double)lh, t.a));
        if (t.b) mul(ipow((double)mh, t.b));
        if (t.c) mul(ipow(q, t.c));
        if (t.rd) mul(rdeltapp);
        if (t.sh) mul(shat);
        mul(trig[t.trig]);
        return v;
    };

    // term 1 and 2
    double acc = eta * hm * std::pow(hq, 2) * hp *
                 (deltap + curvature_aver * R + rdeltapp + deltap * shat);
    acc -= 2 * epsilon_h_t * (eta - eta_k) * lh * hm * hq * hp * shat *
           std::cos(eta * lh - alpha_0 * mh - eta * mh * q);
    // sin(eta) block, without and with shat
    for (int pass = 0; pass < 2; ++pass)
        for (const Pat& p : kS1)
            acc += p.sA * mono(Mono{1, p.coef, false, false, p.a, p.b, p.c, false, pass == 1, 0});
    // cos(eta) * ( ... ) term
    {
        const double poly = (-std::pow(lh, 2) + std::pow(lh, 4) - std::pow(mh, 2) * std::pow(q, 2) +
                             std::pow(mh, 4) * std::pow(q, 4));
        acc += std::cos(eta) * (-2 * (eta - eta_k) * hm * std::pow(hq, 2) * hp * shat -
                                poly * (deltap + rdeltapp + deltap * shat) * s1);
    }
    // sin(2 eta) block: deltap-, rdeltapp-, deltap*shat- weighted
    for (const Pat& p : kS2)
        acc += p.sA * mono(Mono{1, p.coef, true, false, p.a, p.b, p.c, false, false, 1});
    for (const Pat& p : kS2)
        acc += p.sA * mono(Mono{1, p.coef, false, false, p.a, p.b, p.c, true, false, 1});
    for (const Pat& p : kS2)
        acc += p.sA * mono(Mono{1, p.coef, true, false, p.a, p.b, p.c, false, true, 1});
    // helical side bands sin(A) then sin(B): deltap-, rdeltapp-, deltap*shat- weighted
    for (int band = 0; band < 2; ++band) {
        for (const Pat& p : kHel)
            acc += (band ? p.sB : p.sA) *
                   mono(Mono{1, p.coef, true, true, p.a, p.b, p.c, false, false, (signed char)(2 + band)});
        for (const Pat& p : kHel)
            acc += (band ? p.sB : p.sA) *
                   mono(Mono{1, p.coef, false, true, p.a, p.b, p.c, true, false, (signed char)(2 + band)});
        for (const Pat& p : kHel)
            acc += (band ? p.sB : p.sA) *
                   mono(Mono{1, p.coef, true, true, p.a, p.b, p.c, false, true, (signed char)(2 + band)});
    }
    // last term
    acc -= 2 * epsilon_h_t * lh * hm * hp * (hq + shat) * std::sin(alpha_0 * mh - eta * hq);
    return acc / (2. * hm * std::pow(hq, 2) * hp);
}

// ------------------------------------------------------------------------- cylinder & co.
double find_zero_point(double a, double tolerance, int max_iterations) {
    auto f = [a](double x) { return std::cos(x) + a * x * std::sin(x); };
    double lo = 0.0, hi = M_PI;
    if (std::fabs(f(lo)) < tolerance) return lo;
    if (std::fabs(f(hi)) < tolerance) return hi;
    double mid = 0.0;
    for (int it = 0; it < max_iterations; ++it) {
        mid = lo + (hi - lo) / 2.0;
        const double fm = f(mid);
        if (std::fabs(fm) < tolerance || (hi - lo) / 2.0 < tolerance) return mid;
        if (f(lo) * fm < 0) hi = mid; else lo = mid;
    }
    return mid;
}

double calculate_average_value(double a) {
    const double x0 = find_zero_point(a);
    const double integral_value = (1.0 + a) * std::sin(x0) - a * x0 * std::cos(x0);
    return integral_value / x0;
}

Cylinder::Cylinder(const json::Value& input)
    : Parameters(input), shat_coeff(calculate_average_value(shat)) {}

double Cylinder::g_integration_f(double eta) const { return eta * shat_coeff; }
double CylinderOld::g_integration_f(double eta) const { return eta; }

// Pade {3,4} approximant of the magnetic-drift integral (src/Parameters.cpp:404-436)
double TaylorMagneticDrift::g_integration_f(double eta) const {
    const double a = alpha, s = shat;
    const double den = (7 + 16 * a + 40 * std::pow(a, 2) - 28 * s - 80 * a * s + 40 * std::pow(s, 2));
    const double n3 = (-31 - 96 * a - 168 * std::pow(a, 2) - 560 * std::pow(a, 3) + 186 * s +
                       672 * a * s + 1680 * std::pow(a, 2) * s - 504 * std::pow(s, 2) -
                       1680 * a * std::pow(s, 2) + 560 * std::pow(s, 3));
    const double d2 = (3 + 19 * a + 56 * std::pow(a, 2) - 18 * s - 84 * a * s + 28 * std::pow(s, 2));
    const double d4 = (11 - 4 * a + 704 * std::pow(a, 2) - 88 * s - 584 * a * s + 216 * std::pow(s, 2));
    return (eta + (std::pow(eta, 3) * n3) / (42. * den)) /
           (1 + (std::pow(eta, 2) * d2) / (7. * den) + (std::pow(eta, 4) * d4) / (840. * den));
}

}  // namespace emme
