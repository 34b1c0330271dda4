// eigen_solver.cpp -- see eigen_solver.hpp.
#include "eigen_solver.hpp"

#include <stdexcept>

namespace emme {

void EigenSolver::check(int rc) const {
    if (rc != 0) throw std::runtime_error(emme_last_error());
}

EigenSolver::EigenSolver(const Parameters& para_input, value_type eigen_init, int device)
    : para(para_input),
      eigen_value(0.99 * eigen_init),
      d_eigen_value(0.01 * eigen_init),
      dim((unsigned)para_input.dim()),
      host_matrix_(0, 0) {
    std::vector<double> eta, g, b;
    para.tables(eta, g, b);
    const emme_params pod = para.to_pod();
    check(emme_create(&pod, para.npoints, eta.data(), g.data(), b.data(), device, &h_));
    check(emme_seed(h_, eigen_init.real(), eigen_init.imag()));
    double wr, wi, dr, di;
    check(emme_get_eigen_value(h_, &wr, &wi, &dr, &di));
    eigen_value = {wr, wi};
    d_eigen_value = {dr, di};
}

EigenSolver::~EigenSolver() { emme_destroy(h_); }

void EigenSolver::matrixAssembler(matrix_type& mat) {
    if (mat.getRows() != dim || mat.getCols() != dim)
        throw std::runtime_error("Matrix dimension and grid length mismatch.");
    check(emme_assemble(h_, eigen_value.real(), eigen_value.imag(), mat.data()));
}

void EigenSolver::newtonTraceSecantIteration() {
    double wr, wi, dr, di;
    const int rc = emme_newton_trace_step(h_, &wr, &wi, &dr, &di);
    // like the reference, eigen_value is updated before the info check throws
    double a, b, c, d;
    if (emme_get_eigen_value(h_, &a, &b, &c, &d) == 0) {
        eigen_value = {a, b};
        d_eigen_value = {c, d};
    }
    check(rc);
}

void EigenSolver::newtonQRSecantIteration() {
    double wr, wi, dr, di;
    const int rc = emme_newton_qr_step(h_, &wr, &wi, &dr, &di);
    double a, b, c, d;
    if (emme_get_eigen_value(h_, &a, &b, &c, &d) == 0) {
        eigen_value = {a, b};
        d_eigen_value = {c, d};
    }
    check(rc);
}

std::vector<EigenSolver::value_type> EigenSolver::nullSpace() {
    std::vector<value_type> v(dim);
    check(emme_null_space(h_, v.data()));
    return v;
}

const EigenSolver::matrix_type& EigenSolver::eigen_matrix() {
    if (host_matrix_.getRows() != dim) host_matrix_ = matrix_type(dim, dim);
    check(emme_copy_matrix(h_, 0, host_matrix_.data()));
    return host_matrix_;
}

emme_stats EigenSolver::stats() const {
    emme_stats st{};
    emme_get_stats(h_, &st);
    return st;
}

}  // namespace emme
