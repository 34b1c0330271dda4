// eigen_solver.hpp -- C++ host mirror of the reference's EigenSolver over the C ABI.
//
// Same constructor arguments, method names and public fields as
// EigenSolver<Matrix<std::complex<double>>> (reference include/solver.h:44-516) so that
// solve_once_eigen (src/main.cpp:19-80) reads the same; the matrices live on the GPU and are
// downloaded on demand.  Errors of the C ABI become the exceptions the reference throws
// (std::runtime_error with the "Linear solve failed." text, include/solver.h:142-153).
#pragma once
#include <complex>
#include <string>
#include <vector>

#include "../../include/emme_b200.h"
#include "parameters.hpp"

namespace emme {

// Row-major dense matrix, the layout contract of the reference's Matrix<T> (include/Matrix.h:43).
template <typename T>
class Matrix {
   public:
    Matrix(std::size_t rows, std::size_t cols) : rows_(rows), cols_(cols), data_(rows * cols) {}
    T& operator()(std::size_t r, std::size_t c) { return data_[r * cols_ + c]; }
    const T& operator()(std::size_t r, std::size_t c) const { return data_[r * cols_ + c]; }
    std::size_t getRows() const { return rows_; }
    std::size_t getCols() const { return cols_; }
    std::size_t size() const { return data_.size(); }
    T* data() { return data_.data(); }
    const T* data() const { return data_.data(); }
    T trace() const {
        T t{};
        for (std::size_t i = 0; i < rows_ && i < cols_; ++i) t += data_[i * cols_ + i];
        return t;
    }

   private:
    std::size_t rows_, cols_;
    std::vector<T> data_;
};

class EigenSolver {
   public:
    using value_type = std::complex<double>;
    using matrix_type = Matrix<value_type>;

    // EigenSolver(para, eigen_init, coeff_matrix, grid_info): the singular-weight matrix is
    // computed on the fly on the device, so only para (and its grid) is needed.  Seeds like the
    // reference constructor (include/solver.h:396-415).
    EigenSolver(const Parameters& para, value_type eigen_init, int device = 0);
    ~EigenSolver();
    EigenSolver(const EigenSolver&) = delete;
    EigenSolver& operator=(const EigenSolver&) = delete;

    void matrixAssembler(matrix_type& mat);   // A(eigen_value) into a host matrix
    void newtonTraceSecantIteration();        // include/solver.h:113-160
    void newtonQRSecantIteration();           // include/solver.h:210-383
    std::vector<value_type> nullSpace();      // include/solver.h:58-112 (inverse iteration)
    const matrix_type& eigen_matrix();        // downloads the current A

    const Parameters& para;
    value_type eigen_value;
    value_type d_eigen_value;
    unsigned int dim;
    emme_stats stats() const;

   private:
    void check(int rc) const;
    emme_solver* h_ = nullptr;
    matrix_type host_matrix_;
};

}  // namespace emme
