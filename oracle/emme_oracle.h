/* oracle/emme_oracle.h -- CPU restatement of the EMME eigen hot path.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.  It is never linked into
 * libemme_b200.so and the product has no CPU fallback.
 *
 * Parity status: PINNED -- checked bit-for-bit against matrices and iterate lists
 * produced by the unmodified reference compiled in the build container
 * (oracle/_ref/ref_driver, fixtures under tests/golden/; see tests/test_oracle.py).
 * The reference's own tests hold no vectors for this path.
 */
#ifndef EMME_ORACLE_H
#define EMME_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

/* The scalars the integrand and the assembler read from the reference's
 * Parameters object (include/Parameters.h:15-43). */
typedef struct emme_oracle_params {
    double q, R, vt, tau, beta_e;
    double eta_i, eta_e;
    double omega_s_i, omega_s_e, omega_d_bar;
    double arc_coeff;
    double tol;      /* integration_precision   (src/Parameters.cpp:178) */
    double prec;     /* integration_accuracy    (src/Parameters.cpp:179) */
    int maxdepth;    /* integration_iteration_limit */
    int order;       /* integration_start_points: 15 or 31 */
} emme_oracle_params;

typedef void (*emme_oracle_fn)(double x, void* ctx, double* re, double* im);

/* util::integrate semi-infinite front end, include/functions.h:305-331.
 * Returns 0, or 1 when order is neither 15 nor 31 (the reference throws). */
int emme_oracle_integrate(emme_oracle_fn f, void* ctx, double tol, double prec, int maxdepth,
                          int order, double* re, double* im, long* evals);

/* util::bessel_i_alter_helper, include/functions.h:381-408; out = {y0,y1,mu+y0,z4} as
 * 8 doubles.  trips[0]/[1] receive forward/backward trip counts when non-NULL. */
void emme_oracle_bessel_i_alter(double zr, double zi, double* out8, int* trips);

/* Parameters::kappa_f_tau, src/Parameters.cpp:113-184.  g/gp = g_integration_f at
 * eta/eta_p, b/bp = bi at eta/eta_p.  stats (may be NULL): {evals, fwd, bwd}. */
void emme_oracle_kappa(const emme_oracle_params* p, unsigned m, double eta, double eta_p,
                       double g, double gp, double b, double bp, double wr, double wi,
                       double* re, double* im, long* stats);

/* Parameters::kappa_f_tau_e, src/Parameters.cpp:186-209 (m in 0..2). */
void emme_oracle_kappa_e(const emme_oracle_params* p, unsigned m, double eta, double eta_p,
                         double g, double gp, double wr, double wi, double* re, double* im);

/* SingularityHandler(n)(i,j), src/singularity_handler.cpp:3-24. */
double emme_oracle_weight(int n, int i, int j);

/* Grid<double>, include/Grid.h:8-14: eta[i] = -len + i*dx; returns dx. */
double emme_oracle_grid(double len, int n, double* eta);

/* EigenSolver::matrixAssembler, include/solver.h:417-515.  out = dim*dim complex128
 * row-major, dim = N (beta_e == 0) or 2N.  row_begin/row_end restrict the pair rows i
 * (for bounded CPU-baseline samples); pass 0, N for the full matrix.
 * stats (may be NULL): {integrals, evals, fwd trips, bwd trips}. */
void emme_oracle_assemble(const emme_oracle_params* p, int N, const double* eta, const double* g,
                          const double* bi, double dx, double wr, double wi, double* out,
                          int row_begin, int row_end, int nthreads, long* stats);

/* The per-iterate dense step of EigenSolver::newtonTraceSecantIteration,
 * include/solver.h:129-140: solve A X = Ad (dim right-hand sides) and return
 * delta = -1/trace(X).  A and Ad (complex128 row-major) are overwritten.  The
 * reference calls LAPACK zsysv; this is an LU with partial pivoting (SURVEY.md A7:
 * same delta to rounding).  Returns 0 or k>0 if pivot k is exactly zero. */
int emme_oracle_trace_step(int dim, double* A, double* Ad, double* dr, double* di);

/* (A - A_old)/delta, include/solver.h:54-57 via include/Arithmetics.h. */
void emme_oracle_secant(long n, const double* A, const double* Aold, double dr, double di,
                        double* out);

#ifdef __cplusplus
}
#endif
#endif
