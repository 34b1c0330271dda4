/* emme_b200.h -- C ABI of the B200-native EMME eigen hot path (and, at the end of the file, of
 * the reference's second method, the PIC initial-value run: emme_pic_*).
 *
 * Drop-in boundary for the one data-parallel path of ssskkkky/EMME: the assembly of
 * the eigenmatrix A(omega) and the dense step of the Newton/secant root find.  The
 * reference has no FFI of its own; the entry points below are what a maintainer
 * would bind in place of the C++ members listed next to each of them (file:line in
 * the reference tree).  Plain pointers and sizes only -- no C++, CUDA or torch types.
 *
 * Conventions
 *   - every function returns an int status in the style of LAPACK's `info`
 *     (include/solver.h:121,142-153 of the reference): 0 = ok, <0 = argument -k had an
 *     illegal value, >0 = numerical or device failure (see EMME_E_*).
 *     emme_last_error() gives the text for the calling thread's last failure.
 *   - matrices are complex128, row-major, dim x dim, dim = npoints (beta_e == 0) or
 *     2*npoints -- the layout of Matrix<std::complex<double>> (include/Matrix.h:43) and of
 *     eigenMatrics/*.bin (src/main.cpp:61-63).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails
 *     with EMME_E_NO_DEVICE.
 *   - a handle is not re-entrant (like the reference's EigenSolver); different handles may
 *     be used from different threads / processes / devices.
 */
#ifndef EMME_B200_H
#define EMME_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define EMME_E_NO_DEVICE 1000   /* no usable CUDA device / CUDA runtime error       */
#define EMME_E_SINGULAR_BASE 0  /* 1..dim: pivot k exactly zero (LAPACK info > 0)   */
#define EMME_E_CUDA 1001        /* a CUDA call failed; text in emme_last_error()    */
#define EMME_E_BAD_ORDER 1002   /* integration_start_points not 15 or 31            */
#define EMME_E_STATE 1003       /* call sequence error (e.g. step before seed)      */
#define EMME_E_INPUT 1004       /* input.json error; text in emme_last_error()      */
#define EMME_E_NONFINITE 1005   /* Newton step -1/trace (or omega) is zero-divided / not finite:
                                   "Linear solve failed." like include/solver.h:142-153; omega and
                                   delta are updated as the reference does, nothing is assembled  */
#define EMME_E_PEER 1006        /* multi-GPU exchange: a peer did not arrive (timeout) / bad mapping */

/* The scalars the hot path reads from the reference's Parameters object
 * (include/Parameters.h:15-43) plus the mesh spacing of Grid<double> (include/Grid.h:11). */
typedef struct emme_params {
    double q, R, vt, tau, beta_e;
    double eta_i, eta_e;
    double omega_s_i, omega_s_e, omega_d_bar; /* derived, src/Parameters.cpp:62-64 */
    double arc_coeff;
    double integration_precision; /* tol,   src/Parameters.cpp:178 */
    double integration_accuracy;  /* prec,  src/Parameters.cpp:179 */
    int integration_iteration_limit; /* max bisection depth */
    int integration_start_points;    /* 15 or 31            */
    double dx;                       /* Grid::dx            */
} emme_params;

/* Work and time counters of the most recent assembly on a handle. */
typedef struct emme_stats {
    unsigned long long integrals; /* adaptive quadratures                         */
    unsigned long long panels;    /* Gauss-Kronrod panels                          */
    unsigned long long evals;     /* integrand evaluations                         */
    unsigned long long fwd_trips; /* Miller forward-recurrence trips               */
    unsigned long long bwd_trips; /* Miller backward-recurrence trips              */
    unsigned long long max_stack; /* deepest interval stack seen                   */
    double assemble_ms;           /* CUDA-event time of the last assembly kernel    */
    double dense_ms;              /* CUDA-event time of the last dense step         */
    unsigned long long launches;  /* kernels launched by this handle since creation */
    unsigned long long pivot_fallbacks; /* dense steps repeated with row interchanges   */
    unsigned long long sym_steps; /* dense steps completed on the symmetric (L D L^T) path */
    double dense_flops;           /* real flops of the last dense step (path that ran)     */
} emme_stats;

typedef struct emme_solver emme_solver; /* opaque; owns device memory */

/* ---- device / library ---- */
int emme_device_count(void);
/* Measured FP64 FMA throughput of `device` in TFLOP/s (a register-resident DFMA chain
 * kernel, the denominator of the assembly kernel's roofline; MEASURED_PEAKS.json has no
 * FP64 figure).  Returns 0 on success. */
int emme_fp64_peak(int device, double* tflops, double* sm_mhz_nominal);
const char* emme_last_error(void);
const char* emme_version(void);

/* ---- solver life cycle --------------------------------------------------------------
 * Replaces the construction of Parameters / Grid / SingularityHandler / EigenSolver in
 * solve_once_eigen (src/main.cpp:27-37).  eta, g, bi are npoints doubles each:
 * eta[i] = Grid::grid[i] (include/Grid.h:13), g[i] = para.g_integration_f(eta[i]) and
 * bi[i] = para.bi(eta[i]) (src/Parameters.cpp:76-100 and the geometry overrides); the
 * singular-diagonal weights of SingularityHandler (src/singularity_handler.cpp:3-24) are
 * computed on the fly.  The arrays are copied; the caller keeps ownership. */
int emme_create(const emme_params* p, int npoints, const double* eta, const double* g,
                const double* bi, int device, emme_solver** out);
int emme_destroy(emme_solver* s);
int emme_dim(const emme_solver* s);
/* Replace the per-node tables (a scan over shat, k_rho, ... changes g and bi but not the
 * mesh size): asynchronous host-to-device copies on the handle's stream. */
int emme_set_tables(emme_solver* s, const double* eta, const double* g, const double* bi);
/* Replace the scalar parameters (npoints and the EM/ES kind must not change). */
int emme_set_params(emme_solver* s, const emme_params* p);

/* ---- A(omega): EigenSolver::matrixAssembler (include/solver.h:417-515) ---------------
 * emme_assemble        : into caller's HOST buffer (dim*dim complex128, row-major).
 * emme_assemble_device : into caller's DEVICE buffer on the handle's device; only the
 *   work items k with k % shard_count == shard_index are computed (pairs i<j in
 *   diagonal-major order); with shard_count == 1 the full matrix incl. the
 *   diagonal is written.  Used for multi-GPU assembly: every rank fills its share of a
 *   zero-initialised buffer and the shares are summed/gathered by the caller. */
int emme_assemble(emme_solver* s, double wr, double wi, void* host_out);
int emme_assemble_device(emme_solver* s, double wr, double wi, void* dev_out, int shard_index,
                         int shard_count);

/* ---- Newton / secant root find (include/solver.h:396-415 and :113-160) ---------------
 * emme_seed: the EigenSolver constructor: omega = 0.99*w0, delta = 0.01*w0, A_old =
 *   A(omega), omega += delta, A = A(omega), A' = (A - A_old)/delta.
 * emme_newton_trace_step: newtonTraceSecantIteration: solve A X = A', delta = -1/tr X,
 *   omega += delta, re-assemble, A' = (A - A_old)/delta.  Outputs the new eigen_value and
 *   d_eigen_value.  Returns k > 0 if pivot k is exactly zero (the reference's
 *   "Linear solve failed" runtime_error, include/solver.h:142-153).
 * emme_trace_delta: only the dense part on caller-provided HOST matrices (for tests):
 *   delta = -1/trace(A^-1 Ad). */
int emme_seed(emme_solver* s, double w0r, double w0i);
int emme_newton_trace_step(emme_solver* s, double* wr, double* wi, double* dr, double* di);
/* emme_newton_qr_step: newtonQRSecantIteration (include/solver.h:210-383), the iterate of
 *   iteration_method != "TraceSecant" (src/main.cpp:45-49): Householder QR with column pivoting
 *   A P = Q R (zgeqp3), x = R[0:n-1,0:n-1]^-1 R[0:n-1,n-1] (ztrtrs), v[jpvt[i]] = -x[i],
 *   v[jpvt[n-1]] = 1, delta = -R[n-1][n-1] / (Q^H A' v)[n-1] (zunmqr), omega += delta,
 *   re-assemble, A' = (A - A_old)/delta.  Returns k > 0 if R[k-1][k-1] is exactly zero (the
 *   reference's ztrtrs runtime_error, include/solver.h:308-316).
 * emme_qr_delta: only the dense part on caller-provided HOST matrices (for tests). */
int emme_newton_qr_step(emme_solver* s, double* wr, double* wi, double* dr, double* di);
int emme_qr_delta(emme_solver* s, const void* host_A, const void* host_Ad, double* dr, double* di);
int emme_get_eigen_value(const emme_solver* s, double* wr, double* wi, double* dr, double* di);
int emme_trace_delta(emme_solver* s, const void* host_A, const void* host_Ad, double* dr,
                     double* di);

/* Multi-GPU variant of the two calls above: the assemblies inside seed/step are split
 * into (begin: launch my shard into the handle's A buffer) and (finish: after the caller
 * has completed the matrix in that buffer, e.g. by an NCCL all-reduce/all-gather on the
 * device pointer returned by emme_matrix_device_ptr).  See emme_b200/parallel.py. */
int emme_shard_config(emme_solver* s, int shard_index, int shard_count);
int emme_seed_begin(emme_solver* s, double w0r, double w0i);   /* assemble shard at 0.99 w0 */
int emme_seed_middle(emme_solver* s);                          /* A_old<-A, omega+=delta, assemble shard */
int emme_seed_finish(emme_solver* s);                          /* A' = (A-A_old)/delta */
int emme_step_begin(emme_solver* s);                           /* dense step, omega+=delta, assemble shard */
int emme_qr_step_begin(emme_solver* s);                        /* same for the QR-secant iterate */
int emme_step_finish(emme_solver* s, double* wr, double* wi, double* dr, double* di);
void* emme_matrix_device_ptr(emme_solver* s, int which);
/* Peer group (up to 8 GPUs of one NVSwitch box) -- fused compute + exchange, no collective library.
 * Every rank exports the CUDA IPC handles (64 bytes each) of the buffers below and imports every
 * peer's (its own rank included, with any bytes); once all are mapped
 *   - the assembly kernel stores each entry it computes straight into the matrix of EVERY GPU over
 *     NVLink, bracketed by two stream-ordered device barriers (release/acquire flags in peer memory):
 *     emme_seed / emme_newton_trace_step / emme_newton_qr_step then work unchanged on every rank and
 *     all ranks hold the full matrices and the same eigen_value;
 *   - emme_shard_dense(s, 1) also shards the dense step (include/solver.h:129-140) by column blocks:
 *     panels travel by peer stores and ready flags, the trace is bitwise the single-GPU one.
 * emme_peer_attach does the same for handles that live in ONE process (plain pointers, peer access
 * enabled when the devices differ): several ranks per device, each driven by its own host thread. */
#define EMME_MAX_PEER_RANKS 8
#define EMME_PEER_BUF_MATRIX0 0 /* first physical buffer of eigen_matrix / eigen_matrix_old  */
#define EMME_PEER_BUF_MATRIX1 1 /* second one (they swap roles every iterate)                */
#define EMME_PEER_BUF_FLAGS 2   /* flag page: barrier epochs, panel-ready flags, mailbox      */
#define EMME_PEER_BUF_W 3       /* factorisation work matrix                                  */
#define EMME_PEER_BUF_Y 4       /* M = L^-1                                                   */
#define EMME_PEER_BUF_WS 5      /* per-tile partial traces                                    */
#define EMME_PEER_BUFS 6
int emme_ipc_export(emme_solver* s, int which, void* handle64);
int emme_ipc_import(emme_solver* s, int peer_rank, int peer_count, int which, const void* handle64);
int emme_peer_attach(emme_solver* s, int peer_rank, int peer_count, emme_solver* peer);
int emme_shard_dense(emme_solver* s, int enable);
/* seconds a device-side wait for a peer may spin before the call fails with EMME_E_PEER (default 20) */
int emme_peer_set_timeout(double seconds);

/* EigenSolver::nullSpace (include/solver.h:58-112): the eigenvector, i.e. the right singular
 * vector of eigen_matrix for its smallest singular value (dim complex128 to host_out), by
 * inverse iteration on one LU factorisation instead of the reference's full SVD (zgesdd).
 * Normalised to unit 2-norm with the largest component real positive (the reference's vector
 * carries LAPACK's arbitrary phase).  Overwrites eigen_matrix_derivative. */
int emme_null_space(emme_solver* s, void* host_out);

/* which: 0 = eigen_matrix, 1 = eigen_matrix_old, 2 = eigen_matrix_derivative
 * (the three public matrices of EigenSolver, include/solver.h:392-394); for tests of the dense
 * step also 3 = the factored work matrix (L\U of the last symmetric step), 4 = M = L^-1. */
int emme_copy_matrix(emme_solver* s, int which, void* host_out);
/* Asynchronous variant of emme_copy_matrix (which = 0 or 1 only) for PINNED host memory: the copy is ordered after
 * the work that produced the matrix and runs on its own stream, so it overlaps the next
 * iterate (eigen_matrix is not overwritten before the iterate after next; the handle makes the
 * overwriting assembly wait for the copy).  emme_copy_wait blocks until the data has landed. */
int emme_copy_matrix_async(emme_solver* s, int which, void* pinned_host_out);
int emme_copy_wait(emme_solver* s);
int emme_get_stats(const emme_solver* s, emme_stats* out);
/* CUDA stream the handle launches on (cudaStream_t as void*), for event timing. */
void* emme_stream(emme_solver* s);
int emme_synchronize(emme_solver* s);

/* ---- host side: input.json -> emme_params + tables -----------------------------------
 * Replaces util::json::parse_file + Parameters::generate + Grid (src/main.cpp:184,27-32;
 * src/Parameters.cpp:10-66; include/Grid.h:8-14) for callers that do not have the
 * reference's objects.  Scan objects {head, step, tail} collapse to their head
 * (filter_input, src/main.cpp:174-180). */
typedef struct emme_input emme_input;
int emme_input_load(const char* path, emme_input** out);
int emme_input_parse(const char* json_text, emme_input** out);
void emme_input_free(emme_input* in);
int emme_input_set_number(emme_input* in, const char* key, double value);
int emme_input_get_number(const emme_input* in, const char* key, double* value);
int emme_input_get_string(const emme_input* in, const char* key, char* buf, int buflen);
int emme_input_params(const emme_input* in, emme_params* p, int* npoints);
int emme_input_tables(const emme_input* in, double* eta, double* g, double* bi);

/* ======================================================================================
 * Row N4: the PIC method -- `"method": "PIC"`, solve_once_pic (src/main.cpp:82-137) on
 * PIC_State<double> + Integrator (include/solver_pic.h).  Delta-f markers follow
 * eta' = v_para/(qR) and carry a complex weight; every Runge-Kutta stage is
 * put_velocity (include/solver_pic.h:76-135) + update (:137-151) + solve_field (:251-354),
 * three stages per Integrator::step (:423-434).  On the device one kernel per stage does
 * gather + velocity + push + Bessel/phase + deposit for every marker into per-CTA partial
 * densities; a second small kernel (pic_field_kernel) sums them in a fixed order and turns the
 * density into the new field.
 * ====================================================================================== */

/* The members of Parameters that PIC_State reads. */
typedef struct emme_pic_params {
    double q, R, vt, tau, shat, b_theta, length;
    double eta_i, omega_s_i, omega_d_bar;            /* derived, src/Parameters.cpp:62-64 */
    double water_bag_weight_vpara, water_bag_weight_vperp;
    int npoints;                                      /* field cells                     */
    int drift_center_transformation_switch;
} emme_pic_params;

typedef struct emme_pic emme_pic; /* opaque; owns device memory */

/* Host: PIC_State::initialize_marker (include/solver_pic.h:186-205): n markers drawn from a
 * std::mt19937 with the reference's distributions in the reference's draw order (eta, v_para,
 * v_perp, weight).  seed < 0 seeds from std::random_device like the reference
 * (include/solver_pic.h:356-359); a seed >= 0 makes the loading reproducible.  weight is n
 * complex128 (re, im). */
int emme_pic_load_markers(const emme_pic_params* p, long n, long long seed, double* eta,
                          double* v_para, double* v_perp, double* weight);
/* PIC_State constructor (include/solver_pic.h:62-67) from given markers (copied to the device):
 * marker extras (:207-238), quasi-neutrality table (:381-399), zero field (:239-243). */
int emme_pic_create(const emme_pic_params* p, long n_markers, const double* eta,
                    const double* v_para, const double* v_perp, const double* weight, int device,
                    emme_pic** out);
int emme_pic_destroy(emme_pic* s);
/* nsteps x Integrator::step(dt) (include/solver_pic.h:423-434).  The field after every step is
 * kept on the device (the diagnostics of src/main.cpp:104-110 read it). */
int emme_pic_step(emme_pic* s, double dt, int nsteps);
long emme_pic_steps_done(const emme_pic* s);
long emme_pic_marker_num(const emme_pic* s);
/* PIC_State::current_field (include/solver_pic.h:170): npoints complex128 to host_out. */
int emme_pic_current_field(emme_pic* s, void* host_out);
/* The fields recorded after steps [first, first+count): count*npoints complex128, the byte
 * stream solve_once_pic writes to eigenMatrics/ *.bin (src/main.cpp:106-108). */
int emme_pic_field_history(emme_pic* s, long first, long count, void* host_out);
/* Marker state for parity checks: eta (n doubles) and weight (n complex128); either may be NULL. */
int emme_pic_markers(emme_pic* s, double* eta, double* weight);
/* Derived tables: omega_dv, omega_st, p_weight (n doubles each), quasi-neutrality table
 * (npoints doubles); any may be NULL. */
int emme_pic_extras(emme_pic* s, double* omega_dv, double* omega_st, double* p_weight, double* coef);
/* Per-step diagnostics of src/main.cpp:110-117 for steps [first, first+count): 3 doubles per
 * step = {mean Re phi, mean Im phi, rms |phi|}, accumulated in the reference's order. */
int emme_pic_field_stats(emme_pic* s, long first, long count, double* stats3);
/* util::calculate_omega (include/solver_pic.h:475-529): growth rate from a least-squares line
 * through log rms over the second half, frequency from the maxima of log |mean Re phi|. */
int emme_pic_calculate_omega(const double* stats3, long n, double dt, double* omega_re, double* omega_im);
/* CUDA-event time of the last emme_pic_step call and kernels launched so far. */
int emme_pic_get_timing(const emme_pic* s, double* last_step_call_ms, unsigned long long* launches);
void* emme_pic_stream(emme_pic* s);
/* Multi-GPU (markers sharded over ranks).  emme_pic_create_shard takes ALL n_markers markers (so
 * that the p_weight normalisation over all markers, include/solver_pic.h:229-235, is formed in
 * the reference's order) and keeps the contiguous block shard_index of shard_count on the
 * device.  With shard_count > 1 a stage stops after the deposit: the caller sums the density
 * (emme_pic_density_ptr: 2*npoints doubles on the device) over the ranks between
 * emme_pic_stage_begin and emme_pic_stage_finish, which applies the quasi-neutrality table
 * (and records the field after stage 2).  emme_pic_step does all of this for one rank only. */
int emme_pic_create_shard(const emme_pic_params* p, long n_markers, const double* eta,
                          const double* v_para, const double* v_perp, const double* weight,
                          int shard_index, int shard_count, int device, emme_pic** out);
/* The same for callers that hold only their own block: first = index of the block's first marker in
 * the global order, the four arrays are the BLOCK (n_local markers), pw_sum = sum over ALL markers of
 * the un-normalised p_weight (emme_pic_pweight_sum per block, added over the ranks by the caller:
 * one 8-byte reduction instead of every rank holding every marker). */
int emme_pic_pweight_sum(const emme_pic_params* p, long n, const double* v_para, const double* v_perp, double* sum);
int emme_pic_create_block(const emme_pic_params* p, long n_total, long first, long n_local, const double* eta,
                          const double* v_para, const double* v_perp, const double* weight, double pw_sum,
                          int shard_index, int shard_count, int device, emme_pic** out);
/* Fused density exchange (no collective library): every rank exports the CUDA IPC handle (64 bytes) of
 * its exchange buffer and imports every peer's (own rank included, any bytes); from then on
 * emme_pic_step works on sharded states -- the field kernel of every stage stores the rank's density
 * into every peer's buffer over NVLink, signals per 32-cell block and adds the P contributions in rank
 * order, so every rank holds the same field, bit for bit.  emme_pic_peer_attach: the same for states
 * inside one process.  A peer that does not arrive within 20 s fails the call with EMME_E_PEER. */
int emme_pic_ipc_export(emme_pic* s, void* handle64);
int emme_pic_ipc_import(emme_pic* s, int peer_rank, int peer_count, const void* handle64);
int emme_pic_peer_attach(emme_pic* s, int peer_rank, int peer_count, emme_pic* peer);
int emme_pic_stage_begin(emme_pic* s, double dt, int stage);
int emme_pic_stage_finish(emme_pic* s, int stage);
void* emme_pic_density_ptr(emme_pic* s);
/* input.json -> PIC parameters (keys marker_per_cell, step_number, time_step: src/main.cpp:89-94) */
int emme_input_pic_params(const emme_input* in, emme_pic_params* p, long* marker_per_cell,
                          long* step_number, double* time_step);

#ifdef __cplusplus
}
#endif
#endif /* EMME_B200_H */
