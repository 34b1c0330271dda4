// dense.h -- host-side launcher interface of kernel 2 (dense.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "peer.h"

#define DENSE_NB 32

namespace emme {
// device scratch needed by launch_trace_solve for a dim x dim system
size_t dense_workspace_bytes(int dim);
// trace(W^-1 B) -> *d_trace (double2 on device); W and B (dim x dim complex128, row-major)
// are destroyed.  *d_info (device int): 0 or k > 0 if pivot k is exactly zero.
// optimistic != 0: factor without row interchanges and raise *d_flag (device int) if partial
// pivoting would have interchanged anything -- the caller must then repeat with optimistic = 0.
cudaError_t launch_trace_solve(void* W, void* B, int dim, void* workspace, void* d_trace,
                               int* d_info, cudaStream_t stream, unsigned long long* n_launches,
                               int optimistic, int* d_flag, int nrhs = -1);
// nrhs < 0 (default): all dim right-hand sides, back substitution and trace.
// nrhs >= 0: factor W and forward-eliminate only the first nrhs columns of B; no back
// substitution -- follow with launch_solve_factored for complete solves against the factors.
cudaError_t launch_solve_factored(const void* W, void* B, int dim, int nrhs, void* workspace,
                                  int optimistic, cudaStream_t stream, unsigned long long* n_launches);
cudaError_t launch_conj_normalise(const void* X, int ld, int dim, void* rhs, int rhs_ld, double* norm_out,
                                  int conjugate, cudaStream_t stream);
// ---- symmetric path (complex symmetric A, no interchanges): trace(A^-1 B) from
// A^-1 = M^T D^-1 M with A = L D L^T, M = L^-1 -- 4 dim^3 real flops instead of 26/3 dim^3.
// *d_flag bits: 1 = partial pivoting would have interchanged rows, 2 = A is not symmetric; either
// one means the result must be discarded and the step repeated with launch_trace_solve.
size_t dense_sym_workspace_bytes(int dim);
cudaError_t launch_sym_copy_check(const void* A, void* W, int dim, int* d_flag, cudaStream_t stream,
                                  unsigned long long* n_launches);
// Column-sharded variant: the ranks of one box share the step.  W, Y, ws are the
// same three buffers on every rank (peer mappings, index = rank; own entry = local pointer), flags
// the flag pages; `serial` is this dense step's number (identical on every rank, increasing), *epoch
// the rank's barrier counter (all ranks pass the same barriers in the same order).
struct DensePeers {
    PeerFlags flags;
    void* W[EMME_MAX_PEERS];
    void* Y[EMME_MAX_PEERS];
    void* ws[EMME_MAX_PEERS];
    int n, me;
    unsigned long long serial;
    unsigned long long* epoch;
};
// a high-priority side stream and two events (owned by the handle): the panel chain of the blocked
// symmetric path runs there, concurrently with the bulk update on the main stream
struct DenseAux {
    cudaStream_t side;
    cudaEvent_t ev_panel, ev_rest;
};
cudaError_t launch_trace_sym(void* W, void* Y, void* YT, const void* B, int dim, void* sym_workspace,
                             void* d_trace, int* d_info, int* d_flag, cudaStream_t stream,
                             unsigned long long* n_launches, const DensePeers* peers = nullptr,
                             const DenseAux* aux = nullptr);
// outer block width the symmetric path uses for this size (NB = single level)
int dense_sym_outer_block(int dim);
void dense_set_pivot_threshold(double tau);
// Ad = (A - Aold)/delta over n complex entries
cudaError_t launch_secant(const void* A, const void* Aold, void* Ad, size_t n, double dr,
                          double di, int sms, cudaStream_t stream);
cudaError_t dense_prepare();
// load all kernels of dense.cu now (lazy loading must not happen under a device-side peer wait)
cudaError_t dense_preload();
void dense_force_grid_panel(bool on);
void dense_set_outer_block(int nbo);
// measured DFMA throughput (TFLOP/s) of the current device
cudaError_t measure_fp64_peak(double* tflops);
}  // namespace emme
