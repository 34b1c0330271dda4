// qr.h -- host-side launcher of the QR-secant dense step (qr.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace emme {
size_t qr_workspace_bytes(int n);
// Householder QR with column pivoting of W (n x n complex128 row-major, destroyed: R above the
// diagonal, reflector vectors below), then the null-vector estimate v and t = Q^H (B v) of
// EigenSolver::newtonQRSecantIteration (reference include/solver.h:210-370).
// d_out2 (two double2 on the device): [0] = R[n-1][n-1], [1] = t[n-1]; the step is -out[0]/out[1].
// *d_info: k > 0 if R[k-1][k-1] is exactly zero (ztrtrs's info).
cudaError_t launch_qr_step(void* W, const void* B, int n, void* workspace, void* d_out2, int* d_info,
                           cudaStream_t stream, unsigned long long* n_launches);
}  // namespace emme
