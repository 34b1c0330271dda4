// qr.cu -- the dense step of EigenSolver::newtonQRSecantIteration (reference
// include/solver.h:210-383), the other `iteration_method` of the eigen solver.
//
// The reference calls LAPACK:  zgeqp3 (A P = Q R, column pivoting, solver.h:246-251), ztrtrs on
// R[0:n-1, 0:n-1] x = R[0:n-1, n-1] (:301-307), forms v[jpvt[i]] = -x[i], v[jpvt[n-1]] = 1
// (:334-337), t = A' v (:343-349), zunmqr t <- Q^H t (:357-363) and takes
//     delta = -R[n-1, n-1] / t[n-1]                                   (:370).
// (A v = R_nn q_n, so delta = -(q_n^H A v)/(q_n^H A' v): the step only depends on WHICH column the
// pivoting leaves last, not on the phase conventions of the reflectors.)
//
// Here the same Householder QR with column pivoting runs on the device, unblocked (LAPACK's
// zlaqp2 recipe: zlarfg reflectors, largest partial column norm first, the partial-norm downdate
// with its sqrt(eps) recomputation safeguard), three launches per column:
//   qr_pivot_kernel   (one CTA)  pivot search, column interchange, reflector of column i
//   qr_vta_kernel     (grid)     partial sums of g = v^H A over row chunks
//   qr_update_kernel  (grid)     A <- A - conj(tau) v g on the trailing columns + norm downdate
// followed by single-CTA kernels for the triangular solve, v, t = A' v and t <- Q^H t.  The matrix
// is row-major like everything else in this library; "column j" is a strided walk.  The path is
// memory bound (the trailing matrix is read twice and written once per column: 16 dim^3 bytes in
// total) and is the secondary iterate of the reference, so no tensor-core blocking is attempted.
#include <cuda_runtime.h>

#include "qr.h"

namespace emme {

typedef double2 z_t;

namespace {

constexpr int QT = 1024;          // threads of the single-CTA kernels
constexpr int CHUNK = 128;        // most rows per CTA in the two grid kernels (32 on small systems)
constexpr int COLS = 128;         // columns per CTA (one thread per column: coalesced along rows)

__device__ __forceinline__ z_t zmul(z_t a, z_t b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ z_t zconj(z_t a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ z_t zdiv(z_t a, z_t b) {   // Smith (LAPACK zladiv without its rescaling)
    if (fabs(b.x) >= fabs(b.y)) {
        const double r = b.y / b.x, d = b.x + b.y * r;
        return make_double2((a.x + a.y * r) / d, (a.y - a.x * r) / d);
    } else {
        const double r = b.x / b.y, d = b.x * r + b.y;
        return make_double2((a.x * r + a.y) / d, (a.y * r - a.x) / d);
    }
}

// block-wide sum of a double over QT threads (result valid in every thread)
__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = 0.;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    return t;
}

// Step i, part 1 (one CTA):
//   * columns whose downdated norm lost its accuracy (flagged by qr_update_kernel with vn1 < 0)
//     get their norm recomputed from rows i.. (zlaqp2: "recompute column norm");
//   * pivot = first column with the largest partial norm among i..n-1 (idamax), interchange of
//     columns i and pvt over ALL rows, of jpvt and of the norms;
//   * zlarfg on column i: beta = -sign(Re alpha) * ||(alpha, x)||, tau = (beta - alpha)/beta,
//     v = x/(alpha - beta) stored below the diagonal, R_ii = beta; v (with v_i = 1) is also written
//     contiguously to vbuf for the two grid kernels.
__global__ void __launch_bounds__(QT)
qr_pivot_kernel(z_t* __restrict__ A, int n, int i, double* __restrict__ vn1, double* __restrict__ vn2,
                int* __restrict__ jpvt, z_t* __restrict__ tau, z_t* __restrict__ vbuf, int* __restrict__ nflag) {
    __shared__ double sh[QT / 32];
    __shared__ int sh_i[QT / 32];
    __shared__ double sh_v[QT / 32];
    __shared__ int s_pvt;
    const int tid = threadIdx.x;
    // ---- flagged columns (rare; counted by qr_update_kernel): recompute ||A[i:, j]|| ----
    const int any_flag = *nflag;
    __syncthreads();
    if (tid == 0) *nflag = 0;
    for (int j = i; j < n && any_flag > 0; ++j) {
        if (vn1[j] >= 0.) continue;           // uniform across the block
        double acc = 0.;
        for (int r = i + tid; r < n; r += QT) {
            const z_t a = A[(size_t)r * n + j];
            acc += a.x * a.x + a.y * a.y;
        }
        const double nrm = sqrt(block_sum(acc, sh));
        __syncthreads();
        if (tid == 0) {
            vn1[j] = nrm;
            vn2[j] = nrm;
        }
        __syncthreads();
    }
    // ---- pivot: first maximum of vn1[i:n] ----
    double bv = -1.;
    int bj = n;
    for (int j = i + tid; j < n; j += QT) {
        const double v = vn1[j];
        if (v > bv) { bv = v; bj = j; }       // ascending j per thread: keeps the first maximum
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double v2 = __shfl_xor_sync(0xffffffffu, bv, o);
        const int j2 = __shfl_xor_sync(0xffffffffu, bj, o);
        if (v2 > bv || (v2 == bv && j2 < bj)) { bv = v2; bj = j2; }
    }
    if ((tid & 31) == 0) { sh_v[tid >> 5] = bv; sh_i[tid >> 5] = bj; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < QT / 32; ++w)
            if (sh_v[w] > bv || (sh_v[w] == bv && sh_i[w] < bj)) { bv = sh_v[w]; bj = sh_i[w]; }
        s_pvt = bj;
    }
    __syncthreads();
    const int pvt = s_pvt;
    if (pvt != i) {
        for (int r = tid; r < n; r += QT) {
            const z_t a = A[(size_t)r * n + i], b = A[(size_t)r * n + pvt];
            A[(size_t)r * n + i] = b;
            A[(size_t)r * n + pvt] = a;
        }
        if (tid == 0) {
            const int t = jpvt[pvt];
            jpvt[pvt] = jpvt[i];
            jpvt[i] = t;
            vn1[pvt] = vn1[i];                // zlaqp2: the norms of column i move to pvt
            vn2[pvt] = vn2[i];
        }
        __syncthreads();
    }
    // ---- zlarfg(n - i, alpha = A[i][i], x = A[i+1:, i]) ----
    double acc = 0.;
    for (int r = i + 1 + tid; r < n; r += QT) {
        const z_t a = A[(size_t)r * n + i];
        acc += a.x * a.x + a.y * a.y;
    }
    const double xnorm = sqrt(block_sum(acc, sh));
    const z_t alpha = A[(size_t)i * n + i];
    z_t t, scale;
    double beta;
    if (xnorm == 0. && alpha.y == 0.) {
        t = make_double2(0., 0.);             // H = I
        scale = make_double2(0., 0.);
        beta = alpha.x;
    } else {
        const double nrm = sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xnorm * xnorm);
        beta = alpha.x >= 0. ? -nrm : nrm;    // -sign(nrm, Re alpha)
        t = make_double2((beta - alpha.x) / beta, -alpha.y / beta);
        scale = zdiv(make_double2(1., 0.), make_double2(alpha.x - beta, alpha.y));
    }
    __syncthreads();                          // everybody has read alpha
    for (int r = i + 1 + tid; r < n; r += QT) {
        const z_t v = zmul(scale, A[(size_t)r * n + i]);
        A[(size_t)r * n + i] = v;
        vbuf[r - i] = v;
    }
    if (tid == 0) {
        A[(size_t)i * n + i] = make_double2(beta, 0.);
        vbuf[0] = make_double2(1., 0.);
        tau[i] = t;
    }
}

// Step i, part 2: gpart[chunk][j] = sum over the chunk's rows r of conj(v_r) A[r][j], j > i.
__global__ void __launch_bounds__(COLS)
qr_vta_kernel(const z_t* __restrict__ A, int n, int i, int chunk, const z_t* __restrict__ vbuf,
              z_t* __restrict__ gpart) {
    __shared__ z_t sv[CHUNK];
    const int j = i + 1 + blockIdx.x * COLS + threadIdx.x;
    const int r0 = i + blockIdx.y * chunk;
    for (int k = threadIdx.x; k < chunk; k += COLS)
        sv[k] = (r0 + k < n) ? vbuf[r0 + k - i] : make_double2(0., 0.);
    __syncthreads();
    if (j >= n) return;
    double gx = 0., gy = 0.;
    const int rend = r0 + chunk < n ? r0 + chunk : n;
#pragma unroll 8
    for (int r = r0; r < rend; ++r) {
        const z_t a = A[(size_t)r * n + j];
        const z_t v = sv[r - r0];
        gx += v.x * a.x + v.y * a.y;          // conj(v) * a
        gy += v.x * a.y - v.y * a.x;
    }
    gpart[(size_t)blockIdx.y * n + j] = make_double2(gx, gy);
}

// Step i, part 3: g_j = sum of the partials (fixed order); A[r][j] -= conj(tau) v_r g_j for the
// chunk's rows; the CTA row that owns row i then downdates the partial norm of column j
// (zlaqp2): temp = max(0, 1 - (|A[i][j]|/vn1_j)^2), temp2 = temp (vn1_j/vn2_j)^2;
// temp2 <= sqrt(eps): the norm must be recomputed (flag: vn1_j = -1, done by the next
// qr_pivot_kernel); otherwise vn1_j *= sqrt(temp).
__global__ void __launch_bounds__(COLS)
qr_update_kernel(z_t* __restrict__ A, int n, int i, int chunk, const z_t* __restrict__ vbuf,
                 const z_t* __restrict__ gpart, int nchunk, const z_t* __restrict__ tau,
                 double* __restrict__ vn1, const double* __restrict__ vn2, int* __restrict__ nflag) {
    __shared__ z_t sv[CHUNK];
    const int j = i + 1 + blockIdx.x * COLS + threadIdx.x;
    const int r0 = i + blockIdx.y * chunk;
    for (int k = threadIdx.x; k < chunk; k += COLS)
        sv[k] = (r0 + k < n) ? vbuf[r0 + k - i] : make_double2(0., 0.);
    __syncthreads();
    if (j >= n) return;
    double gx = 0., gy = 0.;
    for (int c = 0; c < nchunk; ++c) {
        const z_t p = gpart[(size_t)c * n + j];
        gx += p.x;
        gy += p.y;
    }
    const z_t ct = zconj(tau[i]);
    const z_t f = zmul(ct, make_double2(gx, gy));          // conj(tau) g_j
    const int rend = r0 + chunk < n ? r0 + chunk : n;
#pragma unroll 4
    for (int r = r0; r < rend; ++r) {
        const z_t v = sv[r - r0];
        z_t a = A[(size_t)r * n + j];
        a.x -= v.x * f.x - v.y * f.y;
        a.y -= v.x * f.y + v.y * f.x;
        A[(size_t)r * n + j] = a;
        if (r == i) {
            const double old = vn1[j];
            if (old != 0.) {
                const double q = sqrt(a.x * a.x + a.y * a.y) / old;
                double temp = 1.0 - q * q;
                temp = temp > 0. ? temp : 0.;
                const double rr = old / vn2[j];
                const double temp2 = temp * (rr * rr);
                const double tol3z = 1.0536712127723509e-08;   // sqrt(dlamch('Epsilon')), Epsilon = 2^-53
                if (temp2 <= tol3z) {
                    vn1[j] = i + 1 < n ? -1.0 : 0.0;           // recompute from rows i+1.. next step
                    if (i + 1 < n) atomicAdd(nflag, 1);
                } else {
                    vn1[j] = old * sqrt(temp);
                }
            }
        }
    }
}

__global__ void __launch_bounds__(QT)
qr_init_kernel(const z_t* __restrict__ A, int n, double* __restrict__ vn1, double* __restrict__ vn2,
               int* __restrict__ jpvt) {
    // one thread per column: coalesced along rows
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double acc = 0.;
    for (int r = 0; r < n; ++r) {
        const z_t a = A[(size_t)r * n + j];
        acc += a.x * a.x + a.y * a.y;
    }
    const double nrm = sqrt(acc);
    vn1[j] = nrm;
    vn2[j] = nrm;
    jpvt[j] = j;
}

// After the factorisation (one CTA):
//   x = R[0:n-1, 0:n-1]^-1 R[0:n-1, n-1]   (ztrtrs 'U','N','N'; *info = k+1 if R_kk == 0)
//   vfull[jpvt[k]] = -x[k], vfull[jpvt[n-1]] = 1
__global__ void __launch_bounds__(QT)
qr_nullvec_kernel(const z_t* __restrict__ R, int n, const int* __restrict__ jpvt, z_t* __restrict__ x,
                  z_t* __restrict__ vfull, int* __restrict__ info) {
    const int tid = threadIdx.x;
    const int m = n - 1;
    __shared__ z_t s_xk;
    __shared__ int s_bad;
    if (tid == 0) s_bad = 0;
    for (int r = tid; r < m; r += QT) x[r] = R[(size_t)r * n + m];
    __syncthreads();
    // singularity check first, like ztrtrs
    for (int k = tid; k < m; k += QT) {
        const z_t d = R[(size_t)k * n + k];
        if (d.x == 0. && d.y == 0.) atomicMax(&s_bad, k + 1);
    }
    __syncthreads();
    if (s_bad) {
        if (tid == 0) {
            // ztrtrs reports the FIRST zero diagonal element
            int first = 0;
            for (int k = 0; k < m && !first; ++k) {
                const z_t d = R[(size_t)k * n + k];
                if (d.x == 0. && d.y == 0.) first = k + 1;
            }
            *info = first;
        }
        return;
    }
    for (int k = m - 1; k >= 0; --k) {
        if (tid == 0) {
            const z_t xk = zdiv(x[k], R[(size_t)k * n + k]);
            x[k] = xk;
            s_xk = xk;
        }
        __syncthreads();
        const z_t xk = s_xk;
        for (int r = tid; r < k; r += QT) {
            const z_t a = R[(size_t)r * n + k];
            z_t b = x[r];
            b.x -= a.x * xk.x - a.y * xk.y;
            b.y -= a.x * xk.y + a.y * xk.x;
            x[r] = b;
        }
        __syncthreads();
    }
    for (int k = tid; k < m; k += QT) vfull[jpvt[k]] = make_double2(-x[k].x, -x[k].y);
    if (tid == 0) vfull[jpvt[m]] = make_double2(1., 0.);
}

// t = B v (row-major B, one warp per row)
__global__ void __launch_bounds__(256)
qr_gemv_kernel(const z_t* __restrict__ B, int n, const z_t* __restrict__ v, z_t* __restrict__ t) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    double sx = 0., sy = 0.;
    for (int c = lane; c < n; c += 32) {
        const z_t b = B[(size_t)row * n + c], x = v[c];
        sx += b.x * x.x - b.y * x.y;
        sy += b.x * x.y + b.y * x.x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
    }
    if (lane == 0) t[row] = make_double2(sx, sy);
}

// t <- Q^H t = H_n^H ... H_1^H t  with  H_i^H t = t - conj(tau_i) v_i (v_i^H t)   (zunmqr 'L','C');
// out[0] = R[n-1][n-1], out[1] = t[n-1].
__global__ void __launch_bounds__(QT)
qr_apply_qh_kernel(const z_t* __restrict__ QR, int n, const z_t* __restrict__ tau, z_t* __restrict__ t,
                   z_t* __restrict__ out) {
    __shared__ double sh[QT / 32];
    const int tid = threadIdx.x;
    for (int i = 0; i < n; ++i) {
        const z_t ti = tau[i];
        if (ti.x == 0. && ti.y == 0.) continue;
        double sx = 0., sy = 0.;
        for (int r = i + tid; r < n; r += QT) {
            const z_t v = r == i ? make_double2(1., 0.) : QR[(size_t)r * n + i];
            const z_t x = t[r];
            sx += v.x * x.x + v.y * x.y;      // conj(v) * t
            sy += v.x * x.y - v.y * x.x;
        }
        const double dx = block_sum(sx, sh);
        const double dy = block_sum(sy, sh);
        const z_t f = zmul(zconj(ti), make_double2(dx, dy));
        for (int r = i + tid; r < n; r += QT) {
            const z_t v = r == i ? make_double2(1., 0.) : QR[(size_t)r * n + i];
            z_t x = t[r];
            x.x -= v.x * f.x - v.y * f.y;
            x.y -= v.x * f.y + v.y * f.x;
            t[r] = x;
        }
        __syncthreads();
    }
    if (tid == 0) {
        out[0] = QR[(size_t)(n - 1) * n + (n - 1)];
        out[1] = t[n - 1];
    }
}

}  // namespace

// rows per CTA: short loops (more CTAs, fewer dependent memory round trips) while the step is
// latency bound, long ones once it is bandwidth bound
static int qr_chunk(int n) { return n <= 2048 ? 32 : CHUNK; }

size_t qr_workspace_bytes(int n) {
    const size_t nchunk = (n + qr_chunk(n) - 1) / qr_chunk(n);
    // vn1, vn2 | tau, vbuf, x, vfull, t | gpart | jpvt | nflag
    return sizeof(double) * 2 * n + sizeof(z_t) * 5 * (size_t)n + sizeof(z_t) * nchunk * n + sizeof(int) * n +
           512;
}

cudaError_t launch_qr_step(void* Wv, const void* Bv, int n, void* workspace, void* d_out2, int* d_info,
                           cudaStream_t stream, unsigned long long* n_launches) {
    z_t* W = (z_t*)Wv;
    char* p = (char*)workspace;
    auto take = [&](size_t bytes) {
        char* q = p;
        p += (bytes + 15) / 16 * 16;
        return q;
    };
    const int chunk = qr_chunk(n);
    const int nchunk_max = (n + chunk - 1) / chunk;
    double* vn1 = (double*)take(sizeof(double) * n);
    double* vn2 = (double*)take(sizeof(double) * n);
    z_t* tau = (z_t*)take(sizeof(z_t) * n);
    z_t* vbuf = (z_t*)take(sizeof(z_t) * n);
    z_t* x = (z_t*)take(sizeof(z_t) * n);
    z_t* vfull = (z_t*)take(sizeof(z_t) * n);
    z_t* t = (z_t*)take(sizeof(z_t) * n);
    z_t* gpart = (z_t*)take(sizeof(z_t) * (size_t)nchunk_max * n);
    int* jpvt = (int*)take(sizeof(int) * n);
    int* nflag = (int*)take(sizeof(int));
    unsigned long long nl = 0;
    cudaError_t e = cudaMemsetAsync(d_info, 0, sizeof(int), stream);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(nflag, 0, sizeof(int), stream);
    if (e != cudaSuccess) return e;
    qr_init_kernel<<<(n + 255) / 256, 256, 0, stream>>>(W, n, vn1, vn2, jpvt);
    ++nl;
    for (int i = 0; i < n; ++i) {
        qr_pivot_kernel<<<1, QT, 0, stream>>>(W, n, i, vn1, vn2, jpvt, tau, vbuf, nflag);
        ++nl;
        const int ncols = n - i - 1;
        if (ncols > 0) {
            const int nchunk = (n - i + chunk - 1) / chunk;
            dim3 g((ncols + COLS - 1) / COLS, nchunk);
            qr_vta_kernel<<<g, COLS, 0, stream>>>(W, n, i, chunk, vbuf, gpart);
            qr_update_kernel<<<g, COLS, 0, stream>>>(W, n, i, chunk, vbuf, gpart, nchunk, tau, vn1, vn2, nflag);
            nl += 2;
        }
    }
    qr_nullvec_kernel<<<1, QT, 0, stream>>>(W, n, jpvt, x, vfull, d_info);
    qr_gemv_kernel<<<(n + 7) / 8, 256, 0, stream>>>((const z_t*)Bv, n, vfull, t);
    qr_apply_qh_kernel<<<1, QT, 0, stream>>>(W, n, tau, t, (z_t*)d_out2);
    nl += 3;
    if (n_launches) *n_launches += nl;
    return cudaGetLastError();
}

}  // namespace emme
