// dense.cu -- kernel 2: the dense step of the Newton/secant root find, on the device.
//
// Replaces the LAPACK call of EigenSolver::newtonTraceSecantIteration
// (reference include/solver.h:129-140): zsysv solves A X = A' for dim right-hand sides and
// the step is delta = -1/trace(X).  Three paths, chosen by capi.cu:
//   symmetric path (launch_trace_sym, second half of this file) -- EMME's matrix is complex
//      symmetric and diagonally strong: A = L D L^T without interchanges, M = L^-1 carried through
//      the elimination, trace(A^-1 A') = sum_ij (M^T D^-1 M)_ij A'_ji contracted tile by tile.
//      4 dim^3 flops; symmetry and the partial-pivoting criterion are verified on the device;
//   LU paths (launch_trace_solve) -- the fallback for anything else:
//   1. blocked right-looking LU of W = A (row-major complex128), the elimination applied on the
//      fly to B = A' (augmented system), so the forward substitution Y = L^-1 P B costs no
//      extra pass;
//      - panels: optimistic (no interchanges, verified against the partial-pivoting criterion)
//        or pivoting (one thread per matrix row holding its NB panel entries in registers, one
//        cluster/grid barrier per column; the block-local pivot candidate travels with its
//        whole row; row interchanges are implicit and materialise when rows are written back);
//      - one fused kernel applies the interchanges to the other columns of W and to B and
//        solves the NB x NB unit-lower triangle for the block row;
//      - trailing update: complex GEMM on the FP64 tensor cores (DMMA.8x8x4), 64 x 64 tiles,
//        3-stage cp.async ring.
//   2. blocked back substitution X = U^-1 Y restricted to the LOWER triangle of X -- the
//      trace needs X_ii only, and X_lc (l >= c) depends on nothing above the diagonal -- which
//      halves this phase;
//   3. deterministic trace reduction; the secant quotient (A - A_old)/delta
//      (include/solver.h:54-57) is an elementwise kernel.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include "dense.h"
#include "peer.h"

namespace cg = cooperative_groups;

namespace emme {

typedef double2 z_t;

__device__ __forceinline__ z_t zmul(z_t a, z_t b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ void zfms(z_t& c, z_t a, z_t b) {  // c -= a*b
    c.x = fma(-a.x, b.x, c.x);
    c.x = fma(a.y, b.y, c.x);
    c.y = fma(-a.x, b.y, c.y);
    c.y = fma(-a.y, b.x, c.y);
}
__device__ __forceinline__ z_t zrecip(z_t a) {
    // Smith's scaled reciprocal: safe against overflow of |a|^2
    if (fabs(a.x) >= fabs(a.y)) {
        const double r = a.y / a.x, d = 1.0 / (a.x + a.y * r);
        return make_double2(d, -r * d);
    } else {
        const double r = a.x / a.y, d = 1.0 / (a.x * r + a.y);
        return make_double2(r * d, -d);
    }
}
__device__ __forceinline__ z_t zdiv(z_t a, z_t b) {
    if (fabs(b.x) >= fabs(b.y)) {
        const double r = b.y / b.x, d = b.x + b.y * r;
        return make_double2((a.x + a.y * r) / d, (a.y - a.x * r) / d);
    } else {
        const double r = b.x / b.y, d = b.x * r + b.y;
        return make_double2((a.x * r + a.y) / d, (a.y * r - a.x) / d);
    }
}

// ------------------------------------------------------------------ panel factorisation
// One thread per matrix row of the (dim-k0) x jb panel.  The thread keeps the NOT YET
// ELIMINATED part of its row in registers as a window a[0..NB) whose first element is always
// the current column: after every column the window shifts left by one, so the column loop is
// a rolled loop with static register indices (an unrolled-by-column version was instruction-
// fetch bound: 32 copies of the body, each executed once).  Multipliers and finished pivot rows
// are stored to W at the row's ORIGINAL position as they are produced; the row interchanges are
// implicit (each thread tracks its row's current position) and materialise in a final
// read-all / barrier / write-all permutation.
// Per column: block-local pivot candidate -> published together with its whole row window ->
// ONE barrier -> every CTA picks the global winner and copies its row window into local shared
// memory -> eliminate.  Two communication back ends:
//   ClusterComm: the CTAs form one thread-block cluster; candidates live in shared memory and
//                are read through DSMEM; barrier = hardware cluster barrier     (rows <= 4096)
//   GridComm   : candidates in global memory, barrier = cooperative grid sync  (any size)
constexpr int NB = DENSE_NB;      // panel width

struct PanelCand {                // what a CTA publishes per column step
    double val;                   // |re| + |im| of its best candidate (-1: none)
    int pos;                      // current row position of that candidate
    int pad;
    z_t row[NB];                  // the candidate's row window (column c+k at index k)
};

struct ClusterComm {
    cg::cluster_group cluster;
    PanelCand* s_cand;            // [2] in this CTA's shared memory
    __device__ __forceinline__ unsigned rank() const { return cluster.block_rank(); }
    __device__ __forceinline__ unsigned size() const { return cluster.num_blocks(); }
    __device__ __forceinline__ PanelCand* mine(int c) { return s_cand + (c & 1); }
    __device__ __forceinline__ const PanelCand* peer(int c, unsigned b) {
        return cluster.map_shared_rank(s_cand + (c & 1), b);
    }
    __device__ __forceinline__ void sync() { cluster.sync(); }
    static __device__ __forceinline__ double ld(const double* p) { return *p; }
    static __device__ __forceinline__ int ld(const int* p) { return *p; }
};

struct GridComm {
    cg::grid_group grid;
    PanelCand* xchg;              // [2][gridDim.x] in global memory
    __device__ __forceinline__ unsigned rank() const { return blockIdx.x; }
    __device__ __forceinline__ unsigned size() const { return gridDim.x; }
    __device__ __forceinline__ PanelCand* mine(int c) { return xchg + (size_t)(c & 1) * gridDim.x + blockIdx.x; }
    __device__ __forceinline__ const PanelCand* peer(int c, unsigned b) {
        return xchg + (size_t)(c & 1) * gridDim.x + b;
    }
    __device__ __forceinline__ void sync() {
        __threadfence();
        grid.sync();
    }
    static __device__ __forceinline__ double ld(const double* p) { return __ldcg(p); }
    static __device__ __forceinline__ int ld(const int* p) { return __ldcg(p); }
};

template <int TPB, class Comm>
__device__ __forceinline__ void panel_body(Comm& comm, z_t* __restrict__ W, int ld, int dim, int k0,
                                           int jb, int* __restrict__ ipiv, int* __restrict__ info) {
    __shared__ PanelCand s_best;
    __shared__ double s_val[TPB / 32];
    __shared__ int s_thr[TPB / 32];
    __shared__ int s_win;
    const int rows = dim - k0;
    const int r = (int)comm.rank() * TPB + threadIdx.x;
    const bool have = r < rows;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned csz = comm.size();
    z_t* const myrow = W + (size_t)(k0 + r) * ld + k0;   // original position of this thread's row
    int my_pos = k0 + r;
    bool done = !have;

    z_t a[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) a[k] = (have && k < jb) ? myrow[k] : make_double2(0., 0.);

#pragma unroll 1
    for (int c = 0; c < jb; ++c) {
        // ---- block-local pivot candidate: max |re|+|im|, ties -> lowest position ----
        // a NaN candidate compares false against everything and would leave the search without a
        // winner (position 0x7fffffff used as an address): order it as +inf instead, so that the
        // search is total, every position is a real row, and the NaN reaches the trace where the
        // host reports it (include/solver.h:142-153 semantics: an error, never a fault)
        double myv = done ? -1.0 : fabs(a[0].x) + fabs(a[0].y);
        if (myv != myv) myv = __longlong_as_double(0x7ff0000000000000LL);
        double v = myv;
        int p = done ? 0x7fffffff : my_pos;
        int t = threadIdx.x;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double v2 = __shfl_xor_sync(0xffffffffu, v, o);
            const int p2 = __shfl_xor_sync(0xffffffffu, p, o);
            const int t2 = __shfl_xor_sync(0xffffffffu, t, o);
            if (v2 > v || (v2 == v && p2 < p)) { v = v2; p = p2; t = t2; }
        }
        if (lane == 0) { s_val[warp] = v; s_thr[warp] = t; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double bv = s_val[0];
            int bt = s_thr[0];
            for (int w = 1; w < TPB / 32; ++w)
                if (s_val[w] > bv) { bv = s_val[w]; bt = s_thr[w]; }
            s_win = bt;
        }
        __syncthreads();
        if (threadIdx.x == s_win) {
            PanelCand* mine = comm.mine(c);
            mine->val = myv;
            mine->pos = my_pos;
#pragma unroll
            for (int k = 0; k < NB; ++k) mine->row[k] = a[k];
        }
        comm.sync();
        // ---- global pivot: warp 0 scans the per-CTA candidates, copies the winner's window ----
        if (warp == 0) {
            double bv = -2.0;
            int bp = 0x7fffffff, bb = 0;
            for (unsigned b = lane; b < csz; b += 32) {
                const PanelCand* cnd = comm.peer(c, b);
                const double cv = Comm::ld(&cnd->val);
                const int cp = Comm::ld(&cnd->pos);
                if (cv > bv || (cv == bv && cp < bp)) { bv = cv; bp = cp; bb = (int)b; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double v2 = __shfl_xor_sync(0xffffffffu, bv, o);
                const int p2 = __shfl_xor_sync(0xffffffffu, bp, o);
                const int b2 = __shfl_xor_sync(0xffffffffu, bb, o);
                if (v2 > bv || (v2 == bv && p2 < bp)) { bv = v2; bp = p2; bb = b2; }
            }
            const PanelCand* best = comm.peer(c, (unsigned)bb);
            if (lane < NB)
                s_best.row[lane] = make_double2(Comm::ld(&best->row[lane].x), Comm::ld(&best->row[lane].y));
            if (lane == 0) { s_best.val = bv; s_best.pos = bp; }
        }
        __syncthreads();
        const double pv = s_best.val;
        const int diag = k0 + c;
        // defensive: a position outside the matrix can never be adopted as a row
        const int ppos = (s_best.pos >= diag && s_best.pos < dim) ? s_best.pos : diag;
        if (comm.rank() == 0 && threadIdx.x == 0) {
            ipiv[diag] = ppos;
            if (pv == 0.0 && *info == 0) *info = diag + 1;   // exactly singular (LAPACK info > 0)
        }
        // ---- implicit interchange ----
        const bool i_am_pivot = !done && my_pos == ppos;
        if (!done && !i_am_pivot && my_pos == diag) my_pos = ppos;
        if (i_am_pivot) {
            // this row is U's row `diag`: its window is final, store it at the original position
            my_pos = diag;
            done = true;
#pragma unroll
            for (int k = 0; k < NB; ++k)
                if (c + k < jb) myrow[c + k] = a[k];
        }
        // ---- eliminate, then shift the window ----
        if (!done) {
            if (pv > 0.0) {
                const z_t l = zmul(a[0], zrecip(s_best.row[0]));
                myrow[c] = l;
#pragma unroll
                for (int k = 1; k < NB; ++k) zfms(a[k], l, s_best.row[k]);
            } else {
                myrow[c] = a[0];
            }
#pragma unroll
            for (int k = 0; k < NB - 1; ++k) a[k] = a[k + 1];
            a[NB - 1] = make_double2(0., 0.);
        }
        __syncthreads();   // s_best / s_val / s_win are reused by the next column
    }
    // ---- materialise the interchanges: read every moved row, barrier, write it to its place ----
    const bool moved = have && my_pos != k0 + r;
    if (moved) {
#pragma unroll
        for (int k = 0; k < NB; ++k) a[k] = k < jb ? myrow[k] : make_double2(0., 0.);
    }
    comm.sync();
    if (moved) {
        z_t* dst = W + (size_t)my_pos * ld + k0;
#pragma unroll
        for (int k = 0; k < NB; ++k)
            if (k < jb) dst[k] = a[k];
    }
}

// ------------------------------------------------------------------ panel, optimistic variant
// Partial pivoting keeps the diagonal whenever |a_cc| >= |a_ic| for all i > c (LAPACK's izamax
// criterion |re|+|im|, ties to the lowest row).  EMME's matrices are diagonally strong (diagonal
// 1+1/tau resp. 2*tau/beta_e*b_i against O(dx) quadrature entries), so on them partial pivoting
// never interchanges rows.  This kernel factors a panel WITHOUT interchanges -- which removes the
// per-column pivot search and with it every inter-CTA dependency -- and records, for every
// multiplier it forms, whether partial pivoting would have accepted the diagonal
// (|a_ic| <= tau*|a_cc|, tau = 32: threshold pivoting, see g_tau).  If the flag stays clear the factorisation IS
// the partial-pivoting one; if it is raised the host discards the result and repeats the dense
// step with the pivoting kernels above.  Every CTA first factors the jb x jb diagonal block
// redundantly in shared memory (one warp), then each thread eliminates its own row against it.
template <int TPB>
__global__ void __launch_bounds__(TPB)
panel_nopiv_kernel(z_t* __restrict__ W, int ld, int dim, int k0, int jb, double tau,
                   int* __restrict__ flag, int* __restrict__ info) {
    __shared__ z_t sD[NB][NB + 1];     // diagonal block: strict lower = L11, upper = U11
    __shared__ z_t sInv[NB];           // 1/u_cc
    __shared__ double sAbs[NB];        // |u_cc| (cabs1)
    for (int e = threadIdx.x; e < NB * NB; e += TPB) {
        const int rr = e / NB, cc = e % NB;
        sD[rr][cc] = (rr < jb && cc < jb) ? W[(size_t)(k0 + rr) * ld + k0 + cc] : make_double2(0., 0.);
    }
    __syncthreads();
    int bad = 0;
    if (threadIdx.x < 32) {
        const int i = threadIdx.x;     // lane = row of the diagonal block
        for (int c = 0; c < jb; ++c) {
            const z_t u = sD[c][c];
            const double ua = fabs(u.x) + fabs(u.y);
            const z_t inv = ua > 0.0 ? zrecip(u) : make_double2(0., 0.);
            if (i == 0) {
                sInv[c] = inv;
                sAbs[c] = ua;
                if (ua == 0.0 && blockIdx.x == 0) atomicCAS(info, 0, k0 + c + 1);
            }
            if (i > c && i < jb) {
                const z_t num = sD[i][c];
                if (fabs(num.x) + fabs(num.y) > tau * ua) bad = 1;
                const z_t l = zmul(num, inv);
                sD[i][c] = l;
                for (int k = c + 1; k < jb; ++k) zfms(sD[i][k], l, sD[c][k]);
            }
            __syncwarp();
        }
    }
    __syncthreads();
    const int r = blockIdx.x * TPB + threadIdx.x;      // row offset below k0
    if (r < dim - k0) {
        z_t* myrow = W + (size_t)(k0 + r) * ld + k0;
        if (r < jb) {
            // rows of the diagonal block: CTA 0 writes the factored block back
            if (blockIdx.x == 0) {
                for (int c = 0; c < jb; ++c) myrow[c] = sD[r][c];
            }
        } else {
            z_t a[NB];
#pragma unroll
            for (int k = 0; k < NB; ++k) a[k] = k < jb ? myrow[k] : make_double2(0., 0.);
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                if (c < jb) {
                    if (fabs(a[c].x) + fabs(a[c].y) > tau * sAbs[c]) bad = 1;
                    const z_t l = zmul(a[c], sInv[c]);
                    a[c] = l;
#pragma unroll
                    for (int k = c + 1; k < NB; ++k) zfms(a[k], l, sD[c][k]);
                }
            }
#pragma unroll
            for (int k = 0; k < NB; ++k)
                if (k < jb) myrow[k] = a[k];
        }
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flag, 1);
}

constexpr int PT = 128;           // rows per CTA, grid-cooperative variant
constexpr int CL_TPB = 256;       // rows per CTA, cluster variant
constexpr int CLUSTER_MAX = 16;   // non-portable cluster size (8 is the portable limit)

__global__ void __launch_bounds__(PT)
panel_kernel(z_t* __restrict__ W, int ld, int dim, int k0, int jb, int* __restrict__ ipiv,
             PanelCand* __restrict__ xchg, int* __restrict__ info) {
    GridComm comm{cg::this_grid(), xchg};
    panel_body<PT>(comm, W, ld, dim, k0, jb, ipiv, info);
}

template <int TPB>
__global__ void __launch_bounds__(TPB)
panel_cluster_kernel(z_t* __restrict__ W, int ld, int dim, int k0, int jb, int* __restrict__ ipiv,
                     int* __restrict__ info) {
    __shared__ PanelCand s_cand[2];
    ClusterComm comm{cg::this_cluster(), s_cand};
    panel_body<TPB>(comm, W, ld, dim, k0, jb, ipiv, info);
    comm.sync();   // no CTA may exit while others can still read its shared memory
}

// largest cluster size (power of two <= CLUSTER_MAX) the device can co-schedule for this kernel
static int panel_cluster_limit() {
    static int limit_on[64];
    static bool known[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    int& limit = limit_on[dev & 63];
    if (known[dev & 63]) return limit;      // per device: the non-portable cluster attribute is, too
    known[dev & 63] = true;
    limit = 0;
    if (cudaFuncSetAttribute(panel_cluster_kernel<CL_TPB>,
                             cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        cudaGetLastError();
    }
    for (int csz = CLUSTER_MAX; csz >= 1; csz >>= 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(csz);
        cfg.blockDim = dim3(CL_TPB);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = csz;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, panel_cluster_kernel<CL_TPB>, &cfg) == cudaSuccess && n >= 1) {
            limit = csz;
            break;
        }
        cudaGetLastError();
    }
    return limit;
}

static cudaError_t launch_panel_cluster(z_t* W, int ld, int dim, int k0, int jb, int* ipiv, int* info,
                                        cudaStream_t stream) {
    const int rows = dim - k0;
    int csz = 1;
    while (csz * CL_TPB < rows) csz <<= 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csz);
    cfg.blockDim = dim3(CL_TPB);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csz;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, panel_cluster_kernel<CL_TPB>, W, ld, dim, k0, jb, ipiv, info);
}

// ------------------------------------------------------------------ interchanges + block-row solve
// Column sets are given as two ranges: columns [0,n1) start at p1, columns [n1,n1+n2) at p2
// (both with leading dimension ld); one thread per column, coalesced across threads.
__device__ __forceinline__ z_t* column_ptr(z_t* p1, int n1, z_t* p2, int col) {
    return col < n1 ? p1 + col : p2 + (col - n1);
}

// Apply row interchanges kb .. kb+count-1 (row k <-> ipiv[k]), in order.
__global__ void __launch_bounds__(128)
laswp_kernel(z_t* __restrict__ p1, int n1, z_t* __restrict__ p2, int n2, int ld, int kb, int count,
             const int* __restrict__ ipiv) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= n1 + n2) return;
    z_t* M = column_ptr(p1, n1, p2, col);
    for (int c = 0; c < count; ++c) {
        const int p = ipiv[kb + c];
        if (p != kb + c) {
            const z_t t1 = M[(size_t)(kb + c) * ld], t2 = M[(size_t)p * ld];
            M[(size_t)(kb + c) * ld] = t2;
            M[(size_t)p * ld] = t1;
        }
    }
}

// Optionally apply the interchanges of panel (k0, jb), then forward-substitute rows
// k0 .. k0+jb-1 with the unit-lower jb x jb block L11 = W[k0:k0+jb, k0:k0+jb].
__global__ void __launch_bounds__(128)
swap_trsm_kernel(const z_t* __restrict__ W, z_t* __restrict__ p1, int n1, z_t* __restrict__ p2, int n2,
                 int ld, int k0, int jb, const int* __restrict__ ipiv, int do_swap) {
    __shared__ z_t sL[NB][NB + 1];
    __shared__ int sP[NB];
    for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
        const int rr = e / NB, cc = e % NB;
        sL[rr][cc] = (rr < jb && cc < rr) ? W[(size_t)(k0 + rr) * ld + k0 + cc] : make_double2(0., 0.);
    }
    if (threadIdx.x < NB) sP[threadIdx.x] = threadIdx.x < jb ? ipiv[k0 + threadIdx.x] : 0;
    __syncthreads();
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= n1 + n2) return;
    z_t* M = column_ptr(p1, n1, p2, col);
    if (do_swap) {
        for (int c = 0; c < jb; ++c) {
            const int p = sP[c];
            if (p != k0 + c) {
                const z_t t1 = M[(size_t)(k0 + c) * ld], t2 = M[(size_t)p * ld];
                M[(size_t)(k0 + c) * ld] = t2;
                M[(size_t)p * ld] = t1;
            }
        }
    }
    z_t x[NB];
#pragma unroll
    for (int rr = 0; rr < NB; ++rr) x[rr] = rr < jb ? M[(size_t)(k0 + rr) * ld] : make_double2(0., 0.);
#pragma unroll
    for (int rr = 1; rr < NB; ++rr) {
#pragma unroll
        for (int cc = 0; cc < rr; ++cc) zfms(x[rr], sL[rr][cc], x[cc]);
    }
#pragma unroll
    for (int rr = 0; rr < NB; ++rr)
        if (rr < jb) M[(size_t)(k0 + rr) * ld] = x[rr];
}

// Back substitution, block row [k0, k0+jb): X = U_kk^-1 Y for columns [0, ncols).
__global__ void __launch_bounds__(128)
utrsm_kernel(const z_t* __restrict__ W, z_t* __restrict__ B, int ld, int k0, int jb, int ncols) {
    __shared__ z_t sU[NB][NB + 1];
    for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
        const int rr = e / NB, cc = e % NB;
        z_t v = make_double2(0., 0.);
        if (rr < jb && cc < jb && cc >= rr) v = W[(size_t)(k0 + rr) * ld + k0 + cc];
        if (rr == cc) v = rr < jb ? zrecip(v) : make_double2(1., 0.);
        sU[rr][cc] = v;   // diagonal holds 1/u_rr
    }
    __syncthreads();
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncols) return;
    z_t* M = B + col;
    z_t x[NB];
#pragma unroll
    for (int rr = 0; rr < NB; ++rr) x[rr] = rr < jb ? M[(size_t)(k0 + rr) * ld] : make_double2(0., 0.);
#pragma unroll
    for (int rr = NB - 1; rr >= 0; --rr) {
#pragma unroll
        for (int cc = NB - 1; cc > rr; --cc) zfms(x[rr], sU[rr][cc], x[cc]);
        x[rr] = zmul(x[rr], sU[rr][rr]);
    }
#pragma unroll
    for (int rr = 0; rr < NB; ++rr)
        if (rr < jb) M[(size_t)(k0 + rr) * ld] = x[rr];
}

// ------------------------------------------------------------------ complex GEMM  C -= A*B
// C[M x N] -= A[M x K] * Bm[K x N], all row-major with their own leading dimensions.
// 64x64 tile per CTA, 256 threads, 4x4 complex accumulators per thread, K slabs of 16 staged
// through shared memory.  FP64-pipe bound by construction (64 DFMA per 8 LDS.128).
constexpr int GM = 64, GN = 64, GK = 16, GSTAGES = 3;
constexpr int G_LDA = GK + 4;    // sA[m][k]: row stride = 4 (mod 8) elements -> conflict-free fragment loads
constexpr int G_LDB = GN + 2;    // sB[k][n]: row stride = 2 (mod 8) elements
constexpr int G_STAGE_ELEMS = GM * G_LDA + GK * G_LDB;                  // z_t per pipeline stage
constexpr size_t G_SMEM_BYTES = sizeof(z_t) * G_STAGE_ELEMS * GSTAGES;  // 110 KB: 2 CTAs per SM

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = pred ? 16 : 0;   // src-size 0: nothing is read, the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// FP64 tensor-core multiply-accumulate: D(8x8) += A(8x4) * B(4x8)   (SASS: DMMA.8x8x4)
// lane T holds A[T/4][T%4], B[T%4][T/4] and D[T/4][2*(T%4) + {0,1}].
__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d[0]), "+d"(d[1])
                 : "d"(a), "d"(b));
}

// C[M x N] -= A[M x K] * Bm[K x N] on one 64x64 complex tile (256 threads = 8 warps, 4 x 2, each
// warp a 16 x 32 sub-tile = 2 x 4 DMMA blocks).  The complex product is four real tensor-core
// products per block (re*re, -im*im, re*im, im*re) on the interleaved (re, im) data: one 16-byte
// shared-memory load delivers both fragment values.  On B200 the FP64 tensor rate equals the DFMA
// rate (37 TFLOP/s measured for both), but a DMMA carries 256 FMAs per issue slot instead of 32,
// so the FP64 pipe stays fed (the DFMA version of this tile stalled at 60-64 % pipe utilisation).
// k-slabs of 16 travel global -> shared memory with cp.async through a 3-stage ring.
struct GemmAcc {
    double re[2][4][2], im[2][4][2];
};

// acc = A[m0:m0+64, 0:K) * Bm[0:K, n0:n0+64) (rows beyond M / columns beyond N read as zero).
// SUB: acc = Cin[tile] - A*B instead -- the accumulators START as the C tile (its loads are issued
// right after the first operand stages, so their HBM latency hides behind the pipeline fill) and the
// products are subtracted as they come; the caller's epilogue is then a plain store.  (The first
// version formed A*B from zero and read-modified-wrote C afterwards: the exposed load latency of
// that epilogue was 10-15 % of a K = 256 tile.)
template <bool SUB>
__device__ __forceinline__ void zgemm_mainloop_t(GemmAcc& acc, const z_t* __restrict__ A, int lda,
                                                 const z_t* __restrict__ Bm, int ldb, int M, int N,
                                                 int K, int m0, int n0, const z_t* __restrict__ Cin, int ldc) {
    extern __shared__ __align__(16) unsigned char g_smem_raw[];
    z_t* smem = reinterpret_cast<z_t*>(g_smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;          // warp grid 4 (M) x 2 (N)
    // staging ownership: A slab = 64 rows x 16 k, B slab = 16 k x 64 cols; 4 + 4 elements/thread
    const int ak = threadIdx.x & 15, am = threadIdx.x >> 4;      // A: rows am, am+16, am+32, am+48
    const int bn = threadIdx.x & 63, bk = threadIdx.x >> 6;      // B: k rows bk, bk+4, bk+8, bk+12
    auto issue = [&](int slab) {
        z_t* sA = smem + (size_t)(slab % GSTAGES) * G_STAGE_ELEMS;   // [GM][G_LDA]
        z_t* sB = sA + GM * G_LDA;                                   // [GK][G_LDB]
        const int kk = slab * GK;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int m = am + 16 * e;
            const bool ok = (m0 + m < M) && (kk + ak < K);
            const z_t* src = ok ? A + (size_t)(m0 + m) * lda + kk + ak : A;
            cp_async16(sA + m * G_LDA + ak, src, ok);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = bk + 4 * e;
            const bool ok = (kk + k < K) && (n0 + bn < N);
            const z_t* src = ok ? Bm + (size_t)(kk + k) * ldb + n0 + bn : Bm;
            cp_async16(sB + k * G_LDB + bn, src, ok);
        }
    };
    const int nslab = (K + GK - 1) / GK;
#pragma unroll
    for (int s0 = 0; s0 < GSTAGES - 1; ++s0) {
        if (s0 < nslab) issue(s0);
        cp_async_commit();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int m = m0 + wm * 16 + i * 8 + g;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int n = n0 + wn * 32 + j * 8 + 2 * q + e;
                z_t c = make_double2(0., 0.);
                if (SUB && m < M && n < N) c = Cin[(size_t)m * ldc + n];
                acc.re[i][j][e] = c.x;
                acc.im[i][j][e] = c.y;
            }
        }
    }
    for (int slab = 0; slab < nslab; ++slab) {
        cp_async_wait<GSTAGES - 2>();
        __syncthreads();   // slab `slab` has landed for every thread; slab-1's buffer is free
        if (slab + GSTAGES - 1 < nslab) issue(slab + GSTAGES - 1);
        cp_async_commit();
        const z_t* sA = smem + (size_t)(slab % GSTAGES) * G_STAGE_ELEMS;
        const z_t* sB = sA + GM * G_LDA;
#pragma unroll
        for (int k4 = 0; k4 < GK; k4 += 4) {
            z_t af[2], bf[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) af[i] = sA[(wm * 16 + i * 8 + g) * G_LDA + k4 + q];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = sB[(k4 + q) * G_LDB + wn * 32 + j * 8 + g];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                // SUB: acc -= a*b, i.e. the a fragment enters negated
                const double ar = SUB ? -af[i].x : af[i].x, ai = SUB ? -af[i].y : af[i].y;
                const double nai = -ai;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dmma(acc.re[i][j], ar, bf[j].x);
                    dmma(acc.im[i][j], ar, bf[j].y);
                    dmma(acc.re[i][j], nai, bf[j].y);
                    dmma(acc.im[i][j], ai, bf[j].x);
                }
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();   // the shared-memory ring may be reused by the caller / the next tile
}

__device__ __forceinline__ void zgemm_mainloop(GemmAcc& acc, const z_t* __restrict__ A, int lda,
                                               const z_t* __restrict__ Bm, int ldb, int M, int N,
                                               int K, int m0, int n0) {
    zgemm_mainloop_t<false>(acc, A, lda, Bm, ldb, M, N, K, m0, n0, nullptr, 0);
}

// element (i, j, e) of a thread's accumulators is tile entry (acc_row(i), acc_col(j, e))
__device__ __forceinline__ int acc_row(int i) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    return (warp & 3) * 16 + i * 8 + (lane >> 2);
}
__device__ __forceinline__ int acc_col(int j, int e) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    return (warp >> 2) * 32 + j * 8 + 2 * (lane & 3) + e;
}

__device__ __forceinline__ void zgemm_sub_tile(z_t* __restrict__ C, int ldc,
                                               const z_t* __restrict__ A, int lda,
                                               const z_t* __restrict__ Bm, int ldb, int M, int N,
                                               int K, int bx, int by) {
    const int m0 = by * GM, n0 = bx * GN;
    GemmAcc acc;
    zgemm_mainloop_t<true>(acc, A, lda, Bm, ldb, M, N, K, m0, n0, C, ldc);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int m = m0 + acc_row(i);
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int n = n0 + acc_col(j, e);
                if (n >= N) continue;
                C[(size_t)m * ldc + n] = make_double2(acc.re[i][j][e], acc.im[i][j][e]);
            }
        }
    }
}

__global__ void __launch_bounds__(256, 2)
zgemm_sub_kernel(z_t* __restrict__ C, int ldc, const z_t* __restrict__ A, int lda,
                 const z_t* __restrict__ Bm, int ldb, int M, int N, int K) {
    zgemm_sub_tile(C, ldc, A, lda, Bm, ldb, M, N, K, blockIdx.x, blockIdx.y);
}

// Trailing update of the augmented system in ONE launch: the same L21 (A) multiplies the
// block row of W (C1 -= A*B1, N1 columns) and the block row of the right-hand sides
// (C2 -= A*B2, N2 columns); column tiles [0, nx1) belong to the first product.
__global__ void __launch_bounds__(256, 2)
zgemm_sub2_kernel(z_t* __restrict__ C1, const z_t* __restrict__ B1, int N1, int nx1,
                  z_t* __restrict__ C2, const z_t* __restrict__ B2, int N2, int ld,
                  const z_t* __restrict__ A, int M, int K) {
    if ((int)blockIdx.x < nx1)
        zgemm_sub_tile(C1, ld, A, ld, B1, ld, M, N1, K, blockIdx.x, blockIdx.y);
    else
        zgemm_sub_tile(C2, ld, A, ld, B2, ld, M, N2, K, blockIdx.x - nx1, blockIdx.y);
}

// ------------------------------------------------------------------ symmetric path: kernels
// EMME's matrix is complex SYMMETRIC (include/solver.h:448-453, :494-504 write every entry together
// with its mirror).  Without interchanges A = L D L^T, i.e. the U factor is D L^T, and
//     trace(A^-1 B) = sum_ij (A^-1)_ij B_ji,   A^-1 = M^T D^-1 M,   M = L^-1,
// which needs 4 dim^3 real flops (factor 4/3, M 4/3, lower triangle of M^T D^-1 M 4/3) instead of
// the (8/3 + 4 + 2) dim^3 of LU + forward elimination of dim right-hand sides + half a back
// substitution.  launch_trace_sym below drives these kernels; symmetry and "partial pivoting
// keeps the diagonal" are both VERIFIED on the device (flag bits 2 and 1) and the caller falls
// back to the general LU path when either fails.

constexpr int PTRACE_MAX_CHUNKS = 8;   // split-K factor of ptrace_kernel on small systems

// W <- A on the lower block triangle (64 x 64 blocks; diagonal blocks completely) and check that A
// equals its transpose bit for bit.  32 x 32 tiles, block (32, 8).
__global__ void __launch_bounds__(256)
sym_copy_check_kernel(const z_t* __restrict__ A, z_t* __restrict__ W, int dim, int* __restrict__ flag) {
    __shared__ z_t t[32][33];
    const int bi = blockIdx.y, bj = blockIdx.x;
    if ((bj >> 1) > (bi >> 1)) return;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int i = bj * 32 + r, j = bi * 32 + threadIdx.x;     // mirror tile
        t[r][threadIdx.x] = (i < dim && j < dim) ? A[(size_t)i * dim + j] : make_double2(0., 0.);
    }
    __syncthreads();
    int bad = 0;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int i = bi * 32 + r, j = bj * 32 + threadIdx.x;
        if (i < dim && j < dim) {
            const z_t v = A[(size_t)i * dim + j];
            const z_t m = t[threadIdx.x][r];                      // A[j][i]
            if (v.x != m.x || v.y != m.y) bad = 1;
            W[(size_t)i * dim + j] = v;
        }
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0 && threadIdx.y == 0) atomicOr(flag, 2);
}

__global__ void set_identity_diag_kernel(z_t* __restrict__ Y, int dim) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < dim) Y[(size_t)i * dim + i] = make_double2(1., 0.);
}

// 1/a for well-scaled a (one division); Smith's form otherwise
__device__ __forceinline__ z_t zrecip_fast(z_t a) {
    const double n = a.x * a.x + a.y * a.y;
    if (n > 1e-280 && n < 1e280) {
        const double d = __drcp_rn(n);
        return make_double2(a.x * d, -a.y * d);
    }
    return zrecip(a);
}

// 1/u of a pivot: hardware seed + two Newton steps on |u|^2 (<= 1 ulp; the special cases of the
// library division cannot occur for an accepted pivot).  ONE definition: the panel kernel and the
// ranks that rebuild a received panel (lpanel_from_u12_kernel) must form bit-identical multipliers.
__device__ __forceinline__ z_t pivot_recip(z_t u, double& n2) {
    n2 = fma(u.x, u.x, u.y * u.y);
    double d;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(d) : "d"(n2));
    double er = fma(-n2, d, 1.0);
    d = fma(d, er, d);
    er = fma(-n2, d, 1.0);
    d = fma(d, er, d);
    return make_double2(u.x * d, -u.y * d);
}

// One launch per NB-wide panel [k0, k0+jb).  Small code, no per-thread register rows (the
// unrolled one-thread-per-row elimination of panel_nopiv_kernel is instruction-fetch bound: every
// instruction executes once), tensor cores for the two block products:
//   1. every CTA loads the jb x jb diagonal block, requests its operand tile with cp.async and
//      factors the block redundantly in shared memory while the tile is in flight: all 128
//      threads, thread (row i = t/4, columns k = t%4 + 4e), one barrier per pivot, no branches on
//      the pivot chain.  In-place Gauss-Jordan: the identity is carried along in the part of the
//      block the elimination has already left, every row i > c is updated over ALL columns
//      (a_ik -= l a_ck; column c itself becomes -l), so the block ends as
//      Mi = L11^-1 (strictly lower) \ U11 (diagonal and above);
//   2. row CTAs [0, n_row_ctas): T = A21 Mi^T on the FP64 tensor cores.  U11 = D L11^T (symmetry)
//      makes T[r][c] the entry a_rc just before column c is eliminated, hence
//         L[r, c] = T[r][c] / u_cc,   U[k0+c, r] = T[r][c]   (U12 = D L21^T: no block-row solve)
//      and |T[r][c]| <= tau |u_cc| is the partial-pivoting check;
//   3. column CTAs: rows k0..ke of Y <- Mi Y (the identity carried along: rows k0..ke of
//      M = L^-1 become final).
// A CTA owns `tile` = 16, 32 or 64 rows (row role) or columns (Y role): small systems get small
// tiles so that the products, which run at one SM's tensor rate, spread over more SMs.
// A zero or badly scaled pivot raises the flag like a failed pivoting test: the caller repeats the
// step with the pivoting LU, which reports exact singularity (info) properly.
constexpr int PS_ROWS = 64;                       // largest tile
constexpr int PS_LD = NB + 4;                     // stride of sG, sL, sM, row tile: 4 (mod 8) elements
constexpr int PS_LDY_MAX = PS_ROWS + 2;           // stride of the Y tile: tile + 2 = 2 (mod 8) elements
constexpr int PS_TILE = PS_ROWS * PS_LD > NB * PS_LDY_MAX ? PS_ROWS * PS_LD : NB * PS_LDY_MAX;
constexpr size_t PS_SMEM_BYTES = sizeof(z_t) * (2 * NB * PS_LD + PS_TILE) + sizeof(z_t) * NB + sizeof(double) * NB;

#ifdef EMME_PS_CLOCKS
__device__ long long g_ps_clocks[8];   // scratch/panel_bench.cu: phase boundaries of CTA 0
#define PS_CLOCK(n) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_ps_clocks[n] = clock64(); } while (0)
#else
#define PS_CLOCK(n) do { } while (0)
#endif

__global__ void __launch_bounds__(128)
panel_sym_kernel(z_t* __restrict__ W, z_t* __restrict__ Y, int ld, int dim, int k0, int jb, int ycol0, int ycols,
                 int n_row_ctas, int tile, double tau, int* __restrict__ flag, z_t* __restrict__ dvec) {
    static_assert(NB == 32, "thread mapping of the diagonal-block factorisation");
    extern __shared__ __align__(16) unsigned char ps_smem_raw[];
    z_t* sG = reinterpret_cast<z_t*>(ps_smem_raw);   // [NB][PS_LD] Gauss-Jordan work array
    z_t* sM = sG + NB * PS_LD;                        // [NB][PS_LD] Mi with explicit unit diagonal
    z_t* sT = sM + NB * PS_LD;                        // operand tile
    z_t* sInv = sT + PS_TILE;                         // 1/u_cc
    double* sAbs = reinterpret_cast<double*>(sInv + NB);   // |u_cc| (cabs1)
    const int tid = threadIdx.x;
    PS_CLOCK(0);
    const bool row_role = (int)blockIdx.x < n_row_ctas;
    const int r0 = jb + (int)blockIdx.x * tile;                         // first row offset below k0
    const int c0 = ycol0 + ((int)blockIdx.x - n_row_ctas) * tile;       // first column of Y
    const int ldy = tile + 2;
    // ---- the diagonal block first (it is needed first), then the operand tile (cp.async) ----
    const int gi = tid >> 2, gq = tid & 3;
    z_t d8[NB / 4];
#pragma unroll
    for (int e = 0; e < NB / 4; ++e) {
        const int k = gq + 4 * e;
        d8[e] = (gi < jb && k < jb) ? W[(size_t)(k0 + gi) * ld + k0 + k] : make_double2(gi == k ? 1. : 0., 0.);
    }
    if (row_role) {
        // rows r0 .. r0+tile-1 of the panel, 32 columns each (one 512-byte row per warp instruction)
        for (int rr = tid >> 5; rr < tile; rr += 4) {
            const int cc = tid & 31;
            const bool ok = (k0 + r0 + rr < dim) && cc < jb;
            const z_t* src = ok ? W + (size_t)(k0 + r0 + rr) * ld + k0 + cc : W;
            cp_async16(sT + rr * PS_LD + cc, src, ok);
        }
    } else {
        // rows k0 .. k0+31 of Y, `tile` columns each
        for (int e = tid; e < NB * tile; e += 128) {
            const int kk = e / tile, nn = e - kk * tile;
            const bool ok = kk < jb && (c0 + nn < ycols);
            const z_t* src = ok ? Y + (size_t)(k0 + kk) * ld + c0 + nn : Y;
            cp_async16(sT + kk * ldy + nn, src, ok);
        }
    }
    cp_async_commit();
#pragma unroll
    for (int e = 0; e < NB / 4; ++e) sG[gi * PS_LD + gq + 4 * e] = d8[e];
    __syncthreads();
    PS_CLOCK(1);
    int bad = 0;
    {
        z_t* const rowi = sG + gi * PS_LD;
        const bool in_block = gi < jb;
        for (int c = 0; c < jb; ++c) {
            const z_t* rowc = sG + c * PS_LD;
            // all shared-memory loads of the step are issued before the reciprocal is needed
            const z_t u = rowc[c];
            const z_t num = rowi[c];
            z_t x[NB / 4], pv[NB / 4];
#pragma unroll
            for (int e = 0; e < NB / 4; ++e) {
                x[e] = rowi[gq + 4 * e];
                pv[e] = rowc[gq + 4 * e];
            }
            // column c: a_ic - l*1 with a_ic counted as 0 gives the carried identity's -l
            const bool mine = gq == (c & 3);
            const int ec = c >> 2;
#pragma unroll
            for (int e = 0; e < NB / 4; ++e)
                if (mine && e == ec) {
                    x[e] = make_double2(0., 0.);
                    pv[e] = make_double2(1., 0.);
                }
            // 1/u: hardware seed + two Newton steps; a zero, non-finite or badly scaled pivot is
            // reported through the flag (integer test on the exponent field)
            double n2;
            const z_t inv = pivot_recip(u, n2);
            const unsigned ex = ((unsigned)__double2hiint(n2) >> 20) & 0x7ffu;
            if (ex < 64u || ex > 1983u) bad = 1;
            const double ua = fabs(u.x) + fabs(u.y);
            if (tid == 0) {
                sInv[c] = inv;
                sAbs[c] = ua;
            }
            __syncwarp();   // the four threads of a row have read a_ic before one of them overwrites it
            if (gi > c && in_block) {
                const z_t l = zmul(num, inv);
                if (mine && fabs(num.x) + fabs(num.y) > tau * ua) bad = 1;
#pragma unroll
                for (int e = 0; e < NB / 4; ++e) zfms(x[e], l, pv[e]);
#pragma unroll
                for (int e = 0; e < NB / 4; ++e) rowi[gq + 4 * e] = x[e];
            }
            __syncthreads();
        }
        // clean copy of Mi for the tensor-core products: strictly lower part of sG, unit diagonal
#pragma unroll
        for (int e = 0; e < NB / 4; ++e) {
            const int k = gq + 4 * e;
            sM[gi * PS_LD + k] = k < gi ? rowi[k] : make_double2(k == gi ? 1. : 0., 0.);
        }
    }
    PS_CLOCK(2);
    cp_async_wait<0>();
    __syncthreads();
    PS_CLOCK(3);
    // Every CTA of this launch reads the UNFACTORED diagonal block from W (above) whenever it gets
    // scheduled -- with more CTAs than fit the device at once, or a device shared with other streams,
    // that can be long after CTA 0 has finished.  The block is therefore never written back in place
    // (round 1 did, a latent race); nothing downstream reads L11 / U11, only the pivots d_k = u_kk,
    // which go to their own vector.
    if (blockIdx.x == 0 && tid < jb) dvec[k0 + tid] = sG[tid * PS_LD + tid];
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int rg = tile >> 4;                     // 16-row groups per tile: 1, 2 or 4
    double acc_re[2][4][2], acc_im[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            acc_re[i][j][0] = acc_re[i][j][1] = 0.0;
            acc_im[i][j][0] = acc_im[i][j][1] = 0.0;
        }
    if (row_role) {
        // T (tile x 32) = tile * Mi^T: warp w owns rows 16 (w % rg).. and `rg` of the four 8-column blocks
        const int wr = warp % rg, cb0 = (warp / rg) * rg;
        for (int k4 = 0; k4 < NB; k4 += 4) {
            z_t af[2], bf[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) af[i] = sT[(wr * 16 + i * 8 + g) * PS_LD + k4 + q];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j < rg) bf[j] = sM[((cb0 + j) * 8 + g) * PS_LD + k4 + q];    // B[k][n] = Mi[n][k]
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const double nai = -af[i].y;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (j < rg) {
                        dmma(acc_re[i][j], af[i].x, bf[j].x);
                        dmma(acc_im[i][j], af[i].x, bf[j].y);
                        dmma(acc_re[i][j], nai, bf[j].y);
                        dmma(acc_im[i][j], af[i].y, bf[j].x);
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = r0 + wr * 16 + i * 8 + g;
            if (k0 + r >= dim) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j >= rg) continue;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = (cb0 + j) * 8 + 2 * q + e;
                    if (c >= jb) continue;
                    const z_t t = make_double2(acc_re[i][j][e], acc_im[i][j][e]);
                    if (fabs(t.x) + fabs(t.y) > tau * sAbs[c]) bad = 1;
                    W[(size_t)(k0 + r) * ld + k0 + c] = zmul(t, sInv[c]);
                    W[(size_t)(k0 + c) * ld + k0 + r] = t;
                }
            }
        }
    } else {
        // rows k0..ke of Y (32 x tile) <- Mi * tile: warp w owns rows 16 (w&1).. and `rg` 8-column blocks
        const int mb = warp & 1, nb0 = (warp >> 1) * rg;
        for (int k4 = 0; k4 < NB; k4 += 4) {
            z_t af[2], bf[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) af[i] = sM[(mb * 16 + i * 8 + g) * PS_LD + k4 + q];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j < rg) bf[j] = sT[(k4 + q) * ldy + (nb0 + j) * 8 + g];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const double nai = -af[i].y;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (j < rg) {
                        dmma(acc_re[i][j], af[i].x, bf[j].x);
                        dmma(acc_im[i][j], af[i].x, bf[j].y);
                        dmma(acc_re[i][j], nai, bf[j].y);
                        dmma(acc_im[i][j], af[i].y, bf[j].x);
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int m = mb * 16 + i * 8 + g;
            if (m >= jb) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j >= rg) continue;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int n = c0 + (nb0 + j) * 8 + 2 * q + e;
                    if (n >= ycols) continue;
                    Y[(size_t)(k0 + m) * ld + n] = make_double2(acc_re[i][j][e], acc_im[i][j][e]);
                }
            }
        }
    }
    PS_CLOCK(4);
    if (__syncthreads_or(bad) && tid == 0) atomicOr(flag, 1);
}

// Trailing update of the symmetric path in ONE launch: C1 -= A*B1 on the LOWER block triangle only
// (tiles with column tile <= row tile; C1's origin lies on the diagonal of W), C2 -= A*B2 (rows of
// Y) everywhere.  The two products share A but may have different row counts.
__global__ void __launch_bounds__(256, 2)
zgemm_sym2_kernel(z_t* __restrict__ C1, const z_t* __restrict__ B1, int M1, int N1, int nx1,
                  z_t* __restrict__ C2, const z_t* __restrict__ B2, int M2, int N2, int ld,
                  const z_t* __restrict__ A, int K) {
    const int by = blockIdx.y;
    if ((int)blockIdx.x < nx1) {
        const int bx = blockIdx.x;
        if (bx > by || by * GM >= M1) return;
        zgemm_sub_tile(C1, ld, A, ld, B1, ld, M1, N1, K, bx, by);
    } else {
        if (by * GM >= M2) return;
        zgemm_sub_tile(C2, ld, A, ld, B2, ld, M2, N2, K, blockIdx.x - nx1, by);
    }
}

// YT(i, k) = Y(k, i) / d_k for k >= i (d = the pivots, dvec), zero below the diagonal (only
// the diagonal 64-blocks are ever read there).  32 x 32 tiles, block (32, 8).
__global__ void __launch_bounds__(256)
transpose_invd_kernel(const z_t* __restrict__ Y, const z_t* __restrict__ dvec, z_t* __restrict__ YT, int dim) {
    __shared__ z_t t[32][33];
    const int bi = blockIdx.y, bk = blockIdx.x;
    if (bk < bi) {
        if ((bk >> 1) == (bi >> 1)) {
            for (int r = threadIdx.y; r < 32; r += 8) {
                const int i = bi * 32 + r, k = bk * 32 + threadIdx.x;
                if (i < dim && k < dim) YT[(size_t)i * dim + k] = make_double2(0., 0.);
            }
        }
        return;
    }
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int k = bk * 32 + r, i = bi * 32 + threadIdx.x;
        z_t v = make_double2(0., 0.);
        if (k < dim && i < dim && k >= i) v = zmul(Y[(size_t)k * dim + i], zrecip(dvec[k]));
        t[r][threadIdx.x] = v;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int i = bi * 32 + r, k = bk * 32 + threadIdx.x;
        if (i < dim && k < dim) YT[(size_t)i * dim + k] = t[threadIdx.x][r];
    }
}

// partial[chunk][t] = sum over tile t = (I, J), I >= J, of P_ij * (I > J ? B_ij + B_ji : B_ji) with
// P = YT * Y = A^-1 (only k >= 64 I contributes: YT is upper, Y lower triangular).
// Sharded dense step: a rank forms the tiles t = t_first + blockIdx.x * t_stride and stores each
// partial at its GLOBAL slot [chunk][t] of the workspace of EVERY rank (peer stores), so that all
// ranks reduce the same array in the same order (bitwise the single-GPU trace).
struct PartialDst {
    z_t* p[EMME_MAX_PEERS];
    int n;
};

__global__ void __launch_bounds__(256, 2)
ptrace_kernel(const z_t* __restrict__ YT, const z_t* __restrict__ Y, const z_t* __restrict__ Bd, int dim,
              int ck, const PartialDst partial, int ntiles, int t_first, int t_stride) {
    __shared__ double red[2][8];
    const int t = t_first + (int)blockIdx.x * t_stride;
    if (t >= ntiles) return;
    int I = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while ((I + 1) * (I + 2) / 2 <= t) ++I;
    while (I * (I + 1) / 2 > t) --I;
    const int J = t - I * (I + 1) / 2;
    // split K: chunk blockIdx.y of the range [64 I, dim) (small systems have too few tiles to fill
    // the GPU, and the longest tile would be the critical path)
    const int kstart = I * GM + (int)blockIdx.y * ck;
    const int kend = blockIdx.y + 1 == gridDim.y ? dim : (kstart + ck < dim ? kstart + ck : dim);
    if (kstart >= dim) {
        if (threadIdx.x == 0)
            for (int r = 0; r < partial.n; ++r) partial.p[r][(size_t)blockIdx.y * ntiles + t] = make_double2(0., 0.);
        return;
    }
    GemmAcc acc;
    zgemm_mainloop(acc, YT + kstart, dim, Y + (size_t)kstart * dim, dim, dim, dim, kend - kstart, I * GM,
                   J * GN);
    double sr = 0., si = 0.;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int m = I * GM + acc_row(i);
        if (m >= dim) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int n = J * GN + acc_col(j, e);
                if (n >= dim) continue;
                z_t b = Bd[(size_t)n * dim + m];
                if (I != J) {
                    const z_t b2 = Bd[(size_t)m * dim + n];
                    b.x += b2.x;
                    b.y += b2.y;
                }
                const double pr = acc.re[i][j][e], pi = acc.im[i][j][e];
                sr += pr * b.x - pi * b.y;
                si += pr * b.y + pi * b.x;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, o);
        si += __shfl_xor_sync(0xffffffffu, si, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red[0][warp] = sr; red[1][warp] = si; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double x = 0., y = 0.;
        for (int w = 0; w < 8; ++w) { x += red[0][w]; y += red[1][w]; }
        for (int r = 0; r < partial.n; ++r) partial.p[r][(size_t)blockIdx.y * ntiles + t] = make_double2(x, y);
    }
}

__global__ void reduce_partials_kernel(const z_t* __restrict__ partial, int n, z_t* __restrict__ out) {
    __shared__ double sx[256], sy[256];
    double x = 0., y = 0.;
    for (int i = threadIdx.x; i < n; i += 256) {
        x += partial[i].x;
        y += partial[i].y;
    }
    sx[threadIdx.x] = x;
    sy[threadIdx.x] = y;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            sx[threadIdx.x] += sx[threadIdx.x + o];
            sy[threadIdx.x] += sy[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = make_double2(sx[0], sy[0]);
}

struct ShardPtrs;
__global__ void ydiag_kernel(const ShardPtrs sp, z_t* __restrict__ S, int ld, int dim, int K0, int JB, int blk0, int P,
                             int nbo);
__global__ void shard_update_kernel(z_t* __restrict__ W, z_t* __restrict__ Y, const z_t* __restrict__ S, int ld, int dim,
                                    int K0, int JB, int wblk0, int nwt, int yblk0, int P, int nbo);

// function attributes are per DEVICE: a process that drives several GPUs (LocalShardedGroup, the
// launcher-free binding of INTEGRATION.md) must set them on each
static cudaError_t gemm_setup() {
    static bool done_on[64] = {};
    int dev = 0;
    cudaError_t e0 = cudaGetDevice(&dev);
    if (e0 != cudaSuccess) return e0;
    bool& done = done_on[dev & 63];
    if (done) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(zgemm_sub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)G_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(zgemm_sub2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)G_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(zgemm_sym2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)G_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(ptrace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(panel_sym_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PS_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(ydiag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(shard_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    done = true;
    return cudaSuccess;
}

// ------------------------------------------------------------------ small kernels
__global__ void trace_kernel(const z_t* __restrict__ B, int ld, int dim, z_t* __restrict__ out) {
    __shared__ double sx[256], sy[256];
    double x = 0., y = 0.;
    for (int i = threadIdx.x; i < dim; i += 256) {
        const z_t v = B[(size_t)i * ld + i];
        x += v.x;
        y += v.y;
    }
    sx[threadIdx.x] = x;
    sy[threadIdx.x] = y;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            sx[threadIdx.x] += sx[threadIdx.x + o];
            sy[threadIdx.x] += sy[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = make_double2(sx[0], sy[0]);
}

// A' = (A - A_old)/delta   (reference include/solver.h:54-57, include/Arithmetics.h)
__global__ void secant_kernel(const z_t* __restrict__ A, const z_t* __restrict__ Aold,
                              z_t* __restrict__ Ad, size_t n, z_t delta) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const z_t a = A[i], b = Aold[i];
        Ad[i] = zdiv(make_double2(a.x - b.x, a.y - b.y), delta);
    }
}

cudaError_t launch_secant(const void* A, const void* Aold, void* Ad, size_t n, double dr,
                          double di, int sms, cudaStream_t stream) {
    secant_kernel<<<sms * 8, 256, 0, stream>>>((const z_t*)A, (const z_t*)Aold, (z_t*)Ad, n,
                                               make_double2(dr, di));
    return cudaGetLastError();
}

// ---- FP64 FMA peak: 8 independent DFMA chains per thread, registers only ----
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
           a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

cudaError_t measure_fp64_peak(double* tflops) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8, iters = 4096;
    double* d = nullptr;
    cudaError_t e = cudaMalloc(&d, sizeof(double) * blocks * 256);
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        dfma_peak_kernel<<<blocks, 256>>>(d, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 8 * 16 * (double)iters * blocks * 256;
        const double tf = fl / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return cudaGetLastError();
}

// test hook: force the grid-cooperative panel kernel for every panel (EMME_DENSE_GRID_PANEL=1)
static bool g_force_grid_panel = false;
void dense_force_grid_panel(bool on) { g_force_grid_panel = on; }

// Solve W X = B for nrhs right-hand sides (the first nrhs columns of the dim x dim buffer B,
// overwritten by X) with the factors left in W by launch_trace_solve(..., nrhs >= 0) and the
// interchanges recorded in the workspace.  Used by the inverse iteration of emme_null_space.
cudaError_t launch_solve_factored(const void* Wv, void* Bv, int dim, int nrhs, void* workspace,
                                  int optimistic, cudaStream_t stream, unsigned long long* n_launches) {
    const z_t* W = (const z_t*)Wv;
    z_t* B = (z_t*)Bv;
    const int ld = dim;
    const int nblk_max = (dim + PT - 1) / PT;
    const int* ipiv = (const int*)((const char*)workspace + sizeof(PanelCand) * 2 * (size_t)nblk_max);
    unsigned long long nl = 0;
    {
        cudaError_t e = gemm_setup();
        if (e != cudaSuccess) return e;
    }
    for (int k0 = 0; k0 < dim; k0 += NB) {
        const int jb = dim - k0 < NB ? dim - k0 : NB;
        const int ke = k0 + jb;
        swap_trsm_kernel<<<(nrhs + 127) / 128, 128, 0, stream>>>(W, B, nrhs, nullptr, 0, ld, k0, jb, ipiv,
                                                                  optimistic ? 0 : 1);
        ++nl;
        const int M = dim - ke;
        if (M > 0) {
            dim3 g((nrhs + GN - 1) / GN, (M + GM - 1) / GM);
            zgemm_sub_kernel<<<g, 256, G_SMEM_BYTES, stream>>>(B + (size_t)ke * ld, ld,
                                                               W + (size_t)ke * ld + k0, ld,
                                                               B + (size_t)k0 * ld, ld, M, nrhs, jb);
            ++nl;
        }
    }
    const int last = ((dim - 1) / NB) * NB;
    for (int k0 = last; k0 >= 0; k0 -= NB) {
        const int jb = dim - k0 < NB ? dim - k0 : NB;
        utrsm_kernel<<<(nrhs + 127) / 128, 128, 0, stream>>>(W, B, ld, k0, jb, nrhs);
        ++nl;
        if (k0 > 0) {
            dim3 g((nrhs + GN - 1) / GN, (k0 + GM - 1) / GM);
            zgemm_sub_kernel<<<g, 256, G_SMEM_BYTES, stream>>>(B, ld, W + k0, ld, B + (size_t)k0 * ld, ld, k0,
                                                               nrhs, jb);
            ++nl;
        }
    }
    if (n_launches) *n_launches += nl;
    return cudaGetLastError();
}

// v <- conj(column 0 of X) scaled to unit 2-norm; X is dim x dim row-major (stride ld).
// Also returns the norm (for convergence monitoring).  One block.
__global__ void conj_normalise_kernel(const z_t* __restrict__ X, int ld, int dim, z_t* __restrict__ rhs,
                                      int rhs_ld, double* __restrict__ norm_out, int conjugate) {
    __shared__ double sh[256];
    double acc = 0.;
    for (int i = threadIdx.x; i < dim; i += 256) {
        const z_t v = X[(size_t)i * ld];
        acc += v.x * v.x + v.y * v.y;
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    const double nrm = sqrt(sh[0]);
    const double inv = nrm > 0. ? 1.0 / nrm : 0.;
    for (int i = threadIdx.x; i < dim; i += 256) {
        const z_t v = X[(size_t)i * ld];
        rhs[(size_t)i * rhs_ld] = make_double2(v.x * inv, conjugate ? -v.y * inv : v.y * inv);
    }
    if (threadIdx.x == 0 && norm_out) *norm_out = nrm;
}

cudaError_t launch_conj_normalise(const void* X, int ld, int dim, void* rhs, int rhs_ld, double* norm_out,
                                  int conjugate, cudaStream_t stream) {
    conj_normalise_kernel<<<1, 256, 0, stream>>>((const z_t*)X, ld, dim, (z_t*)rhs, rhs_ld, norm_out,
                                                 conjugate);
    return cudaGetLastError();
}

size_t dense_workspace_bytes(int dim) {
    const int nblk = (dim + PT - 1) / PT;
    return sizeof(PanelCand) * 2 * (size_t)nblk + sizeof(int) * (size_t)dim + 64;
}

cudaError_t dense_prepare() { return gemm_setup(); }   // one-time function attributes (before any capture)

static int g_nbo = 0;     // outer block width (multiple of NB); 0 = choose by size
void dense_set_outer_block(int nbo) {
    g_nbo = nbo <= 0 ? 0 : (nbo < NB ? NB : (nbo / NB) * NB);
}
// The interchange-free paths accept the diagonal pivot while |a_ic| <= tau |a_cc| for every row below
// it (threshold pivoting: element growth per elimination step bounded by tau).  tau = 32 is the
// relative pivot threshold 0.03 of sparse direct solvers.  Round 1 used 4, which EMME's own matrices
// exceed close to a root: the Newton iterates approach det A = 0, but the singularity goes into the
// LAST pivot (no rows below it) while the ratios elsewhere stay modest -- measured on C1 (N = 128,
// 256) along 7 iterates down to cond A = 1e16: max ratio 0.6, 1.9, 4.3, 3.8, 7.0, 8.4, 13.4.  With
// tau = 4 two of twenty bench iterates at N = 8192 fell back to the pivoting LU (200 ms instead of 79).
static double g_tau = 32.0;
void dense_set_pivot_threshold(double tau) { g_tau = tau; }

// W (dim x dim, destroyed) and B (dim x dim, destroyed): trace(W^-1 B) -> *d_trace (device).
//
// Two-level right-looking LU: outer blocks of g_nbo columns are factored by NB-wide panels whose
// updates stay inside the outer block; the rest of the matrix and the right-hand sides see one
// interchange pass, a block-row solve and ONE rank-g_nbo GEMM per outer block (K = 128 keeps the
// FP64 pipe busy; K = 32 was dominated by the C-tile read-modify-write).
cudaError_t launch_trace_solve(void* Wv, void* Bv, int dim, void* workspace, void* d_trace,
                               int* d_info, cudaStream_t stream, unsigned long long* n_launches,
                               int optimistic, int* d_flag, int nrhs) {
    unsigned long long nl = 0;
    const bool want_trace = nrhs < 0;      // nrhs < 0: all dim right-hand sides + trace
    if (nrhs < 0) nrhs = dim;
    z_t* W = (z_t*)Wv;
    z_t* B = (z_t*)Bv;
    const int ld = dim;
    const int nblk_max = (dim + PT - 1) / PT;
    PanelCand* xchg = (PanelCand*)workspace;
    int* ipiv = (int*)((char*)workspace + sizeof(PanelCand) * 2 * (size_t)nblk_max);
    cudaError_t e = gemm_setup();
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(d_info, 0, sizeof(int), stream);
    if (e != cudaSuccess) return e;
    // small systems are latency bound: single-level blocking means fewer launches
    const int NBO = g_nbo > 0 ? g_nbo : (dim <= 2048 ? NB : 128);
    if (optimistic) {
        e = cudaMemsetAsync(d_flag, 0, sizeof(int), stream);
        if (e != cudaSuccess) return e;
    }
    auto at = [&](z_t* M, int r, int c) { return M + (size_t)r * ld + c; };
    auto gemm1 = [&](z_t* C, const z_t* A, const z_t* Bm, int M, int N, int K) {
        if (M <= 0 || N <= 0) return;
        dim3 g((N + GN - 1) / GN, (M + GM - 1) / GM);
        zgemm_sub_kernel<<<g, 256, G_SMEM_BYTES, stream>>>(C, ld, A, ld, Bm, ld, M, N, K);
        ++nl;
    };
    auto gemm2 = [&](z_t* C1, const z_t* B1, int N1, z_t* C2, const z_t* B2, int N2, const z_t* A, int M,
                     int K) {
        if (M <= 0) return;
        const int nx1 = (N1 + GN - 1) / GN, nx2 = (N2 + GN - 1) / GN;
        if (nx1 + nx2 == 0) return;
        dim3 g(nx1 + nx2, (M + GM - 1) / GM);
        zgemm_sub2_kernel<<<g, 256, G_SMEM_BYTES, stream>>>(C1, B1, N1, nx1, C2, B2, N2, ld, A, M, K);
        ++nl;
    };

    for (int K0 = 0; K0 < dim; K0 += NBO) {
        const int JB = dim - K0 < NBO ? dim - K0 : NBO;
        const int KE = K0 + JB;
        // ---- factor the outer block column by NB-wide panels ----
        for (int k0 = K0; k0 < KE; k0 += NB) {
            int jb = KE - k0 < NB ? KE - k0 : NB;
            const int ke = k0 + jb;
            int rows = dim - k0;
            int nblk = (rows + PT - 1) / PT;
            if (optimistic) {
                panel_nopiv_kernel<128><<<(rows + 127) / 128, 128, 0, stream>>>(W, ld, dim, k0, jb, g_tau,
                                                                              d_flag, d_info);
                e = cudaGetLastError();
            } else if (rows <= CL_TPB * panel_cluster_limit() && !g_force_grid_panel) {
                e = launch_panel_cluster(W, ld, dim, k0, jb, ipiv, d_info, stream);
            } else {
                int k0v = k0, dimv = dim, ldv = ld;
                void* args[] = {&W, &ldv, &dimv, &k0v, &jb, &ipiv, &xchg, &d_info};
                e = cudaLaunchCooperativeKernel((void*)panel_kernel, dim3(nblk), dim3(PT), args, 0, stream);
            }
            if (e != cudaSuccess) return e;
            ++nl;
            if (k0 > K0 && !optimistic) {   // multipliers of the earlier panels follow their rows
                laswp_kernel<<<(k0 - K0 + 127) / 128, 128, 0, stream>>>(at(W, 0, K0), k0 - K0, nullptr, 0,
                                                                        ld, k0, jb, ipiv);
                ++nl;
            }
            const int nin = KE - ke;
            if (nin > 0) {
                swap_trsm_kernel<<<(nin + 127) / 128, 128, 0, stream>>>(W, at(W, 0, ke), nin, nullptr, 0,
                                                                        ld, k0, jb, ipiv, optimistic ? 0 : 1);
                ++nl;
                gemm1(at(W, ke, ke), at(W, ke, k0), at(W, k0, ke), dim - ke, nin, jb);
            }
        }
        // ---- the rest of W and all right-hand sides: interchanges, block-row solve, rank-JB update
        const int nrest = dim - KE;
        const int ncol = nrest + nrhs;
        if (ncol == 0) continue;
        if (!optimistic) {
            laswp_kernel<<<(ncol + 127) / 128, 128, 0, stream>>>(at(W, 0, KE), nrest, B, nrhs, ld, K0, JB, ipiv);
            ++nl;
        }
        for (int k0 = K0; k0 < KE; k0 += NB) {
            const int jb = KE - k0 < NB ? KE - k0 : NB;
            const int ke = k0 + jb;
            swap_trsm_kernel<<<(ncol + 127) / 128, 128, 0, stream>>>(W, at(W, 0, KE), nrest, B, nrhs, ld, k0,
                                                                     jb, ipiv, 0);
            ++nl;
            gemm2(at(W, ke, KE), at(W, k0, KE), nrest, at(B, ke, 0), at(B, k0, 0), nrhs, at(W, ke, k0),
                  KE - ke, jb);
        }
        gemm2(at(W, KE, KE), at(W, K0, KE), nrest, at(B, KE, 0), at(B, K0, 0), nrhs, at(W, KE, K0),
              dim - KE, JB);
    }
    if (!want_trace) {
        if (n_launches) *n_launches += nl;
        return cudaGetLastError();
    }
    // ---- back substitution X = U^-1 Y, lower triangle of X only (the trace needs X_ii) ----
    const int lastK = ((dim - 1) / NBO) * NBO;
    for (int K0 = lastK; K0 >= 0; K0 -= NBO) {
        const int JB = dim - K0 < NBO ? dim - K0 : NBO;
        const int KE = K0 + JB;
        const int lastk = K0 + ((JB - 1) / NB) * NB;
        for (int k0 = lastk; k0 >= K0; k0 -= NB) {
            const int jb = KE - k0 < NB ? KE - k0 : NB;
            const int ke = k0 + jb;
            utrsm_kernel<<<(ke + 127) / 128, 128, 0, stream>>>(W, B, ld, k0, jb, ke);
            ++nl;
            // rows of this outer block above the panel, columns [0, k0)
            gemm1(at(B, K0, 0), at(W, K0, k0), at(B, k0, 0), k0 - K0, k0, jb);
        }
        // rows above the outer block, columns [0, K0)
        gemm1(B, at(W, 0, K0), at(B, K0, 0), K0, K0, JB);
    }
    trace_kernel<<<1, 256, 0, stream>>>(B, ld, dim, (z_t*)d_trace);
    if (n_launches) *n_launches += nl + 1;
    return cudaGetLastError();
}


// ------------------------------------------------------------------ symmetric path: blocked kernels
// Two-level (outer block NBO = 256 > NB) form of the symmetric path, written so that it SHARDS over
// the GPUs of one box by column blocks (outer block K belongs to rank K mod P, for W and for Y):
//   * the owner factors panel K (NB-wide panels + updates inside the outer block) carrying only the
//     DIAGONAL block of the identity, so that it ends with M_KK = (L_KK)^-1 next to the L panel and
//     the mirrored block row U12 = D L21^T;
//   * panel K = {L panel, U12, M_KK} is stored into the W / Y buffers of every other rank through
//     their peer mappings and announced with a release store into their flag pages;
//   * every rank then transforms rows K of ITS columns of Y with one GEMM (Y_K <- M_KK Y_K, final
//     rows of M = L^-1, written straight into the Y of every rank from the epilogue) and applies the
//     rank-NBO update to ITS column blocks of W (lower block triangle) and Y;
//   * look-ahead: the owner of panel K+1 updates that block first, factors and publishes it, and only
//     then does the rest of its update -- the panel factorisation leaves the critical path.
// With P = 1 the same code is the single-GPU path; every matrix element sees the same operations in
// the same order for every P, so the trace is bitwise independent of the number of GPUs.

struct ShardPtrs {
    z_t* W[EMME_MAX_PEERS];
    z_t* Y[EMME_MAX_PEERS];
    z_t* dvec[EMME_MAX_PEERS];
    int n, me;
};

// Copy panel K to the ranks in `mask`: the mirrored block row U12 = W[K0:KE, KE:dim) (= T^T, the
// entries of the panel just before their column was eliminated), M_KK = Y[K0:KE, K0:KE) and the
// pivots d[K0:KE).  The L panel itself is NOT sent: every receiver rebuilds it from U12 and the
// pivots with the owner's own operation (lpanel_from_u12_kernel), which halves the NVLink volume.
// Row-contiguous 16-byte accesses.
__global__ void __launch_bounds__(256)
publish_panel_kernel(const ShardPtrs sp, int ld, int dim, int K0, int JB, unsigned mask) {
    const int KE = K0 + JB;
    const size_t n2 = (size_t)JB * (dim - KE);          // U12
    const size_t n3 = (size_t)JB * JB;                  // M_KK
    const size_t total = n2 + n3;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < (size_t)JB) {
        const z_t v = sp.dvec[sp.me][K0 + gid];
#pragma unroll
        for (int r = 0; r < EMME_MAX_PEERS; ++r)
            if (r < sp.n && r != sp.me && ((mask >> r) & 1u)) sp.dvec[r][K0 + gid] = v;
    }
    for (size_t e = gid; e < total; e += stride) {
        size_t off;
        bool in_y = false;
        if (e < n2) {
            const int w = dim - KE;
            off = (size_t)(K0 + e / w) * ld + KE + e % w;
        } else {
            const size_t f = e - n2;
            off = (size_t)(K0 + f / JB) * ld + K0 + f % JB;
            in_y = true;
        }
        const z_t v = in_y ? sp.Y[sp.me][off] : sp.W[sp.me][off];
#pragma unroll
        for (int r = 0; r < EMME_MAX_PEERS; ++r)
            if (r < sp.n && r != sp.me && ((mask >> r) & 1u)) (in_y ? sp.Y[r] : sp.W[r])[off] = v;
    }
}

// Receiver side: L[KE + r][K0 + c] = U12[K0 + c][KE + r] * (1/d_c) -- the very product the owner's
// panel kernel formed (same t, same pivot_recip), so the rebuilt panel is bit-identical to the
// owner's.  32 x 32 tiles through shared memory, block (32, 8).
__global__ void __launch_bounds__(256)
lpanel_from_u12_kernel(z_t* __restrict__ W, const z_t* __restrict__ dvec, int ld, int dim, int K0, int JB) {
    __shared__ z_t t[32][33];
    __shared__ z_t sinv[32];
    const int KE = K0 + JB;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    if (threadIdx.y == 0) {
        double n2;
        const int c = c0 + threadIdx.x;
        sinv[threadIdx.x] = c < JB ? pivot_recip(dvec[K0 + c], n2) : make_double2(0., 0.);
    }
    for (int cc = threadIdx.y; cc < 32; cc += 8) {
        const int c = c0 + cc, r = r0 + threadIdx.x;
        t[cc][threadIdx.x] = (c < JB && KE + r < dim) ? W[(size_t)(K0 + c) * ld + KE + r] : make_double2(0., 0.);
    }
    __syncthreads();
    for (int rr = threadIdx.y; rr < 32; rr += 8) {
        const int r = r0 + rr, c = c0 + threadIdx.x;
        if (c < JB && KE + r < dim) W[(size_t)(KE + r) * ld + K0 + c] = zmul(t[threadIdx.x][rr], sinv[threadIdx.x]);
    }
}

// column tile t of a block-cyclic list of outer blocks: blocks blk0, blk0+P, ... each NBO/64 tiles wide
__device__ __forceinline__ int shard_col0(int t, int blk0, int P, int nbo) {
    const int tpb = nbo / GN;
    return (blk0 + (t / tpb) * P) * nbo + (t % tpb) * GN;
}

// S[0:JB, cols] = M_KK * Y[K0:KE, cols] for this rank's Y column blocks blk0, blk0+P, ... < K
// (M_KK = Y[K0:KE, K0:KE), unit lower triangular: row tile `by` only needs k < 64 (by + 1)).
// The result -- final rows K0..KE of M = L^-1 -- goes to the local scratch S (the B operand of the
// update that follows; Y itself is still being read by other CTAs) and into Y of every OTHER rank.
__global__ void __launch_bounds__(256, 2)
ydiag_kernel(const ShardPtrs sp, z_t* __restrict__ S, int ld, int dim, int K0, int JB, int blk0, int P, int nbo) {
    const int col0 = shard_col0(blockIdx.x, blk0, P, nbo);
    if (col0 >= K0) return;
    const int m0 = blockIdx.y * GM;
    if (m0 >= JB) return;
    const z_t* Yl = sp.Y[sp.me];
    const int N = K0 - col0 < GN ? K0 - col0 : GN;
    const int Kk = m0 + GM < JB ? m0 + GM : JB;
    GemmAcc acc;
    zgemm_mainloop(acc, Yl + (size_t)K0 * ld + K0, ld, Yl + (size_t)K0 * ld + col0, ld, JB, N, Kk, m0, 0);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int m = m0 + acc_row(i);
        if (m >= JB) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int n = acc_col(j, e);
                if (n >= N) continue;
                const z_t v = make_double2(acc.re[i][j][e], acc.im[i][j][e]);
                S[(size_t)m * ld + col0 + n] = v;
                const size_t off = (size_t)(K0 + m) * ld + col0 + n;
#pragma unroll
                for (int r = 0; r < EMME_MAX_PEERS; ++r)
                    if (r < sp.n && r != sp.me) sp.Y[r][off] = v;
            }
        }
    }
}

// Y[K0:KE, cols] <- S[0:JB, cols] for the same column list (the local copy of the final rows)
__global__ void __launch_bounds__(256)
ycopy_kernel(z_t* __restrict__ Y, const z_t* __restrict__ S, int ld, int K0, int JB, int blk0, int P, int nbo) {
    const int col0 = shard_col0(blockIdx.x, blk0, P, nbo);
    if (col0 >= K0) return;
    const int c = threadIdx.x & 63;
    if (col0 + c >= K0) return;
    for (int m = (threadIdx.x >> 6) + 4 * blockIdx.y; m < JB; m += 4 * gridDim.y)
        Y[(size_t)(K0 + m) * ld + col0 + c] = S[(size_t)m * ld + col0 + c];
}

// Rank-JB update with panel K of this rank's column blocks, one launch:
//   W tiles (x < nwt): blocks wblk0, wblk0+P, ...:  W[KE:, c] -= L[KE:, K] U12[K, c], lower block triangle only
//   Y tiles          : blocks yblk0, yblk0+P, ... <= K:  Y[KE:, c] -= L[KE:, K] S[:, c]
__global__ void __launch_bounds__(256, 2)
shard_update_kernel(z_t* __restrict__ W, z_t* __restrict__ Y, const z_t* __restrict__ S, int ld, int dim, int K0,
                    int JB, int wblk0, int nwt, int yblk0, int P, int nbo) {
    const int KE = K0 + JB;
    const int by = blockIdx.y;
    const int M = dim - KE;
    if (by * GM >= M) return;
    const z_t* A = W + (size_t)KE * ld + K0;
    if ((int)blockIdx.x < nwt) {
        const int col0 = shard_col0(blockIdx.x, wblk0, P, nbo);
        if (col0 >= dim || KE + by * GM < col0) return;          // outside / above the block diagonal
        const int N = dim - col0 < GN ? dim - col0 : GN;
        zgemm_sub_tile(W + (size_t)KE * ld + col0, ld, A, ld, W + (size_t)K0 * ld + col0, ld, M, N, JB, 0, by);
    } else {
        const int col0 = shard_col0((int)blockIdx.x - nwt, yblk0, P, nbo);
        if (col0 >= KE) return;
        const int N = KE - col0 < GN ? KE - col0 : GN;
        zgemm_sub_tile(Y + (size_t)KE * ld + col0, ld, A, ld, S + col0, ld, M, N, JB, 0, by);
    }
}

// ------------------------------------------------------------------ symmetric path: driver
// per-tile partial traces [PTRACE_MAX_CHUNKS][ntiles], then the pivots d_k (dim entries)
static size_t sym_ws_partials(int dim) {
    const size_t nt = (dim + GM - 1) / GM;
    return (nt * (nt + 1) / 2) * PTRACE_MAX_CHUNKS;
}
size_t dense_sym_workspace_bytes(int dim) { return sizeof(z_t) * (sym_ws_partials(dim) + (size_t)dim) + 64; }

// W <- A (lower block triangle) + bitwise symmetry check (raises bit 2 of *d_flag).  Kept apart
// from launch_trace_sym because A alternates between two buffers while the rest of the step
// works on fixed buffers and is replayed as a CUDA graph.
cudaError_t launch_sym_copy_check(const void* A, void* W, int dim, int* d_flag, cudaStream_t stream,
                                  unsigned long long* n_launches) {
    const int nt = (dim + 31) / 32;
    sym_copy_check_kernel<<<dim3(nt, nt), dim3(32, 8), 0, stream>>>((const z_t*)A, (z_t*)W, dim, d_flag);
    if (n_launches) ++*n_launches;
    return cudaGetLastError();
}

int dense_sym_outer_block(int dim) { return g_nbo > 0 ? g_nbo : (dim <= 2048 ? NB : 256); }

#define DCHK(call)                        \
    do {                                  \
        cudaError_t e__ = (call);         \
        if (e__ != cudaSuccess) return e__; \
    } while (0)

// trace(A^-1 B) for complex symmetric A, W holding A's lower block triangle (destroyed: on return
// W holds L below and U = D L^T right of the diagonal blocks; the blocks themselves are left
// unfactored, the pivots live in the workspace), Y and YT dim x dim scratch, B read only.  Raises bit 1 of
// *d_flag when partial pivoting would have interchanged rows (the caller then repeats the step with
// the pivoting LU).  *d_flag is NOT cleared here.
// `peers` (may be null = one GPU): the W / Y / workspace / flag-page mappings of the ranks that share
// the step; *peers->epoch and peers->serial drive the barriers and panel flags.
cudaError_t launch_trace_sym(void* Wv, void* Yv, void* YTv, const void* Bv, int dim, void* sym_workspace,
                             void* d_trace, int* d_info, int* d_flag, cudaStream_t stream,
                             unsigned long long* n_launches, const DensePeers* peers, const DenseAux* aux) {
    unsigned long long nl = 0;
    z_t* W = (z_t*)Wv;
    z_t* Y = (z_t*)Yv;
    z_t* YT = (z_t*)YTv;
    const int ld = dim;
    const int P = peers ? peers->n : 1, me = peers ? peers->me : 0;
    DCHK(gemm_setup());
    DCHK(cudaMemsetAsync(d_info, 0, sizeof(int), stream));
    DCHK(cudaMemsetAsync(Y, 0, sizeof(z_t) * (size_t)dim * dim, stream));
    set_identity_diag_kernel<<<(dim + 255) / 256, 256, 0, stream>>>(Y, dim);
    ++nl;
    // outer block: single level while the step is launch bound; 256 columns beyond (measured at
    // dim 8192: 84.0 ms with 128, 78.5 with 256, 78.8 with 512; at 2048 all within 1 %)
    int NBO = dense_sym_outer_block(dim);
    if (P > 1 && NBO < GN) NBO = GN;           // sharding needs tile-aligned column blocks
    auto at = [&](z_t* M, int r, int c) { return M + (size_t)r * ld + c; };
    // tile = rows (columns of Y) per CTA: small systems get small tiles (more SMs share the products)
    const int tile = dim <= 4736 ? 16 : (dim <= 9472 ? 32 : 64);
    // C1 = W[r1:, c1:c1+N1) (lower block triangle), C2 = Y[r1:r1+M2, yc:yc+N2), A = W[r1:, kb:kb+K)
    // C1 = W[r1:r1+M1, c1:c1+N1) (lower block triangle), C2 = Y[r1:r1+M2, yc:yc+N2), A = W[r1:, kb:kb+K),
    // on stream `st`; returns whether anything was launched
    auto update_on = [&](cudaStream_t st, int r1, int M1, int c1, int N1, int M2, int yc, int N2, int kb, int K) -> bool {
        const int M = M1 > M2 ? M1 : M2;
        if (M <= 0) return false;
        const int nx1 = M1 > 0 ? (N1 + GN - 1) / GN : 0, nx2 = M2 > 0 ? (N2 + GN - 1) / GN : 0;
        if (nx1 + nx2 == 0) return false;
        dim3 g(nx1 + nx2, (M + GM - 1) / GM);
        zgemm_sym2_kernel<<<g, 256, G_SMEM_BYTES, st>>>(at(W, r1, c1), at(W, kb, c1), M1, N1, nx1, at(Y, r1, yc),
                                                        at(Y, kb, yc), M2, N2, ld, at(W, r1, kb), K);
        ++nl;
        return true;
    };
    auto update = [&](int r1, int M1, int c1, int N1, int M2, int yc, int N2, int kb, int K) {
        update_on(stream, r1, M1, c1, N1, M2, yc, N2, kb, K);
    };
    ShardPtrs sp{};
    PartialDst pd{};
    sp.n = P;
    sp.me = me;
    pd.n = P;
    for (int r = 0; r < P; ++r) {
        sp.W[r] = peers ? (z_t*)peers->W[r] : W;
        sp.Y[r] = peers ? (z_t*)peers->Y[r] : Y;
        pd.p[r] = peers ? (z_t*)peers->ws[r] : (z_t*)sym_workspace;
        sp.dvec[r] = pd.p[r] + sym_ws_partials(dim);
    }
    z_t* dvec = (z_t*)sym_workspace + sym_ws_partials(dim);
    if (NBO % GN != 0 || NBO == NB) {
        // ---- single level / unaligned outer block: one GPU only (small systems, launch bound) ----
        if (P > 1) return cudaErrorInvalidValue;
        for (int K0 = 0; K0 < dim; K0 += NBO) {
            const int JB = dim - K0 < NBO ? dim - K0 : NBO;
            const int KE = K0 + JB;
            for (int k0 = K0; k0 < KE; k0 += NB) {
                const int jb = KE - k0 < NB ? KE - k0 : NB;
                const int ke = k0 + jb;
                const int n_row = (dim - ke + tile - 1) / tile, n_col = (ke + tile - 1) / tile;
                panel_sym_kernel<<<n_row + n_col, 128, PS_SMEM_BYTES, stream>>>(W, Y, ld, dim, k0, jb, 0, ke, n_row,
                                                                                tile, g_tau, d_flag, dvec);
                ++nl;
                // inside the outer block: columns [ke, KE) of W below the panel, rows [ke, KE) of Y
                if (ke < KE) update(ke, dim - ke, ke, KE - ke, KE - ke, 0, ke, k0, jb);
            }
            // everything below the outer block: rank-JB update
            if (KE < dim) update(KE, dim - KE, KE, dim - KE, dim - KE, 0, KE, K0, JB);
        }
    } else {
        // ---- blocked, column-block-cyclic over P ranks (P = 1: the single-GPU path) ----
        z_t* S = YT;                       // scratch rows 0..NBO of YT (YT is written at the very end)
        const int nb = (dim + NBO - 1) / NBO;
        if (P > 1 && nb > PEER_MAX_PANELS) return cudaErrorInvalidValue;
        if (P > 1) {
            // nobody stores into a peer's W / Y before that peer has initialised them
            DCHK(launch_peer_barrier(peers->flags, ++*peers->epoch, stream));
            ++nl;
        }
        auto owner = [&](int K) { return K % P; };
        auto first_own = [&](int from) {   // smallest own block index >= from
            int c = from + ((me - from) % P + P) % P;
            return c;
        };
        // The dependent chain of the step -- bring block K+1 up to date, factor it, publish it -- runs
        // on its own HIGH-priority stream, concurrently with the bulk update of step K on the main
        // stream: the panel kernels are latency bound (a 32 x 32 block factorisation per launch) and
        // need few SMs, so on one GPU the chain (13 % of the step at dim 8192) disappears behind the
        // bulk, and on several GPUs the owner of a panel is not held up by its own factorisation.
        // Fork: the chain waits for `ev_ready` (panel K available and every earlier update done);
        // join: the next step's first use of panel K+1 waits for `ev_chain`.
        cudaStream_t chain = (aux && aux->side) ? aux->side : stream;
        const bool two_streams = chain != stream;
        auto factor_publish = [&](int K, cudaStream_t st) -> cudaError_t {
            const int K0 = K * NBO;
            const int JB = dim - K0 < NBO ? dim - K0 : NBO;
            const int KE = K0 + JB;
            for (int k0 = K0; k0 < KE; k0 += NB) {
                const int jb = KE - k0 < NB ? KE - k0 : NB;
                const int ke = k0 + jb;
                const int n_row = (dim - ke + tile - 1) / tile, n_col = (ke - K0 + tile - 1) / tile;
                panel_sym_kernel<<<n_row + n_col, 128, PS_SMEM_BYTES, st>>>(W, Y, ld, dim, k0, jb, K0, ke, n_row,
                                                                            tile, g_tau, d_flag, dvec);
                ++nl;
                if (ke < KE) update_on(st, ke, dim - ke, ke, KE - ke, KE - ke, K0, ke - K0, k0, jb);
            }
            if (P > 1) {
                // the owner of the NEXT panel is on the critical path (it must update its block with
                // this panel and factor it): it is served and signalled first, the others follow
                const size_t total = (size_t)(dim - KE) * JB + (size_t)JB * JB;
                int blocks = (int)((total + 1023) / 1024);
                if (blocks > 1184) blocks = 1184;
                const unsigned all = (P >= 32 ? 0xffffffffu : ((1u << P) - 1u)) & ~(1u << me);
                const int next = (K + 1 < nb) ? owner(K + 1) : -1;
                unsigned first = (next >= 0 && next != me) ? (1u << next) : 0u;
                if (first) {
                    publish_panel_kernel<<<blocks, 256, 0, st>>>(sp, ld, dim, K0, JB, first);
                    DCHK(launch_peer_signal(peers->flags, PEER_W_PANEL + K, peers->serial, st, first));
                    nl += 2;
                }
                if (all & ~first) {
                    publish_panel_kernel<<<blocks, 256, 0, st>>>(sp, ld, dim, K0, JB, all & ~first);
                    DCHK(launch_peer_signal(peers->flags, PEER_W_PANEL + K, peers->serial, st, all & ~first));
                    nl += 2;
                }
            }
            return cudaGetLastError();
        };
        auto shard_update = [&](cudaStream_t st, int K0, int JB, int wblk0, int wblk_end, int yblk0, int yblk_end) {
            // W blocks wblk0, wblk0+P, ... < wblk_end; Y blocks yblk0, yblk0+P, ... < yblk_end
            const int KE = K0 + JB;
            const int M = dim - KE;
            if (M <= 0) return;
            const int tpb = NBO / GN;
            const int nwb = wblk0 < wblk_end ? (wblk_end - wblk0 + P - 1) / P : 0;
            const int nyb = yblk0 < yblk_end ? (yblk_end - yblk0 + P - 1) / P : 0;
            if (nwb + nyb == 0) return;
            dim3 g((nwb + nyb) * tpb, (M + GM - 1) / GM);
            shard_update_kernel<<<g, 256, G_SMEM_BYTES, st>>>(W, Y, S, ld, dim, K0, JB, wblk0, nwb * tpb, yblk0,
                                                              P, NBO);
            ++nl;
        };
        if (owner(0) == me) DCHK(factor_publish(0, stream));
        bool chain_pending = false;
        for (int K = 0; K < nb; ++K) {
            const int K0 = K * NBO;
            const int JB = dim - K0 < NBO ? dim - K0 : NBO;
            const int KE = K0 + JB;
            if (owner(K) != me) {
                DCHK(launch_peer_wait(peers->flags, PEER_W_PANEL + K, peers->serial, stream));
                ++nl;
                if (KE < dim) {   // rebuild the L panel from the received U12 and pivots
                    lpanel_from_u12_kernel<<<dim3((JB + 31) / 32, (dim - KE + 31) / 32), dim3(32, 8), 0, stream>>>(
                        W, dvec, ld, dim, K0, JB);
                    ++nl;
                }
            } else if (chain_pending) {
                DCHK(cudaStreamWaitEvent(stream, aux->ev_rest, 0));      // join: my chain factored panel K
                chain_pending = false;
            }
            // look-ahead (the critical path of the whole step runs through the panels): the owner of
            // panel K+1 brings that block up to date, factors and publishes it -- on the chain stream
            const bool ahead = KE < dim && K + 1 < nb && owner(K + 1) == me;
            if (ahead) {
                if (two_streams) {
                    DCHK(cudaEventRecord(aux->ev_panel, stream));        // fork: panel K + all earlier updates
                    DCHK(cudaStreamWaitEvent(chain, aux->ev_panel, 0));
                }
                shard_update(chain, K0, JB, K + 1, K + 2, 0, 0);
                DCHK(factor_publish(K + 1, chain));
                if (two_streams) {
                    DCHK(cudaEventRecord(aux->ev_rest, chain));
                    chain_pending = true;
                }
            }
            // rows K of my columns of Y become final: Y_K <- M_KK Y_K (blocks < K), via the scratch S
            const int y0 = first_own(0);
            if (y0 < K) {
                const int nyb = (K - y0 + P - 1) / P;
                dim3 g(nyb * (NBO / GN), (JB + GM - 1) / GM);
                ydiag_kernel<<<g, 256, G_SMEM_BYTES, stream>>>(sp, S, ld, dim, K0, JB, y0, P, NBO);
                dim3 gc(nyb * (NBO / GN), (JB + 15) / 16 < 8 ? (JB + 15) / 16 : 8);
                ycopy_kernel<<<gc, 256, 0, stream>>>(Y, S, ld, K0, JB, y0, P, NBO);
                nl += 2;
            }
            if (owner(K) == me) {
                // the owner's own block: S_K = M_KK (B operand of the update of Y[KE:, block K])
                DCHK(cudaMemcpy2DAsync(S + K0, sizeof(z_t) * ld, at(Y, K0, K0), sizeof(z_t) * ld, sizeof(z_t) * JB, JB,
                                       cudaMemcpyDeviceToDevice, stream));
            }
            if (KE >= dim) continue;
            if (ahead) shard_update(stream, K0, JB, K + 1 + P, nb, y0, K + 1);
            else shard_update(stream, K0, JB, first_own(K + 1), nb, y0, K + 1);
        }
        if (chain_pending) DCHK(cudaStreamWaitEvent(stream, aux->ev_rest, 0));
        if (P > 1) {
            // every rank's final rows of Y have arrived everywhere
            DCHK(launch_peer_barrier(peers->flags, ++*peers->epoch, stream));
            ++nl;
        }
    }
    // A^-1 = M^T D^-1 M on the lower block triangle, contracted with B on the fly
    {
        const int nt32 = (dim + 31) / 32;
        transpose_invd_kernel<<<dim3(nt32, nt32), dim3(32, 8), 0, stream>>>(Y, dvec, YT, dim);
        ++nl;
        const int nt = (dim + GM - 1) / GM;
        const int ntiles = nt * (nt + 1) / 2;
        int nchunks = (600 + ntiles - 1) / ntiles;
        if (nchunks > PTRACE_MAX_CHUNKS) nchunks = PTRACE_MAX_CHUNKS;
        if (nchunks > (dim + 127) / 128) nchunks = (dim + 127) / 128;
        if (nchunks < 1) nchunks = 1;
        const int ck = (((dim + nchunks - 1) / nchunks) + GK - 1) / GK * GK;
        const int my_tiles = (ntiles - me + P - 1) / P;
        if (my_tiles > 0) {
            ptrace_kernel<<<dim3(my_tiles, nchunks), 256, G_SMEM_BYTES, stream>>>(YT, Y, (const z_t*)Bv, dim, ck, pd,
                                                                                  ntiles, me, P);
            ++nl;
        }
        if (P > 1) {
            // "my partials are stored everywhere" + flags; wait for everybody's, OR the flags
            DCHK(launch_peer_post(peers->flags, peers->serial, (const double2*)d_trace, d_flag, d_info, stream));
            DCHK(launch_peer_collect(peers->flags, peers->serial, (double2*)d_trace, d_flag, d_info, stream));
            nl += 2;
        }
        reduce_partials_kernel<<<1, 256, 0, stream>>>((const z_t*)sym_workspace, ntiles * nchunks,
                                                      (z_t*)d_trace);
        ++nl;
    }
    if (n_launches) *n_launches += nl;
    return cudaGetLastError();
}

// Load every kernel of this translation unit now.  CUDA loads kernels lazily, on their first launch,
// and that load may need the context to drain -- which never happens while a device-side wait of the
// peer protocol is spinning for exactly the work whose launch is being loaded (a deadlock until the
// wait gives up; seen with several ranks on one device).  Called once, before any wait can exist.
#define EMME_TOUCH(k)                                                        \
    do {                                                                     \
        cudaFuncAttributes a__;                                              \
        cudaError_t e__ = cudaFuncGetAttributes(&a__, (const void*)(k));     \
        if (e__ != cudaSuccess) return e__;                                  \
    } while (0)
cudaError_t dense_preload() {
    EMME_TOUCH(panel_nopiv_kernel<128>);
    EMME_TOUCH(panel_kernel);
    EMME_TOUCH(panel_cluster_kernel<CL_TPB>);
    EMME_TOUCH(laswp_kernel);
    EMME_TOUCH(swap_trsm_kernel);
    EMME_TOUCH(utrsm_kernel);
    EMME_TOUCH(zgemm_sub_kernel);
    EMME_TOUCH(zgemm_sub2_kernel);
    EMME_TOUCH(sym_copy_check_kernel);
    EMME_TOUCH(set_identity_diag_kernel);
    EMME_TOUCH(panel_sym_kernel);
    EMME_TOUCH(zgemm_sym2_kernel);
    EMME_TOUCH(transpose_invd_kernel);
    EMME_TOUCH(ptrace_kernel);
    EMME_TOUCH(reduce_partials_kernel);
    EMME_TOUCH(trace_kernel);
    EMME_TOUCH(secant_kernel);
    EMME_TOUCH(conj_normalise_kernel);
    EMME_TOUCH(publish_panel_kernel);
    EMME_TOUCH(lpanel_from_u12_kernel);
    EMME_TOUCH(ydiag_kernel);
    EMME_TOUCH(ycopy_kernel);
    EMME_TOUCH(shard_update_kernel);
    return cudaSuccess;
}

}  // namespace emme
