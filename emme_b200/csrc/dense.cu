// dense.cu -- kernel 2: the dense step of the Newton/secant root find, on the device.
//
// Replaces the LAPACK call of EigenSolver::newtonTraceSecantIteration
// (reference include/solver.h:129-140): zsysv solves A X = A' for dim right-hand sides and
// the step is delta = -1/trace(X).  Here:
//   1. blocked right-looking LU with partial pivoting of W = A (row-major complex128), the
//      elimination applied on the fly to B = A' (augmented system), so the forward
//      substitution Y = L^-1 P B costs no extra pass;
//      - panel (dim-k0 rows x NB columns): ONE cooperative kernel per panel, one thread per
//        matrix row holding its NB panel entries in registers, one grid barrier per column
//        (the block-local pivot candidate travels with its whole row, so a second barrier
//        for "publish the pivot row" is not needed); row interchanges are implicit
//        (threads track their row's position) and materialise when rows are written back;
//      - one fused kernel applies the interchanges to the other columns of W and to B and
//        solves the NB x NB unit-lower triangle for the block row;
//      - trailing update: register-tiled complex DGEMM (DFMA), K = NB.
//   2. blocked back substitution X = U^-1 Y restricted to the LOWER triangle of X -- the
//      trace needs X_ii only, and X_lc (l >= c) depends on nothing above the diagonal -- which
//      halves this phase;
//   3. deterministic trace reduction; the secant quotient (A - A_old)/delta
//      (include/solver.h:54-57) is an elementwise kernel.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include "dense.h"

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
        const double myv = done ? -1.0 : fabs(a[0].x) + fabs(a[0].y);
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
        const int ppos = s_best.pos;
        const int diag = k0 + c;
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
    static int limit = -1;
    if (limit >= 0) return limit;
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
// One thread per column of [W(:, k0+jb:) | B(:, :)]: apply the jb row interchanges of the
// panel, then forward-substitute with the unit-lower NB x NB block.
__global__ void __launch_bounds__(128)
swap_trsm_kernel(z_t* __restrict__ W, z_t* __restrict__ B, int ld, int dim, int k0, int jb,
                 const int* __restrict__ ipiv) {
    __shared__ z_t sL[NB][NB + 1];
    __shared__ int sP[NB];
    for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
        const int rr = e / NB, cc = e % NB;
        sL[rr][cc] = (rr < jb && cc < rr) ? W[(size_t)(k0 + rr) * ld + k0 + cc] : make_double2(0., 0.);
    }
    if (threadIdx.x < NB) sP[threadIdx.x] = threadIdx.x < jb ? ipiv[k0 + threadIdx.x] : 0;
    __syncthreads();
    const int nW = dim - (k0 + jb);                 // columns of W right of the panel
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= nW + dim) return;
    z_t* M = col < nW ? W + (k0 + jb + col) : B + (col - nW);
    // interchanges, in order
    for (int c = 0; c < jb; ++c) {
        const int p = sP[c];
        if (p != k0 + c) {
            const z_t t1 = M[(size_t)(k0 + c) * ld], t2 = M[(size_t)p * ld];
            M[(size_t)(k0 + c) * ld] = t2;
            M[(size_t)p * ld] = t1;
        }
    }
    // unit-lower solve on rows k0 .. k0+jb-1
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
constexpr int GM = 64, GN = 64, GK = 16;

__device__ __forceinline__ void zgemm_sub_tile(z_t* __restrict__ C, int ldc,
                                               const z_t* __restrict__ A, int lda,
                                               const z_t* __restrict__ Bm, int ldb, int M, int N,
                                               int K, int bx, int by) {
    __shared__ z_t sA[GK][GM + 1];
    __shared__ z_t sB[GK][GN];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = by * GM, n0 = bx * GN;
    z_t acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = make_double2(0., 0.);

    for (int kk = 0; kk < K; kk += GK) {
        // A tile: GM rows x GK cols -> sA[k][m]; 1024 elements / 256 threads
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = threadIdx.x + e * 256;
            const int m = idx / GK, k = idx % GK;
            z_t v = make_double2(0., 0.);
            if (m0 + m < M && kk + k < K) v = A[(size_t)(m0 + m) * lda + kk + k];
            sA[k][m] = v;
        }
        // B tile: GK rows x GN cols -> sB[k][n]
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = threadIdx.x + e * 256;
            const int k = idx / GN, n = idx % GN;
            z_t v = make_double2(0., 0.);
            if (kk + k < K && n0 + n < N) v = Bm[(size_t)(kk + k) * ldb + n0 + n];
            sB[k][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            z_t av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = sA[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = sB[k][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][j].x = fma(av[i].x, bv[j].x, acc[i][j].x);
                    acc[i][j].x = fma(-av[i].y, bv[j].y, acc[i][j].x);
                    acc[i][j].y = fma(av[i].x, bv[j].y, acc[i][j].y);
                    acc[i][j].y = fma(av[i].y, bv[j].x, acc[i][j].y);
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx + 16 * j;
            if (n >= N) continue;
            z_t c = C[(size_t)m * ldc + n];
            c.x -= acc[i][j].x;
            c.y -= acc[i][j].y;
            C[(size_t)m * ldc + n] = c;
        }
    }
}

__global__ void __launch_bounds__(256)
zgemm_sub_kernel(z_t* __restrict__ C, int ldc, const z_t* __restrict__ A, int lda,
                 const z_t* __restrict__ Bm, int ldb, int M, int N, int K) {
    zgemm_sub_tile(C, ldc, A, lda, Bm, ldb, M, N, K, blockIdx.x, blockIdx.y);
}

// Trailing update of the augmented system in ONE launch: the same L21 (A) multiplies the
// block row of W (C1 -= A*B1, N1 columns) and the block row of the right-hand sides
// (C2 -= A*B2, N2 columns); column tiles [0, nx1) belong to the first product.
__global__ void __launch_bounds__(256)
zgemm_sub2_kernel(z_t* __restrict__ C1, const z_t* __restrict__ B1, int N1, int nx1,
                  z_t* __restrict__ C2, const z_t* __restrict__ B2, int N2, int ld,
                  const z_t* __restrict__ A, int M, int K) {
    if ((int)blockIdx.x < nx1)
        zgemm_sub_tile(C1, ld, A, ld, B1, ld, M, N1, K, blockIdx.x, blockIdx.y);
    else
        zgemm_sub_tile(C2, ld, A, ld, B2, ld, M, N2, K, blockIdx.x - nx1, blockIdx.y);
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

size_t dense_workspace_bytes(int dim) {
    const int nblk = (dim + PT - 1) / PT;
    return sizeof(PanelCand) * 2 * (size_t)nblk + sizeof(int) * (size_t)dim + 64;
}

// W (dim x dim, destroyed) and B (dim x dim, destroyed): trace(W^-1 B) -> *d_trace (device).
cudaError_t launch_trace_solve(void* Wv, void* Bv, int dim, void* workspace, void* d_trace,
                               int* d_info, cudaStream_t stream, unsigned long long* n_launches) {
    unsigned long long nl = 0;
    z_t* W = (z_t*)Wv;
    z_t* B = (z_t*)Bv;
    const int ld = dim;
    const int nblk_max = (dim + PT - 1) / PT;
    PanelCand* xchg = (PanelCand*)workspace;
    int* ipiv = (int*)((char*)workspace + sizeof(PanelCand) * 2 * (size_t)nblk_max);
    cudaError_t e = cudaMemsetAsync(d_info, 0, sizeof(int), stream);
    if (e != cudaSuccess) return e;

    for (int k0 = 0; k0 < dim; k0 += NB) {
        int jb = dim - k0 < NB ? dim - k0 : NB;
        int rows = dim - k0;
        int nblk = (rows + PT - 1) / PT;
        if (rows <= CL_TPB * panel_cluster_limit() && !g_force_grid_panel) {
            e = launch_panel_cluster(W, ld, dim, k0, jb, ipiv, d_info, stream);
        } else {
            void* args[] = {&W, (void*)&ld, &dim, &k0, &jb, &ipiv, &xchg, &d_info};
            e = cudaLaunchCooperativeKernel((void*)panel_kernel, dim3(nblk), dim3(PT), args, 0, stream);
        }
        if (e != cudaSuccess) return e;
        nl += 2;
        const int ncol = (dim - (k0 + jb)) + dim;
        swap_trsm_kernel<<<(ncol + 127) / 128, 128, 0, stream>>>(W, B, ld, dim, k0, jb, ipiv);
        const int M = dim - (k0 + jb);
        if (M > 0) {
            nl += 1;
            const int nx1 = (M + GN - 1) / GN, nx2 = (dim + GN - 1) / GN;
            dim3 g(nx1 + nx2, (M + GM - 1) / GM);
            zgemm_sub2_kernel<<<g, 256, 0, stream>>>(
                W + (size_t)(k0 + jb) * ld + (k0 + jb), W + (size_t)k0 * ld + (k0 + jb), M, nx1,
                B + (size_t)(k0 + jb) * ld, B + (size_t)k0 * ld, dim, ld,
                W + (size_t)(k0 + jb) * ld + k0, M, jb);
        }
    }
    // back substitution, lower triangle of X only
    const int last = ((dim - 1) / NB) * NB;
    for (int k0 = last; k0 >= 0; k0 -= NB) {
        const int jb = dim - k0 < NB ? dim - k0 : NB;
        const int ncols = k0 + jb;   // columns c <= last row of this block
        utrsm_kernel<<<(ncols + 127) / 128, 128, 0, stream>>>(W, B, ld, k0, jb, ncols);
        nl += 1;
        if (k0 > 0) {
            nl += 1;
            dim3 g((k0 + GN - 1) / GN, (k0 + GM - 1) / GM);
            zgemm_sub_kernel<<<g, 256, 0, stream>>>(B, ld, W + k0, ld, B + (size_t)k0 * ld, ld, k0,
                                                    k0, jb);
        }
    }
    trace_kernel<<<1, 256, 0, stream>>>(B, ld, dim, (z_t*)d_trace);
    if (n_launches) *n_launches += nl + 1;
    return cudaGetLastError();
}

}  // namespace emme
