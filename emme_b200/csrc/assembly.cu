// assembly.cu -- kernel 1: adaptive Gauss-Kronrod assembly of A(omega), one lane per quadrature.
//
// Replaces EigenSolver::matrixAssembler (reference include/solver.h:417-515) together with
// everything it calls per matrix element: Parameters::kappa_f_tau / kappa_f_tau_e
// (src/Parameters.cpp:113-209), util::integrate + gauss_kronrod_adaptive/basic
// (include/functions.h:181-251,305-331), util::bessel_i_alter_helper (:381-408),
// SingularityHandler (src/singularity_handler.cpp:3-24) and the DedicatedThreadPool fan-out.
//
// Mapping (round-1 profile r1a motivated the change from "lane = node"):
//   * a work item is one adaptive quadrature (pair i<j, mode m); items are ordered diagonal-major
//     (d = j-i), the modes of an EM run inside each diagonal, so consecutive items are
//     near-identical integrals (same |i-j|, same m, neighbouring eta); the expensive diagonals run
//     first: the far ones (long Miller recurrences) where the geometry makes them costly, then
//     d = 1, 2, ... (decode_item);
//   * a LANE owns one item and walks its panels; the 32 lanes of a warp walk the 15 (31)
//     Kronrod nodes of their current panels in lockstep (node index j is warp-uniform), so
//     - the Miller recurrences of neighbouring integrals at the same node have nearly the
//       same trip count (little divergence),
//     - the exp(<-40) underflow guard skips the Bessel part for the whole warp at once,
//     - the panel sums are accumulated sequentially in the reference's own order (centre,
//       then f(+a_i)+f(-a_i), i = 1..) with no shuffles;
//   * the LIFO interval stack of gauss_kronrod_adaptive is per lane: the left child stays in
//     registers, right siblings go to shared memory and, past STACK_SMEM entries, to a global
//     spill area sized for the contract "depth <= integration_iteration_limit";
//   * persistent grid; at every panel boundary the idle lanes of a warp (finished integrals)
//     refill TOGETHER from a global atomic counter once at least `refill_min` of them are idle,
//     so a warp always consists of a few cohorts of consecutive items;
//   * everything an evaluation needs that does not depend on the pair (eta, eta') -- the contour
//     rotation, its jacobian, i*tau~*omega -- comes from a table of the first bisection levels that
//     node_table_kernel rebuilds at the start of every assembly (NodeConst, emme_eval.cuh);
//   * the kernel is ISSUE-SLOT bound (an FP64 instruction occupies a scheduler for two cycles,
//     nothing issues in its shadow), so per-lane state that is only touched at
//     item or panel boundaries lives in shared memory and the evaluation is written for the
//     smallest weighted instruction count.
#include <cuda_runtime.h>
#include <stdint.h>

#include "assembly.h"
#include "emme_eval.cuh"
#include "gk_tables.h"

namespace emme {

__constant__ GKTables c_gk15 = EMME_GK15_INIT;
__constant__ GKTables c_gk31 = EMME_GK31_INIT;

constexpr int STACK_SMEM = 8;    // right-sibling intervals kept in shared memory per lane
constexpr int BLOCK = 128;
constexpr int TRIG_DEPTH = 8;                           // node table covers bisection levels 0..8
constexpr int TRIG_PANELS = (1 << (TRIG_DEPTH + 1)) - 1;   // heap-ordered panels
#ifndef EMME_ASM_MIN_BLOCKS
#define EMME_ASM_MIN_BLOCKS 5
#endif
constexpr int MIN_BLOCKS = EMME_ASM_MIN_BLOCKS;

// item index k -> (i, j, m).  Diagonal-major (d = j-i); pairs on diagonals before d in the plain
// order: T(d) = (d-1)*(2N-d)/2.  The nm modes of an electromagnetic run (3 integrals per pair) sit
// INSIDE the diagonal -- all pairs of diagonal d for m = 0, then m = 1, then m = 2 -- so that cohorts
// stay uniform in m and the near-singular diagonals of every mode run first.  (Mode-major, the order
// of rounds 1-2, started the deep bisections of m = 1, 2 one and two thirds into the queue: in C3 one
// such cohort lasts longer than the balanced share of a warp, so they set the length of the launch.)
// d_split = 0: d = 1, 2, ..., N-1.  d_split > 0: the far diagonals first, d = N-1 down to d_split
// (lengths 1, 2, ...), then d = 1 .. d_split-1.  Where the Bessel argument grows with |eta| (every
// geometry: b ~ k_rho^2 (1 + s^2 eta^2)) the far pairs carry the longest Miller recurrences and are
// the costliest items after the near-singular ones; left at the end of the queue they are the tail of
// the launch (C1: the last cohorts cost 1.5x the mean, 5.5 cohorts per warp).  The host decides
// (capi.cu::choose_item_order); the entries do not depend on the order.
__device__ __forceinline__ void decode_item(unsigned long long k, int N, int nm, int d_split, int& i, int& j,
                                            int& m) {
    unsigned long long p = k / (unsigned)nm;     // a pair on the item's diagonal
    long long d;
    unsigned long long base;                      // pairs on the diagonals that come before d in this order
    const unsigned long long n_far = d_split > 0 ? (unsigned long long)(N - d_split) * (N - d_split + 1) / 2 : 0ULL;
    if (p < n_far) {
        long long L = (long long)floor((1.0 + sqrt(1.0 + 8.0 * (double)p)) * 0.5);
        if (L < 1) L = 1;
        while (L > 1 && (unsigned long long)(L * (L - 1) / 2) > p) --L;
        while ((unsigned long long)(L * (L + 1) / 2) <= p) ++L;
        d = N - L;
        base = (unsigned long long)(L * (L - 1) / 2);
    } else {
        p -= n_far;
        const double tn = 2.0 * N + 1.0;
        double disc = tn * tn - 8.0 * ((double)N + (double)p);
        if (disc < 0.) disc = 0.;
        d = (long long)floor((tn - sqrt(disc)) * 0.5);
        if (d < 1) d = 1;
        if (d > N - 1) d = N - 1;
        // fix up rounding of the closed form
        while (d > 1 && (unsigned long long)((d - 1) * (2LL * N - d) / 2) > p) --d;
        while (d < N - 1 && (unsigned long long)(d * (2LL * N - d - 1) / 2) <= p) ++d;
        base = n_far + (unsigned long long)((d - 1) * (2LL * N - d) / 2);
    }
    const int L = N - (int)d;
    const unsigned r = (unsigned)(k - (unsigned long long)nm * base);    // position among the nm*L items of d
    m = (int)(r / (unsigned)L);
    const int t = (int)(r - (unsigned)m * (unsigned)L);
    // Position t on the diagonal -> row i, alternating between the two ends (0, L-1, 1, L-2, ...):
    // a cohort of 32 consecutive items then holds 16 neighbouring pairs and their 16 mirror images
    // (eta -> -eta), which cost the same on the symmetric geometries, instead of 32 neighbours -- the
    // spread of Miller trip counts inside a warp halves (it matters on small grids, where 32
    // neighbours span a wide range of eta: N = 1024 ran at 28.2 of 32 lanes per instruction).
    i = (t & 1) ? (L - 1 - (t >> 1)) : (t >> 1);
    j = i + (int)d;
}

// Destination(s) of the assembled entries: the local matrix, or -- pair-sharded multi-GPU
// assembly -- the same matrix on every GPU of the node (peer pointers mapped over NVLink): each
// rank stores its entries straight into all peers while it computes, so no collective follows.
__device__ __forceinline__ void store_c(const PeerSet& dst, size_t idx, cplx v) {
    const double2 e = make_double2(v.re, v.im);
#pragma unroll
    for (int r = 0; r < EMME_MAX_PEERS; ++r)
        if (r < dst.n) dst.p[r][idx] = e;
}

// Write the entries produced by item (i, j, m) -- include/solver.h:448-453 and :476-504.
__device__ __forceinline__ void scatter(const RunConst& rc, const PeerSet& A, int i, int j, int m,
                                        cplx kappa_all) {
    const int N = rc.N;
    const size_t dim = rc.em ? 2 * (size_t)N : (size_t)N;
    if (m == 0) {
        const double w = sing_weight(N, i, j);
        cplx a = -kappa_all;
        a = mk(a.re * w, a.im * w);
        a = mk(a.re * rc.dx, a.im * rc.dx);
        store_c(A, (size_t)i * dim + j, a);
        store_c(A, (size_t)j * dim + i, a);
    } else if (m == 1) {
        const cplx a = mk(kappa_all.re * rc.dx, kappa_all.im * rc.dx);
        store_c(A, (size_t)i * dim + (j + N), a);
        store_c(A, (size_t)j * dim + (i + N), -a);
        store_c(A, (size_t)(i + N) * dim + j, -a);
        store_c(A, (size_t)(j + N) * dim + i, a);
    } else {
        const cplx a = mk(kappa_all.re * rc.dx, kappa_all.im * rc.dx);
        store_c(A, (size_t)(i + N) * dim + (j + N), a);
        store_c(A, (size_t)(j + N) * dim + (i + N), a);
    }
}

// Per-lane state that is only touched at item or panel boundaries lives in shared memory instead
// of registers ([field][thread]: conflict-free): the pair constants of the integrand (7 doubles,
// one load each per evaluation), the current interval and the running sum.  This is what lets five
// CTAs (20 warps) stay resident per SM without spills (r1j).
enum { PS_DV, PS_CL, PS_S, PS_TWO_OVER_S, PS_HB, PS_C1, PS_CA, PS_CB, PS_CAA, PS_FIELDS };
struct PairSmem {
    const volatile double* p;     // &s_pair[0][threadIdx.x]
    __device__ __forceinline__ double f_Dv() const { return p[PS_DV * BLOCK]; }
    __device__ __forceinline__ double f_ca() const { return p[PS_CA * BLOCK]; }
    __device__ __forceinline__ double f_cb() const { return p[PS_CB * BLOCK]; }
    __device__ __forceinline__ double f_cA() const { return p[PS_CAA * BLOCK]; }
    __device__ __forceinline__ double f_cl() const { return p[PS_CL * BLOCK]; }
    __device__ __forceinline__ double f_s() const { return p[PS_S * BLOCK]; }
    __device__ __forceinline__ double f_two_over_s() const { return p[PS_TWO_OVER_S * BLOCK]; }
    __device__ __forceinline__ double f_hb() const { return p[PS_HB * BLOCK]; }
    __device__ __forceinline__ double f_c1() const { return p[PS_C1 * BLOCK]; }
};
enum { LS_L, LS_R, LS_ABS_TOL, LS_SUM_RE, LS_SUM_IM, LS_FIELDS };

template <int ORDER>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS)
assemble_kernel(const RunConst rc, const double* __restrict__ eta, const double* __restrict__ gt,
                const double* __restrict__ bt, const PeerSet A,
                unsigned long long n_items_local, unsigned long long shard_index,
                unsigned long long shard_count, unsigned long long* __restrict__ counter,
                double2* __restrict__ spill, int spill_cap, unsigned long long* __restrict__ stats,
                int refill_min, const NodeConst* __restrict__ table, int d_split) {
    constexpr int H = (ORDER - 1) / 2;          // 7 or 15 symmetric node pairs
    const GKTables& T = ORDER == 15 ? c_gk15 : c_gk31;

    __shared__ double2 s_stack[STACK_SMEM][BLOCK];   // [slot][thread]: conflict-free
    __shared__ int s_pid[STACK_SMEM][BLOCK];         // heap index of the stacked panel (-1: untabulated)
    __shared__ double s_pair[PS_FIELDS][BLOCK];
    __shared__ double s_lane[LS_FIELDS][BLOCK];
    volatile double* const lane_state = &s_lane[0][threadIdx.x];
    const PairSmem pc{&s_pair[0][threadIdx.x]};

    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    double2* my_spill =
        spill ? spill + ((size_t)blockIdx.x * BLOCK + threadIdx.x) * (size_t)spill_cap : nullptr;

    bool active = false;
    bool warp_exhausted = false;
    int it_i = 0, it_j = 0, it_m = 0, top = 0, pid = 0;
    EvalCounters cnt{0u, 0u};
    unsigned int n_eval32 = 0, n_panel32 = 0, n_int32 = 0;   // per thread: far below 2^32
    int max_top = 0;

    for (;;) {
        // ---- refill: idle lanes fetch consecutive items together ----
        const unsigned idle = __ballot_sync(0xffffffffu, !active);
        if (idle != 0u && !warp_exhausted && (__popc(idle) >= refill_min || idle == 0xffffffffu)) {
            const int leader = __ffs(idle) - 1;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, leader);
            bool miss = false;
            if (!active) {
                const unsigned long long k = base + (unsigned)__popc(idle & lt_mask);
                if (k >= n_items_local) {
                    miss = true;
                } else {
                    // shards take chunks of 32 consecutive items round-robin
                    const unsigned long long kg = (k >> 5) * (shard_count << 5) + (shard_index << 5) + (k & 31);
                    decode_item(kg, rc.N, rc.em ? 3 : 1, d_split, it_i, it_j, it_m);
                    {
                        const PairConst c = make_pair(rc, eta[it_i], eta[it_j], gt[it_i], gt[it_j], bt[it_i], bt[it_j]);
                        volatile double* ps = &s_pair[0][threadIdx.x];
                        ps[PS_DV * BLOCK] = c.Dv;
                        ps[PS_CA * BLOCK] = c.ca;
                        ps[PS_CB * BLOCK] = c.cb;
                        ps[PS_CAA * BLOCK] = c.cA;
                        ps[PS_CL * BLOCK] = c.cl;
                        ps[PS_S * BLOCK] = c.s;
                        ps[PS_TWO_OVER_S * BLOCK] = c.two_over_s;
                        ps[PS_HB * BLOCK] = c.hb;
                        ps[PS_C1 * BLOCK] = c.c1;
                    }
                    lane_state[LS_SUM_RE * BLOCK] = 0.;
                    lane_state[LS_SUM_IM * BLOCK] = 0.;
                    lane_state[LS_ABS_TOL * BLOCK] = 0.;
                    lane_state[LS_L * BLOCK] = 0.0;
                    lane_state[LS_R * BLOCK] = rc.half_pi;
                    top = 0;
                    pid = 0;
                    active = true;
                }
            }
            if (__any_sync(0xffffffffu, miss)) warp_exhausted = true;
        }
        if (__ballot_sync(0xffffffffu, active) == 0u) {
            if (warp_exhausted) break;
            continue;
        }
        // ---- one Gauss-Kronrod panel per active lane, nodes in lockstep ----
        cplx K = mk(0., 0.), G = mk(0., 0.), fplus = mk(0., 0.);
#pragma unroll 1
        for (int j = 0; j <= 2 * H; ++j) {
            const int ni = (j + 1) >> 1;                       // node index 0..H
            if (active) {
                NodeConst nc;
                if (pid >= 0) {
                    // tabulated panel: the node constants of this (panel, node) for this omega, built
                    // by node_table_kernel with the very same node_const() and node position arithmetic
                    const double2* e = reinterpret_cast<const double2*>(table + (size_t)pid * (2 * H + 1) + j);
                    const double2 e0 = __ldg(e), e1 = __ldg(e + 1), e2 = __ldg(e + 2), e3 = __ldg(e + 3),
                                  e4 = __ldg(e + 4), e5 = __ldg(e + 5);
                    nc.taut = mk(e0.x, e0.y);
                    nc.itaut = mk(e1.x, e1.y);
                    nc.h = mk(e2.x, e2.y);
                    nc.M = mk(e3.x, e3.y);
                    nc.pj = mk(e4.x, e4.y);
                    nc.icsq = e5.x;
                } else {
                    // node position exactly as the reference forms it: scale*x + mid, no FMA
                    const double node = (j & 1) ? T.a[ni] : -T.a[ni];  // j = 0: -0*scale + mid = mid
                    const double l = lane_state[LS_L * BLOCK], r = lane_state[LS_R * BLOCK];
                    const double mid = (r + l) / 2, scale = (r - l) / 2;
                    nc = node_const(rc, __dadd_rn(__dmul_rn(scale, node), mid));
                }
                const cplx fx = eval_node(rc, pc, it_m, nc, cnt);
                ++n_eval32;
                if (j == 0) {
                    K = mk(T.kw[0] * fx.re, T.kw[0] * fx.im);
                    G = mk(T.gw[0] * fx.re, T.gw[0] * fx.im);
                } else if (j & 1) {
                    fplus = fx;
                } else {
                    const cplx f = mk(fplus.re + fx.re, fplus.im + fx.im);
                    const double kw = T.kw[ni], gw = T.gw[ni];
                    if ((ni & 1) == 0) {
                        G.re += gw * f.re;
                        G.im += gw * f.im;
                    }
                    K.re += kw * f.re;
                    K.im += kw * f.im;
                }
            }
        }
        if (active) {
            ++n_panel32;
            // ---- accept / bisect (include/functions.h:233-247) ----
            const double l = lane_state[LS_L * BLOCK], r = lane_state[LS_R * BLOCK];
            const double mid = (r + l) / 2, scale = (r - l) / 2;
            double abs_tol = lane_state[LS_ABS_TOL * BLOCK];
            const cplx integral = mk(K.re * scale, K.im * scale);
            const double e0 = fmax(hypot(K.re - G.re, K.im - G.im),
                                   hypot(K.re, K.im) * 2.220446049250313e-16 * 2);
            const double err = e0 * scale;
            const double rel = hypot(rc.tol * integral.re, rc.tol * integral.im);
            if (abs_tol == 0.) {
                abs_tol = rel;
                lane_state[LS_ABS_TOL * BLOCK] = rel;
            }
            const bool split = ldexp(scale, rc.maxdepth) > rc.thr_len &&
                               err > abs_tol * rc.inv_scale + rc.prec && err > rel + rc.prec;
            if (split) {
                // push the right half, continue with the left half (LIFO: left is next)
                const double2 e = make_double2(mid, r);
                // children of heap node p are 2p+1 (left) and 2p+2 (right)
                const int right = (pid >= 0 && 2 * pid + 2 < TRIG_PANELS) ? 2 * pid + 2 : -1;
                if (top < STACK_SMEM) {
                    s_stack[top][threadIdx.x] = e;
                    s_pid[top][threadIdx.x] = right;
                } else {
                    my_spill[top - STACK_SMEM] = e;   // deeper than the table: pid is -1 by construction
                }
                ++top;
                max_top = max(max_top, top);
                lane_state[LS_R * BLOCK] = mid;
                pid = right >= 0 ? right - 1 : -1;
            } else {
                const cplx sum = mk(lane_state[LS_SUM_RE * BLOCK] + integral.re,
                                    lane_state[LS_SUM_IM * BLOCK] + integral.im);
                if (top > 0) {
                    lane_state[LS_SUM_RE * BLOCK] = sum.re;
                    lane_state[LS_SUM_IM * BLOCK] = sum.im;
                    --top;
                    const double2 e = top < STACK_SMEM ? s_stack[top][threadIdx.x] : my_spill[top - STACK_SMEM];
                    pid = top < STACK_SMEM ? s_pid[top][threadIdx.x] : -1;
                    lane_state[LS_L * BLOCK] = e.x;
                    lane_state[LS_R * BLOCK] = e.y;
                } else {
                    // ---- finalize: kappa = -i*pref*sum (+ electron part), scatter ----
                    cplx kap = mk(rc.kappa_pref * sum.im, -rc.kappa_pref * sum.re);
                    if (it_m > 0) kap = kap + kappa_e(rc, it_m, eta[it_i] - eta[it_j], gt[it_i] - gt[it_j]);
                    scatter(rc, A, it_i, it_j, it_m, kap);
                    ++n_int32;
                    active = false;
                }
            }
        }
    }
    // ---- counters ----
    unsigned long long n_fwd = cnt.fwd, n_bwd = cnt.bwd;
    unsigned long long n_eval = n_eval32, n_panel = n_panel32, n_int = n_int32;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_eval += __shfl_xor_sync(0xffffffffu, n_eval, o);
        n_fwd += __shfl_xor_sync(0xffffffffu, n_fwd, o);
        n_bwd += __shfl_xor_sync(0xffffffffu, n_bwd, o);
        n_panel += __shfl_xor_sync(0xffffffffu, n_panel, o);
        n_int += __shfl_xor_sync(0xffffffffu, n_int, o);
        max_top = max(max_top, __shfl_xor_sync(0xffffffffu, max_top, o));
    }
    if (lane == 0) {
        atomicAdd(&stats[0], n_int);
        atomicAdd(&stats[1], n_panel);
        atomicAdd(&stats[2], n_eval);
        atomicAdd(&stats[3], n_fwd);
        atomicAdd(&stats[4], n_bwd);
        atomicMax(&stats[5], (unsigned long long)(max_top + 1));
    }
}

// Node table: one thread per (panel, node) walks from the root panel [0, pi/2] down to its panel
// with the bisection arithmetic of gauss_kronrod_adaptive (mid = (r+l)/2), forms the node
// position like the quadrature does and stores node_const() of it.  The constants depend on omega
// and arc_coeff, so the table is rebuilt at the start of every assembly (7,665 / 15,841 entries).
template <int ORDER>
__global__ void node_table_kernel(NodeConst* __restrict__ table, const RunConst rc) {
    constexpr int H = (ORDER - 1) / 2;
    const GKTables& T = ORDER == 15 ? c_gk15 : c_gk31;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= TRIG_PANELS * (2 * H + 1)) return;
    const int p = e / (2 * H + 1), j = e % (2 * H + 1);
    int depth = 0;
    while ((1 << (depth + 1)) - 1 <= p) ++depth;
    const int idx = p - ((1 << depth) - 1);
    double l = 0.0, r = rc.half_pi;
    for (int b = depth - 1; b >= 0; --b) {
        const double mid = (r + l) / 2;
        if ((idx >> b) & 1) l = mid; else r = mid;
    }
    const double mid = (r + l) / 2, scale = (r - l) / 2;
    const int ni = (j + 1) >> 1;
    const double node = (j & 1) ? T.a[ni] : -T.a[ni];
    table[e] = node_const(rc, __dadd_rn(__dmul_rn(scale, node), mid));
}

size_t assembly_node_table_bytes(int order) { return sizeof(NodeConst) * (size_t)TRIG_PANELS * order; }

static cudaError_t build_node_table(const RunConst& rc, void* table, cudaStream_t stream) {
    const int n = TRIG_PANELS * rc.order;
    if (rc.order == 15)
        node_table_kernel<15><<<(n + 127) / 128, 128, 0, stream>>>((NodeConst*)table, rc);
    else
        node_table_kernel<31><<<(n + 127) / 128, 128, 0, stream>>>((NodeConst*)table, rc);
    return cudaGetLastError();
}

// Diagonal entries (include/solver.h:443 and :465-470).
__global__ void diagonal_kernel(const RunConst rc, const double* __restrict__ bt, const PeerSet A) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rc.N) return;
    const size_t N = rc.N;
    const size_t dim = rc.em ? 2 * N : N;
    store_c(A, (size_t)i * dim + i, mk(rc.diag_es, 0.));
    if (rc.em) {
        store_c(A, (size_t)i * dim + (i + N), mk(0., 0.));
        store_c(A, (size_t)(i + N) * dim + i, mk(0., 0.));
        store_c(A, (size_t)(i + N) * dim + (i + N), mk(rc.diag_em * bt[i], 0.));
    }
}

static int g_blocks_per_sm[2] = {0, 0};

int assembly_grid_blocks(int order, int device) {
    int idx = order == 15 ? 0 : 1;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (g_blocks_per_sm[idx] == 0) {
        int nb = 0;
        if (order == 15)
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, assemble_kernel<15>, BLOCK, 0);
        else
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, assemble_kernel<31>, BLOCK, 0);
        g_blocks_per_sm[idx] = nb > 0 ? nb : 1;
    }
    return sms * g_blocks_per_sm[idx];
}

int assembly_groups_per_block(int order) { (void)order; return BLOCK; }  // stacks are per lane
int assembly_stack_smem() { return STACK_SMEM; }

cudaError_t launch_assembly(const RunConst& rc, const double* eta, const double* g,
                            const double* bi, const PeerSet& A, int shard_index, int shard_count,
                            unsigned long long* counter, void* spill, int spill_cap,
                            unsigned long long* stats, int grid_blocks, cudaStream_t stream,
                            unsigned long long* n_launches, int refill_min, void* table, int d_split) {
    const unsigned long long N = rc.N;
    const unsigned long long n_items = N * (N - 1) / 2 * (rc.em ? 3ULL : 1ULL);
    const unsigned long long sc = shard_count, si = shard_index;
    // shards own chunks of 32 consecutive items, round-robin
    const unsigned long long n_chunks = (n_items + 31) / 32;
    unsigned long long n_local = 0;
    if (n_chunks > si) {
        const unsigned long long my_chunks = (n_chunks - si + sc - 1) / sc;
        n_local = my_chunks * 32;
        const unsigned long long last_chunk = (my_chunks - 1) * sc + si;   // global index
        if (last_chunk == n_chunks - 1) n_local -= n_chunks * 32 - n_items;
    }
    if (d_split < 2 || d_split > rc.N - 1) d_split = 0;
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(stats, 0, 8 * sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    e = build_node_table(rc, table, stream);
    if (e != cudaSuccess) return e;
    if (n_launches) ++*n_launches;
    if (shard_index == 0) {
        diagonal_kernel<<<(rc.N + 127) / 128, 128, 0, stream>>>(rc, bi, A);
        if (n_launches) ++*n_launches;
    }
    if (n_local > 0) {
        if (n_launches) ++*n_launches;
        if (rc.order == 15) {
            assemble_kernel<15><<<grid_blocks, BLOCK, 0, stream>>>(
                rc, eta, g, bi, A, n_local, si, sc, counter, (double2*)spill, spill_cap,
                stats, refill_min, (const NodeConst*)table, d_split);
        } else {
            assemble_kernel<31><<<grid_blocks, BLOCK, 0, stream>>>(
                rc, eta, g, bi, A, n_local, si, sc, counter, (double2*)spill, spill_cap,
                stats, refill_min, (const NodeConst*)table, d_split);
        }
    }
    return cudaGetLastError();
}

cudaError_t assembly_preload() {
    cudaFuncAttributes a;
    const void* ks[] = {(const void*)assemble_kernel<15>, (const void*)assemble_kernel<31>,
                        (const void*)node_table_kernel<15>, (const void*)node_table_kernel<31>,
                        (const void*)diagonal_kernel};
    for (const void* k : ks) {
        cudaError_t e = cudaFuncGetAttributes(&a, k);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace emme
