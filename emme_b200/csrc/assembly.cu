// assembly.cu -- kernel 1: warp-cooperative adaptive Gauss-Kronrod assembly of A(omega).
//
// Replaces EigenSolver::matrixAssembler (reference include/solver.h:417-515) together with
// everything it calls per matrix element: Parameters::kappa_f_tau / kappa_f_tau_e
// (src/Parameters.cpp:113-209), util::integrate + gauss_kronrod_adaptive/basic
// (include/functions.h:181-251,305-331), util::bessel_i_alter_helper (:381-408),
// SingularityHandler (src/singularity_handler.cpp:3-24) and the DedicatedThreadPool fan-out.
//
// Mapping:
//   * a work item is one adaptive quadrature: (pair i<j, mode m); pairs are enumerated
//     diagonal-major (d = j-i ascending) so that neighbouring items cost about the same and
//     the most expensive ones (small d) are issued first;
//   * a GROUP of GS lanes owns one item at a time: GS = 16 for GK15 (two groups per warp),
//     GS = 32 for GK31; lane g evaluates Kronrod node g of the current panel;
//   * the panel sums are accumulated in the reference's order (centre, then +-a_1, +-a_2 ...)
//     by broadcasting each lane's weighted value with warp shuffles, so every lane of the
//     group holds identical K, G and takes identical accept/bisect decisions;
//   * the LIFO interval stack of gauss_kronrod_adaptive lives in shared memory (one slot
//     array per group, leader-owned), spilling to global memory beyond STACK_SMEM entries so
//     that the contract "depth <= integration_iteration_limit" holds for any input;
//   * the grid is persistent (SM count x resident CTAs); groups pull items from a global
//     atomic counter and refill independently, so a group never waits for its warp sibling.
#include <cuda_runtime.h>
#include <stdint.h>

#include "assembly.h"
#include "emme_eval.cuh"
#include "gk_tables.h"

namespace emme {

__constant__ GKTables c_gk15 = EMME_GK15_INIT;
__constant__ GKTables c_gk31 = EMME_GK31_INIT;

constexpr int STACK_SMEM = 24;  // interval-stack entries kept in shared memory per group
constexpr int BLOCK = 128;

// item k (local to this shard) -> global item, pair (i, j) and mode m.
// pairs with diagonal < d: T(d) = (d-1)*(2N-d)/2
__device__ __forceinline__ void decode_item(unsigned long long kg, int N, int nm, int& i, int& j,
                                            int& m) {
    const unsigned long long p = kg / (unsigned)nm;
    m = (int)(kg - p * (unsigned)nm);
    const double tn = 2.0 * N + 1.0;
    double disc = tn * tn - 8.0 * ((double)N + (double)p);
    if (disc < 0.) disc = 0.;
    long long d = (long long)floor((tn - sqrt(disc)) * 0.5);
    if (d < 1) d = 1;
    if (d > N - 1) d = N - 1;
    // fix up rounding of the closed form
    while (d > 1 && (unsigned long long)((d - 1) * (2LL * N - d) / 2) > p) --d;
    while (d < N - 1 && (unsigned long long)(d * (2LL * N - d - 1) / 2) <= p) ++d;
    const unsigned long long base = (unsigned long long)((d - 1) * (2LL * N - d) / 2);
    i = (int)(p - base);
    j = i + (int)d;
}

__device__ __forceinline__ void store_c(double2* A, size_t idx, cplx v) {
    A[idx] = make_double2(v.re, v.im);
}

// Write the entries produced by item (i, j, m) -- include/solver.h:448-453 and :476-504.
__device__ __forceinline__ void scatter(const RunConst& rc, double2* A, int i, int j, int m,
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

template <int ORDER>
__global__ void __launch_bounds__(BLOCK, 4)
assemble_kernel(const RunConst rc, const double* __restrict__ eta, const double* __restrict__ gt,
                const double* __restrict__ bt, double2* __restrict__ A,
                unsigned long long n_items_local, unsigned long long shard_index,
                unsigned long long shard_count, unsigned long long* __restrict__ counter,
                double2* __restrict__ spill, int spill_cap, unsigned long long* __restrict__ stats) {
    constexpr int GS = ORDER == 15 ? 16 : 32;   // lanes per group
    constexpr int H = (ORDER - 1) / 2;          // 7 or 15 symmetric node pairs
    constexpr int GPB = BLOCK / GS;             // groups per block
    const GKTables& T = ORDER == 15 ? c_gk15 : c_gk31;

    __shared__ double2 s_stack[GPB][STACK_SMEM];

    const int lane = threadIdx.x & 31;
    const int gl = lane & (GS - 1);             // lane within group
    const int grp = threadIdx.x / GS;           // group within block
    const unsigned gmask = GS == 32 ? 0xffffffffu : (0xffffu << (lane & 16));
    double2* my_spill =
        spill ? spill + ((size_t)blockIdx.x * GPB + grp) * (size_t)spill_cap : nullptr;

    // node owned by this lane: 0 centre, 1..H -> +a, H+1..2H -> -a; spare lane mirrors centre
    int nidx = gl == 0 ? 0 : (gl <= H ? gl : (gl <= 2 * H ? gl - H : 0));
    const double node = gl <= H ? T.a[nidx] : -T.a[nidx];
    const double kw = T.kw[nidx];
    const double gw = T.gw[nidx];
    const bool counted = gl <= 2 * H;

    const int nm = rc.em ? 3 : 1;
    bool active = false, exhausted = false;
    int it_i = 0, it_j = 0, it_m = 0, top = 0;
    PairConst pc;
    cplx sum = mk(0., 0.);
    double abs_tol = 0.;
    EvalCounters cnt{0u, 0u};
    unsigned long long n_eval = 0, n_panel = 0, n_int = 0, n_fwd = 0, n_bwd = 0;
    int max_top = 0;

    for (;;) {
        if (!active && !exhausted) {
            unsigned long long k = 0;
            if (gl == 0) k = atomicAdd(counter, 1ULL);
            k = __shfl_sync(gmask, k, 0, GS);
            if (k >= n_items_local) {
                exhausted = true;
            } else {
                decode_item(k * shard_count + shard_index, rc.N, nm, it_i, it_j, it_m);
                pc = make_pair(rc, eta[it_i], eta[it_j], gt[it_i], gt[it_j], bt[it_i], bt[it_j]);
                sum = mk(0., 0.);
                abs_tol = 0.;
                if (gl == 0) s_stack[grp][0] = make_double2(0.0, rc.half_pi);
                top = 1;
                active = true;
            }
        }
        if (__all_sync(0xffffffffu, !active)) break;
        if (active) {
            // ---- pop (leader) and broadcast ----
            --top;
            double l = 0., r = 0.;
            if (gl == 0) {
                const double2 e = top < STACK_SMEM ? s_stack[grp][top] : my_spill[top - STACK_SMEM];
                l = e.x;
                r = e.y;
            }
            l = __shfl_sync(gmask, l, 0, GS);
            r = __shfl_sync(gmask, r, 0, GS);
            const double mid = (r + l) / 2;
            const double scale = (r - l) / 2;
            // node position exactly as the reference forms it: scale*x + mid, no FMA
            const double x = __dadd_rn(__dmul_rn(scale, node), mid);
            const cplx fx = eval_node(rc, pc, it_m, x, cnt);
            if (counted) ++n_eval;
            // f(+a_i) + f(-a_i) on lanes 1..H
            cplx f = fx;
            {
                const double pr = __shfl_down_sync(gmask, fx.re, H, GS);
                const double pi = __shfl_down_sync(gmask, fx.im, H, GS);
                if (gl >= 1 && gl <= H) f = mk(fx.re + pr, fx.im + pi);
            }
            const cplx kf = mk(kw * f.re, kw * f.im);
            const cplx gf = mk(gw * f.re, gw * f.im);
            // sequential sums in node order (include/functions.h:189-201)
            cplx K = mk(__shfl_sync(gmask, kf.re, 0, GS), __shfl_sync(gmask, kf.im, 0, GS));
            cplx G = mk(__shfl_sync(gmask, gf.re, 0, GS), __shfl_sync(gmask, gf.im, 0, GS));
#pragma unroll
            for (int n = 1; n <= H; ++n) {
                K.re += __shfl_sync(gmask, kf.re, n, GS);
                K.im += __shfl_sync(gmask, kf.im, n, GS);
                if ((n & 1) == 0) {
                    G.re += __shfl_sync(gmask, gf.re, n, GS);
                    G.im += __shfl_sync(gmask, gf.im, n, GS);
                }
            }
            ++n_panel;
            // ---- accept / bisect (include/functions.h:233-247) ----
            const cplx integral = mk(K.re * scale, K.im * scale);
            const double e0 = fmax(hypot(K.re - G.re, K.im - G.im),
                                   hypot(K.re, K.im) * 2.220446049250313e-16 * 2);
            const double err = e0 * scale;
            const double rel = hypot(rc.tol * integral.re, rc.tol * integral.im);
            if (abs_tol == 0.) abs_tol = rel;
            const bool split = ldexp(scale, rc.maxdepth) > rc.thr_len &&
                               err > abs_tol * rc.inv_scale + rc.prec && err > rel + rc.prec;
            if (split) {
                if (gl == 0) {
                    const double2 e1 = make_double2(mid, r), e2 = make_double2(l, mid);
                    if (top < STACK_SMEM) s_stack[grp][top] = e1; else my_spill[top - STACK_SMEM] = e1;
                    if (top + 1 < STACK_SMEM) s_stack[grp][top + 1] = e2; else my_spill[top + 1 - STACK_SMEM] = e2;
                }
                top += 2;
                max_top = max(max_top, top);
            } else {
                sum = sum + integral;
            }
            if (top == 0) {
                // ---- finalize: kappa = -i*pref*sum (+ electron part), scatter ----
                if (gl == 0) {
                    cplx kap = mk(rc.kappa_pref * sum.im, -rc.kappa_pref * sum.re);
                    if (it_m > 0) kap = kap + kappa_e(rc, it_m, pc.deta, gt[it_i] - gt[it_j]);
                    scatter(rc, A, it_i, it_j, it_m, kap);
                }
                ++n_int;
                active = false;
            }
        }
    }
    // ---- counters ----
    n_fwd = counted ? cnt.fwd : 0;
    n_bwd = counted ? cnt.bwd : 0;
    if (gl != 0) { n_panel = 0; n_int = 0; }
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
        atomicMax(&stats[5], (unsigned long long)max_top);
    }
}

// Diagonal entries (include/solver.h:443 and :465-470).
__global__ void diagonal_kernel(const RunConst rc, const double* __restrict__ bt,
                                double2* __restrict__ A) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rc.N) return;
    const size_t N = rc.N;
    const size_t dim = rc.em ? 2 * N : N;
    A[(size_t)i * dim + i] = make_double2(rc.diag_es, 0.);
    if (rc.em) {
        A[(size_t)i * dim + (i + N)] = make_double2(0., 0.);
        A[(size_t)(i + N) * dim + i] = make_double2(0., 0.);
        A[(size_t)(i + N) * dim + (i + N)] = make_double2(rc.diag_em * bt[i], 0.);
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

int assembly_groups_per_block(int order) { return order == 15 ? BLOCK / 16 : BLOCK / 32; }
int assembly_stack_smem() { return STACK_SMEM; }

cudaError_t launch_assembly(const RunConst& rc, const double* eta, const double* g,
                            const double* bi, void* A, int shard_index, int shard_count,
                            unsigned long long* counter, void* spill, int spill_cap,
                            unsigned long long* stats, int grid_blocks, cudaStream_t stream,
                            unsigned long long* n_launches) {
    const unsigned long long N = rc.N;
    const unsigned long long n_items = N * (N - 1) / 2 * (rc.em ? 3ULL : 1ULL);
    const unsigned long long sc = shard_count, si = shard_index;
    const unsigned long long n_local = n_items > si ? (n_items - si + sc - 1) / sc : 0;
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(stats, 0, 8 * sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    if (shard_index == 0) {
        diagonal_kernel<<<(rc.N + 127) / 128, 128, 0, stream>>>(rc, bi, (double2*)A);
        if (n_launches) ++*n_launches;
    }
    if (n_local > 0) {
        if (n_launches) ++*n_launches;
        if (rc.order == 15) {
            assemble_kernel<15><<<grid_blocks, BLOCK, 0, stream>>>(
                rc, eta, g, bi, (double2*)A, n_local, si, sc, counter, (double2*)spill, spill_cap,
                stats);
        } else {
            assemble_kernel<31><<<grid_blocks, BLOCK, 0, stream>>>(
                rc, eta, g, bi, (double2*)A, n_local, si, sc, counter, (double2*)spill, spill_cap,
                stats);
        }
    }
    return cudaGetLastError();
}

}  // namespace emme
