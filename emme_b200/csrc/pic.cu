// pic.cu -- row N4: the reference's PIC method (`"method": "PIC"`, src/main.cpp:82-137) on the
// GPU: PIC_State<double> + Integrator of include/solver_pic.h behind the emme_pic_* C ABI.
//
// The reference runs, per Runge-Kutta stage, three passes over the markers with a field solve
// in between: put_velocity (include/solver_pic.h:76-135, one cyl_bessel_j(1, .) per marker),
// update (:137-151) and solve_field (:251-354, one cyl_bessel_j(0, .) and one complex
// exponential per marker, 256 thread-pool batches, a serial sum over the batches).
//
// Here a stage is ONE kernel (`pic_stage_kernel`), one marker per thread:
//   gather   phi, dphi from the field (shared-memory copy, linear weights, :93-100)
//   velocity vs = A phi + B dphi (+ the drift term without the pull-back transformation)
//   combine  the stage's linear combination of velocities (Integrator::coef, :466-470: k0 has
//            coefficient 0 after stage 0, so only k1 is ever stored)
//   push     eta <- bound(eta + v_para h/(qR)), weight += v h  -- the reference's operation
//            order, so eta is bit-identical to the reference's
//   deposit  at the new eta: J0, J1 from ONE Miller backward recurrence (`bessel_j01`), the
//            pull-back phase exp(-i omega_d_integral omega_dv) from one sincos, density
//            j0 w dc_pb into per-CTA shared-memory cells (shared atomics); every CTA then
//            stores its cells as one partial density (plain coalesced stores)
//   next     the marker's velocity coefficients A, B for the NEXT stage are formed now, while
//            j0, dj0, dc_pb, omega_d(eta) are in registers: what crosses the field solve per
//            marker is (eta, w, A, B) = 56 bytes instead of the reference's j0/dc_pb extras
//            plus a second Bessel evaluation
//   field    `pic_field_kernel` sums the per-CTA partials in a fixed order and turns the density
//            into the new field (quasi-neutrality table, :349-351); after stage 2 it appends the
//            field to the on-device history that the diagnostics read.
// Six launches per Integrator::step, captured once per dt into a CUDA graph and replayed.
// (The first version flushed the cells with red.global.add.f64 and let the last CTA -- a ticket --
// form the field in the same launch.  At equal block size it measured within 1 us per stage of this
// one; the partial-sum form is kept because its reduction order is fixed.  What did matter is the
// CTA size: one 1024-thread CTA per SM instead of four 256-thread ones, 49 -> 39 us per stage at
// 1M markers.  See DESIGN.md section 4c.)
//
// Differences to the reference that are visible in the numbers: deposits are summed in a
// different (and run-to-run varying) order, Bessel J comes from the Miller recurrence (abs. error
// < 1e-15; libstdc++'s cyl_bessel_j is within 7e-15 of it) and the velocity is evaluated in
// factored form; fields agree with the reference to 4e-15 of the largest field value, positions bit
// for bit (tests/test_pic_gpu.py).
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <numeric>
#include <random>
#include <string>
#include <vector>

#include "../../include/emme_b200.h"
#include "../host/json.hpp"
#include "../host/parameters.hpp"
#include "common.h"
#include "peer.h"
#include "pic_eval.cuh"

namespace {

namespace cg = cooperative_groups;
using emme::d2;
using emme::mk2;
using emme::PicConst;

// Multi-GPU density exchange without a collective library: every rank owns an exchange buffer
// [2 parities][P ranks][nf] of partial densities followed by [P ranks][field-kernel blocks] ready
// flags (one cudaMalloc block, mapped by every peer).  See pic_field_kernel, mode 3.
struct PicPeers {
    d2* xch[EMME_MAX_PEER_RANKS];                   // exchange buffer of rank r (own entry = local pointer)
    unsigned long long* flags[EMME_MAX_PEER_RANKS]; // its flag area
    int n, me;
    unsigned long long timeout_ns;
    unsigned long long* err;                         // local: a wait gave up
};

struct PicDev {
    PicPeers peers;
    PicConst k;
    long n;
    int nf;
    int use_smem;
    double* eta;
    const double *vpar, *vperp, *pw;
    d2 *w, *A, *B, *k1;
    double* c;  // omega_d(eta) * omega_dv, only without the pull-back transformation
    d2 *field, *dens;   // dens: summed density (the buffer a multi-GPU exchange reduces)
    d2* part;           // per-CTA partial densities [nparts][nf] (shared-memory cells only)
    int nparts;
    const double* coef;
    d2* hist;
    unsigned long long* step;   // Integrator::step calls started
};

__device__ __forceinline__ void deposit(const PicDev& d, d2* cells, double eta, d2 den) {
    int idx;
    double wt;
    emme::pic_locate(d.k, eta, idx, wt);
    const int i1 = (idx + 1 == d.nf) ? 0 : idx + 1;
    const double w0 = 1.0 - wt;
    atomicAdd(&cells[idx].x, den.x * w0);
    atomicAdd(&cells[idx].y, den.y * w0);
    atomicAdd(&cells[i1].x, den.x * wt);
    atomicAdd(&cells[i1].y, den.y * wt);
}

// Density -> field (include/solver_pic.h:339-351).  blockDim = (32 cells, 32 chunks): chunk c sums
// the partials c, c+32, ... of its cell (all loads of a thread independent: one L2 round trip),
// then the chunks are added in index order, so the result is deterministic for given partials.
// mode 0: field = coef * sum(partials); mode 1 (multi-GPU, before the exchange):
// dens = sum(partials); mode 2 (after the exchange): field = coef * dens.  Without shared-memory
// cells the stage kernel has already accumulated into dens, which is cleared once it is consumed.
// mode 3 (multi-GPU, fused exchange): the block sums this rank's partials of its 32 cells, stores
// them into slot `me` of EVERY rank's exchange buffer over NVLink, announces them with one release
// store per peer (flag [me][block] = stage serial), waits for the same block of every other rank
// (acquire loads of local memory) and adds the P slots in rank order -- every rank forms the same
// field, bit for bit, with no collective call and only block-local synchronisation.  The serial is
// 3 * (steps started) + stage, derived from the on-device step counter, so the launch is the same
// every step (CUDA graph); the two parities of the buffer keep a fast rank's next stage out of the
// slots a slow rank is still reading.
#define PIC_FIELD_CHUNKS 32
__global__ void __launch_bounds__(32 * PIC_FIELD_CHUNKS) pic_field_kernel(PicDev d, int mode, int record, int stage) {
    __shared__ d2 red[PIC_FIELD_CHUNKS][32];
    const int cell = blockIdx.x * 32 + threadIdx.x, ch = threadIdx.y;
    d2 acc = mk2(0.0, 0.0);
    const bool summed = (mode == 2) || !d.use_smem;
    cudaGridDependencySynchronize();             // the stage kernel's partials are complete and visible
    cudaTriggerProgrammaticLaunchCompletion();   // the next stage kernel may start its prologue
    if (cell < d.nf) {
        if (summed) {
            if (ch == 0) acc = d.dens[cell];
        } else {
#pragma unroll 5
            for (int g = ch; g < d.nparts; g += PIC_FIELD_CHUNKS) {
                const d2 v = d.part[(size_t)g * d.nf + cell];
                acc.x += v.x;
                acc.y += v.y;
            }
        }
    }
    red[ch][threadIdx.x] = acc;
    __syncthreads();
    if (mode == 3) {
        const PicPeers& pp = d.peers;
        const unsigned long long seq = 3ull * (*d.step) + (unsigned long long)stage;   // *d.step >= 1 here
        const size_t slot = (size_t)(seq & 1ull) * pp.n;
        const bool mine = ch == 0 && cell < d.nf;
        if (mine) {
            for (int c = 1; c < PIC_FIELD_CHUNKS; ++c) {
                acc.x += red[c][threadIdx.x].x;
                acc.y += red[c][threadIdx.x].y;
            }
            if (summed) d.dens[cell] = mk2(0.0, 0.0);
            for (int r = 0; r < pp.n; ++r) pp.xch[r][(slot + pp.me) * d.nf + cell] = acc;
        }
        __syncthreads();
        if (ch == 0 && threadIdx.x < pp.n) {
            // one thread per peer: publish my block, then wait for that peer's
            const int r = threadIdx.x;
            __threadfence_system();
            emme::st_release_sys(pp.flags[r] + (size_t)pp.me * gridDim.x + blockIdx.x, seq);
            if (!emme::spin_until(pp.flags[pp.me] + (size_t)r * gridDim.x + blockIdx.x, seq, pp.timeout_ns))
                atomicMax(pp.err, seq);
        }
        __syncthreads();
        if (!mine) return;
        acc = mk2(0.0, 0.0);
        for (int r = 0; r < pp.n; ++r) {          // rank order: the same sum on every rank
            const double2 v = __ldcg(reinterpret_cast<const double2*>(&pp.xch[pp.me][(slot + r) * d.nf + cell]));
            acc.x += v.x;
            acc.y += v.y;
        }
        const double cf = d.coef[cell];
        const d2 f = mk2(acc.x * cf, acc.y * cf);
        d.field[cell] = f;
        if (record) d.hist[(*d.step - 1) * (unsigned long long)d.nf + cell] = f;
        return;
    }
    if (ch != 0 || cell >= d.nf) return;
    for (int c = 1; c < PIC_FIELD_CHUNKS; ++c) {
        acc.x += red[c][threadIdx.x].x;
        acc.y += red[c][threadIdx.x].y;
    }
    if (mode == 1) {
        d.dens[cell] = acc;
        return;
    }
    const double cf = d.coef[cell];
    const d2 f = mk2(acc.x * cf, acc.y * cf);
    d.field[cell] = f;
    if (summed) d.dens[cell] = mk2(0.0, 0.0);
    if (record) d.hist[(*d.step - 1) * (unsigned long long)d.nf + cell] = f;
}

// First deposit-independent state: A = B = 0 (the reference's marker extras start with
// j0 = dc_pb = 0, include/solver_pic.h:207-227, and the field with 0), c from the loaded eta.
template <bool SWITCH>
__global__ void pic_init_kernel(PicDev d) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < d.n; i += (long)gridDim.x * blockDim.x) {
        d.A[i] = mk2(0.0, 0.0);
        d.B[i] = mk2(0.0, 0.0);
        d.k1[i] = mk2(0.0, 0.0);
        if (!SWITCH) {
            d.c[i] = emme::pic_initial_c(d.k, d.eta[i], d.vpar[i], d.vperp[i]);
        }
    }
}

// register budget of the stage kernel: (1024, 1) allows 64 registers and any block size;
// tuning builds pass -DPIC_LB_THREADS=256 -DPIC_LB_BLOCKS=5|6 (block size is then fixed at 256)
#ifndef PIC_LB_THREADS
#define PIC_LB_THREADS 1024
#define PIC_LB_BLOCKS 1
#endif

// L1 prefetch of the marker a thread handles in its NEXT trip: costs no registers, and the loads at
// the top of the next trip (the kernel's largest stall in the first profile) find their lines
#ifndef PIC_PREFETCH
#define PIC_PREFETCH 1
#endif
__device__ __forceinline__ void prefetch_l1(const void* p) {
#if PIC_PREFETCH
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#endif
}

// The marker loop of one Runge-Kutta stage (grid-stride, one marker per thread and trip).
template <bool SWITCH>
__device__ __forceinline__ void stage_markers(const PicDev& d, const d2* fld, d2* cells, int stage, double h,
                                              double c1, double c2) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < d.n; i += stride) {
        double eta = d.eta[i];
        d2 w = d.w[i];
        const d2 A = d.A[i], B = d.B[i];
        const double vpar = d.vpar[i], vperp = d.vperp[i], pw = d.pw[i];
        if (i + stride < d.n) {
            const long nx = i + stride;
            prefetch_l1(&d.eta[nx]);
            prefetch_l1(&d.w[nx]);
            prefetch_l1(&d.A[nx]);
            prefetch_l1(&d.B[nx]);
            prefetch_l1(&d.vpar[nx]);
            prefetch_l1(&d.vperp[nx]);
            prefetch_l1(&d.pw[nx]);
            if (stage == 2) prefetch_l1(&d.k1[nx]);
        }
        // gather + velocity of this stage (include/solver_pic.h:91-121)
        d2 vs = emme::pic_velocity(d.k, fld, eta, A, B);
        if (!SWITCH) {  // -weight omega_d omega_dv i (include/solver_pic.h:112-114)
            const double c = d.c[i];
            vs.x += c * w.y;
            vs.y -= c * w.x;
        }
        d2 v = vs;
        if (stage == 1) {
            d.k1[i] = vs;
        } else if (stage == 2) {
            const d2 k1 = d.k1[i];
            v = mk2(c1 * k1.x + c2 * vs.x, c1 * k1.y + c2 * vs.y);
        }
        // update (include/solver_pic.h:141-145), the reference's operation order for eta
        eta = emme::pic_push(d.k, eta, vpar, h);
        w.x = fma(v.x, h, w.x);
        w.y = fma(v.y, h, w.y);
        // deposit at the new position + coefficients of the next stage
        d2 den, An, Bn;
        double cn;
        emme::pic_marker_at<SWITCH>(d.k, eta, vpar, vperp, pw, w, den, An, Bn, cn);
        deposit(d, cells, eta, den);
        d.eta[i] = eta;
        d.w[i] = w;
        d.A[i] = An;
        d.B[i] = Bn;
        if (!SWITCH) d.c[i] = cn;
    }
}

// One Runge-Kutta stage for every marker.  h = coef[stage][stage+1] * dt; the stage's velocity
// combination is v = k0 (stage 0), k1 (stage 1), c1 k1 + c2 k2 (stage 2).
template <bool SWITCH, bool SMEM>
__global__ void __launch_bounds__(PIC_LB_THREADS, PIC_LB_BLOCKS) pic_stage_kernel(PicDev d, int stage, double h, double c1,
                                                        double c2) {
    extern __shared__ d2 smem[];
    d2* s_field = smem;
    d2* s_dens = smem + d.nf;
    // Programmatic dependent launch: this grid may start while pic_field_kernel of the previous
    // stage is still running; everything before the dependency sync touches shared memory only.
    if (SMEM)
        for (int i = threadIdx.x; i < d.nf; i += blockDim.x) s_dens[i] = mk2(0.0, 0.0);
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();   // lets pic_field_kernel get resident early; it waits itself
    if (stage == 0 && blockIdx.x == 0 && threadIdx.x == 0) *d.step += 1;
    if (SMEM) {
        for (int i = threadIdx.x; i < d.nf; i += blockDim.x) s_field[i] = d.field[i];
        __syncthreads();
    }
    const d2* fld = SMEM ? s_field : d.field;
    d2* cells = SMEM ? s_dens : d.dens;
    const int nf = d.nf;
    stage_markers<SWITCH>(d, fld, cells, stage, h, c1, c2);
    if (SMEM) {
        __syncthreads();
        d2* mine = d.part + (size_t)blockIdx.x * nf;
        for (int i = threadIdx.x; i < nf; i += blockDim.x) mine[i] = s_dens[i];
    }
}

// nsteps x Integrator::step in ONE cooperative launch (shared-memory cells only): the grid-wide
// barrier replaces the kernel boundaries.  Per stage: marker loop -> partial store -> grid.sync ->
// every CTA reduces its slice of cells over all partials (32 chunks per cell, fixed order), applies
// the quasi-neutrality table, writes the field (and the history after stage 2) -> grid.sync.
template <bool SWITCH>
__global__ void __launch_bounds__(PIC_LB_THREADS, PIC_LB_BLOCKS) pic_persistent_kernel(PicDev d, double dt, int nsteps,
                                                                        unsigned long long first_slot) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ d2 smem[];
    d2* s_field = smem;
    d2* s_dens = smem + d.nf;
    __shared__ d2 red[PIC_FIELD_CHUNKS][32];
    const int nf = d.nf;
    const double RK[4][4] = EMME_PIC_RK_COEF;   // only constant indices below: folded at compile time
    const int cpc = (nf + gridDim.x - 1) / gridDim.x;             // cells per CTA in the reduce
    const int c_begin = blockIdx.x * cpc, c_end = min(c_begin + cpc, nf);
    const int lane_cell = threadIdx.x & 31, ch = threadIdx.x >> 5;   // 32 cells x up to 32 chunks
    const int nch = min((int)(blockDim.x >> 5), PIC_FIELD_CHUNKS);
    for (int step = 0; step < nsteps; ++step) {
        for (int stage = 0; stage < 3; ++stage) {
            for (int i = threadIdx.x; i < nf; i += blockDim.x) {
                const double2 f = __ldcg(reinterpret_cast<const double2*>(&d.field[i]));
                s_field[i] = mk2(f.x, f.y);
                s_dens[i] = mk2(0.0, 0.0);
            }
            __syncthreads();
            const double hfac = stage == 0 ? RK[0][1] : (stage == 1 ? RK[1][2] : RK[2][3]);
            stage_markers<SWITCH>(d, s_field, s_dens, stage, hfac * dt, RK[2][1], RK[2][2]);
            __syncthreads();
            d2* mine = d.part + (size_t)blockIdx.x * nf;
            for (int i = threadIdx.x; i < nf; i += blockDim.x) mine[i] = s_dens[i];
            grid.sync();
            for (int c0 = c_begin; c0 < c_end; c0 += 32) {
                const int cell = c0 + lane_cell;
                d2 acc = mk2(0.0, 0.0);
                if (cell < c_end && ch < nch) {
                    for (int g = ch; g < (int)gridDim.x; g += nch) {
                        const double2 v = __ldcg(reinterpret_cast<const double2*>(&d.part[(size_t)g * nf + cell]));
                        acc.x += v.x;
                        acc.y += v.y;
                    }
                }
                if (ch < PIC_FIELD_CHUNKS) red[ch][lane_cell] = acc;
                __syncthreads();
                if (ch == 0 && cell < c_end) {
                    for (int c = 1; c < nch; ++c) {
                        acc.x += red[c][lane_cell].x;
                        acc.y += red[c][lane_cell].y;
                    }
                    const double cf = d.coef[cell];
                    const d2 f = mk2(acc.x * cf, acc.y * cf);
                    d.field[cell] = f;
                    if (stage == 2) d.hist[(first_slot + step) * (unsigned long long)nf + cell] = f;
                }
                __syncthreads();
            }
            grid.sync();
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *d.step = first_slot + nsteps;
}

const double RK_COEF[4][4] = EMME_PIC_RK_COEF;

}  // namespace

struct emme_pic {
    int device = 0, sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    emme_pic_params p{};
    long n = 0, n_total = 0, first = 0;
    PicDev d{};
    std::vector<double> h_coef;  // host copy for emme_pic_extras
    std::vector<long> perm;  // device slot j holds marker perm[j] of this shard (sorted by v_perp); lazily downloaded
    long* d_perm = nullptr;
    int block = 256;
    void* d_vpar = nullptr;
    void* d_vperp = nullptr;
    void* d_pw = nullptr;
    void* d_coef = nullptr;
    long hist_cap = 0;
    long steps_done = 0;
    int shard_count = 1, shard_index = 0;
    // fused peer exchange (pic_field_kernel mode 3): one cudaMalloc block = exchange slots + flags
    void* xch = nullptr;
    size_t xch_flag_offset = 0;
    void* peer_xch[EMME_MAX_PEER_RANKS] = {};
    bool peer_ipc[EMME_MAX_PEER_RANKS] = {};
    int peer_count = 0;       // > 1 once every rank's buffer is mapped: emme_pic_step then works sharded
    int next_call = 0;        // sharded protocol: 2*stage = begin(stage) expected, 2*stage+1 = finish(stage)
    int grid = 0;
    size_t smem = 0;
    cudaGraphExec_t graph = nullptr;
    double graph_dt = 0;
    int use_graph = 1;
    int use_pdl = 1;          // programmatic dependent launch between the kernels of a step
    // one cooperative launch per emme_pic_step call (grid.sync instead of kernel boundaries):
    // measured slower than the graph (40.2 vs 38.9 us/stage at 1M markers, 21.3 vs 10.2 at 64K), kept
    // as a tested alternative behind EMME_PIC_PERSISTENT=1
    int use_persistent = 0;
    int pgrid = 0;
    double last_ms = 0;
    unsigned long long launches = 0;
};

namespace {

using emme::capi_fail;

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return capi_fail(EMME_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));    \
    } while (0)

// Device memory comes from the device's stream-ordered pool (cudaMallocAsync) with a release
// threshold of 8 GiB: a scan creates and destroys one PIC state per scan point, and plain
// cudaMalloc/cudaFree of the ~15 buffers cost 0.15-0.6 s per state (measured, EMME_PIC_TIMING=1)
// against 24 ms for the 180 steps themselves.
template <typename T>
cudaError_t dev_alloc(T** p, size_t count, cudaStream_t stream) {
    return cudaMallocAsync(reinterpret_cast<void**>(p), sizeof(T) * (count ? count : 1), stream);
}

cudaError_t configure_pool(int device) {
    cudaMemPool_t pool = nullptr;
    cudaError_t e = cudaDeviceGetDefaultMemPool(&pool, device);
    if (e != cudaSuccess) return e;
    unsigned long long cur = 0;
    e = cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &cur);
    if (e != cudaSuccess) return e;
    unsigned long long want = 8ULL << 30;
    if (cur >= want) return cudaSuccess;
    return cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &want);
}

int ensure_history(emme_pic* s, long steps_total) {
    if (steps_total <= s->hist_cap) return 0;
    long cap = s->hist_cap ? s->hist_cap : 256;
    while (cap < steps_total) cap *= 2;
    d2* nh = nullptr;
    CU(dev_alloc(&nh, (size_t)cap * s->d.nf, s->stream));
    if (s->d.hist && s->steps_done > 0) {
        CU(cudaMemcpyAsync(nh, s->d.hist, sizeof(d2) * (size_t)s->steps_done * s->d.nf,
                           cudaMemcpyDeviceToDevice, s->stream));
        CU(cudaStreamSynchronize(s->stream));
    }
    if (s->d.hist) CU(cudaFreeAsync(s->d.hist, s->stream));
    s->d.hist = nh;
    s->hist_cap = cap;
    if (s->graph) {  // the captured launches hold the old pointer
        cudaGraphExecDestroy(s->graph);
        s->graph = nullptr;
    }
    return 0;
}

// Launch with the programmatic-stream-serialisation attribute (s->use_pdl): the kernel may become
// resident before its predecessor in the stream has finished; its cudaGridDependencySynchronize()
// is the real dependency.  Captured into the step graph as a programmatic edge.
template <typename... Params, typename... Args>
cudaError_t launch_pdl(emme_pic* s, void (*kernel)(Params...), dim3 grid, dim3 block, size_t smem, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = s->use_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    s->launches++;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<Params>(args)...);
}

cudaError_t launch_stage(emme_pic* s, double dt, int stage) {
    const double h = RK_COEF[stage][stage + 1] * dt;
    const double c1 = RK_COEF[2][1], c2 = RK_COEF[2][2];
    const bool sw = s->p.drift_center_transformation_switch != 0;
    dim3 grid(s->grid), block(s->block);
    if (s->d.use_smem) {
        if (sw) return launch_pdl(s, pic_stage_kernel<true, true>, grid, block, s->smem, s->d, stage, h, c1, c2);
        return launch_pdl(s, pic_stage_kernel<false, true>, grid, block, s->smem, s->d, stage, h, c1, c2);
    }
    if (sw) return launch_pdl(s, pic_stage_kernel<true, false>, grid, block, 0, s->d, stage, h, c1, c2);
    return launch_pdl(s, pic_stage_kernel<false, false>, grid, block, 0, s->d, stage, h, c1, c2);
}

cudaError_t launch_field(emme_pic* s, int mode, int record, int stage = 0) {
    return launch_pdl(s, pic_field_kernel, dim3((s->d.nf + 31) / 32), dim3(32, PIC_FIELD_CHUNKS), 0, s->d, mode, record,
                      stage);
}

}  // namespace

extern "C" {

int emme_pic_load_markers(const emme_pic_params* p, long n, long long seed, double* eta,
                          double* v_para, double* v_perp, double* weight) {
    if (!p) return capi_fail(-1, "null params");
    if (n < 0) return capi_fail(-2, "negative marker count");
    if (!eta || !v_para || !v_perp || !weight) return capi_fail(-4, "null output array");
    // initialize_marker (include/solver_pic.h:186-205): same engine, same distributions, same
    // order of draws per marker
    std::mt19937 gen(seed < 0 ? std::random_device{}() : (unsigned)seed);
    std::uniform_real_distribution<double> uniform_eta(-p->length, p->length);
    std::uniform_real_distribution<double> uniform_w(0, 0.001);
    std::normal_distribution<double> normal_vpara(0, p->vt / std::sqrt(p->water_bag_weight_vpara));
    std::normal_distribution<double> normal_vperp(0, p->vt / std::sqrt(p->water_bag_weight_vperp));
    for (long i = 0; i < n; ++i) {
        eta[i] = uniform_eta(gen);
        v_para[i] = normal_vpara(gen);
        v_perp[i] = std::abs(normal_vperp(gen));
        weight[2 * i] = uniform_w(gen);
        weight[2 * i + 1] = 0.0;
    }
    return 0;
}

}  // extern "C"

namespace {

// ---- device order: markers sorted by v_perp (counting sort on PIC_SORT_BUCKETS buckets) ----
// The Miller recurrence of bessel_j01 runs ~ x + 12.6 x^(1/3) trips with x = (v_perp/vt) sb(eta);
// lanes of a warp that share v_perp differ only through sb(eta), which halves the spread of trip
// counts inside a warp.  The order of markers is otherwise free (deposits commute);
// emme_pic_markers undoes it.  Round 1 sorted and permuted on the host (one core, 68 ms per million
// markers with the staging copies); here the raw arrays are uploaded as they are and the device
// does histogram -> scan -> scatter (and forms p_weight = pw * 2L/sum on the way).
constexpr int PIC_SORT_BUCKETS = 4096;

__device__ __forceinline__ int pic_bucket(double v, double scale) {
    const double b = v * scale;
    return b > 0 ? (b < PIC_SORT_BUCKETS - 1 ? (int)b : PIC_SORT_BUCKETS - 1) : 0;
}

__global__ void pic_sort_count_kernel(const double* __restrict__ vperp, long n, double scale,
                                      unsigned* __restrict__ hist, unsigned* __restrict__ pos) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        pos[i] = atomicAdd(&hist[pic_bucket(vperp[i], scale)], 1u);
}

// exclusive scan of the PIC_SORT_BUCKETS counters, one block of 1024 threads (4 per thread)
__global__ void __launch_bounds__(1024) pic_sort_scan_kernel(unsigned* __restrict__ hist) {
    __shared__ unsigned sh[1024];
    const int t = threadIdx.x;
    unsigned v[PIC_SORT_BUCKETS / 1024], sum = 0;
#pragma unroll
    for (int k = 0; k < PIC_SORT_BUCKETS / 1024; ++k) {
        v[k] = hist[t * (PIC_SORT_BUCKETS / 1024) + k];
        sum += v[k];
    }
    sh[t] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const unsigned add = t >= o ? sh[t - o] : 0u;
        __syncthreads();
        sh[t] += add;
        __syncthreads();
    }
    unsigned run = sh[t] - sum;
#pragma unroll
    for (int k = 0; k < PIC_SORT_BUCKETS / 1024; ++k) {
        hist[t * (PIC_SORT_BUCKETS / 1024) + k] = run;
        run += v[k];
    }
}

// raw (caller's order) -> device order; pw_raw == nullptr: unit water-bag weights, pw = v_perp
__global__ void pic_sort_scatter_kernel(long n, double scale, int identity, const unsigned* __restrict__ start,
                                        const unsigned* __restrict__ pos, const double* __restrict__ eta_raw,
                                        const double* __restrict__ vpar_raw, const double* __restrict__ vperp_raw,
                                        const double* __restrict__ pw_raw, const d2* __restrict__ w_raw, double inn,
                                        double* __restrict__ eta, double* __restrict__ vpar, double* __restrict__ vperp,
                                        double* __restrict__ pw, d2* __restrict__ w, long* __restrict__ perm) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const double vq = vperp_raw[i];
        const long j = identity ? i : (long)start[pic_bucket(vq, scale)] + pos[i];
        eta[j] = eta_raw[i];
        vpar[j] = vpar_raw[i];
        vperp[j] = vq;
        pw[j] = (pw_raw ? pw_raw[i] : vq) * inn;      // include/solver_pic.h:229-235
        w[j] = w_raw[i];
        perm[j] = i;
    }
}

cudaError_t pic_preload();

// un-normalised p_weight of one marker (include/solver_pic.h:229-232)
inline double pweight_raw(const emme_pic_params* p, double vp, double vq, double wa, double wb) {
    return vq * std::exp(-(vp * vp * wa + vq * vq * wb) / (2 * p->vt * p->vt));
}

// The PIC_State constructor for the contiguous block [first, first + n) of n_total markers; the
// arrays point at the BLOCK.  pw_sum is the sum of the un-normalised p_weight over ALL markers.
int create_impl(const emme_pic_params* p, long n_total, long first, long n, const double* eta, const double* v_para,
                const double* v_perp, const double* weight, double pw_sum, int shard_index, int shard_count,
                int device, emme_pic** out) {
    if (p->npoints < 4) return capi_fail(-1, "npoints must be at least 4");
    if (emme_device_count() <= 0)
        return capi_fail(EMME_E_NO_DEVICE, "no CUDA device: emme_b200 has no CPU fallback");
    CU(cudaSetDevice(device));
    // released by the guard on every early return below
    struct Guard {
        emme_pic* h;
        ~Guard() { if (h) emme_pic_destroy(h); }
    } guard{new emme_pic()};
    emme_pic* s = guard.h;
    s->device = device;
    s->p = *p;
    s->n_total = n_total;
    s->shard_count = shard_count;
    s->shard_index = shard_index;
    s->first = first;
    s->n = n;
    const int nf = p->npoints;
    CU(cudaDeviceGetAttribute(&s->sms, cudaDevAttrMultiProcessorCount, device));
    CU(configure_pool(device));
    CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&s->ev0));
    CU(cudaEventCreate(&s->ev1));

    // EMME_PIC_TIMING=1 prints the host phases of this call to stderr
    const bool timing = std::getenv("EMME_PIC_TIMING") != nullptr;
    auto tnow = [] { return std::chrono::steady_clock::now(); };
    auto tprev = tnow();
    auto lap = [&](const char* what) {
        if (!timing) return;
        const auto t = tnow();
        std::fprintf(stderr, "[emme_pic_create] %-28s %8.3f ms\n", what,
                     std::chrono::duration<double, std::milli>(t - tprev).count());
        tprev = t;
    };
    lap("stream + events");
    const double wa = 1 - p->water_bag_weight_vpara, wb = 1 - p->water_bag_weight_vperp;
    const bool unit = wa == 0 && wb == 0;
    const double inn = 2 * p->length / pw_sum;
    // cal_quasi_neutrality_coef (include/solver_pic.h:381-399)
    const double cell_width = 2 * p->length / p->npoints;
    s->h_coef.resize(nf);
    for (int idx = 0; idx < nf; ++idx) {
        const double b = p->b_theta * (1. + std::pow(p->shat * (idx * cell_width - p->length), 2));
        const double g0 = std::cyl_bessel_i(0, b) * std::exp(-b);
        s->h_coef[idx] = 1. / ((1. + 1. / p->tau - g0) * cell_width);
    }
    double vmax = 0;
    for (long i = 0; i < n; ++i) vmax = std::max(vmax, v_perp[i]);
    std::vector<double> pw_host;
    if (!unit) {   // glibc's exp, like the reference: formed on the host, normalised on the device
        pw_host.resize(n);
        for (long i = 0; i < n; ++i) pw_host[i] = pweight_raw(p, v_para[i], v_perp[i], wa, wb);
    }
    lap("table + v_perp range (host)");
    PicDev& d = s->d;
    d.n = n;
    d.nf = nf;
    PicConst& k = d.k;
    k.nf = nf;
    k.L = p->length;
    k.cw = cell_width;
    k.inv_2cw = 1.0 / (2. * cell_width);
    k.qR = p->q * p->R;
    k.inv_qR = 1.0 / k.qR;
    k.inv_vt = 1.0 / p->vt;
    k.shat = p->shat;
    k.b_theta = p->b_theta;
    k.omega_d_bar = p->omega_d_bar;
    k.omega_s_i = p->omega_s_i;
    k.eta_i = p->eta_i;
    k.inv_2vt2 = 1.0 / (2. * p->vt * p->vt);
    double *dvpar, *dvperp, *dpw, *dcoef;
    CU(dev_alloc(&d.eta, n, s->stream));
    CU(dev_alloc(&dvpar, n, s->stream));
    CU(dev_alloc(&dvperp, n, s->stream));
    CU(dev_alloc(&dpw, n, s->stream));
    CU(dev_alloc(&d.w, n, s->stream));
    CU(dev_alloc(&d.A, n, s->stream));
    CU(dev_alloc(&d.B, n, s->stream));
    CU(dev_alloc(&d.k1, n, s->stream));
    CU(dev_alloc(&d.c, p->drift_center_transformation_switch ? 1 : n, s->stream));
    CU(dev_alloc(&d.field, nf, s->stream));
    CU(dev_alloc(&d.dens, nf, s->stream));
    CU(dev_alloc(&dcoef, nf, s->stream));
    CU(dev_alloc(&d.step, 1, s->stream));
    CU(dev_alloc(&s->d_perm, n, s->stream));
    d.vpar = dvpar; d.vperp = dvperp; d.pw = dpw; d.coef = dcoef;
    s->d_vpar = dvpar; s->d_vperp = dvperp; s->d_pw = dpw; s->d_coef = dcoef;
    lap("cudaMalloc");
    {
        // raw arrays in the caller's order (released to the pool once the scatter has run)
        double *r_eta, *r_vpar, *r_vperp, *r_pw = nullptr;
        d2* r_w;
        unsigned *hist, *pos;
        CU(dev_alloc(&r_eta, n, s->stream));
        CU(dev_alloc(&r_vpar, n, s->stream));
        CU(dev_alloc(&r_vperp, n, s->stream));
        CU(dev_alloc(&r_w, n, s->stream));
        CU(dev_alloc(&hist, PIC_SORT_BUCKETS, s->stream));
        CU(dev_alloc(&pos, n, s->stream));
        if (!unit) CU(dev_alloc(&r_pw, n, s->stream));
        CU(cudaMemsetAsync(hist, 0, sizeof(unsigned) * PIC_SORT_BUCKETS, s->stream));
        CU(cudaMemcpyAsync(r_vperp, v_perp, sizeof(double) * n, cudaMemcpyHostToDevice, s->stream));
        const bool identity = std::getenv("EMME_PIC_NOSORT") != nullptr;
        const double scale = vmax > 0 ? (PIC_SORT_BUCKETS - 1) / vmax : 0.0;
        const int sg = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
        pic_sort_count_kernel<<<sg, 256, 0, s->stream>>>(r_vperp, n, scale, hist, pos);   // overlaps the other uploads
        pic_sort_scan_kernel<<<1, 1024, 0, s->stream>>>(hist);
        CU(cudaMemcpyAsync(r_eta, eta, sizeof(double) * n, cudaMemcpyHostToDevice, s->stream));
        CU(cudaMemcpyAsync(r_vpar, v_para, sizeof(double) * n, cudaMemcpyHostToDevice, s->stream));
        CU(cudaMemcpyAsync(r_w, weight, sizeof(d2) * n, cudaMemcpyHostToDevice, s->stream));
        if (!unit) CU(cudaMemcpyAsync(r_pw, pw_host.data(), sizeof(double) * n, cudaMemcpyHostToDevice, s->stream));
        pic_sort_scatter_kernel<<<sg, 256, 0, s->stream>>>(n, scale, identity ? 1 : 0, hist, pos, r_eta, r_vpar, r_vperp,
                                                          r_pw, r_w, inn, d.eta, dvpar, dvperp, dpw, d.w, s->d_perm);
        s->launches += 3;
        CU(cudaGetLastError());
        for (void* b : {(void*)r_eta, (void*)r_vpar, (void*)r_vperp, (void*)r_w, (void*)hist, (void*)pos, (void*)r_pw})
            if (b) CU(cudaFreeAsync(b, s->stream));
        CU(cudaStreamSynchronize(s->stream));   // pw_host and the caller's arrays may go away
        lap("upload + device sort");
    }
    CU(cudaMemcpyAsync(dcoef, s->h_coef.data(), sizeof(double) * nf, cudaMemcpyHostToDevice, s->stream));
    CU(cudaMemsetAsync(d.field, 0, sizeof(d2) * nf, s->stream));
    CU(cudaMemsetAsync(d.dens, 0, sizeof(d2) * nf, s->stream));
    CU(cudaMemsetAsync(d.step, 0, sizeof(unsigned long long), s->stream));

    // launch geometry: field + density cells in shared memory when they fit
    s->smem = sizeof(d2) * 2 * (size_t)nf;
    d.use_smem = s->smem <= 200 * 1024;
    const bool sw = p->drift_center_transformation_switch != 0;
    // One 1024-thread CTA per SM once there are enough markers to fill them: the per-CTA costs
    // (field copy, cell clearing, partial store, one more partial for pic_field_kernel to sum)
    // are paid 148 instead of 592 times, and meshes whose cells take > 56 KB keep 32 warps per SM
    s->block = (PIC_LB_THREADS >= 1024 && (n >= (long)s->sms * 1024 || s->smem > 56 * 1024)) ? 1024 : 256;
    if (const char* e = std::getenv("EMME_PIC_BLOCK")) s->block = std::atoi(e);
    int per_sm = 1;
    {
        const void* fn = d.use_smem ? (sw ? (const void*)pic_stage_kernel<true, true> : (const void*)pic_stage_kernel<false, true>)
                                    : (sw ? (const void*)pic_stage_kernel<true, false> : (const void*)pic_stage_kernel<false, false>);
        if (d.use_smem) CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem));
        else s->smem = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, s->block, s->smem));
    }
    if (per_sm < 1) per_sm = 1;
    const long want = (n + s->block - 1) / s->block;
    const long cap = (long)s->sms * per_sm;
    s->grid = (int)(want < cap ? want : cap);
    if (s->grid < 1) s->grid = 1;
    if (const char* e = std::getenv("EMME_PIC_PERSISTENT")) s->use_persistent = std::atoi(e);
    int coop = 0;
    CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
    if (!d.use_smem || !coop || shard_count > 1) s->use_persistent = 0;
    if (s->use_persistent) {
        const void* pfn = sw ? (const void*)pic_persistent_kernel<true> : (const void*)pic_persistent_kernel<false>;
        CU(cudaFuncSetAttribute(pfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem));
        int pper = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pper, pfn, s->block, s->smem));
        if (pper < 1) {
            s->use_persistent = 0;
        } else {
            const long pcap = (long)s->sms * pper;
            s->pgrid = (int)(want < pcap ? want : pcap);
            if (s->pgrid > s->grid) s->grid = s->pgrid;   // one partial buffer serves both paths
        }
    }
    d.nparts = d.use_smem ? s->grid : 0;
    CU(dev_alloc(&d.part, (size_t)d.nparts * nf, s->stream));
    if (const char* e = std::getenv("EMME_PIC_GRAPH")) s->use_graph = std::atoi(e);
    if (const char* e = std::getenv("EMME_PIC_PDL")) s->use_pdl = std::atoi(e);
    if (shard_count > 1) {
        s->use_pdl = 0;   // the exchange (fused, or the caller's collective) sits between the kernels of a stage
        // exchange slots [2][P][nf] + ready flags [P][field blocks]: ONE cudaMalloc block (at least
        // 2 MiB, so that the exported allocation holds nothing else), mapped by every peer
        const size_t nblk = (size_t)(nf + 31) / 32;
        s->xch_flag_offset = sizeof(d2) * 2 * (size_t)shard_count * nf;
        s->xch_flag_offset = (s->xch_flag_offset + 255) / 256 * 256;
        size_t bytes = s->xch_flag_offset + sizeof(unsigned long long) * ((size_t)shard_count * nblk + 1);
        if (bytes < (2u << 20)) bytes = 2u << 20;
        CU(pic_preload());
        CU(cudaMalloc(&s->xch, bytes));
        CU(cudaMemsetAsync(s->xch, 0, bytes, s->stream));
        d.peers.n = shard_count;
        d.peers.me = shard_index;
        d.peers.timeout_ns = emme::peer_timeout_ns();   // emme_peer_set_timeout, 20 s by default
        d.peers.err = reinterpret_cast<unsigned long long*>((char*)s->xch + s->xch_flag_offset) + (size_t)shard_count * nblk;
    }

    const int ig = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
    if (sw) pic_init_kernel<true><<<ig > 0 ? ig : 1, 256, 0, s->stream>>>(d);
    else pic_init_kernel<false><<<ig > 0 ? ig : 1, 256, 0, s->stream>>>(d);
    s->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s->stream));
    lap("launch geometry + init kernel");
    guard.h = nullptr;
    *out = s;
    return 0;
}

// make the host copy of the device order available (emme_pic_markers / emme_pic_extras)
int ensure_perm(emme_pic* s) {
    if ((long)s->perm.size() == s->n) return 0;
    s->perm.resize(s->n);
    CU(cudaMemcpyAsync(s->perm.data(), s->d_perm, sizeof(long) * s->n, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}

// load every kernel a sharded step can launch before the first device-side wait exists (CUDA loads
// kernels lazily, and a load may have to wait for the context to drain; see dense_preload)
cudaError_t pic_preload() {
    cudaFuncAttributes a;
    const void* ks[] = {(const void*)pic_stage_kernel<true, true>, (const void*)pic_stage_kernel<false, true>,
                        (const void*)pic_stage_kernel<true, false>, (const void*)pic_stage_kernel<false, false>,
                        (const void*)pic_field_kernel, (const void*)pic_init_kernel<true>,
                        (const void*)pic_init_kernel<false>};
    for (const void* k : ks) {
        cudaError_t e = cudaFuncGetAttributes(&a, k);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

void set_peer(emme_pic* s, int r, void* base) {
    s->peer_xch[r] = base;
    s->d.peers.xch[r] = reinterpret_cast<d2*>(base);
    s->d.peers.flags[r] = reinterpret_cast<unsigned long long*>((char*)base + s->xch_flag_offset);
    bool all = true;
    for (int q = 0; q < s->shard_count; ++q) all = all && s->peer_xch[q] != nullptr;
    if (all) {
        s->peer_count = s->shard_count;
        if (s->graph) { cudaGraphExecDestroy(s->graph); s->graph = nullptr; }
    }
}

}  // namespace

extern "C" {

int emme_pic_pweight_sum(const emme_pic_params* p, long n, const double* v_para, const double* v_perp, double* sum) {
    if (!p) return capi_fail(-1, "null params");
    if (n < 0) return capi_fail(-2, "negative marker count");
    if (!v_para || !v_perp) return capi_fail(-3, "null velocity array");
    if (!sum) return capi_fail(-5, "null output");
    const double wa = 1 - p->water_bag_weight_vpara, wb = 1 - p->water_bag_weight_vperp;
    double acc = 0;
    // the exponent is -0/(2 vt^2) for unit weights: exp gives exactly 1, v_perp * 1 = v_perp
    if (wa == 0 && wb == 0) for (long i = 0; i < n; ++i) acc += v_perp[i];
    else for (long i = 0; i < n; ++i) acc += pweight_raw(p, v_para[i], v_perp[i], wa, wb);
    *sum = acc;
    return 0;
}

int emme_pic_create_block(const emme_pic_params* p, long n_total, long first, long n_local, const double* eta,
                          const double* v_para, const double* v_perp, const double* weight, double pw_sum,
                          int shard_index, int shard_count, int device, emme_pic** out) {
    if (!p) return capi_fail(-1, "null params");
    if (n_total <= 0) return capi_fail(-2, "marker count must be positive");
    if (first < 0 || n_local < 0 || first + n_local > n_total) return capi_fail(-3, "block outside the marker range");
    if (!eta || !v_para || !v_perp || !weight) return capi_fail(-5, "null marker array");
    if (!(pw_sum > 0)) return capi_fail(-9, "the p_weight sum over all markers must be positive");
    if (shard_count < 1 || shard_count > EMME_MAX_PEER_RANKS || shard_index < 0 || shard_index >= shard_count)
        return capi_fail(-10, "bad shard");
    if (!out) return capi_fail(-13, "null output handle");
    return create_impl(p, n_total, first, n_local, eta, v_para, v_perp, weight, pw_sum, shard_index, shard_count,
                       device, out);
}

int emme_pic_create_shard(const emme_pic_params* p, long n_total, const double* eta,
                          const double* v_para, const double* v_perp, const double* weight,
                          int shard_index, int shard_count, int device, emme_pic** out) {
    if (!p) return capi_fail(-1, "null params");
    if (n_total <= 0) return capi_fail(-2, "marker count must be positive");
    if (!eta) return capi_fail(-3, "null eta");
    if (!v_para) return capi_fail(-4, "null v_para");
    if (!v_perp) return capi_fail(-5, "null v_perp");
    if (!weight) return capi_fail(-6, "null weight");
    if (shard_count < 1 || shard_count > EMME_MAX_PEER_RANKS || shard_index < 0 || shard_index >= shard_count)
        return capi_fail(-7, "bad shard");
    if (!out) return capi_fail(-10, "null output handle");
    // initialize_marker_extras (include/solver_pic.h:207-238): p_weight normalised over ALL markers,
    // summed in the reference's order
    double sum = 0;
    if (int rc = emme_pic_pweight_sum(p, n_total, v_para, v_perp, &sum)) return rc;
    const long per = n_total / shard_count, rem = n_total % shard_count;
    const long first = shard_index * per + (shard_index < rem ? shard_index : rem);
    const long n = per + (shard_index < rem ? 1 : 0);
    return create_impl(p, n_total, first, n, eta + first, v_para + first, v_perp + first, weight + 2 * first, sum,
                       shard_index, shard_count, device, out);
}

int emme_pic_create(const emme_pic_params* p, long n_markers, const double* eta, const double* v_para,
                    const double* v_perp, const double* weight, int device, emme_pic** out) {
    return emme_pic_create_shard(p, n_markers, eta, v_para, v_perp, weight, 0, 1, device, out);
}

// ---- peer mapping of the exchange buffer (CUDA IPC, or plain pointers inside one process) ----
int emme_pic_ipc_export(emme_pic* s, void* handle64) {
    if (!s) return capi_fail(-1, "null handle");
    if (!handle64) return capi_fail(-2, "null output");
    if (!s->xch) return capi_fail(EMME_E_STATE, "emme_pic_ipc_export: not a sharded state");
    CU(cudaSetDevice(s->device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, s->xch));
    std::memcpy(handle64, &h, 64);
    return 0;
}

int emme_pic_ipc_import(emme_pic* s, int peer_rank, int peer_count, const void* handle64) {
    if (!s) return capi_fail(-1, "null handle");
    if (!s->xch || peer_count != s->shard_count) return capi_fail(-3, "emme_pic_ipc_import: peer count != shard count");
    if (peer_rank < 0 || peer_rank >= peer_count) return capi_fail(-2, "bad peer rank");
    CU(cudaSetDevice(s->device));
    if (peer_rank == s->shard_index) {
        set_peer(s, peer_rank, s->xch);
        return 0;
    }
    if (!handle64) return capi_fail(-4, "null handle bytes");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    void* ptr = nullptr;
    CU(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    s->peer_ipc[peer_rank] = true;
    set_peer(s, peer_rank, ptr);
    return 0;
}

int emme_pic_peer_attach(emme_pic* s, int peer_rank, int peer_count, emme_pic* peer) {
    if (!s) return capi_fail(-1, "null handle");
    if (!peer || !peer->xch) return capi_fail(-4, "emme_pic_peer_attach: the peer is not a sharded state");
    if (!s->xch || peer_count != s->shard_count || peer->shard_count != s->shard_count || peer->d.nf != s->d.nf)
        return capi_fail(-3, "emme_pic_peer_attach: shard count / mesh mismatch");
    if (peer_rank < 0 || peer_rank >= peer_count) return capi_fail(-2, "bad peer rank");
    CU(cudaSetDevice(s->device));
    if (peer->device != s->device) {
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, s->device, peer->device));
        if (!can) return capi_fail(EMME_E_PEER, "emme_pic_peer_attach: no peer access between the two devices");
        cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
        cudaGetLastError();
    }
    set_peer(s, peer_rank, peer->xch);
    return 0;
}

int emme_pic_destroy(emme_pic* s) {
    if (!s) return 0;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->graph) cudaGraphExecDestroy(s->graph);
    PicDev& d = s->d;
    if (s->stream) {   // back to the pool, stream-ordered
        void* bufs[] = {d.eta, s->d_vpar, s->d_vperp, s->d_pw, d.w, d.A, d.B, d.k1, d.c, d.field, d.dens,
                        s->d_coef, d.hist, d.part, d.step, s->d_perm};
        for (void* b : bufs)
            if (b) cudaFreeAsync(b, s->stream);
        cudaStreamSynchronize(s->stream);
    }
    for (int r = 0; r < EMME_MAX_PEER_RANKS; ++r)
        if (s->peer_xch[r] && s->peer_ipc[r]) cudaIpcCloseMemHandle(s->peer_xch[r]);
    if (s->xch) cudaFree(s->xch);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
    return 0;
}

int emme_pic_step(emme_pic* s, double dt, int nsteps) {
    if (!s) return capi_fail(-1, "null handle");
    if (nsteps < 0) return capi_fail(-3, "negative step count");
    if (s->shard_count > 1 && s->peer_count != s->shard_count)
        return capi_fail(EMME_E_STATE, "sharded PIC state without peer mappings: map the ranks (emme_pic_ipc_export/"
                                       "import or emme_pic_peer_attach), or use emme_pic_stage_begin/finish around "
                                       "your own density exchange");
    CU(cudaSetDevice(s->device));
    const int fmode = s->shard_count > 1 ? 3 : 0;   // 3: fused peer exchange inside the field kernel
    if (int rc = ensure_history(s, s->steps_done + nsteps)) return rc;
    if (s->use_persistent && nsteps > 0) {
        const bool sw = s->p.drift_center_transformation_switch != 0;
        const void* pfn = sw ? (const void*)pic_persistent_kernel<true> : (const void*)pic_persistent_kernel<false>;
        unsigned long long first_slot = (unsigned long long)s->steps_done;
        void* args[] = {&s->d, &dt, &nsteps, &first_slot};
        CU(cudaEventRecord(s->ev0, s->stream));
        CU(cudaLaunchCooperativeKernel(pfn, dim3(s->pgrid), dim3(s->block), args, s->smem, s->stream));
        s->launches++;
        CU(cudaEventRecord(s->ev1, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        float pms = 0;
        CU(cudaEventElapsedTime(&pms, s->ev0, s->ev1));
        s->last_ms = pms;
        s->steps_done += nsteps;
        return 0;
    }
    if (s->use_graph && nsteps > 0 && (!s->graph || s->graph_dt != dt)) {
        if (s->graph) { cudaGraphExecDestroy(s->graph); s->graph = nullptr; }
        cudaGraph_t g = nullptr;
        CU(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
        const unsigned long long before = s->launches;
        cudaError_t e = cudaSuccess;
        for (int st = 0; st < 3 && e == cudaSuccess; ++st) {
            e = launch_stage(s, dt, st);
            if (e == cudaSuccess) e = launch_field(s, fmode, st == 2, st);
        }
        s->launches = before;
        cudaError_t e2 = cudaStreamEndCapture(s->stream, &g);
        if (e == cudaSuccess) e = e2;
        if (e == cudaSuccess) e = cudaGraphInstantiate(&s->graph, g, 0);
        if (g) cudaGraphDestroy(g);          // on every path: a failed capture must not leak the graph
        if (e != cudaSuccess) {
            cudaGetLastError();
            s->graph = nullptr;
            return capi_fail(EMME_E_CUDA, std::string("step graph: ") + cudaGetErrorString(e));
        }
        s->graph_dt = dt;
    }
    CU(cudaEventRecord(s->ev0, s->stream));
    for (int k = 0; k < nsteps; ++k) {
        if (s->use_graph) {
            CU(cudaGraphLaunch(s->graph, s->stream));
            s->launches += 6;
        } else {
            for (int st = 0; st < 3; ++st) {
                CU(launch_stage(s, dt, st));
                CU(launch_field(s, fmode, st == 2, st));
            }
        }
    }
    CU(cudaEventRecord(s->ev1, s->stream));
    unsigned long long perr = 0;
    if (fmode == 3) CU(cudaMemcpyAsync(&perr, s->d.peers.err, sizeof perr, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    s->last_ms = ms;
    s->steps_done += nsteps;
    if (perr != 0) {
        CU(cudaMemsetAsync(s->d.peers.err, 0, sizeof perr, s->stream));
        return capi_fail(EMME_E_PEER, "PIC density exchange: a peer did not deliver its density in time (stage serial " +
                                          std::to_string(perr) + "); the ranks are out of step");
    }
    return 0;
}

int emme_pic_stage_begin(emme_pic* s, double dt, int stage) {
    if (!s) return capi_fail(-1, "null handle");
    if (stage < 0 || stage > 2) return capi_fail(-3, "stage must be 0, 1 or 2");
    if (s->next_call != 2 * stage)
        return capi_fail(EMME_E_STATE, "stage sequence is begin(0), finish(0), begin(1), finish(1), begin(2), finish(2)");
    CU(cudaSetDevice(s->device));
    if (stage == 0)
        if (int rc = ensure_history(s, s->steps_done + 1)) return rc;
    CU(launch_stage(s, dt, stage));
    if (s->d.use_smem) CU(launch_field(s, 1, 0));   // dens = sum of this rank's partials
    s->next_call = 2 * stage + 1;
    return 0;
}

int emme_pic_stage_finish(emme_pic* s, int stage) {
    if (!s) return capi_fail(-1, "null handle");
    if (stage < 0 || stage > 2) return capi_fail(-2, "stage must be 0, 1 or 2");
    if (s->next_call != 2 * stage + 1)
        return capi_fail(EMME_E_STATE, "emme_pic_stage_finish without the matching emme_pic_stage_begin");
    CU(cudaSetDevice(s->device));
    CU(launch_field(s, 2, stage == 2));
    s->next_call = stage == 2 ? 0 : 2 * stage + 2;
    if (stage == 2) s->steps_done++;
    return 0;
}

void* emme_pic_density_ptr(emme_pic* s) { return s ? (void*)s->d.dens : nullptr; }
void* emme_pic_stream(emme_pic* s) { return s ? (void*)s->stream : nullptr; }
long emme_pic_steps_done(const emme_pic* s) { return s ? s->steps_done : -1; }
long emme_pic_marker_num(const emme_pic* s) { return s ? s->n : -1; }

int emme_pic_current_field(emme_pic* s, void* host_out) {
    if (!s) return capi_fail(-1, "null handle");
    if (!host_out) return capi_fail(-2, "null output");
    CU(cudaSetDevice(s->device));
    CU(cudaMemcpyAsync(host_out, s->d.field, sizeof(d2) * s->d.nf, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}

int emme_pic_field_history(emme_pic* s, long first, long count, void* host_out) {
    if (!s) return capi_fail(-1, "null handle");
    if (first < 0 || count < 0 || first + count > s->steps_done) return capi_fail(-2, "step range outside the recorded history");
    if (!host_out) return capi_fail(-4, "null output");
    if (count == 0) return 0;
    CU(cudaSetDevice(s->device));
    CU(cudaMemcpyAsync(host_out, s->d.hist + (size_t)first * s->d.nf, sizeof(d2) * (size_t)count * s->d.nf,
                       cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}

int emme_pic_markers(emme_pic* s, double* eta, double* weight) {
    if (!s) return capi_fail(-1, "null handle");
    CU(cudaSetDevice(s->device));
    std::vector<double> t_eta(eta ? s->n : 0), t_w(weight ? 2 * (size_t)s->n : 0);
    if (eta) CU(cudaMemcpyAsync(t_eta.data(), s->d.eta, sizeof(double) * s->n, cudaMemcpyDeviceToHost, s->stream));
    if (weight) CU(cudaMemcpyAsync(t_w.data(), s->d.w, sizeof(d2) * s->n, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if (int rc = ensure_perm(s)) return rc;
    for (long j = 0; j < s->n; ++j) {   // back to the caller's marker order
        const long i = s->perm[j];
        if (eta) eta[i] = t_eta[j];
        if (weight) {
            weight[2 * i] = t_w[2 * j];
            weight[2 * i + 1] = t_w[2 * j + 1];
        }
    }
    return 0;
}

int emme_pic_extras(emme_pic* s, double* omega_dv, double* omega_st, double* p_weight, double* coef) {
    if (!s) return capi_fail(-1, "null handle");
    const emme_pic_params& p = s->p;
    if (omega_dv || omega_st || p_weight) {
        // the marker constants live on the device only (in device order): fetch and un-permute
        CU(cudaSetDevice(s->device));
        std::vector<double> t_vpar(s->n), t_vperp(s->n), t_pw(s->n);
        CU(cudaMemcpyAsync(t_vpar.data(), s->d_vpar, sizeof(double) * s->n, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaMemcpyAsync(t_vperp.data(), s->d_vperp, sizeof(double) * s->n, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaMemcpyAsync(t_pw.data(), s->d_pw, sizeof(double) * s->n, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        if (int rc = ensure_perm(s)) return rc;
        for (long j = 0; j < s->n; ++j) {
            const long i = s->perm[j];
            const double vp = t_vpar[j], vq = t_vperp[j];
            if (omega_dv) omega_dv[i] = (vp * vp + .5 * vq * vq) / (2. * p.vt * p.vt);
            if (omega_st)
                omega_st[i] = p.omega_s_i * (1. + p.eta_i * ((vp * vp + vq * vq) / (2. * p.vt * p.vt) - 1.5));
            if (p_weight) p_weight[i] = t_pw[j];
        }
    }
    if (coef)
        for (int i = 0; i < p.npoints; ++i) coef[i] = s->h_coef[i];
    return 0;
}

int emme_pic_field_stats(emme_pic* s, long first, long count, double* stats3) {
    if (!s) return capi_fail(-1, "null handle");
    if (!stats3) return capi_fail(-4, "null output");
    const int nf = s->d.nf;
    std::vector<std::complex<double>> h((size_t)count * nf);
    if (int rc = emme_pic_field_history(s, first, count, h.data())) return rc;
    for (long k = 0; k < count; ++k) {  // src/main.cpp:110-117
        double real = 0, imag = 0, norm = 0;
        for (int i = 0; i < nf; ++i) {
            const std::complex<double> val = h[(size_t)k * nf + i];
            real = real + std::real(val);
            imag = imag + std::imag(val);
            norm = norm + std::real(val * std::conj(val));
        }
        stats3[3 * k] = real / nf;
        stats3[3 * k + 1] = imag / nf;
        stats3[3 * k + 2] = std::sqrt(norm / nf);
    }
    return 0;
}

int emme_pic_calculate_omega(const double* stats3, long size, double dt, double* omega_re, double* omega_im) {
    if (!stats3) return capi_fail(-1, "null stats");
    if (size < 4) return capi_fail(-2, "need at least four steps");
    if (!omega_re || !omega_im) return capi_fail(-4, "null output");
    // util::calculate_omega (include/solver_pic.h:475-529)
    const std::size_t n = (std::size_t)size / 2;
    double t = 0, weighted_sum = 0, sum = 0;
    for (std::size_t i = n; i < (std::size_t)size; ++i) {
        const double val = std::log(stats3[3 * i + 2]);
        weighted_sum += val * t;
        sum += val;
        t += dt;
    }
    const double gamma = 6 * (2 * weighted_sum - dt * sum * (n + 1)) / (dt * dt * n * (n * n - 1));
    std::vector<double> real_log;
    real_log.reserve(size - n);
    for (std::size_t i = n; i < (std::size_t)size; ++i) real_log.push_back(std::log(std::abs(stats3[3 * i])));
    std::vector<double> max_pts;
    for (std::size_t i = 1; i + 1 < real_log.size(); ++i)
        if (real_log[i] > real_log[i - 1] && real_log[i] > real_log[i + 1]) max_pts.push_back(i * dt);
    double omega = 0;
    if (max_pts.size() > 1) omega = M_PI * (max_pts.size() - 1) / (max_pts.back() - max_pts.front());
    *omega_re = omega;
    *omega_im = gamma;
    return 0;
}

int emme_pic_get_timing(const emme_pic* s, double* last_step_call_ms, unsigned long long* launches) {
    if (!s) return capi_fail(-1, "null handle");
    if (last_step_call_ms) *last_step_call_ms = s->last_ms;
    if (launches) *launches = s->launches;
    return 0;
}

}  // extern "C"
