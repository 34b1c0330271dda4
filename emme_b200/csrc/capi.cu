// capi.cu -- the C ABI of include/emme_b200.h: handle management, the Newton/secant state
// machine of EigenSolver (reference include/solver.h:396-415, 113-160) around kernels 1 and 2,
// and the host-side input helpers.  No CPU fallback: every compute entry point needs a device.
#include <cuda_runtime.h>

#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/emme_b200.h"
#include "../host/json.hpp"
#include "../host/parameters.hpp"
#include "assembly.h"
#include "common.h"
#include "dense.h"
#include "peer.h"
#include "qr.h"
#include "run_const.h"

static_assert(EMME_MAX_PEERS == EMME_MAX_PEER_RANKS, "peer group size: csrc/assembly.h and include/emme_b200.h disagree");

using emme::RunConst;
typedef std::complex<double> zc;

static thread_local std::string g_err;

static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

int emme::capi_fail(int code, const std::string& msg) { return fail(code, msg); }

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return fail(EMME_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));    \
    } while (0)

struct emme_solver {
    int device = 0, sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t evd0 = nullptr, evd1 = nullptr;   // dense-step timing (created once, not per step)
    emme::DenseAux aux{};             // high-priority chain stream + fork/join events of the dense step
    int use_lookahead = 1;            // EMME_DENSE_LOOKAHEAD=0: panel chain on the main stream (one stream)
    // emme_copy_matrix_async: device->host copies overlap the next iterate on their own stream
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copy_src = nullptr, ev_copy_done = nullptr;
    bool copy_pending = false;
    emme_params p{};
    int N = 0, dim = 0;
    double *d_eta = nullptr, *d_g = nullptr, *d_bi = nullptr;
    // the three public matrices of EigenSolver + LU work copy
    void *A = nullptr, *Aold = nullptr, *Ad = nullptr, *W = nullptr;
    unsigned long long *d_counter = nullptr, *d_stats = nullptr;
    void* d_spill = nullptr;
    void* d_trig = nullptr;           // node table of kernel 1
    int spill_cap = 0, grid_blocks = 0;
    int launch_blocks = 0;            // CTAs actually launched (<= grid_blocks, which sizes the spill area)
    void* d_dense_ws = nullptr;
    // symmetric dense path: M = L^-1, its scaled transpose, per-tile partial traces
    void *Y = nullptr, *YT = nullptr, *d_sym_ws = nullptr;
    void* d_qr_ws = nullptr;          // QR-secant iterate: norms, reflector scalars, vectors
    double2* d_qr_out = nullptr;      // [R_nn, (Q^H A' v)_n]
    double2* d_trace = nullptr;
    int* d_info = nullptr;
    // Newton state
    bool seeded = false;
    zc w{0, 0}, dw{0, 0};
    int shard_index = 0, shard_count = 1;
    // Peer group (one NVSwitch box).  Buffers every rank exposes to the others (CUDA IPC mappings, or
    // plain pointers for ranks inside one process): 0, 1 = the two physical matrix buffers that A and
    // A_old alternate between (pair-sharded assembly stores straight into them), 2 = flag page
    // (peer.h), 3 = W, 4 = Y, 5 = partial-trace workspace (column-sharded dense step).
    void* phys[2] = {nullptr, nullptr};
    unsigned long long* flag_page = nullptr;
    void* peer_buf[EMME_PEER_BUFS][EMME_MAX_PEERS] = {};
    bool peer_ipc[EMME_PEER_BUFS][EMME_MAX_PEERS] = {};
    int peer_count = 0;               // 0: local stores only; > 1 once every mapping is present
    int peer_pending = 0;             // group size announced by emme_ipc_import / emme_peer_attach
    unsigned long long peer_epoch = 0;    // barriers passed (all ranks call them in the same order)
    unsigned long long dense_serial = 0;  // sharded dense steps started
    int dense_sharded = 0;
    emme_stats stats{};
    unsigned long long launches = 0;
    int refill_min = 32;
    int d_split = 0;                  // kernel 1's item order: diagonals >= d_split first (0: d = 1, 2, ...)
    int optimistic = 1;               // try the interchange-free factorisation first
    int null_optimistic = 1;
    int use_graph = 1;                // replay the optimistic dense step as a CUDA graph
    int use_sym = 1;                  // try the symmetric (L D L^T, explicit inverse) path first
    cudaGraphExec_t dense_graph = nullptr, sym_graph = nullptr, qr_graph = nullptr;
    unsigned long long dense_graph_launches = 0, sym_graph_launches = 0, qr_graph_launches = 0;
    unsigned long long pivot_fallbacks = 0, sym_steps = 0;
    int* d_flag = nullptr;
    size_t bytes() const { return sizeof(double) * 2 * (size_t)dim * dim; }
};

extern "C" {

const char* emme_last_error(void) { return g_err.c_str(); }
const char* emme_version(void) { return "emme_b200 0.1 (sm_100a)"; }

int emme_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int emme_dim(const emme_solver* s) { return s ? s->dim : -1; }

int emme_fp64_peak(int device, double* tflops, double* sm_mhz_nominal) {
    if (!tflops) return fail(-2, "null output");
    if (emme_device_count() <= 0) return fail(EMME_E_NO_DEVICE, "no CUDA device");
    CU(cudaSetDevice(device));
    CU(emme::measure_fp64_peak(tflops));
    if (sm_mhz_nominal) {
        int khz = 0;
        CU(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device));
        *sm_mhz_nominal = khz / 1000.0;
    }
    return 0;
}

// Item order of kernel 1 (assembly.cu::decode_item).  The Miller recurrence of a pair starts at
// floor(|z|) + 1 with |z| = sqrt(b b') / |lambda|.  Where sqrt(b b') of the pairs that span the mesh is
// large, the recurrence outweighs the rest of an evaluation and the far diagonals are the costliest
// items after the near-singular ones, so they go first.  Measured (profiles/r2_order_sweep.txt): C1
// physics, sqrt(b_0 b_N-1) = 24.7: N = 512 0.77 -> 0.68 ms, N = 1024 1.94 -> 1.84, N = 2048 6.71 -> 6.59,
// neutral from 4096 up; C3 (6.5: |z| <= 9, far pairs are the cheapest items) 0.881 -> 0.894 ms, so it
// keeps the plain order.  The switch sits between the two.
// EMME_ASM_FAR_SPLIT=f overrides: diagonals d >= f*N first, 0 = plain order.
static void choose_item_order(emme_solver* s, const double* bi) {
    double frac = std::sqrt(std::fabs(bi[0] * bi[s->N - 1])) >= 16.0 ? 0.5 : 0.0;
    if (const char* e = std::getenv("EMME_ASM_FAR_SPLIT")) frac = std::atof(e);
    s->d_split = (frac > 0.0 && frac < 1.0) ? (int)(frac * s->N) : 0;
    if (s->d_split < 2) s->d_split = 0;
}

int emme_set_tables(emme_solver* s, const double* eta, const double* g, const double* bi) {
    if (!s) return fail(-1, "null handle");
    if (!eta) return fail(-2, "null eta");
    if (!g) return fail(-3, "null g");
    if (!bi) return fail(-4, "null bi");
    CU(cudaSetDevice(s->device));
    const size_t tb = sizeof(double) * s->N;
    CU(cudaMemcpyAsync(s->d_eta, eta, tb, cudaMemcpyHostToDevice, s->stream));
    CU(cudaMemcpyAsync(s->d_g, g, tb, cudaMemcpyHostToDevice, s->stream));
    CU(cudaMemcpyAsync(s->d_bi, bi, tb, cudaMemcpyHostToDevice, s->stream));
    choose_item_order(s, bi);
    return 0;
}

int emme_set_params(emme_solver* s, const emme_params* p) {
    if (!s) return fail(-1, "null handle");
    if (!p) return fail(-2, "null params");
    const int dim = std::fpclassify(p->beta_e) == FP_ZERO ? s->N : 2 * s->N;
    if (dim != s->dim) return fail(-2, "emme_set_params: beta_e changes the matrix dimension");
    if (p->integration_start_points != s->p.integration_start_points ||
        p->integration_iteration_limit != s->p.integration_iteration_limit)
        return fail(-2, "emme_set_params: quadrature order / depth must not change");
    s->p = *p;
    return 0;
}

int emme_destroy(emme_solver* s) {
    if (!s) return 0;
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();          // no peer may be mid-store into buffers that are about to go
    for (int w = 0; w < EMME_PEER_BUFS; ++w)
        for (int r = 0; r < EMME_MAX_PEERS; ++r)
            if (s->peer_buf[w][r] && s->peer_ipc[w][r]) cudaIpcCloseMemHandle(s->peer_buf[w][r]);
    cudaFree(s->flag_page);
    cudaFree(s->d_eta);
    cudaFree(s->d_g);
    cudaFree(s->d_bi);
    cudaFree(s->A);
    cudaFree(s->Aold);
    cudaFree(s->Ad);
    cudaFree(s->W);
    cudaFree(s->d_counter);
    cudaFree(s->d_stats);
    cudaFree(s->d_spill);
    cudaFree(s->d_trig);
    cudaFree(s->d_dense_ws);
    cudaFree(s->Y);
    cudaFree(s->YT);
    cudaFree(s->d_sym_ws);
    cudaFree(s->d_qr_ws);
    cudaFree(s->d_qr_out);
    cudaFree(s->d_trace);
    cudaFree(s->d_info);
    cudaFree(s->d_flag);
    if (s->dense_graph) cudaGraphExecDestroy(s->dense_graph);
    if (s->sym_graph) cudaGraphExecDestroy(s->sym_graph);
    if (s->qr_graph) cudaGraphExecDestroy(s->qr_graph);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->evd0) cudaEventDestroy(s->evd0);
    if (s->evd1) cudaEventDestroy(s->evd1);
    if (s->aux.ev_panel) cudaEventDestroy(s->aux.ev_panel);
    if (s->aux.ev_rest) cudaEventDestroy(s->aux.ev_rest);
    if (s->aux.side) cudaStreamDestroy(s->aux.side);
    if (s->ev_copy_src) cudaEventDestroy(s->ev_copy_src);
    if (s->ev_copy_done) cudaEventDestroy(s->ev_copy_done);
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
    return 0;
}

int emme_create(const emme_params* p, int npoints, const double* eta, const double* g,
                const double* bi, int device, emme_solver** out) {
    if (!p) return fail(-1, "emme_create: params is null");
    if (npoints < 2) return fail(-2, "emme_create: npoints < 2");
    if (!eta) return fail(-3, "emme_create: eta is null");
    if (!g) return fail(-4, "emme_create: g is null");
    if (!bi) return fail(-5, "emme_create: bi is null");
    if (!out) return fail(-7, "emme_create: out is null");
    if (p->integration_start_points != 15 && p->integration_start_points != 31)
        return fail(EMME_E_BAD_ORDER, "integration_start_points should be 15 or 31");
    int ndev = emme_device_count();
    if (ndev <= 0) return fail(EMME_E_NO_DEVICE, "no CUDA device: emme_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(-6, "emme_create: device out of range");
    CU(cudaSetDevice(device));
    std::unique_ptr<emme_solver, int (*)(emme_solver*)> s(new emme_solver, emme_destroy);
    s->device = device;
    s->p = *p;
    s->N = npoints;
    s->dim = std::fpclassify(p->beta_e) == FP_ZERO ? npoints : 2 * npoints;
    CU(cudaDeviceGetAttribute(&s->sms, cudaDevAttrMultiProcessorCount, device));
    {
        // the side stream carries the dependent chain of the dense step (panel factorisations): its
        // short kernels take SM slots ahead of the bulk updates queued on the handle's stream
        int prio_low = 0, prio_high = 0;
        CU(cudaDeviceGetStreamPriorityRange(&prio_low, &prio_high));
        CU(cudaStreamCreateWithPriority(&s->stream, cudaStreamNonBlocking, prio_low));
        CU(cudaStreamCreateWithPriority(&s->aux.side, cudaStreamNonBlocking, prio_high));
    }
    CU(cudaEventCreate(&s->ev0));
    CU(cudaEventCreate(&s->ev1));
    CU(cudaEventCreate(&s->evd0));
    CU(cudaEventCreate(&s->evd1));
    CU(cudaEventCreateWithFlags(&s->aux.ev_panel, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&s->aux.ev_rest, cudaEventDisableTiming));
    if (const char* e = std::getenv("EMME_DENSE_LOOKAHEAD")) s->use_lookahead = std::atoi(e) != 0;
    const size_t tb = sizeof(double) * npoints;
    CU(cudaMalloc(&s->d_eta, tb));
    CU(cudaMalloc(&s->d_g, tb));
    CU(cudaMalloc(&s->d_bi, tb));
    CU(cudaMemcpy(s->d_eta, eta, tb, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(s->d_g, g, tb, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(s->d_bi, bi, tb, cudaMemcpyHostToDevice));
    choose_item_order(s.get(), bi);
    CU(cudaMalloc(&s->A, s->bytes()));
    CU(cudaMalloc(&s->d_counter, sizeof(unsigned long long)));
    CU(cudaMalloc(&s->d_stats, 8 * sizeof(unsigned long long)));
    s->grid_blocks = emme::assembly_grid_blocks(p->integration_start_points, device);
    s->launch_blocks = s->grid_blocks;
    if (const char* e = std::getenv("EMME_ASM_BLOCKS_PER_SM")) {   // tuning: resident CTAs per SM
        const int v = std::atoi(e) * s->sms;
        if (v >= 1 && v <= s->grid_blocks) s->launch_blocks = v;
    }
    // interval stack: at most integration_iteration_limit right siblings
    if (const char* e = std::getenv("EMME_REFILL_MIN")) s->refill_min = std::atoi(e);
    if (const char* e = std::getenv("EMME_DENSE_GRID_PANEL")) emme::dense_force_grid_panel(std::atoi(e) != 0);
    if (const char* e = std::getenv("EMME_DENSE_NBO")) emme::dense_set_outer_block(std::atoi(e));
    if (s->refill_min < 1) s->refill_min = 1;
    if (s->refill_min > 32) s->refill_min = 32;
    s->spill_cap = p->integration_iteration_limit + 1 - emme::assembly_stack_smem();
    if (s->spill_cap < 0) s->spill_cap = 0;
    if (s->spill_cap > 0) {
        const size_t groups = (size_t)s->grid_blocks *
                              emme::assembly_groups_per_block(p->integration_start_points);
        CU(cudaMalloc(&s->d_spill, groups * s->spill_cap * sizeof(double2)));
    }
    CU(cudaMalloc(&s->d_trig, emme::assembly_node_table_bytes(p->integration_start_points)));
    CU(cudaMalloc(&s->d_trace, sizeof(double2)));
    CU(cudaMalloc(&s->d_info, sizeof(int)));
    CU(cudaMalloc(&s->d_flag, sizeof(int)));
    if (const char* e = std::getenv("EMME_DENSE_OPTIMISTIC")) s->optimistic = std::atoi(e) != 0;
    if (const char* e = std::getenv("EMME_DENSE_TAU")) emme::dense_set_pivot_threshold(std::atof(e));
    if (const char* e = std::getenv("EMME_DENSE_GRAPH")) s->use_graph = std::atoi(e) != 0;
    if (const char* e = std::getenv("EMME_DENSE_SYM")) s->use_sym = std::atoi(e) != 0;
    *out = s.release();
    return 0;
}

static int ensure_newton_buffers(emme_solver* s) {
    CU(emme::dense_prepare());
    if (!s->Aold) CU(cudaMalloc(&s->Aold, s->bytes()));
    if (!s->Ad) CU(cudaMalloc(&s->Ad, s->bytes()));
    if (!s->W) CU(cudaMalloc(&s->W, s->bytes()));
    if (!s->d_dense_ws) CU(cudaMalloc(&s->d_dense_ws, emme::dense_workspace_bytes(s->dim)));
    if (s->use_sym) {
        if (!s->Y) CU(cudaMalloc(&s->Y, s->bytes()));
        if (!s->YT) CU(cudaMalloc(&s->YT, s->bytes()));
        if (!s->d_sym_ws) {   // peer-visible in the sharded dense step: at least one 2 MiB block of its own
            size_t wsb = emme::dense_sym_workspace_bytes(s->dim);
            if (wsb < (2u << 20)) wsb = 2u << 20;
            CU(cudaMalloc(&s->d_sym_ws, wsb));
        }
    }
    return 0;
}

// enqueue one assembly of A(w) into `dst` (device), this handle's shard only
static emme::PeerFlags peer_flags(const emme_solver* s) {
    emme::PeerFlags f{};
    f.n = s->peer_count;
    f.me = s->shard_index;
    for (int r = 0; r < s->peer_count; ++r) f.p[r] = (unsigned long long*)s->peer_buf[EMME_PEER_BUF_FLAGS][r];
    return f;
}

// stream-ordered barrier over the peer group (a kernel: no host round trip, no NCCL)
static int peer_barrier(emme_solver* s) {
    if (s->peer_count <= 1) return 0;
    CU(emme::launch_peer_barrier(peer_flags(s), ++s->peer_epoch, s->stream));
    ++s->launches;
    return 0;
}

// after a stream synchronisation: did a wait on the flag page give up?
static int peer_check(emme_solver* s) {
    if (s->peer_count <= 1 || !s->flag_page) return 0;
    unsigned long long err = 0;
    CU(cudaMemcpyAsync(&err, s->flag_page + emme::PEER_W_ERROR, sizeof err, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if (err != 0) {
        CU(cudaMemsetAsync(s->flag_page + emme::PEER_W_ERROR, 0, sizeof err, s->stream));
        return fail(EMME_E_PEER, "multi-GPU exchange: a peer did not arrive in time (flag word " +
                                     std::to_string(err - 1) + "); the ranks are out of step");
    }
    return 0;
}

// to_peers: store the entries into the matrix of EVERY rank (pair-sharded Newton path only; the
// caller brackets the launch with peer barriers)
static int enqueue_assembly(emme_solver* s, zc w, void* dst, int shard_index, int shard_count,
                            bool to_peers = false) {
    if (!std::isfinite(w.real()) || !std::isfinite(w.imag()))
        return fail(EMME_E_NONFINITE, "matrixAssembler: omega is not finite");
    RunConst rc = emme::make_run_const(s->p, s->N, w.real(), w.imag());
    emme::PeerSet ps{};
    ps.n = 1;
    ps.p[0] = (double2*)dst;
    if (to_peers && s->peer_count > 1 && (dst == s->phys[0] || dst == s->phys[1])) {
        const int which = dst == s->phys[0] ? 0 : 1;
        ps.n = s->peer_count;
        for (int r = 0; r < s->peer_count; ++r) ps.p[r] = (double2*)s->peer_buf[which][r];
    }
    if (s->copy_pending) {
        // an asynchronous download may still be reading the buffer this assembly overwrites
        CU(cudaStreamWaitEvent(s->stream, s->ev_copy_done, 0));
        s->copy_pending = false;
    }
    CU(cudaEventRecord(s->ev0, s->stream));
    CU(emme::launch_assembly(rc, s->d_eta, s->d_g, s->d_bi, ps, shard_index, shard_count,
                             s->d_counter, s->d_spill, s->spill_cap, s->d_stats, s->launch_blocks,
                             s->stream, &s->launches, s->refill_min, s->d_trig, s->d_split));
    CU(cudaEventRecord(s->ev1, s->stream));
    return 0;
}

static int collect_stats(emme_solver* s) {
    unsigned long long h[8];
    CU(cudaMemcpyAsync(h, s->d_stats, sizeof h, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    s->stats.integrals = h[0];
    s->stats.panels = h[1];
    s->stats.evals = h[2];
    s->stats.fwd_trips = h[3];
    s->stats.bwd_trips = h[4];
    s->stats.max_stack = h[5];
    s->stats.assemble_ms = ms;
    return 0;
}

int emme_assemble_device(emme_solver* s, double wr, double wi, void* dev_out, int shard_index,
                         int shard_count) {
    if (!s) return fail(-1, "emme_assemble_device: null handle");
    if (!dev_out) return fail(-4, "emme_assemble_device: null output");
    if (shard_count < 1 || shard_index < 0 || shard_index >= shard_count)
        return fail(-5, "emme_assemble_device: bad shard");
    CU(cudaSetDevice(s->device));
    int rc = enqueue_assembly(s, zc(wr, wi), dev_out, shard_index, shard_count);
    if (rc) return rc;
    return collect_stats(s);
}

int emme_assemble(emme_solver* s, double wr, double wi, void* host_out) {
    if (!s) return fail(-1, "emme_assemble: null handle");
    if (!host_out) return fail(-4, "emme_assemble: null output");
    CU(cudaSetDevice(s->device));
    int rc = enqueue_assembly(s, zc(wr, wi), s->A, 0, 1);
    if (rc) return rc;
    CU(cudaMemcpyAsync(host_out, s->A, s->bytes(), cudaMemcpyDeviceToHost, s->stream));
    return collect_stats(s);
}

}  // extern "C" (templates below)

// ---- dense step on (A, Ad): delta = -1/trace(A^-1 Ad); A is preserved ----
// Three paths, each verified on the device, tried in this order:
//   0  symmetric: A = L D L^T without interchanges, trace from the explicit inverse (4 dim^3
//      flops; Ad is only read).  Abandoned if A is not bitwise symmetric (flag bit 2) or partial
//      pivoting would have interchanged rows (flag bit 1);
//   1  optimistic LU of the augmented system (no interchanges, verified; Ad destroyed);
//   2  LU with partial pivoting.
// `restore_rhs` rebuilds Ad after a path that destroyed it.

// Capture a fixed launch sequence on fixed buffers once, then replay it (the launch-bound loop).
template <class Enqueue>
static int replay_graph(emme_solver* s, cudaGraphExec_t* exec, unsigned long long* n_in_graph,
                        Enqueue enqueue) {
    if (!*exec && s->use_graph) {
        cudaGraph_t g = nullptr;
        unsigned long long nl = 0;
        CU(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
        cudaError_t le = enqueue(&nl);
        cudaError_t ce = cudaStreamEndCapture(s->stream, &g);
        if (le != cudaSuccess || ce != cudaSuccess || !g) {
            cudaGetLastError();
            s->use_graph = 0;          // capture not possible: fall back to plain launches
            if (g) cudaGraphDestroy(g);
        } else {
            cudaError_t ie = cudaGraphInstantiate(exec, g, 0);
            cudaGraphDestroy(g);
            if (ie != cudaSuccess) {
                cudaGetLastError();
                *exec = nullptr;
                s->use_graph = 0;
            }
            *n_in_graph = nl;
        }
    }
    if (s->use_graph && *exec) {
        CU(cudaGraphLaunch(*exec, s->stream));
        s->launches += *n_in_graph;
    } else {
        CU(enqueue(&s->launches));
    }
    return 0;
}

template <class RestoreRhs>
static int dense_delta(emme_solver* s, zc* delta, RestoreRhs restore_rhs) {
    const cudaEvent_t e0 = s->evd0, e1 = s->evd1;
    CU(cudaEventRecord(e0, s->stream));
    double tr[2] = {0., 0.};
    int info = 0, flag = 0;
    bool rhs_intact = true;
    int path = (s->use_sym && s->Y) ? 0 : (s->optimistic ? 1 : 2);
    const double d3 = (double)s->dim * s->dim * s->dim;
    for (;;) {
        if (path == 0) {
            CU(cudaMemsetAsync(s->d_flag, 0, sizeof(int), s->stream));
            CU(emme::launch_sym_copy_check(s->A, s->W, s->dim, s->d_flag, s->stream, &s->launches));
            if (s->dense_sharded && s->peer_count > 1) {
                // column-block-cyclic over the ranks: panels travel by peer stores + flags
                emme::DensePeers dp{};
                dp.flags = peer_flags(s);
                dp.n = s->peer_count;
                dp.me = s->shard_index;
                for (int r = 0; r < s->peer_count; ++r) {
                    dp.W[r] = s->peer_buf[EMME_PEER_BUF_W][r];
                    dp.Y[r] = s->peer_buf[EMME_PEER_BUF_Y][r];
                    dp.ws[r] = s->peer_buf[EMME_PEER_BUF_WS][r];
                }
                dp.serial = ++s->dense_serial;
                dp.epoch = &s->peer_epoch;
                // one stream: with several ranks the step is bound by the panel chain itself, and chain
                // kernels that wait for bulk CTAs to retire make it longer (N = 8192, chain stream on /
                // off: 43.6 / 43.3 ms on 2 GPUs, 28.4 / 25.1 on 4, 20.8 / 17.9 on 8; one GPU 73.9 / 78.7)
                CU(emme::launch_trace_sym(s->W, s->Y, s->YT, s->Ad, s->dim, s->d_sym_ws, s->d_trace, s->d_info,
                                          s->d_flag, s->stream, &s->launches, &dp, nullptr));
            } else {
                auto enqueue = [&](unsigned long long* nl) {
                    return emme::launch_trace_sym(s->W, s->Y, s->YT, s->Ad, s->dim, s->d_sym_ws, s->d_trace,
                                                  s->d_info, s->d_flag, s->stream, nl, nullptr,
                                                  s->use_lookahead ? &s->aux : nullptr);
                };
                // Large systems run the blocked path with the panel chain on a high-priority stream;
                // replayed from a CUDA graph the chain's kernels lose most of that priority (dim 8192:
                // 76.3 ms replayed, 73.9 ms launched directly, 78.7 ms on one stream) and the ~600
                // launches of a step hide behind 70 ms of work anyway: launched directly from 4096 up.
                const int nbo = emme::dense_sym_outer_block(s->dim);
                if (s->use_lookahead && nbo >= 64 && nbo % 64 == 0 && s->dim >= 4096) {
                    CU(enqueue(&s->launches));
                } else {
                    int rc = replay_graph(s, &s->sym_graph, &s->sym_graph_launches, enqueue);
                    if (rc) return rc;
                }
            }
        } else {
            if (!rhs_intact) {
                int rc = restore_rhs();
                if (rc) return rc;
            }
            rhs_intact = false;
            CU(cudaMemcpyAsync(s->W, s->A, s->bytes(), cudaMemcpyDeviceToDevice, s->stream));
            auto enqueue = [&](unsigned long long* nl) {
                return emme::launch_trace_solve(s->W, s->Ad, s->dim, s->d_dense_ws, s->d_trace, s->d_info,
                                                s->stream, nl, path == 1, s->d_flag);
            };
            if (path == 1) {
                int rc = replay_graph(s, &s->dense_graph, &s->dense_graph_launches, enqueue);
                if (rc) return rc;
            } else {
                CU(enqueue(&s->launches));
            }
        }
        CU(cudaEventRecord(e1, s->stream));
        CU(cudaMemcpyAsync(tr, s->d_trace, sizeof tr, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaMemcpyAsync(&info, s->d_info, sizeof info, cudaMemcpyDeviceToHost, s->stream));
        flag = 0;
        if (path < 2) CU(cudaMemcpyAsync(&flag, s->d_flag, sizeof flag, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        if (path == 0) {
            if (s->dense_sharded && s->peer_count > 1) {
                int rc = peer_check(s);
                if (rc) return rc;
            }
            if (flag == 0) {
                ++s->sym_steps;
                s->stats.dense_flops = 4.0 * d3 / ((s->dense_sharded && s->peer_count > 1) ? s->peer_count : 1);
                break;
            }
            if ((flag & 1) || !s->optimistic) {
                if (flag & 1) ++s->pivot_fallbacks;
                path = 2;
            } else {
                path = 1;
            }
            continue;
        }
        s->stats.dense_flops = (8.0 / 3.0 + 4.0 + 2.0) * d3;
        if (path == 1 && flag != 0) {
            ++s->pivot_fallbacks;
            path = 2;
            continue;
        }
        break;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    s->stats.dense_ms = ms;
    // d_eigen_value = -1.0 / trace   (include/solver.h:139)
    *delta = -1.0 / zc(tr[0], tr[1]);
    if (info != 0) {
        char buf[256];
        std::snprintf(buf, sizeof buf,
                      "Linear solve failed. The factorization has been completed, but the pivot "
                      "is exactly singular at %d, so the solution could not be computed.", info);
        return fail(info, buf);
    }
    // The reference divides by the trace unchecked (include/solver.h:139): iterating past
    // convergence makes A == A_old, A' = 0, trace = 0 and omega non-finite, after which its next
    // zsysv fails or returns garbage.  Here a zero or non-finite trace is reported as a failed
    // linear solve right away, and no assembly is ever launched at a non-finite omega.
    if ((tr[0] == 0.0 && tr[1] == 0.0) || !std::isfinite(tr[0]) || !std::isfinite(tr[1]) ||
        !std::isfinite(delta->real()) || !std::isfinite(delta->imag())) {
        char buf[256];
        std::snprintf(buf, sizeof buf,
                      "Linear solve failed. trace(A^-1 A') = (%g, %g): the Newton step -1/trace is not "
                      "finite (iterating past convergence, or a non-finite matrix).", tr[0], tr[1]);
        return fail(EMME_E_NONFINITE, buf);
    }
    return 0;
}

extern "C" {

static int secant(emme_solver* s) {
    CU(emme::launch_secant(s->A, s->Aold, s->Ad, (size_t)s->dim * s->dim, s->dw.real(),
                           s->dw.imag(), s->sms, s->stream));
    ++s->launches;
    return 0;
}

int emme_shard_config(emme_solver* s, int shard_index, int shard_count) {
    if (!s) return fail(-1, "null handle");
    if (shard_count < 1 || shard_index < 0 || shard_index >= shard_count)
        return fail(-2, "emme_shard_config: bad shard");
    s->shard_index = shard_index;
    s->shard_count = shard_count;
    return 0;
}

// When sharded, the caller completes the matrix between begin/middle/finish; the buffer is
// zeroed first so that shares can be summed.
static int assemble_current(emme_solver* s) {
    // with peer stores every rank writes every entry of every GPU's matrix: nothing to zero or sum
    if (s->shard_count > 1 && s->peer_count <= 1) CU(cudaMemsetAsync(s->A, 0, s->bytes(), s->stream));
    const bool p2p = s->shard_count > 1 && s->peer_count > 1;
    if (p2p) {
        // The buffer about to be overwritten on EVERY rank is the one each of them may still be
        // reading as eigen_matrix_old (secant of the previous iterate, restore_rhs of a fallback dense
        // step, a pending download): nobody stores before everybody has arrived here.
        if (s->copy_pending) {
            CU(cudaStreamWaitEvent(s->stream, s->ev_copy_done, 0));
            s->copy_pending = false;
        }
        int rc = peer_barrier(s);
        if (rc) return rc;
    }
    int rc = enqueue_assembly(s, s->w, s->A, s->shard_index, s->shard_count, p2p);
    if (rc) return rc;
    if (p2p) {
        // ... and nobody reads the matrix before everybody's stores have landed
        rc = peer_barrier(s);
        if (rc) return rc;
    }
    rc = collect_stats(s);
    if (rc) return rc;
    return p2p ? peer_check(s) : 0;
}

// ---- peer group plumbing: see emme_b200/parallel.py::ShardedEigenSolver ----
static int ensure_peer_buffers(emme_solver* s) {
    int rc = ensure_newton_buffers(s);
    if (rc) return rc;
    // every kernel the peer protocol can launch is loaded before the first device-side wait exists
    CU(emme::assembly_preload());
    CU(emme::dense_preload());
    CU(emme::peer_preload());
    if (!s->phys[0]) {
        s->phys[0] = s->A;
        s->phys[1] = s->Aold;
    }
    if (!s->flag_page) {
        // a 2 MiB allocation of its own: small cudaMalloc blocks share pages, and an exported page
        // should expose nothing but the flags
        static_assert(sizeof(unsigned long long) * emme::PEER_PAGE_WORDS <= (2u << 20), "flag page size");
        CU(cudaMalloc(&s->flag_page, 2u << 20));
        CU(cudaMemset(s->flag_page, 0, 2u << 20));
    }
    return 0;
}

static void* local_peer_buf(emme_solver* s, int which) {
    switch (which) {
        case EMME_PEER_BUF_MATRIX0: return s->phys[0];
        case EMME_PEER_BUF_MATRIX1: return s->phys[1];
        case EMME_PEER_BUF_FLAGS: return s->flag_page;
        case EMME_PEER_BUF_W: return s->W;
        case EMME_PEER_BUF_Y: return s->Y;
        case EMME_PEER_BUF_WS: return s->d_sym_ws;
    }
    return nullptr;
}

static void peer_mapping_added(emme_solver* s, int peer_count) {
    s->peer_pending = peer_count;
    bool all = true;
    for (int w = 0; w < EMME_PEER_BUFS; ++w) {
        if (!local_peer_buf(s, w)) continue;      // W / Y / workspace only exist on the symmetric path
        for (int r = 0; r < peer_count; ++r) all = all && s->peer_buf[w][r] != nullptr;
    }
    if (all) s->peer_count = peer_count;   // switch on once every mapping is present
}

int emme_ipc_export(emme_solver* s, int which, void* handle64) {
    if (!s) return fail(-1, "null handle");
    if (which < 0 || which >= EMME_PEER_BUFS) return fail(-2, "emme_ipc_export: which must be 0..5");
    if (!handle64) return fail(-3, "null output");
    CU(cudaSetDevice(s->device));
    int rc = ensure_peer_buffers(s);
    if (rc) return rc;
    void* buf = local_peer_buf(s, which);
    if (!buf) return fail(EMME_E_STATE, "emme_ipc_export: this buffer does not exist on this handle");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, buf));
    std::memcpy(handle64, &h, 64);
    return 0;
}

int emme_ipc_import(emme_solver* s, int peer_rank, int peer_count, int which, const void* handle64) {
    if (!s) return fail(-1, "null handle");
    if (peer_count < 1 || peer_count > EMME_MAX_PEERS) return fail(-3, "emme_ipc_import: 1..8 peers");
    if (peer_rank < 0 || peer_rank >= peer_count) return fail(-2, "emme_ipc_import: bad peer rank");
    if (which < 0 || which >= EMME_PEER_BUFS) return fail(-4, "emme_ipc_import: which must be 0..5");
    CU(cudaSetDevice(s->device));
    if (!s->phys[0]) return fail(EMME_E_STATE, "emme_ipc_import before emme_ipc_export");
    if (peer_rank == s->shard_index) {
        s->peer_buf[which][peer_rank] = local_peer_buf(s, which);
    } else {
        if (!handle64) return fail(-5, "null handle bytes");
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handle64, 64);
        void* ptr = nullptr;
        CU(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        s->peer_buf[which][peer_rank] = ptr;
        s->peer_ipc[which][peer_rank] = true;
    }
    peer_mapping_added(s, peer_count);
    return 0;
}

int emme_peer_attach(emme_solver* s, int peer_rank, int peer_count, emme_solver* peer) {
    if (!s) return fail(-1, "null handle");
    if (peer_count < 1 || peer_count > EMME_MAX_PEERS) return fail(-3, "emme_peer_attach: 1..8 peers");
    if (peer_rank < 0 || peer_rank >= peer_count) return fail(-2, "emme_peer_attach: bad peer rank");
    if (!peer) return fail(-4, "emme_peer_attach: null peer");
    if (peer->dim != s->dim) return fail(-4, "emme_peer_attach: peers must have the same dimension");
    CU(cudaSetDevice(peer->device));
    int rc = ensure_peer_buffers(peer);
    if (rc) return rc;
    CU(cudaSetDevice(s->device));
    rc = ensure_peer_buffers(s);
    if (rc) return rc;
    if (peer->device != s->device) {
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, s->device, peer->device));
        if (!can) return fail(EMME_E_PEER, "emme_peer_attach: no peer access between the two devices");
        cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
        cudaGetLastError();
    }
    for (int w = 0; w < EMME_PEER_BUFS; ++w) s->peer_buf[w][peer_rank] = local_peer_buf(peer, w);
    peer_mapping_added(s, peer_count);
    return 0;
}

int emme_shard_dense(emme_solver* s, int enable) {
    if (!s) return fail(-1, "null handle");
    if (enable) {
        if (s->peer_count <= 1) return fail(EMME_E_STATE, "emme_shard_dense: map the peers first");
        if (!s->use_sym || !s->Y) return fail(EMME_E_STATE, "emme_shard_dense: the symmetric path is switched off");
        const int nbo = emme::dense_sym_outer_block(s->dim);
        if (nbo < 64 || nbo % 64 != 0)
            return fail(EMME_E_STATE, "emme_shard_dense: outer block " + std::to_string(nbo) +
                                          " (single level below dim 2049): set EMME_DENSE_NBO to a multiple of 64");
    }
    s->dense_sharded = enable ? 1 : 0;
    return 0;
}

int emme_peer_set_timeout(double seconds) {
    emme::peer_set_timeout(seconds);
    return 0;
}

int emme_seed_begin(emme_solver* s, double w0r, double w0i) {
    if (!s) return fail(-1, "null handle");
    CU(cudaSetDevice(s->device));
    int rc = ensure_newton_buffers(s);
    if (rc) return rc;
    const zc w0(w0r, w0i);
    s->w = 0.99 * w0;   // include/solver.h:401
    s->dw = 0.01 * w0;  // include/solver.h:402
    s->seeded = false;
    return assemble_current(s);  // eigen_matrix_old = A(0.99 w0), :411 (swapped in seed_middle)
}

int emme_seed_middle(emme_solver* s) {
    if (!s) return fail(-1, "null handle");
    CU(cudaSetDevice(s->device));
    std::swap(s->A, s->Aold);
    s->w += s->dw;  // :412
    return assemble_current(s);  // :413
}

int emme_seed_finish(emme_solver* s) {
    if (!s) return fail(-1, "null handle");
    CU(cudaSetDevice(s->device));
    int rc = secant(s);  // :414
    if (rc) return rc;
    CU(cudaStreamSynchronize(s->stream));
    s->seeded = true;
    return 0;
}

int emme_seed(emme_solver* s, double w0r, double w0i) {
    int rc = emme_seed_begin(s, w0r, w0i);
    if (rc) return rc;
    rc = emme_seed_middle(s);
    if (rc) return rc;
    return emme_seed_finish(s);
}

int emme_step_begin(emme_solver* s) {
    if (!s) return fail(-1, "null handle");
    if (!s->seeded) return fail(EMME_E_STATE, "emme_newton_trace_step before emme_seed");
    CU(cudaSetDevice(s->device));
    zc delta(std::nan(""), std::nan(""));
    const zc dw_prev = s->dw;
    int rc = dense_delta(s, &delta, [&]() -> int {   // include/solver.h:130-139
        // A' = (A - A_old)/delta_prev is rebuilt from the intact A and A_old
        CU(emme::launch_secant(s->A, s->Aold, s->Ad, (size_t)s->dim * s->dim, dw_prev.real(),
                               dw_prev.imag(), s->sms, s->stream));
        ++s->launches;
        return 0;
    });
    // :140 -- the reference updates omega before it checks info; a CUDA failure computed no delta
    if (rc == 0 || (rc > 0 && rc <= s->dim) || rc == EMME_E_NONFINITE) {
        s->dw = delta;
        s->w += delta;
    }
    if (rc) return rc;
    std::swap(s->A, s->Aold);  // eigen_matrix_old = eigen_matrix (:114) without a copy
    return assemble_current(s);  // :157
}

int emme_step_finish(emme_solver* s, double* wr, double* wi, double* dr, double* di) {
    if (!s) return fail(-1, "null handle");
    CU(cudaSetDevice(s->device));
    int rc = secant(s);  // :159
    if (rc) return rc;
    CU(cudaStreamSynchronize(s->stream));
    if (wr) *wr = s->w.real();
    if (wi) *wi = s->w.imag();
    if (dr) *dr = s->dw.real();
    if (di) *di = s->dw.imag();
    return 0;
}

int emme_newton_trace_step(emme_solver* s, double* wr, double* wi, double* dr, double* di) {
    int rc = emme_step_begin(s);
    if (rc) return rc;
    return emme_step_finish(s, wr, wi, dr, di);
}

// newtonQRSecantIteration (include/solver.h:210-383): W <- A, Householder QR with column
// pivoting, null-vector estimate v from R, delta = -R_nn / (Q^H A' v)_n.  A and A' are only read.
static int qr_delta(emme_solver* s, zc* delta) {
    if (!s->d_qr_ws) CU(cudaMalloc(&s->d_qr_ws, emme::qr_workspace_bytes(s->dim)));
    if (!s->d_qr_out) CU(cudaMalloc(&s->d_qr_out, 2 * sizeof(double2)));
    const cudaEvent_t e0 = s->evd0, e1 = s->evd1;
    CU(cudaEventRecord(e0, s->stream));
    CU(cudaMemcpyAsync(s->W, s->A, s->bytes(), cudaMemcpyDeviceToDevice, s->stream));
    // 3*dim + 4 short launches on fixed buffers: captured once, replayed (launch bound otherwise)
    {
        int rc = replay_graph(s, &s->qr_graph, &s->qr_graph_launches, [&](unsigned long long* nl) {
            return emme::launch_qr_step(s->W, s->Ad, s->dim, s->d_qr_ws, s->d_qr_out, s->d_info, s->stream, nl);
        });
        if (rc) return rc;
    }
    CU(cudaEventRecord(e1, s->stream));
    double out[4];
    int info = 0;
    CU(cudaMemcpyAsync(out, s->d_qr_out, sizeof out, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaMemcpyAsync(&info, s->d_info, sizeof info, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    s->stats.dense_ms = ms;
    s->stats.dense_flops = (16.0 / 3.0) * (double)s->dim * s->dim * s->dim;   // Householder QR, complex
    if (info != 0) {
        // the reference's message for ztrtrs info > 0 (include/solver.h:312-314)
        return fail(info, "\xe7\x9f\xa9\xe9\x98\xb5\xe7\xac\xac " + std::to_string(info) +
                              " \xe4\xb8\xaa\xe5\xaf\xb9\xe8\xa7\x92\xe7\xba\xbf\xe5\x85\x83\xe7\xb4\xa0\xe4\xb8\xba\xe9\x9b\xb6\xef\xbc\x8c"
                              "\xe6\x97\xa0\xe6\xb3\x95\xe6\xb1\x82\xe8\xa7\xa3 (R has a zero diagonal element: cannot solve)");
    }
    *delta = -zc(out[0], out[1]) / zc(out[2], out[3]);   // :370
    if (!std::isfinite(delta->real()) || !std::isfinite(delta->imag()))
        return fail(EMME_E_NONFINITE, "QR-secant step is not finite: (Q^H A' v)_n = (" + std::to_string(out[2]) +
                                          ", " + std::to_string(out[3]) + ") (iterating past convergence, or a "
                                          "non-finite matrix)");
    return 0;
}

int emme_qr_step_begin(emme_solver* s) {
    if (!s) return fail(-1, "null handle");
    if (!s->seeded) return fail(EMME_E_STATE, "emme_newton_qr_step before emme_seed");
    CU(cudaSetDevice(s->device));
    zc delta(std::nan(""), std::nan(""));
    int rc = qr_delta(s, &delta);
    if (rc == EMME_E_NONFINITE) {
        s->dw = delta;
        s->w += delta;
    }
    if (rc) return rc;
    s->dw = delta;
    s->w += delta;               // :371
    std::swap(s->A, s->Aold);    // eigen_matrix_old = eigen_matrix (:211) without a copy
    return assemble_current(s);  // :380
}

int emme_newton_qr_step(emme_solver* s, double* wr, double* wi, double* dr, double* di) {
    int rc = emme_qr_step_begin(s);
    if (rc) return rc;
    return emme_step_finish(s, wr, wi, dr, di);   // :382
}

int emme_qr_delta(emme_solver* s, const void* host_A, const void* host_Ad, double* dr, double* di) {
    if (!s) return fail(-1, "null handle");
    if (!host_A) return fail(-2, "null A");
    if (!host_Ad) return fail(-3, "null Ad");
    CU(cudaSetDevice(s->device));
    int rc = ensure_newton_buffers(s);
    if (rc) return rc;
    CU(cudaMemcpyAsync(s->A, host_A, s->bytes(), cudaMemcpyHostToDevice, s->stream));
    CU(cudaMemcpyAsync(s->Ad, host_Ad, s->bytes(), cudaMemcpyHostToDevice, s->stream));
    zc delta;
    rc = qr_delta(s, &delta);
    if (dr) *dr = delta.real();
    if (di) *di = delta.imag();
    return rc;
}

int emme_get_eigen_value(const emme_solver* s, double* wr, double* wi, double* dr, double* di) {
    if (!s) return fail(-1, "null handle");
    if (wr) *wr = s->w.real();
    if (wi) *wi = s->w.imag();
    if (dr) *dr = s->dw.real();
    if (di) *di = s->dw.imag();
    return 0;
}

int emme_trace_delta(emme_solver* s, const void* host_A, const void* host_Ad, double* dr,
                     double* di) {
    if (!s) return fail(-1, "null handle");
    if (!host_A) return fail(-2, "null A");
    if (!host_Ad) return fail(-3, "null Ad");
    CU(cudaSetDevice(s->device));
    int rc = ensure_newton_buffers(s);
    if (rc) return rc;
    CU(cudaMemcpyAsync(s->A, host_A, s->bytes(), cudaMemcpyHostToDevice, s->stream));
    CU(cudaMemcpyAsync(s->Ad, host_Ad, s->bytes(), cudaMemcpyHostToDevice, s->stream));
    zc delta;
    rc = dense_delta(s, &delta, [&]() -> int {
        CU(cudaMemcpyAsync(s->Ad, host_Ad, s->bytes(), cudaMemcpyHostToDevice, s->stream));
        return 0;
    });
    if (dr) *dr = delta.real();
    if (di) *di = delta.imag();
    return rc;
}

// nullSpace (include/solver.h:58-112): the reference takes the right singular vector of the
// smallest singular value from a full SVD (zgesdd) and conjugates it.  Here: inverse iteration
// on A^H A with ONE LU factorisation of A.  A is complex symmetric, so A^H = conj(A) and
//     (A^H A)^-1 v = A^-1 conj(A^-1 conj(v)).
// The vector is defined up to a complex phase; it is normalised to unit 2-norm and rotated so
// that its largest component is real positive.
int emme_null_space(emme_solver* s, void* host_out) {
    if (!s) return fail(-1, "null handle");
    if (!host_out) return fail(-2, "null output");
    CU(cudaSetDevice(s->device));
    int rc = ensure_newton_buffers(s);
    if (rc) return rc;
    const int dim = s->dim;
    // factor a copy of A (W) once; Ad serves as the right-hand-side / solution buffer
    int info = 0, flag = 0;
    for (int attempt = s->optimistic ? 0 : 1; attempt < 2; ++attempt) {
        const int optimistic = attempt == 0;
        CU(cudaMemcpyAsync(s->W, s->A, s->bytes(), cudaMemcpyDeviceToDevice, s->stream));
        CU(emme::launch_trace_solve(s->W, s->Ad, dim, s->d_dense_ws, s->d_trace, s->d_info, s->stream,
                                    &s->launches, optimistic, s->d_flag, 0));
        CU(cudaMemcpyAsync(&info, s->d_info, sizeof info, cudaMemcpyDeviceToHost, s->stream));
        if (optimistic)
            CU(cudaMemcpyAsync(&flag, s->d_flag, sizeof flag, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        s->null_optimistic = optimistic;
        if (!optimistic || flag == 0) break;
    }
    if (info != 0) return fail(info, "nullSpace: matrix is exactly singular");
    // start vector: all ones (deterministic), stored in column 0 of Ad
    std::vector<zc> v(dim, zc(1.0 / std::sqrt((double)dim), 0.0));
    CU(cudaMemcpy2DAsync(s->Ad, sizeof(zc) * dim, v.data(), sizeof(zc), sizeof(zc), dim,
                         cudaMemcpyHostToDevice, s->stream));
    double* d_norm = reinterpret_cast<double*>(s->d_trace);
    double prev = 0.0;
    for (int it = 0; it < 12; ++it) {
        double nrm[2] = {0, 0};
        for (int half = 0; half < 2; ++half) {
            // rhs <- conj(rhs) happens inside the normalise kernel of the previous solve
            CU(emme::launch_solve_factored(s->W, s->Ad, dim, 1, s->d_dense_ws, s->null_optimistic,
                                           s->stream, &s->launches));
            CU(emme::launch_conj_normalise(s->Ad, dim, dim, s->Ad, dim, d_norm, 1, s->stream));
            ++s->launches;
            CU(cudaMemcpyAsync(&nrm[half], d_norm, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        }
        CU(cudaStreamSynchronize(s->stream));
        // after two conj-solves the growth factor estimates 1/sigma_min^2
        const double growth = nrm[0] * nrm[1];
        if (it > 0 && std::fabs(growth - prev) <= 1e-13 * growth) break;
        prev = growth;
    }
    // after an even number of conj-normalise passes column 0 holds conj(v); undo and fix the phase
    CU(cudaMemcpy2DAsync(v.data(), sizeof(zc), s->Ad, sizeof(zc) * dim, sizeof(zc), dim,
                         cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    size_t imax = 0;
    for (size_t i = 0; i < v.size(); ++i) {
        v[i] = std::conj(v[i]);
        if (std::abs(v[i]) > std::abs(v[imax])) imax = i;
    }
    const zc phase = std::conj(v[imax]) / std::abs(v[imax]);
    for (auto& x : v) x *= phase;
    std::memcpy(host_out, v.data(), sizeof(zc) * dim);
    return 0;
}

void* emme_matrix_device_ptr(emme_solver* s, int which) {
    if (!s) return nullptr;
    return which == 0 ? s->A : which == 1 ? s->Aold : which == 2 ? s->Ad : which == 3 ? s->W : which == 4 ? s->Y : nullptr;
}

int emme_copy_matrix(emme_solver* s, int which, void* host_out) {
    if (!s) return fail(-1, "null handle");
    if (which < 0 || which > 4) return fail(-2, "emme_copy_matrix: which must be 0 .. 4");
    if (!host_out) return fail(-3, "null output");
    void* src = emme_matrix_device_ptr(s, which);
    if (!src) return fail(EMME_E_STATE, "matrix not allocated yet (call emme_seed first)");
    CU(cudaSetDevice(s->device));
    CU(cudaMemcpyAsync(host_out, src, s->bytes(), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}

int emme_copy_matrix_async(emme_solver* s, int which, void* pinned_host_out) {
    if (!s) return fail(-1, "null handle");
    // eigen_matrix_derivative (2) is rewritten by the secant and destroyed by the LU paths without
    // waiting for a download; only the two matrices the assemblies guard are offered here
    if (which < 0 || which > 1) return fail(-2, "emme_copy_matrix_async: which must be 0 or 1");
    if (!pinned_host_out) return fail(-3, "null output");
    void* src = emme_matrix_device_ptr(s, which);
    if (!src) return fail(EMME_E_STATE, "matrix not allocated yet (call emme_seed first)");
    CU(cudaSetDevice(s->device));
    if (!s->copy_stream) {
        CU(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&s->ev_copy_src, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&s->ev_copy_done, cudaEventDisableTiming));
    }
    CU(cudaEventRecord(s->ev_copy_src, s->stream));            // everything that produced the matrix
    CU(cudaStreamWaitEvent(s->copy_stream, s->ev_copy_src, 0));
    CU(cudaMemcpyAsync(pinned_host_out, src, s->bytes(), cudaMemcpyDeviceToHost, s->copy_stream));
    CU(cudaEventRecord(s->ev_copy_done, s->copy_stream));
    s->copy_pending = true;
    return 0;
}

int emme_copy_wait(emme_solver* s) {
    if (!s) return fail(-1, "null handle");
    CU(cudaSetDevice(s->device));
    if (s->copy_stream) CU(cudaStreamSynchronize(s->copy_stream));
    return 0;
}

int emme_get_stats(const emme_solver* s, emme_stats* out) {
    if (!s) return fail(-1, "null handle");
    if (!out) return fail(-2, "null output");
    *out = s->stats;
    out->launches = s->launches;
    out->pivot_fallbacks = s->pivot_fallbacks;
    out->sym_steps = s->sym_steps;
    return 0;
}

void* emme_stream(emme_solver* s) { return s ? (void*)s->stream : nullptr; }

int emme_synchronize(emme_solver* s) {
    if (!s) return fail(-1, "null handle");
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}

// ------------------------------------------------------------------ host-side input helpers
struct emme_input {
    emme::json::Value json;  // scan objects already collapsed to their head
};

static int finish_input(emme::json::Value all, emme_input** out) {
    // filter_input (src/main.cpp:174-180)
    emme::json::Value in = all.clone();
    for (auto& [key, val] : in.as_object()) {
        if (val.is_object() && val.contains("head")) {
            emme::json::Value head = val.at("head");
            val = head;
        }
    }
    *out = new emme_input{std::move(in)};
    return 0;
}

int emme_input_load(const char* path, emme_input** out) {
    if (!path) return fail(-1, "null path");
    if (!out) return fail(-2, "null out");
    try {
        return finish_input(emme::json::parse_file(path), out);
    } catch (const std::exception& e) {
        return fail(EMME_E_INPUT, e.what());
    }
}

int emme_input_parse(const char* text, emme_input** out) {
    if (!text) return fail(-1, "null text");
    if (!out) return fail(-2, "null out");
    try {
        return finish_input(emme::json::parse(std::string(text)), out);
    } catch (const std::exception& e) {
        return fail(EMME_E_INPUT, e.what());
    }
}

void emme_input_free(emme_input* in) { delete in; }

int emme_input_set_number(emme_input* in, const char* key, double value) {
    if (!in) return fail(-1, "null input");
    if (!key) return fail(-2, "null key");
    in->json[std::string(key)] = emme::json::Value(value);
    return 0;
}

int emme_input_get_number(const emme_input* in, const char* key, double* value) {
    if (!in) return fail(-1, "null input");
    if (!key) return fail(-2, "null key");
    try {
        // "name[k]" addresses element k of an array value, e.g. initial_guess[0]
        std::string k(key);
        double v;
        const size_t lb = k.find('[');
        if (lb != std::string::npos && k.back() == ']') {
            const size_t idx = (size_t)std::atoi(k.substr(lb + 1, k.size() - lb - 2).c_str());
            v = in->json.at(k.substr(0, lb)).at(idx).number();
        } else {
            v = in->json.at(k).number();
        }
        if (value) *value = v;
        return 0;
    } catch (const std::exception& e) {
        return fail(EMME_E_INPUT, e.what());
    }
}

int emme_input_get_string(const emme_input* in, const char* key, char* buf, int buflen) {
    if (!in) return fail(-1, "null input");
    if (!key) return fail(-2, "null key");
    try {
        const std::string& v = in->json.at(key).as_string();
        if (buf && buflen > 0) {
            std::strncpy(buf, v.c_str(), buflen - 1);
            buf[buflen - 1] = 0;
        }
        return 0;
    } catch (const std::exception& e) {
        return fail(EMME_E_INPUT, e.what());
    }
}

int emme_input_params(const emme_input* in, emme_params* p, int* npoints) {
    if (!in) return fail(-1, "null input");
    try {
        auto para = emme::Parameters::generate(in->json);
        if (p) *p = para->to_pod();
        if (npoints) *npoints = para->npoints;
        return 0;
    } catch (const std::exception& e) {
        return fail(EMME_E_INPUT, e.what());
    }
}

int emme_input_pic_params(const emme_input* in, emme_pic_params* p, long* marker_per_cell,
                          long* step_number, double* time_step) {
    if (!in) return fail(-1, "null input");
    try {
        auto para = emme::Parameters::generate(in->json);
        if (p) {
            p->q = para->q; p->R = para->R; p->vt = para->vt; p->tau = para->tau;
            p->shat = para->shat; p->b_theta = para->b_theta; p->length = para->length;
            p->eta_i = para->eta_i; p->omega_s_i = para->omega_s_i; p->omega_d_bar = para->omega_d_bar;
            p->water_bag_weight_vpara = para->water_bag_weight_vpara;
            p->water_bag_weight_vperp = para->water_bag_weight_vperp;
            p->npoints = para->npoints;
            p->drift_center_transformation_switch = para->drift_center_transformation_switch ? 1 : 0;
        }
        // src/main.cpp:89,93-94
        if (marker_per_cell) *marker_per_cell = (long)in->json.at("marker_per_cell").number();
        if (step_number) *step_number = (long)in->json.at("step_number").number();
        if (time_step) *time_step = in->json.at("time_step").number();
        return 0;
    } catch (const std::exception& e) {
        return fail(EMME_E_INPUT, e.what());
    }
}

int emme_input_tables(const emme_input* in, double* eta, double* g, double* bi) {
    if (!in) return fail(-1, "null input");
    try {
        auto para = emme::Parameters::generate(in->json);
        std::vector<double> e, gg, b;
        para->tables(e, gg, b);
        const size_t nb = sizeof(double) * e.size();
        if (eta) std::memcpy(eta, e.data(), nb);
        if (g) std::memcpy(g, gg.data(), nb);
        if (bi) std::memcpy(bi, b.data(), nb);
        return 0;
    } catch (const std::exception& e) {
        return fail(EMME_E_INPUT, e.what());
    }
}

}  // extern "C"
