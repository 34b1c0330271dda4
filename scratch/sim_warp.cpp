// scratch/sim_warp.cpp -- CPU model of kernel 1's warp schedule (analysis tool, never shipped).
//
// Runs the SAME per-node arithmetic and accept/bisect rule as assemble_kernel (emme_eval.cuh compiled
// for the host, like tests/emul/emul_assembly.cpp) for blocks of 32*K items in lockstep, the way a warp
// does: every round each busy lane evaluates node j of its current panel; the round costs what the
// slowest lane costs (SIMT loops run to the largest trip count, the underflow skip only pays when every
// busy lane takes it).  Issue-slot weights from DESIGN.md section 3 (r1j profile).  Output: lanes per
// instruction, block times, and the makespan of a greedy assignment to the resident warps.
//
// g++ -O2 -std=c++17 -fopenmp -ffp-contract=off -shared -fPIC -o /tmp/libsimwarp.so scratch/sim_warp.cpp
#define _GNU_SOURCE 1
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <queue>
#include <vector>

#include "../emme_b200/csrc/emme_eval.cuh"
#include "../emme_b200/csrc/gk_tables.h"
#include "../emme_b200/csrc/run_const.h"
#include "../include/emme_b200.h"

using namespace emme;
static const GKTables T15 = EMME_GK15_INIT;
static const GKTables T31 = EMME_GK31_INIT;

struct Lane {
    PairConst pc;
    std::vector<double> stk;
    cplx sum, K, G, fplus;
    double abs_tol, l, r;
    int m;
    bool busy;
    int next;   // index of this lane's next item inside the block
};

// slot weights (warp-instruction issue slots, FP64 counted twice)
static double W_UNDER = 90., W_FULL = 540., W_FWD = 9., W_BWD = 20., W_PANEL = 490.;

extern "C" void sim_set_weights(double u, double f, double fw, double bw, double pn) {
    W_UNDER = u; W_FULL = f; W_FWD = fw; W_BWD = bw; W_PANEL = pn;
}

// items: n triples (i, j, m) in schedule order.  K: cohorts per lane-private queue (1 = shipped kernel).
// out[0] = useful lane-slots, out[1] = 32 * warp slots, out[2] = makespan, out[3] = balanced time,
// out[4] = evals (check against the kernel's counter)
extern "C" int sim_schedule(const emme_params* p, int N, const double* eta, const double* g,
                            const double* bi, double wr, double wi, const int* items, long n, int K,
                            int warps, double* out, double* blk_time_out) {
    RunConst rc = make_run_const(*p, N, wr, wi);
    const GKTables& T = rc.order == 15 ? T15 : T31;
    const int H = (rc.order - 1) / 2;
    const long B = 32L * K;
    const long nb = (n + B - 1) / B;
    std::vector<double> blk_time(nb), blk_useful(nb);
    unsigned long long n_eval = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : n_eval)
    for (long b = 0; b < nb; ++b) {
        const long lo = b * B, hi = std::min(n, lo + B);
        Lane ln[32];
        auto start = [&](Lane& L, long k) {
            const int i = items[3 * k], j = items[3 * k + 1];
            L.m = items[3 * k + 2];
            L.pc = make_pair(rc, eta[i], eta[j], g[i], g[j], bi[i], bi[j]);
            L.stk.clear();
            L.sum = mk(0., 0.);
            L.abs_tol = 0.;
            L.l = 0.;
            L.r = rc.half_pi;
            L.busy = true;
        };
        for (int l = 0; l < 32; ++l) {
            ln[l].busy = false;
            ln[l].next = l;
            if (lo + l < hi) {
                start(ln[l], lo + l);
                ln[l].next = l + 32;
            }
        }
        double t_warp = 0., t_use = 0.;
        for (;;) {
            bool any = false;
            for (int l = 0; l < 32; ++l) any = any || ln[l].busy;
            if (!any) break;
            // one panel per busy lane, nodes in lockstep (same order as the kernel: centre, +a, -a, ...)
            for (int j = 0; j <= 2 * H; ++j) {
                const int ni = (j + 1) >> 1;
                double mx_f = 0., mx_b = 0.;
                bool any_full = false, any_under = false;
                for (int l = 0; l < 32; ++l) {
                    Lane& L = ln[l];
                    if (!L.busy) continue;
                    const double node = (j & 1) ? T.a[ni] : -T.a[ni];
                    const double mid = (L.r + L.l) / 2, scale = (L.r - L.l) / 2;
                    volatile double prod = scale * node;
                    const double x = prod + mid;
                    EvalCounters c{0u, 0u};
                    const cplx fx = eval_node(rc, L.pc, L.m, node_const(rc, x), c);
                    ++n_eval;
                    const bool full = (c.fwd + c.bwd) != 0;
                    any_full = any_full || full;
                    any_under = any_under || !full;
                    mx_f = std::max(mx_f, (double)c.fwd);
                    mx_b = std::max(mx_b, (double)c.bwd);
                    t_use += full ? W_FULL + W_FWD * c.fwd + W_BWD * c.bwd : W_UNDER;
                    if (j == 0) {
                        L.K = mk(T.kw[0] * fx.re, T.kw[0] * fx.im);
                        L.G = mk(T.gw[0] * fx.re, T.gw[0] * fx.im);
                    } else if (j & 1) {
                        L.fplus = fx;
                    } else {
                        const cplx f = mk(L.fplus.re + fx.re, L.fplus.im + fx.im);
                        if ((ni & 1) == 0) {
                            L.G.re += T.gw[ni] * f.re;
                            L.G.im += T.gw[ni] * f.im;
                        }
                        L.K.re += T.kw[ni] * f.re;
                        L.K.im += T.kw[ni] * f.im;
                    }
                }
                t_warp += any_full ? W_FULL + W_FWD * mx_f + W_BWD * mx_b : W_UNDER;
            }
            t_warp += W_PANEL;
            for (int l = 0; l < 32; ++l) {
                Lane& L = ln[l];
                if (!L.busy) continue;
                t_use += W_PANEL;
                const double mid = (L.r + L.l) / 2, scale = (L.r - L.l) / 2;
                const cplx integral = mk(L.K.re * scale, L.K.im * scale);
                const double e0 = std::fmax(std::hypot(L.K.re - L.G.re, L.K.im - L.G.im),
                                            std::hypot(L.K.re, L.K.im) * 2.220446049250313e-16 * 2);
                const double err = e0 * scale;
                const double rel = std::hypot(rc.tol * integral.re, rc.tol * integral.im);
                if (L.abs_tol == 0.) L.abs_tol = rel;
                const bool split = std::ldexp(scale, rc.maxdepth) > rc.thr_len &&
                                   err > L.abs_tol * rc.inv_scale + rc.prec && err > rel + rc.prec;
                if (split) {
                    L.stk.push_back(mid);
                    L.stk.push_back(L.r);
                    L.r = mid;
                } else {
                    L.sum = L.sum + integral;
                    if (!L.stk.empty()) {
                        L.r = L.stk.back();
                        L.stk.pop_back();
                        L.l = L.stk.back();
                        L.stk.pop_back();
                    } else {
                        L.busy = false;
                        if (lo + L.next < hi) {       // lane-private queue: next cohort's item of this lane
                            start(L, lo + L.next);
                            L.next += 32;
                        }
                    }
                }
            }
        }
        blk_time[b] = t_warp;
        blk_useful[b] = t_use;
    }
    double tot = 0., use = 0.;
    for (long b = 0; b < nb; ++b) {
        tot += blk_time[b];
        use += blk_useful[b];
        if (blk_time_out) blk_time_out[b] = blk_time[b];
    }
    std::priority_queue<double, std::vector<double>, std::greater<double>> h;
    for (int w = 0; w < warps; ++w) h.push(0.);
    double makespan = 0.;
    for (long b = 0; b < nb; ++b) {
        const double t = h.top() + blk_time[b];
        h.pop();
        h.push(t);
        makespan = std::max(makespan, t);
    }
    out[0] = use;
    out[1] = 32. * tot;
    out[2] = makespan;
    out[3] = tot / warps;
    out[4] = (double)n_eval;
    return 0;
}
