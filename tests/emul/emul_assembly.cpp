// tests/emul/emul_assembly.cpp -- TEST INFRASTRUCTURE: CPU emulation of kernel 1's group logic.
//
// Compiles the SAME per-node arithmetic the CUDA kernel uses (emme_b200/csrc/emme_eval.cuh,
// run_const.h, gk_tables.h) with g++ and replays, lane by lane, what one group of
// assemble_kernel does for each work item (emme_b200/csrc/assembly.cu): node placement,
// reference-order panel sums, accept/bisect rule, LIFO stack, scatter.  It lets the CPU test
// suite check the kernel's algebra (hoisted constants, the 2*lambda and conj(e)/t identities,
// diagonal-major item decoding) against the oracle without a GPU.  It uses glibc's libm, not
// CUDA's, so it says nothing about device rounding; the -m gpu tests do that.
// Never linked into libemme_b200.so.
#define _GNU_SOURCE 1
#include <cmath>
#include <cstddef>
#include <vector>

#include "../../emme_b200/csrc/emme_eval.cuh"
#include "../../emme_b200/csrc/gk_tables.h"
#include "../../emme_b200/csrc/run_const.h"

using namespace emme;

static const GKTables T15 = EMME_GK15_INIT;
static const GKTables T31 = EMME_GK31_INIT;

static void decode_item(unsigned long long kg, int N, int nm, int& i, int& j, int& m) {
    const unsigned long long p = kg / (unsigned)nm;
    m = (int)(kg - p * (unsigned)nm);
    const double tn = 2.0 * N + 1.0;
    double disc = tn * tn - 8.0 * ((double)N + (double)p);
    if (disc < 0.) disc = 0.;
    long long d = (long long)std::floor((tn - std::sqrt(disc)) * 0.5);
    if (d < 1) d = 1;
    if (d > N - 1) d = N - 1;
    while (d > 1 && (unsigned long long)((d - 1) * (2LL * N - d) / 2) > p) --d;
    while (d < N - 1 && (unsigned long long)(d * (2LL * N - d - 1) / 2) <= p) ++d;
    const unsigned long long base = (unsigned long long)((d - 1) * (2LL * N - d) / 2);
    i = (int)(p - base);
    j = i + (int)d;
}

extern "C" int emul_assemble(const emme_params* p, int N, const double* eta, const double* g,
                             const double* bi, double wr, double wi, double* out,
                             unsigned long long* stats) {
    RunConst rc = make_run_const(*p, N, wr, wi);
    if (rc.order != 15 && rc.order != 31) return 1;
    const GKTables& T = rc.order == 15 ? T15 : T31;
    const int H = (rc.order - 1) / 2;
    const size_t dim = rc.em ? 2 * (size_t)N : (size_t)N;
    const int nm = rc.em ? 3 : 1;
    auto put = [&](size_t r, size_t c, cplx v) {
        out[2 * (r * dim + c)] = v.re;
        out[2 * (r * dim + c) + 1] = v.im;
    };
    for (int i = 0; i < N; ++i) {
        put(i, i, mk(rc.diag_es, 0.));
        if (rc.em) {
            put(i, i + N, mk(0., 0.));
            put(i + N, i, mk(0., 0.));
            put(i + N, i + N, mk(rc.diag_em * bi[i], 0.));
        }
    }
    const unsigned long long n_items = (unsigned long long)N * (N - 1) / 2 * nm;
    unsigned long long n_eval = 0, n_panel = 0, n_fwd = 0, n_bwd = 0, max_top = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : n_eval, n_panel, n_fwd, n_bwd) reduction(max : max_top)
    for (unsigned long long k = 0; k < n_items; ++k) {
        int it_i, it_j, it_m;
        decode_item(k, N, nm, it_i, it_j, it_m);
        PairConst pc = make_pair(rc, eta[it_i], eta[it_j], g[it_i], g[it_j], bi[it_i], bi[it_j]);
        std::vector<double> stk;
        stk.push_back(0.0);
        stk.push_back(rc.half_pi);
        cplx sum = mk(0., 0.);
        double abs_tol = 0.;
        EvalCounters cnt{0u, 0u};
        while (!stk.empty()) {
            const double r = stk.back();
            stk.pop_back();
            const double l = stk.back();
            stk.pop_back();
            const double mid = (r + l) / 2, scale = (r - l) / 2;
            cplx fx[31];
            for (int gl = 0; gl <= 2 * H; ++gl) {
                const int nidx = gl == 0 ? 0 : (gl <= H ? gl : gl - H);
                const double node = gl <= H ? T.a[nidx] : -T.a[nidx];
                volatile double prod = scale * node;  // no FMA, as __dmul_rn/__dadd_rn
                const double x = prod + mid;
                fx[gl] = eval_node(rc, pc, it_m, node_const(rc, x), cnt);
                ++n_eval;
            }
            cplx K = mk(T.kw[0] * fx[0].re, T.kw[0] * fx[0].im);
            cplx G = mk(T.gw[0] * fx[0].re, T.gw[0] * fx[0].im);
            for (int n = 1; n <= H; ++n) {
                const cplx f = mk(fx[n].re + fx[n + H].re, fx[n].im + fx[n + H].im);
                K.re += T.kw[n] * f.re;
                K.im += T.kw[n] * f.im;
                if ((n & 1) == 0) {
                    G.re += T.gw[n] * f.re;
                    G.im += T.gw[n] * f.im;
                }
            }
            ++n_panel;
            const cplx integral = mk(K.re * scale, K.im * scale);
            const double e0 = std::fmax(std::hypot(K.re - G.re, K.im - G.im),
                                        std::hypot(K.re, K.im) * 2.220446049250313e-16 * 2);
            const double err = e0 * scale;
            const double rel = std::hypot(rc.tol * integral.re, rc.tol * integral.im);
            if (abs_tol == 0.) abs_tol = rel;
            const bool split = std::ldexp(scale, rc.maxdepth) > rc.thr_len &&
                               err > abs_tol * rc.inv_scale + rc.prec && err > rel + rc.prec;
            if (split) {
                stk.push_back(mid);
                stk.push_back(r);
                stk.push_back(l);
                stk.push_back(mid);
                if (stk.size() / 2 > max_top) max_top = stk.size() / 2;
            } else {
                sum = sum + integral;
            }
        }
        n_fwd += cnt.fwd;
        n_bwd += cnt.bwd;
        cplx kap = mk(rc.kappa_pref * sum.im, -rc.kappa_pref * sum.re);
        if (it_m > 0) kap = kap + kappa_e(rc, it_m, pc.deta, g[it_i] - g[it_j]);
        const int i = it_i, j = it_j;
        if (it_m == 0) {
            const double w = sing_weight(N, i, j);
            cplx a = -kap;
            a = mk(a.re * w, a.im * w);
            a = mk(a.re * rc.dx, a.im * rc.dx);
            put(i, j, a);
            put(j, i, a);
        } else if (it_m == 1) {
            const cplx a = mk(kap.re * rc.dx, kap.im * rc.dx);
            put(i, j + N, a);
            put(j, i + N, -a);
            put(i + N, j, -a);
            put(j + N, i, a);
        } else {
            const cplx a = mk(kap.re * rc.dx, kap.im * rc.dx);
            put(i + N, j + N, a);
            put(j + N, i + N, a);
        }
    }
    if (stats) {
        stats[0] = n_items;
        stats[1] = n_panel;
        stats[2] = n_eval;
        stats[3] = n_fwd;
        stats[4] = n_bwd;
        stats[5] = max_top;
    }
    return 0;
}

// exp(a + i b) through the kernel's own cexp_lean (emme_eval.cuh), for the ulp test
extern "C" void emul_cexp(const double* a, const double* b, int n, double* out) {
    for (int i = 0; i < n; ++i) {
        const cplx v = cexp_lean(a[i], b[i]);
        out[2 * i] = v.re;
        out[2 * i + 1] = v.im;
    }
}
