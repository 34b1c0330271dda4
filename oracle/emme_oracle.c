/* oracle/emme_oracle.c -- plain-C CPU restatement of the EMME eigen hot path.
 *
 * TEST INFRASTRUCTURE (see emme_oracle.h).  Every function cites the reference
 * file:line it follows.  The arithmetic is sequenced exactly as the reference's
 * C++ evaluates it (libstdc++ std::complex scalar/complex overloads are
 * component-wise; complex*complex and anything/complex go through the same
 * libgcc __muldc3/__divdc3 that GCC emits for C99 _Complex), so that with the same
 * compiler and glibc the results are bit-identical to oracle/_ref (tests/test_oracle.py).
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fopenmp (oracle/Makefile).
 */
#define _GNU_SOURCE
#include "emme_oracle.h"

#include <complex.h>
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cplx;

/* component-wise helpers = libstdc++ operator*(T, complex<T>) etc. (bits/complex) */
static inline cplx sc_mul(double s, cplx z) { return CMPLX(s * creal(z), s * cimag(z)); }
static inline cplx c_divs(cplx z, double s) { return CMPLX(creal(z) / s, cimag(z) / s); }
static inline cplx s_divc(double s, cplx z) { return CMPLX(s, 0.0) / z; } /* T / complex<T> */
static inline cplx s_addc(double s, cplx z) { return CMPLX(s + creal(z), cimag(z)); }
static inline cplx s_subc(double s, cplx z) { return CMPLX(-creal(z) + s, -cimag(z)); }
static inline cplx c_subs(cplx z, double s) { return CMPLX(creal(z) - s, cimag(z)); }

/* ---- Gauss-Kronrod tables, include/functions.h:92-122 (15) and :125-162 (31) ---- */
static const double A15[8] = {0.,
                              0.20778495500789847,
                              0.40584515137739717,
                              0.58608723546769113,
                              0.74153118559939444,
                              0.86486442335976907,
                              0.94910791234275852,
                              0.99145537112081264};
static const double G15[4] = {0.41795918367346939, 0.38183005050511894, 0.27970539148927667,
                              0.12948496616886969};
static const double K15[8] = {2.09482141084727828e-01, 2.04432940075298892e-01,
                              1.90350578064785410e-01, 1.69004726639267903e-01,
                              1.40653259715525919e-01, 1.04790010322250184e-01,
                              6.30920926299785533e-02, 2.29353220105292250e-02};
static const double A31[16] = {0.0,
                               0.1011420669187175,
                               0.20119409399743452,
                               0.29918000715316881,
                               0.39415134707756337,
                               0.48508186364023968,
                               0.57097217260853885,
                               0.65099674129741697,
                               0.72441773136017005,
                               0.79041850144246593,
                               0.84820658341042722,
                               0.8972645323440819,
                               0.9372733924007059,
                               0.96773907567913913,
                               0.98799251802048543,
                               0.99800229869339706};
static const double G31[8] = {0.20257824192556112, 0.19843148532711152, 0.18616100001556193,
                              0.1662692058169939,  0.1395706779261542,  0.10715922046717143,
                              0.07036604748810768, 0.030753241996119};
static const double K31[16] = {0.10133000701479155,   0.100769845523875595,  0.099173598721791959,
                               0.0966427269836236785, 0.093126598170825321,  0.0885644430562117706,
                               0.083080502823133021,  0.0768496807577203789, 0.069854121318728259,
                               0.0620095678006706403, 0.053481524690928087,  0.0445897513247648766,
                               0.035346360791375846,  0.0254608473267153202, 0.0150079473293161225,
                               0.00537747987292334899};

typedef struct {
    emme_oracle_fn f;
    void* ctx;
    long evals;
} wrapped_fn;

/* the [0,inf) -> [0,pi/2] wrapper lambda, include/functions.h:315-318 */
static cplx eval_wrapped(wrapped_fn* w, double x) {
    const double c = cos(x);
    double re, im;
    w->f(tan(x), w->ctx, &re, &im);
    w->evals++;
    return c_divs(CMPLX(re, im), c * c);
}

/* gauss_kronrod_basic on the panel (mid, scale), include/functions.h:181-209 with the
 * normalize_func of :232.  Returns the un-scaled Kronrod sum and error estimate. */
static cplx gk_basic(wrapped_fn* w, int order, double mid, double scale, double* err) {
    const int nabs = order == 15 ? 8 : 16;
    const double* A = order == 15 ? A15 : A31;
    const double* G = order == 15 ? G15 : G31;
    const double* K = order == 15 ? K15 : K31;
    const int gauss_order = (order - 1) / 2;
    cplx f0 = eval_wrapped(w, scale * 0.0 + mid);
    cplx gi = (gauss_order & 1) ? sc_mul(G[0], f0) : CMPLX(0.0, 0.0);
    cplx ki = sc_mul(K[0], f0);
    for (int i = 1; i < nabs; ++i) {
        cplx fp = eval_wrapped(w, scale * A[i] + mid);
        cplx fm = eval_wrapped(w, scale * (-A[i]) + mid);
        cplx f = fp + fm;
        gi += ((gauss_order - i) & 1) ? sc_mul(G[i / 2], f) : CMPLX(0.0, 0.0);
        ki += sc_mul(K[i], f);
    }
    *err = fmax(cabs(ki - gi), cabs(ki) * DBL_EPSILON * 2);
    return ki;
}

/* gauss_kronrod_adaptive, include/functions.h:211-251, called as :322-327 with
 * a = 0, b = pi/2, abs_tol = 0. */
int emme_oracle_integrate(emme_oracle_fn f, void* ctx, double tol, double prec, int maxdepth,
                          int order, double* re, double* im, long* evals) {
    if (order != 15 && order != 31) return 1;
    wrapped_fn w = {f, ctx, 0};
    const double a = 0.0, b = M_PI / 2.0; /* std::numbers::pi / 2.0 */
    size_t cap = 64, top = 0;
    double(*stack)[2] = malloc(cap * sizeof *stack);
    cplx sum = CMPLX(0.0, 0.0);
    double abs_tol = 0.0;
    const double inv_scale = 2. / (b - a);
    stack[top][0] = a;
    stack[top][1] = b;
    ++top;
    while (top) {
        --top;
        const double l = stack[top][0], r = stack[top][1];
        const double mid = (r + l) / 2;
        const double scale = (r - l) / 2;
        double e;
        cplx k = gk_basic(&w, order, mid, scale, &e);
        cplx integral = CMPLX(creal(k) * scale, cimag(k) * scale); /* complex * double */
        double err = e * scale;
        if (fpclassify(abs_tol) == FP_ZERO) { abs_tol = cabs(sc_mul(tol, integral)); }
        if (ldexp(scale, maxdepth) > 0.99 * (b - a) && err > abs_tol * inv_scale + prec &&
            err > cabs(sc_mul(tol, integral)) + prec) {
            if (top + 2 > cap) {
                cap *= 2;
                stack = realloc(stack, cap * sizeof *stack);
            }
            stack[top][0] = mid;
            stack[top][1] = r;
            ++top;
            stack[top][0] = l;
            stack[top][1] = mid;
            ++top;
        } else {
            sum += integral;
        }
    }
    free(stack);
    *re = creal(sum);
    *im = cimag(sum);
    if (evals) *evals = w.evals;
    return 0;
}

/* util::bessel_i_alter_helper, include/functions.h:381-408 */
static void bessel_i_alter(cplx z, cplx out[4], int* trips) {
    const double THRESHOLD = 2.e+7;
    int n = (int)(floor(cabs(z)) + 1);
    cplx p0 = 0., p1 = 1., p_tmp;
    int fwd = 0, bwd = 0;
    double test_1 =
        fmax(sqrt(THRESHOLD * cabs(p1) * cabs(p0 - s_divc(2.0 * n, z) * p1)), THRESHOLD);
    for (; cabs(p1) <= test_1; ++n) {
        p_tmp = p0 - s_divc(2.0 * n, z) * p1;
        p0 = p1;
        p1 = p_tmp;
        ++fwd;
    }
    cplx y0 = s_divc(1.0, p1), y1 = 0., y_tmp;
    cplx mu = 0.;
    for (n--; n > 0; --n) {
        y_tmp = s_divc(2. * n, z) * y0 + y1;
        y1 = y0;
        y0 = y_tmp;
        mu += sc_mul(2. * (creal(z) < 0 ? 1 - 2 * (n & 1) : 1), y1);
        ++bwd;
    }
    out[0] = y0;
    out[1] = y1;
    out[2] = mu + y0;
    out[3] = creal(z) < 0 ? z : -z;
    if (trips) {
        trips[0] = fwd;
        trips[1] = bwd;
    }
}

void emme_oracle_bessel_i_alter(double zr, double zi, double* out8, int* trips) {
    cplx o[4];
    bessel_i_alter(CMPLX(zr, zi), o, trips);
    for (int k = 0; k < 4; ++k) {
        out8[2 * k] = creal(o[k]);
        out8[2 * k + 1] = cimag(o[k]);
    }
}

/* std::pow(const complex<double>&, const double&), libstdc++ <complex>:
 * real positive base -> pow(); else polar(exp(y*log|x|), y*arg x) via clog. */
static cplx cpow_real_exp(cplx x, double y) {
    if (cimag(x) == 0.0 && creal(x) > 0.0) return CMPLX(pow(creal(x), y), 0.0);
    cplx t = clog(x);
    double rho = exp(y * creal(t)), theta = y * cimag(t);
    return CMPLX(rho * cos(theta), rho * sin(theta));
}

typedef struct {
    const emme_oracle_params* p;
    unsigned m;
    double eta, eta_p, g, gp, b, bp;
    cplx omega;
    long fwd, bwd;
} integrand_ctx;

/* Parameters::beta_1, src/Parameters.cpp:87-90 (g_integration_f values are inputs) */
static double beta_1(const emme_oracle_params* p, double g, double gp) {
    return (p->q * p->R) / p->vt * (p->omega_d_bar) * (g - gp);
}
/* Parameters::beta_1_e, src/Parameters.cpp:92-95 */
static double beta_1_e(const emme_oracle_params* p, double g, double gp) {
    return (p->q * p->R) / p->vt * (p->omega_d_bar * p->omega_s_e / p->omega_s_i) * (g - gp);
}

/* the integrand lambda, src/Parameters.cpp:120-176 */
static void integrand(double t, void* vctx, double* re, double* im) {
    integrand_ctx* c = vctx;
    const emme_oracle_params* p = c->p;
    const double q = p->q, R = p->R, vt = p->vt, arc = p->arc_coeff;
    const double eta = c->eta, eta_p = c->eta_p;
    const cplx omega = c->omega;

    const double omi = -copysign(1, creal(omega));                       /* :121 */
    const cplx exp_arg = cexp(sc_mul(atan(t / arc), CMPLX(-omi * 0.0, -omi * 1.0))); /* :122-123 */
    const cplx taut = sc_mul(t, exp_arg);                                /* :124 */

    /* :126-129  1.i * exp_arg is a complex*complex product */
    const cplx ie = CMPLX(0.0, 1.0) * exp_arg;
    const cplx num = sc_mul(t, sc_mul(omi, ie));
    const cplx jacob = exp_arg - c_divs(num, arc * (1.0 + pow((t / arc), 2)));

    /* lambda_f_tau, :101-106 */
    const double beta_1_val = beta_1(p, c->g, c->gp);
    cplx lam = sc_mul(0.5, CMPLX(0.0, 1.0)) * sc_mul(vt, taut);
    lam = c_divs(lam, q * R * (eta - eta_p));
    lam = sc_mul(beta_1_val, lam);
    const cplx lambda = s_addc(1.0, lam);
    const double bi_eta = c->b, bi_eta_p = c->bp;                        /* :132-133 */

    cplx bes[4];
    int trips[2];
    bessel_i_alter(s_divc(sqrt(bi_eta * bi_eta_p), lambda), bes, trips); /* :135-136 */
    c->fwd += trips[0];
    c->bwd += trips[1];
    const cplx y0 = bes[0], y1 = bes[1], mu = bes[2], z = bes[3];

    const cplx lambda_cubic_inv = cpow_real_exp(lambda, -3.);            /* :138-139 */
    const cplx norm_vel = s_divc(q * R * (eta - eta_p), sc_mul(vt, taut)); /* :140 */

    /* :142-147 */
    cplx nv2 = sc_mul(0.5, norm_vel) * norm_vel;
    cplx inner = s_addc(1.0, sc_mul(p->eta_i, c_subs(nv2, 1.5)));
    cplx i0_coef = (omega - sc_mul(p->omega_s_i, inner)) / lambda +
                   sc_mul(p->omega_s_i * p->eta_i, s_subc(.5 * (bi_eta + bi_eta_p), lambda)) *
                       lambda_cubic_inv;
    /* :149-151 */
    cplx i1_coef = sc_mul(-p->omega_s_i * p->eta_i * sqrt(bi_eta * bi_eta_p), lambda_cubic_inv);

    /* :157-164 */
    const cplx log_norm_vel = sc_mul(-0.5, norm_vel) * norm_vel;
    const cplx log_i_beta = sc_mul(beta_1_val, CMPLX(-0.0, -.5)) * norm_vel;
    const cplx log_hf_tau = (CMPLX(0.0, 1.0) * taut) * omega;
    const cplx log_exp_term =
        s_divc(-(bi_eta + bi_eta_p), s_addc(2.0, sc_mul(beta_1_val, CMPLX(0.0, 1.0)) / norm_vel));
    const cplx log_coef = log_norm_vel + log_i_beta + log_hf_tau + log_exp_term;

    /* :167-173 */
    cplx arg = log_coef - z;
    cplx se = creal(arg) < -40. ? CMPLX(0., 0.) : cexp(arg);
    /* :174-175 */
    cplx res = cpow_real_exp(norm_vel, (double)c->m) / taut * jacob * se *
               (i0_coef * y0 + i1_coef * y1) / mu;
    *re = creal(res);
    *im = cimag(res);
}

void emme_oracle_kappa(const emme_oracle_params* p, unsigned m, double eta, double eta_p,
                       double g, double gp, double b, double bp, double wr, double wi,
                       double* re, double* im, long* stats) {
    integrand_ctx c = {p, m, eta, eta_p, g, gp, b, bp, CMPLX(wr, wi), 0, 0};
    double rr = 0, ri = 0;
    long evals = 0;
    emme_oracle_integrate(integrand, &c, p->tol, p->prec, p->maxdepth, p->order, &rr, &ri, &evals);
    /* src/Parameters.cpp:182-183 */
    cplx k = c_divs(sc_mul(p->q * p->R, CMPLX(-0.0, -1.0)), p->vt * sqrt((2.0 * M_PI))) *
             CMPLX(rr, ri);
    *re = creal(k);
    *im = cimag(k);
    if (stats) {
        stats[0] += evals;
        stats[1] += c.fwd;
        stats[2] += c.bwd;
    }
}

void emme_oracle_kappa_e(const emme_oracle_params* p, unsigned m, double eta, double eta_p,
                         double g, double gp, double wr, double wi, double* re, double* im) {
    const double q = p->q, R = p->R, vt = p->vt, tau = p->tau;
    const cplx omega = CMPLX(wr, wi);
    cplx k = CMPLX(0.0, 0.0);
    if (m == 1) { /* src/Parameters.cpp:195-198 */
        k = c_divs(sc_mul(q * R, CMPLX(-0.0, -1.0)), 2.0 * vt * tau) * c_subs(omega, p->omega_s_e);
        k = c_divs(sc_mul(eta - eta_p, k), fabs(eta - eta_p));
    } else if (m == 2) { /* :199-204 */
        double pref = (q * q * R * R) / (2.0 * vt * vt * tau) * (eta - eta_p) / (fabs(eta - eta_p));
        cplx t1 = sc_mul(eta - eta_p, omega * c_subs(omega, p->omega_s_e));
        cplx t2 = sc_mul(beta_1_e(p, g, gp) * vt / (q * R),
                         c_subs(omega, p->omega_s_e * (1.0 + p->eta_e)));
        k = sc_mul(pref, t1 - t2);
    }
    *re = creal(k);
    *im = cimag(k);
}

double emme_oracle_weight(int n, int i, int j) {
    static const double coeff[6] = {0.0,
                                    2.951388888888883,
                                    -2.4305555555555305,
                                    4.166666666667441,
                                    -0.3472222222224549,
                                    1.159722222222284};
    int diff = abs(i - j);
    double w = diff <= 5 ? coeff[diff] : 1.0;
    if (j == 0 || j == n - 1) w -= 0.5;
    return w;
}

double emme_oracle_grid(double len, int n, double* eta) {
    double dx = (2 * len) / (unsigned)(n - 1);
    for (int i = 0; i < n; ++i) eta[i] = -len + (unsigned)i * dx;
    return dx;
}

static inline void put(double* out, long dim, long r, long c, cplx v) {
    out[2 * (r * dim + c)] = creal(v);
    out[2 * (r * dim + c) + 1] = cimag(v);
}

void emme_oracle_assemble(const emme_oracle_params* p, int N, const double* eta, const double* g,
                          const double* bi, double dx, double wr, double wi, double* out,
                          int row_begin, int row_end, int nthreads, long* stats) {
    const int em = fpclassify(p->beta_e) != FP_ZERO;
    const long dim = em ? 2L * N : N;
    long s_int = 0, s_ev = 0, s_f = 0, s_b = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    /* diagonal, include/solver.h:443 and :465-470 */
    for (int i = row_begin; i < row_end; ++i) {
        put(out, dim, i, i, CMPLX(1.0 + 1.0 / p->tau, 0.0));
        if (em) {
            put(out, dim, i, i + N, CMPLX(0.0, 0.0));
            put(out, dim, i + N, i, CMPLX(0.0, 0.0));
            put(out, dim, i + N, i + N, CMPLX((2.0 * p->tau) / p->beta_e * bi[i], 0.0));
        }
    }
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : s_int, s_ev, s_f, s_b)
    for (int i = row_begin; i < row_end; ++i) {
        for (int j = i + 1; j < N; ++j) {
            long st[3] = {0, 0, 0};
            double kr, ki, er, ei;
            /* include/solver.h:448-453 / :476-479 */
            emme_oracle_kappa(p, 0, eta[i], eta[j], g[i], g[j], bi[i], bi[j], wr, wi, &kr, &ki, st);
            emme_oracle_kappa_e(p, 0, eta[i], eta[j], g[i], g[j], wr, wi, &er, &ei);
            cplx k0 = CMPLX(kr, ki) + CMPLX(er, ei);
            cplx a = sc_mul(dx, sc_mul(emme_oracle_weight(N, i, j), -k0));
            put(out, dim, i, j, a);
            put(out, dim, j, i, a);
            s_int += 1;
            if (em) {
                emme_oracle_kappa(p, 1, eta[i], eta[j], g[i], g[j], bi[i], bi[j], wr, wi, &kr, &ki, st);
                emme_oracle_kappa_e(p, 1, eta[i], eta[j], g[i], g[j], wr, wi, &er, &ei);
                cplx a1 = sc_mul(dx, CMPLX(kr, ki) + CMPLX(er, ei)); /* :480-484 */
                emme_oracle_kappa(p, 2, eta[i], eta[j], g[i], g[j], bi[i], bi[j], wr, wi, &kr, &ki, st);
                emme_oracle_kappa_e(p, 2, eta[i], eta[j], g[i], g[j], wr, wi, &er, &ei);
                cplx a2 = sc_mul(dx, CMPLX(kr, ki) + CMPLX(er, ei)); /* :486-490 */
                put(out, dim, i, j + N, a1);
                put(out, dim, i + N, j + N, a2);
                put(out, dim, j, i + N, -a1);    /* :494-495 */
                put(out, dim, j + N, i + N, a2); /* :496-498 */
                put(out, dim, i + N, j, -a1);    /* :500-501 */
                put(out, dim, j + N, i, a1);     /* :503-504 */
                s_int += 2;
            }
            s_ev += st[0];
            s_f += st[1];
            s_b += st[2];
        }
    }
    if (stats) {
        stats[0] = s_int;
        stats[1] = s_ev;
        stats[2] = s_f;
        stats[3] = s_b;
    }
}

void emme_oracle_secant(long n, const double* A, const double* Aold, double dr, double di,
                        double* out) {
    const cplx d = CMPLX(dr, di);
    for (long k = 0; k < n; ++k) {
        cplx v = (CMPLX(A[2 * k], A[2 * k + 1]) - CMPLX(Aold[2 * k], Aold[2 * k + 1])) / d;
        out[2 * k] = creal(v);
        out[2 * k + 1] = cimag(v);
    }
}

int emme_oracle_trace_step(int dim, double* Ad_, double* Bd_, double* dr, double* di) {
    cplx* A = (cplx*)Ad_;
    cplx* B = (cplx*)Bd_;
    const long n = dim;
    /* LU with partial pivoting applied to A and, row-wise, to all right-hand sides */
    for (long k = 0; k < n; ++k) {
        long piv = k;
        double best = cabs(A[k * n + k]);
        for (long r = k + 1; r < n; ++r) {
            double v = cabs(A[r * n + k]);
            if (v > best) {
                best = v;
                piv = r;
            }
        }
        if (best == 0.0) return (int)(k + 1);
        if (piv != k) {
            for (long c = 0; c < n; ++c) {
                cplx t = A[k * n + c];
                A[k * n + c] = A[piv * n + c];
                A[piv * n + c] = t;
                t = B[k * n + c];
                B[k * n + c] = B[piv * n + c];
                B[piv * n + c] = t;
            }
        }
        const cplx inv = 1.0 / A[k * n + k];
#pragma omp parallel for schedule(static)
        for (long r = k + 1; r < n; ++r) {
            cplx l = A[r * n + k] * inv;
            if (l == 0.0) continue;
            A[r * n + k] = l;
            for (long c = k + 1; c < n; ++c) A[r * n + c] -= l * A[k * n + c];
            for (long c = 0; c < n; ++c) B[r * n + c] -= l * B[k * n + c];
        }
    }
    /* back substitution, all right-hand sides */
    for (long k = n - 1; k >= 0; --k) {
        const cplx inv = 1.0 / A[k * n + k];
        for (long c = 0; c < n; ++c) B[k * n + c] *= inv;
#pragma omp parallel for schedule(static)
        for (long r = 0; r < k; ++r) {
            cplx u = A[r * n + k];
            for (long c = 0; c < n; ++c) B[r * n + c] -= u * B[k * n + c];
        }
    }
    cplx tr = 0.0;
    for (long k = 0; k < n; ++k) tr += B[k * n + k];
    cplx d = -1.0 / tr; /* include/solver.h:139 */
    *dr = creal(d);
    *di = cimag(d);
    return 0;
}
