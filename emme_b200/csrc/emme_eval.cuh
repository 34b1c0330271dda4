// emme_eval.cuh -- per-node integrand evaluation of the EMME ion kernel (fp64).
//
// One call = one evaluation of the wrapped integrand g(x) = f(tan x)/cos^2 x that the
// reference's util::integrate hands to Gauss-Kronrod (include/functions.h:315-318), with f
// the lambda of Parameters::kappa_f_tau (src/Parameters.cpp:120-176) and the Miller
// recurrence of util::bessel_i_alter_helper (include/functions.h:381-408) inlined.
//
// This is NOT a transcription: everything that depends only on the pair (eta, eta') is
// hoisted into PairConst, everything that depends only on the node (and omega) into NodeConst,
// divisions by complex numbers are replaced by one reciprocal of lambda and one of mu, and two
// algebraic identities remove work per node:
//     2 + i*beta_1/nu  ==  2*lambda            (nu = qR*deta/(vt*tau~))
//     1/tau~           ==  conj(e)/t           (tau~ = t*e, |e| = 1)
// so results agree with the reference to rounding (a few ulp per factor), not bit-for-bit.
//
// The functions are __host__ __device__ so that tests/emul (test infrastructure) can run the
// same arithmetic on the CPU against the oracle before any GPU time is spent; the product
// only ever calls them from kernels.
#pragma once
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define EMME_HD __host__ __device__ __forceinline__
#else
#define EMME_HD inline
#endif

namespace emme {

struct cplx {
    double re, im;
};

EMME_HD cplx mk(double r, double i) { return cplx{r, i}; }
EMME_HD cplx operator+(cplx a, cplx b) { return mk(a.re + b.re, a.im + b.im); }
EMME_HD cplx operator-(cplx a, cplx b) { return mk(a.re - b.re, a.im - b.im); }
EMME_HD cplx operator-(cplx a) { return mk(-a.re, -a.im); }
EMME_HD cplx operator*(cplx a, cplx b) {
    return mk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
EMME_HD cplx operator*(double s, cplx a) { return mk(s * a.re, s * a.im); }
EMME_HD cplx conj(cplx a) { return mk(a.re, -a.im); }
EMME_HD double norm2(cplx a) { return a.re * a.re + a.im * a.im; }
// 1/x and 1/sqrt(x) for positive, normal x: hardware seed (MUFU, >= 20 bits) + two Newton steps in
// FMAs -- within 1 ulp, 9 issue slots instead of the ~25 of the correctly rounded library division
// (whose special-case handling this kernel never needs: x is |lambda|^2, |mu|^2 or 1 + u^2).  The
// kernel is issue-slot bound, so these count.
EMME_HD double rcp_pos(double x) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
#else
    return 1.0 / x;         // host build (tests/emul) only
#endif
}
EMME_HD cplx recip(cplx a) {
    const double d = rcp_pos(norm2(a));
    return mk(a.re * d, -a.im * d);
}
// i*a
EMME_HD cplx mul_i(cplx a) { return mk(-a.im, a.re); }
EMME_HD double rsqrt_(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    double e = fma(-hx * y, y, 0.5);     // (1 - x y^2)/2
    y = fma(y, e, y);
    e = fma(-hx * y, y, 0.5);
    return fma(y, e, y);
#else
    return 1.0 / sqrt(x);   // host build (tests/emul) only
#endif
}

// Scalars of the run (one per assembly), precomputed on the host in the reference's own
// association order so that they are bit-identical to what the reference multiplies by.
struct RunConst {
    double qR;          // q*R
    double vt;
    double arc;         // arc_coeff
    double inv_arc;     // 1/arc_coeff
    double omega_s_i;
    double eta_i;
    double wsi_etai;    // omega_s_i*eta_i
    double c_beta;      // (q*R)/vt*omega_d_bar          -> beta_1   = c_beta  *(g-g')
    double c_beta_e;    // (q*R)/vt*(omega_d_bar*omega_s_e/omega_s_i)
    double kappa_pref;  // (q*R)/(vt*sqrt(2*pi))         (src/Parameters.cpp:182-183)
    double ke1_pref;    // (q*R)/(2*vt*tau)              (src/Parameters.cpp:196)
    double ke2_pref;    // (q*q*R*R)/(2*vt*vt*tau)       (src/Parameters.cpp:200)
    double a0;          // Re(omega) - omega_s_i*(1 - 1.5*eta_i): constant part of i0_coef's numerator
    double omega_s_e;
    double eta_e;
    double diag_es;     // 1 + 1/tau                     (include/solver.h:443)
    double diag_em;     // (2*tau)/beta_e                (include/solver.h:469)
    double tol, prec;   // integration_precision / integration_accuracy
    double thr_len;     // 0.99*(b-a), b-a = pi/2        (include/functions.h:240)
    double inv_scale;   // 2/(b-a)                       (include/functions.h:219)
    double half_pi;     // b = pi/2
    double dx;
    double wr, wi;      // omega
    double omi;         // -copysign(1, Re omega)        (src/Parameters.cpp:121)
    int maxdepth;
    int order;
    int N;
    int em;             // beta_e != 0
};

// Everything that depends on the pair (eta_i, eta_j) only.
struct PairConst {
    double deta;    // eta - eta'
    double D;       // q*R*(eta-eta')
    double Dv;      // D/vt
    double beta1;   // beta_1(eta, eta')
    double cl;      // 0.5*vt*beta1/D : lambda = 1 + i*cl*tau~
    double s;       // sqrt(b*b')
    double two_over_s;
    double hb;      // 0.5*(b+b')
    double bsum;    // b+b'
    double c1;      // -omega_s_i*eta_i*s
    double ca;      // Dv^2                      : nu^2 = ca/tau~^2
    double cb;      // 0.5*beta1*Dv              : 0.5*beta1*nu = cb/tau~
    double cA;      // omega_s_i*eta_i*Dv^2      : omega_s_i*eta_i*nu^2 = cA/tau~^2
    // accessors: eval_node is generic over where the pair constants live (this struct on the host
    // emulation, shared memory in the kernel -- see PairSmem in assembly.cu)
    EMME_HD double f_Dv() const { return Dv; }
    EMME_HD double f_beta1() const { return beta1; }
    EMME_HD double f_cl() const { return cl; }
    EMME_HD double f_s() const { return s; }
    EMME_HD double f_two_over_s() const { return two_over_s; }
    EMME_HD double f_hb() const { return hb; }
    EMME_HD double f_c1() const { return c1; }
    EMME_HD double f_ca() const { return ca; }
    EMME_HD double f_cb() const { return cb; }
    EMME_HD double f_cA() const { return cA; }
};

EMME_HD PairConst make_pair(const RunConst& rc, double eta, double etap, double g, double gp,
                            double b, double bp) {
    PairConst pc;
    pc.deta = eta - etap;
    pc.D = rc.qR * pc.deta;
    pc.Dv = pc.D / rc.vt;
    pc.beta1 = rc.c_beta * (g - gp);
    pc.cl = 0.5 * rc.vt * pc.beta1 / pc.D;
    pc.s = sqrt(b * bp);
    pc.two_over_s = 2.0 / pc.s;
    pc.bsum = b + bp;
    pc.hb = 0.5 * pc.bsum;
    pc.c1 = -rc.omega_s_i * rc.eta_i * pc.s;
    pc.ca = pc.Dv * pc.Dv;
    pc.cb = 0.5 * pc.beta1 * pc.Dv;
    pc.cA = rc.wsi_etai * pc.ca;
    return pc;
}

struct EvalCounters {
    unsigned int fwd, bwd;
};

// Scaled modified Bessel ratios by Miller's algorithm (include/functions.h:381-408).
// z = s/lambda, zc = 2/z = (2/s)*lambda.  Returns y0, y1, mu(+y0); the 4th element of the
// reference's array (-z or z) is formed by the caller.
// exact-order helpers with explicit fused multiply-adds (4 DFMA per complex multiply-add)
EMME_HD cplx cfma(cplx a, cplx b, cplx c) {    // a*b + c
    return mk(fma(a.re, b.re, fma(-a.im, b.im, c.re)), fma(a.re, b.im, fma(a.im, b.re, c.im)));
}
EMME_HD cplx cfms(cplx a, cplx b, cplx c) {    // c - a*b
    return mk(fma(-a.re, b.re, fma(a.im, b.im, c.re)), fma(-a.re, b.im, fma(-a.im, b.re, c.im)));
}
// |p|^2 <= thr for non-negative doubles, on the integer pipe (NaN compares "greater": loop ends)
EMME_HD bool le_nonneg(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(a) <= __double_as_longlong(b);
#else
    return a <= b;
#endif
}

EMME_HD void bessel_i_alter(cplx z, cplx zc, cplx& y0, cplx& y1, cplx& mu, EvalCounters& cnt) {
    const double THRESHOLD = 2.e+7;
    // n0 = floor(|z|) + 1 (include/functions.h:384) without a double-precision square root: FP32
    // estimate, then the exact integer fix-up against |z|^2 (n*n is exact in FP64).  floor of the
    // exact root and floor of the rounded root differ only if |z| lies within an ulp of an integer.
    const double n2 = norm2(z);
#if defined(__CUDA_ARCH__)
    float sq;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"((float)n2));   // +-1 is repaired below
    int n0 = (int)sq;
#else
    int n0 = (int)sqrtf((float)n2);
#endif
    {
        const double d0 = (double)n0;
        if (d0 * d0 > n2) --n0;
        else if (fma(d0, d0, 2.0 * d0 + 1.0) <= n2) ++n0;
    }
    n0 += 1;
    // the order n is carried as an int (loop control, integer pipe) and as a double (exact)
    int n = n0;
    double dn = (double)n0;
    // Forward recurrence p_{k+1} = p_{k-1} - (2n/z) p_k until |p|^2 exceeds the threshold.  Only the
    // STOPPING INDEX N is used below (the backward pass restarts from y_N = 1; the VALUE p_N only
    // scales the returned triple): the search therefore runs in FP32 on the otherwise idle FP32
    // pipe (relative error ~1e-6 in |p|^2 can move N by one only when |p_N|^2 sits within 1e-6 of
    // the threshold).  N itself must be the reference's: stopping at |p| > 2e7 leaves the
    // reference's ratios with a truncation error of ~1e-9, so a different start index -- even a
    // larger, more accurate one -- moves them by 1e-9..1e-11, beyond the 1e-10 parity bar (tried in
    // round 2 with a tabulated upper bound).
    if (n2 >= 1e-10) {
        const float thr2 = (float)(THRESHOLD * THRESHOLD);
        const float zr = (float)zc.re, zi = (float)zc.im;
        float fn = (float)n0;
        float par = 0.f, pai = 0.f, pbr = 1.f, pbi = 0.f;   // pa = p_{k-1}, pb = p_k, k = n0 + 2*rounds
        // Two trips per round with ONE threshold test (the kernel is issue bound: the test is a
        // third of a trip).  When p_k fails the test the round before may already have produced
        // p_{k-1} above the threshold: checked once on exit.  |p| <= 2e7 grows by at most
        // (n |2/z|)^2 <= 1e13 per untested round, far below FLT_MAX.
        for (;;) {
            if (!(fmaf(pbr, pbr, pbi * pbi) <= thr2)) break;
            {
                const float cr = fn * zr, ci = fn * zi;
                par = fmaf(-cr, pbr, fmaf(ci, pbi, par));     // pa <- p_{k+1}
                pai = fmaf(-cr, pbi, fmaf(-ci, pbr, pai));
            }
            fn += 1.0f;
            {
                const float cr = fn * zr, ci = fn * zi;
                pbr = fmaf(-cr, par, fmaf(ci, pai, pbr));     // pb <- p_{k+2}
                pbi = fmaf(-cr, pai, fmaf(-ci, par, pbi));
            }
            fn += 1.0f;
        }
        if (!(fmaf(par, par, pai * pai) <= thr2)) fn -= 1.0f;  // p_{k-1} was the first above the threshold
        n = (int)fn;
        dn = (double)n;
    } else {
        // test_1 = max(sqrt(T*|p1|*|p0 - 2n/z*p1|), T) with p0 = 0, p1 = 1; compared squared:
        // max(T*n*|2/z|, T^2).  The first term only wins for |z| < n*1e-7 (FP64 here: rare, and the
        // products overflow FP32).
        const double thr2 = fmax(THRESHOLD * (dn * (2.0 / sqrt(n2))), THRESHOLD * THRESHOLD);
        cplx pa = mk(0., 0.), pb = mk(1., 0.);   // pa = p_{k-1}, pb = p_k
        for (;;) {
            if (!le_nonneg(norm2(pb), thr2)) break;
            pa = cfms(dn * zc, pb, pa);          // pa <- p_{k+1}
            dn += 1.0;
            ++n;
            if (!le_nonneg(norm2(pa), thr2)) break;
            pb = cfms(dn * zc, pa, pb);          // pb <- p_{k+2}
            dn += 1.0;
            ++n;
        }
    }
    cnt.fwd += (unsigned)(n - n0);
    --n;
    dn -= 1.0;
    cnt.bwd += (unsigned)n;
    // backward recurrence y_{k-1} = (2k/z) y_k + y_{k+1} from y_N = 1 (the reference starts from
    // 1/p_N; the returned ratios y/mu do not depend on that scale, so the reciprocal is dropped),
    // mu += 2*(Re z < 0 ? 1 - 2*(k & 1) : 1) * y_k.  Two trips per round: ya/yb swap roles instead
    // of being moved, and the alternating sign is a pair of constants.
    const bool neg = z.re < 0;
    const double sgA = (neg && (n & 1)) ? -2.0 : 2.0;   // sign of the first (and every odd) trip
    const double sgB = neg ? -sgA : sgA;
    cplx ya = mk(1., 0.), yb = mk(0., 0.);   // ya = y_k ("y0"), yb = y_{k+1} ("y1")
    mu = mk(0., 0.);
    for (; n >= 4; n -= 4) {                 // four trips per round: a quarter of the loop control
        yb = cfma(dn * zc, ya, yb);
        mu = mk(fma(sgA, ya.re, mu.re), fma(sgA, ya.im, mu.im));
        dn -= 1.0;
        ya = cfma(dn * zc, yb, ya);
        mu = mk(fma(sgB, yb.re, mu.re), fma(sgB, yb.im, mu.im));
        dn -= 1.0;
        yb = cfma(dn * zc, ya, yb);
        mu = mk(fma(sgA, ya.re, mu.re), fma(sgA, ya.im, mu.im));
        dn -= 1.0;
        ya = cfma(dn * zc, yb, ya);
        mu = mk(fma(sgB, yb.re, mu.re), fma(sgB, yb.im, mu.im));
        dn -= 1.0;
    }
    for (; n >= 2; n -= 2) {
        yb = cfma(dn * zc, ya, yb);          // yb <- y_{k-1}; ya is now "y1"
        mu = mk(fma(sgA, ya.re, mu.re), fma(sgA, ya.im, mu.im));
        dn -= 1.0;
        ya = cfma(dn * zc, yb, ya);          // ya <- y_{k-2}; yb is now "y1"
        mu = mk(fma(sgB, yb.re, mu.re), fma(sgB, yb.im, mu.im));
        dn -= 1.0;
    }
    if (n == 1) {
        yb = cfma(dn * zc, ya, yb);
        mu = mk(fma(sgA, ya.re, mu.re), fma(sgA, ya.im, mu.im));
        y0 = yb;
        y1 = ya;
    } else {
        y0 = ya;
        y1 = yb;
    }
    mu = mu + y0;
}

// What the integrand needs of a quadrature node x in (0, pi/2): t = tan x, 1/t and 1/cos^2 x
// (the [0,inf) -> [0,pi/2] map of include/functions.h:315-318).  It depends on the panel and the
// node only, not on the integrand, so the kernel tabulates it for the first levels of the
// bisection tree (node_trig is the single definition both the table and the direct path use).
struct NodeTrig {
    double t, it, icsq, x;
};

EMME_HD NodeTrig node_trig(double x) {
    double sx, c;
    sincos(x, &sx, &c);
    const double ic = 1.0 / c;
    NodeTrig n;
    n.t = sx * ic;
    n.it = c / sx;
    n.icsq = ic * ic;
    n.x = x;
    return n;
}

// Everything the integrand needs of a quadrature node that does NOT depend on the pair (eta, eta'):
// the contour rotation e = exp(-i*omi*atan(u)), u = t/arc (src/Parameters.cpp:121-124), and what is
// built from it and from omega.  One definition for the per-assembly node table
// (assembly.cu::node_table_kernel), the direct path below the table and the host emulation.
struct NodeConst {
    cplx taut;    // tau~ = t*e
    cplx itaut;   // 1/tau~ = conj(e)/t
    cplx h;       // -0.5*(1/tau~)^2
    cplx M;       // i*tau~*omega                       (:160)
    cplx pj;      // jacobian/tau~                      (:126-129, :174)
    double icsq;  // 1/cos^2 x                          (include/functions.h:317)
    double pad;
};

EMME_HD NodeConst node_const(const RunConst& rc, double x) {
    const NodeTrig nt = node_trig(x);
    const double t = nt.t, it = nt.it;
    // cos(atan u) = 1/sqrt(1+u^2), sin(atan u) = u/sqrt(1+u^2) -- no atan, no second sincos
    const double u = t * rc.inv_arc;
    const double w1 = 1.0 + u * u;
    const double rs = rsqrt_(w1);
    const cplx e = mk(rs, -rc.omi * (u * rs));
    // jacobian: e - i*e*omi*t/(arc*(1+u^2)) = e - i*e*omi*u/(1+u^2), 1/(1+u^2) = rs^2
    const double jd = rc.omi * u * (rs * rs);
    const cplx jacob = e - jd * mul_i(e);
    NodeConst nc;
    nc.taut = t * e;
    nc.itaut = it * conj(e);
    nc.h = (-0.5) * (nc.itaut * nc.itaut);
    nc.M = mul_i(nc.taut * mk(rc.wr, rc.wi));
    nc.pj = nc.itaut * jacob;
    nc.icsq = nt.icsq;
    nc.pad = 0.;
    return nc;
}

// exp(a + i b) for the integrand's exponential factor: a is known to lie in [-40, ~60] (the
// underflow guard has already removed a < -40), so the general-purpose exp()/sincos() of the CUDA
// math library -- 48 + 74 instructions per call, two thirds of them range/special-case handling on
// the integer pipe -- are replaced by straight-line code: Cody-Waite reduction with FMAs, fdlibm's
// minimax kernels for sin/cos on [-pi/4, pi/4] (error < 2^-58) and a degree-13 Taylor kernel for
// exp on [-ln2/2, ln2/2] (truncation 4e-18).  |b| >= 2^19 (never seen in practice) and a outside
// [-700, 700] fall back to the library.  tests/test_emul.py::test_lean_cexp pins the product to <= 3 ulp per component.
// Measured on B200 (N = 8192).  Round 1, before the kernel became issue-slot bound: 122.2 ms with
// it against 121.0 ms with the library calls (kept OFF).  Round 2, on the shipped kernel: 100.95 ms
// with it against 102.21 ms (C1, C3 at N = 1024 unchanged) -- ON by default; -DEMME_LEAN_CEXP=0
// restores the library calls.
#ifndef EMME_LEAN_CEXP
#define EMME_LEAN_CEXP 1
#endif

// coefficients of cexp_lean: in __constant__ memory on the device so that they are direct
// c[bank][offset] operands of the DFMAs (as immediates each 64-bit constant costs two UMOVs)
#define EMME_CEXP_COEFFS                                                                            \
    {6.36619772367581382433e-01,  /*  0 2/pi                        */                               \
     1.57079632673412561417e+00,  /*  1 pi/2, first 33 bits         */                               \
     6.07710050630396597660e-11,  /*  2 pi/2, next 33 bits          */                               \
     2.02226624879595063154e-21,  /*  3 pi/2, tail                  */                               \
     1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,            \
     -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01, /* 4-9 S6..S1 */ \
     -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,           \
     2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02, /* 10-15 C6..C1 */ \
     1.44269504088896338700e+00,  /* 16 log2(e)                     */                               \
     6.93147180369123816490e-01,  /* 17 ln2, first 33 bits          */                               \
     1.90821492927058770002e-10,  /* 18 ln2, tail                   */                               \
     1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07,     \
     2.7557319223985893e-06, 2.48015873015873e-05, 1.984126984126984e-04, 1.388888888888889e-03,     \
     8.333333333333333e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, /* 19-29: 1/13! .. 1/3! */ \
     6755399441055744.0}          /* 30 1.5*2^52: round-to-nearest-integer trick */
#if defined(__CUDACC__)
static __constant__ double c_cexp[31] = EMME_CEXP_COEFFS;
#endif
static const double h_cexp[31] = EMME_CEXP_COEFFS;
#if defined(__CUDA_ARCH__)
#define EMME_CX(i) c_cexp[i]
#else
#define EMME_CX(i) h_cexp[i]
#endif

EMME_HD int hi_word(double x) {
#if defined(__CUDA_ARCH__)
    return __double2hiint(x);
#else
    long long b;
    memcpy(&b, &x, 8);
    return (int)(b >> 32);
#endif
}
EMME_HD int lo_word(double x) {
#if defined(__CUDA_ARCH__)
    return __double2loint(x);
#else
    long long b;
    memcpy(&b, &x, 8);
    return (int)(b & 0xffffffffLL);
#endif
}
EMME_HD double from_words(int hi, int lo) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(hi, lo);
#else
    const long long b = ((long long)hi << 32) | (unsigned int)lo;
    double x;
    memcpy(&x, &b, 8);
    return x;
#endif
}

EMME_HD cplx cexp_lib(double a, double b) {
    double es, ec;
    sincos(b, &es, &ec);
    const double er = exp(a);
    return mk(er * ec, er * es);
}

EMME_HD cplx cexp_lean(double a, double b) {
    // |b| < 2^19 and |a| < 512, tested on the exponent fields (integer pipe)
    if ((hi_word(b) & 0x7fffffff) < 0x41200000 && (hi_word(a) & 0x7fffffff) < 0x40800000) {
        const double MAGIC = EMME_CX(30);
        // ---- sin b, cos b: q = rint(b*2/pi) sits in the low word of b*2/pi + 1.5*2^52 ----
        const double qm = fma(b, EMME_CX(0), MAGIC);
        const int q = lo_word(qm);
        const double qd = qm - MAGIC;
        double r = fma(-qd, EMME_CX(1), b);        // exact product (33-bit constant, |q| < 2^20)
        r = fma(-qd, EMME_CX(2), r);
        r = fma(-qd, EMME_CX(3), r);
        const double z = r * r;
        double ps = fma(z, EMME_CX(4), EMME_CX(5));
        ps = fma(z, ps, EMME_CX(6));
        ps = fma(z, ps, EMME_CX(7));
        ps = fma(z, ps, EMME_CX(8));
        ps = fma(z, ps, EMME_CX(9));
        const double sn = fma(z * r, ps, r);
        double pc = fma(z, EMME_CX(10), EMME_CX(11));
        pc = fma(z, pc, EMME_CX(12));
        pc = fma(z, pc, EMME_CX(13));
        pc = fma(z, pc, EMME_CX(14));
        pc = fma(z, pc, EMME_CX(15));
        const double hz = 0.5 * z, w = 1.0 - hz;
        const double cs = w + (((1.0 - w) - hz) + z * (z * pc));
        // quadrant: swap for odd q, signs from bit 1 of q (sin) and of q+1 (cos), applied to the sign bits
        const double s0 = (q & 1) ? cs : sn, c0 = (q & 1) ? sn : cs;
        const double s_ = from_words(hi_word(s0) ^ ((q & 2) << 30), lo_word(s0));
        const double c_ = from_words(hi_word(c0) ^ (((q + 1) & 2) << 30), lo_word(c0));
        // ---- exp a = 2^k * exp(t), k = rint(a*log2 e), |t| <= ln2/2: Taylor to t^13 (4e-18) ----
        const double km = fma(a, EMME_CX(16), MAGIC);
        const int k = lo_word(km);
        const double kd = km - MAGIC;
        double t = fma(-kd, EMME_CX(17), a);
        t = fma(-kd, EMME_CX(18), t);
        double pe = fma(t, EMME_CX(19), EMME_CX(20));
        pe = fma(t, pe, EMME_CX(21));
        pe = fma(t, pe, EMME_CX(22));
        pe = fma(t, pe, EMME_CX(23));
        pe = fma(t, pe, EMME_CX(24));
        pe = fma(t, pe, EMME_CX(25));
        pe = fma(t, pe, EMME_CX(26));
        pe = fma(t, pe, EMME_CX(27));
        pe = fma(t, pe, EMME_CX(28));
        pe = fma(t, pe, EMME_CX(29));
        pe = fma(t, pe, 0.5);
        pe = fma(t, pe, 1.0);
        pe = fma(t, pe, 1.0);
        // scale by 2^k on the exponent field (pe in [0.7, 1.42], |k| < 740: always a normal number)
        const double er = from_words(hi_word(pe) + (k << 20), lo_word(pe));
        return mk(er * c_, er * s_);
    }
    return cexp_lib(a, b);
}

// g(x) for mode m (0, 1, 2) at the node described by nc.
template <class PC>
EMME_HD cplx eval_node(const RunConst& rc, const PC& pc, int m, const NodeConst& nc,
                       EvalCounters& cnt) {
    // lambda = 1 + i*cl*tau~ (:101-106, :131)
    const double cl = pc.f_cl();
    const cplx lambda = mk(1.0 - cl * nc.taut.im, cl * nc.taut.re);
    const cplx il = recip(lambda);
    const cplx z = pc.f_s() * il;                  // sqrt(b b')/lambda  (:135-136)
    const cplx zc = pc.f_two_over_s() * lambda;    // 2/z
    const cplx z4 = z.re < 0 ? z : -z;         // include/functions.h:407
    // log of the exponential factor (:157-164) with nu = Dv/tau~ (:140) and 2 + i*beta1/nu == 2*lambda:
    //   -nu^2/2 - i*beta1*nu/2 + i*tau~*omega - (b+b')/(2*lambda)  =  ca*h + cb*(-i/tau~) + M - hb/lambda
    const double hb = pc.f_hb(), ca = pc.f_ca(), cb = pc.f_cb();
    const cplx L = mk(fma(ca, nc.h.re, fma(cb, nc.itaut.im, fma(-hb, il.re, nc.M.re))),
                      fma(ca, nc.h.im, fma(-cb, nc.itaut.re, fma(-hb, il.im, nc.M.im))));
    const cplx arg = L - z4;
    // safe_exp underflow guard (:167-173): the exponential factor is an exact zero and the node
    // contributes nothing -- decided BEFORE the Bessel recurrence, which is skipped (a third of all
    // evaluations of C1).  One corner is kept faithful: the reference still multiplies its zero by
    // (i0 y0 + i1 y1)/mu, and once the backward recurrence of bessel_i_alter_helper has overflowed
    // (|z| beyond ~700: lambda ~ 0, only reached at Im omega < 0 far from any root) that product is
    // 0 * inf = NaN, the matrix entry is NaN, zsysv fails and the scan records the point as "NaN"
    // (src/main.cpp:311-318; C5 point 63, tests/golden/c5.json).  Below |z| = 500 no overflow is
    // possible (growth <= e^|z|), above it the recurrence runs and decides.
    const bool under = arg.re < -40.;
    if (under && norm2(z) < 2.5e5) return mk(0., 0.);

    cplx y0, y1, mu;
    bessel_i_alter(z, zc, y0, y1, mu, cnt);

    // i0_coef*y0 + i1_coef*y1 (:142-151) with pow(lambda, -3.) = il^3 factored as il * il^2:
    //   i0 = il*(A + B*il^2),  i1 = il*(c1*il^2),  A = omega - omega_s_i*inner,  B = wsi_etai*(hb - lambda)
    const cplx il2 = il * il;
    //   A = omega - omega_s_i*(1 + eta_i*(nu^2/2 - 1.5)) = a0 + cA*h  (+ i Im omega)
    const double cA = pc.f_cA();
    const cplx A = mk(fma(cA, nc.h.re, rc.a0), fma(cA, nc.h.im, rc.wi));
    const cplx B = rc.wsi_etai * (mk(hb, 0.) - lambda);
    const cplx S = (A + B * il2) * y0 + (pc.f_c1() * il2) * y1;
    if (under) {
        const double probe = (S.re + S.im) + (mu.re + mu.im);      // non-finite iff any part is
        return (probe - probe == 0.0) ? mk(0., 0.) : mk(probe - probe, probe - probe);   // 0, or NaN like 0 * inf
    }

#if EMME_LEAN_CEXP
    const cplx se = cexp_lean(arg.re, arg.im);
#else
    const cplx se = cexp_lib(arg.re, arg.im);
#endif

    cplx pw = nc.pj;                           // nu^m / tau~ * jacobian   (:174)
    if (m >= 1) {
        const cplx nu = pc.f_Dv() * nc.itaut;
        pw = pw * nu;
        if (m >= 2) pw = pw * nu;
    }
    // f = pw * se * il * S / mu, 1/mu = conj(mu)/|mu|^2 with the real factor applied last
    const cplx f = (pw * se) * ((il * S) * conj(mu));
    const double sc = nc.icsq * rcp_pos(norm2(mu));
    return sc * f;                             // f(tan x)/cos^2 x, include/functions.h:317
}

// Closed-form electron part (src/Parameters.cpp:186-209), m = 1, 2 (m = 0 is zero).
EMME_HD cplx kappa_e(const RunConst& rc, int m, double deta, double dg) {
    const cplx w = mk(rc.wr, rc.wi);
    const double sgn_num = deta, sgn_den = fabs(deta);
    if (m == 1) {
        const cplx a = mk(rc.wr - rc.omega_s_e, rc.wi);
        cplx k = mk(rc.ke1_pref * a.im, -rc.ke1_pref * a.re);  // -i*pref*(omega-omega_s_e)
        k = sgn_num * k;
        return mk(k.re / sgn_den, k.im / sgn_den);
    }
    if (m == 2) {
        const double pref = rc.ke2_pref * sgn_num / sgn_den;
        const cplx t1 = deta * (w * mk(rc.wr - rc.omega_s_e, rc.wi));
        const double b1e = rc.c_beta_e * dg;
        const double c2 = b1e * rc.vt / rc.qR;
        const cplx t2 = c2 * mk(rc.wr - rc.omega_s_e * (1.0 + rc.eta_e), rc.wi);
        return pref * (t1 - t2);
    }
    return mk(0., 0.);
}

// SingularityHandler(n)(i,j) on the fly (src/singularity_handler.cpp:3-24).
EMME_HD double sing_weight(int n, int i, int j) {
    const int diff = i > j ? i - j : j - i;
    double w;
    switch (diff) {
        case 0: w = 0.0; break;
        case 1: w = 2.951388888888883; break;
        case 2: w = -2.4305555555555305; break;
        case 3: w = 4.166666666667441; break;
        case 4: w = -0.3472222222224549; break;
        case 5: w = 1.159722222222284; break;
        default: w = 1.0; break;
    }
    if (j == 0 || j == n - 1) w -= 0.5;
    return w;
}

// ---- Gauss-Kronrod tables (include/functions.h:92-162).  Index 0 is the centre node. ----
struct GKTables {
    double a[16];   // abscissae
    double kw[16];  // Kronrod weights
    double gw[16];  // Gauss weight of node i (0 for pure Kronrod nodes)
};

}  // namespace emme
