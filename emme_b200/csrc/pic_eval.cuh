// pic_eval.cuh -- per-marker arithmetic of the PIC stage kernel (row N4, fp64).
//
// What one marker does in one Runge-Kutta stage of the reference's PIC method
// (include/solver_pic.h: put_velocity :76-135, update :137-151, cal_density in solve_field
// :257-288), restated for pic.cu's fused stage kernel.  NOT a transcription:
//   * the velocity is kept in factored form vs = A phi + B dphi (+ c-term), with
//       A = p_weight conj(dc_pb) (i (omega_st - omega_d omega_dv) j0 - v_para/(qR) dj0)
//       B = -p_weight conj(dc_pb) v_para/(qR) j0
//     formed once per marker position, right after the deposit that needs the same j0, dc_pb;
//   * J0 and J1 come from ONE Miller backward recurrence instead of two library calls;
//   * omega_dv, omega_st are recomputed from (v_para, v_perp) instead of being loaded.
// Only the position update keeps the reference's exact operation order (eta is bit-identical).
//
// The functions are __host__ __device__ so that tests/emul (test infrastructure) can replay the
// stage on the CPU against the reference's field dumps before any GPU time is spent; the
// product only ever calls them from kernels.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define PIC_HD __host__ __device__ __forceinline__
#else
#define PIC_HD inline
#endif

namespace emme {

struct alignas(16) d2 {
    double x, y;
};
PIC_HD d2 mk2(double x, double y) {
    d2 r;
    r.x = x;
    r.y = y;
    return r;
}
PIC_HD d2 cmul2(d2 a, d2 b) { return mk2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x)); }

// exactly rounded single operations (no FMA contraction) for the position update
#if defined(__CUDA_ARCH__)
#define PIC_ADD(a, b) __dadd_rn((a), (b))
#define PIC_MUL(a, b) __dmul_rn((a), (b))
#define PIC_DIV(a, b) __ddiv_rn((a), (b))
#define PIC_RCP(a) __drcp_rn(a)
#define PIC_RSQRT(a) rsqrt(a)
#define PIC_CBRTF(a) __powf((a), 0.33333334f)   /* only sizes the recurrence */
#else
#define PIC_ADD(a, b) ((a) + (b))   /* host builds use -ffp-contract=off */
#define PIC_MUL(a, b) ((a) * (b))
#define PIC_DIV(a, b) ((a) / (b))
#define PIC_RCP(a) (1.0 / (a))
#define PIC_RSQRT(a) (1.0 / sqrt(a))
#define PIC_CBRTF(a) cbrtf(a)
#endif

// run constants of a PIC state
struct PicConst {
    int nf;
    double L, cw, inv_2cw, qR, inv_qR, inv_vt, shat, b_theta, omega_d_bar, omega_s_i, eta_i, inv_2vt2;
};

// J0(x), J1(x), x >= 0, by Miller's backward recurrence J_{k-1} = (2k/x) J_k - J_{k+1} from an
// even start index M with J_M(x) < 3e-18 (M ~ x + 12.6 x^(1/3) + 5, calibrated against scipy),
// normalised with 1 = J0 + 2 sum_{k>=1} J_2k.  Absolute error < 1e-15 for x <= 60.
PIC_HD void bessel_j01(double x, double& j0, double& j1) {
    if (x < 1e-3) {
        const double q = 0.25 * x * x;
        j0 = 1.0 - q + 0.25 * q * q;
        j1 = 0.5 * x * (1.0 - 0.5 * q + q * q * (1.0 / 12.0));
        return;
    }
    const float xf = (float)x;
    const int M = 2 * (int)ceilf(0.5f * (xf + 12.6f * PIC_CBRTF(xf) + 5.2f));
    const double t = 2.0 * PIC_RCP(x);
    double jp = 0.0, jk = 1.0, s = 0.0, kd = (double)M;
#pragma unroll 2
    for (int k = M; k > 2; k -= 2) {
        const double a = kd * t;          // 2k/x
        const double jm1 = fma(a, jk, -jp);
        const double jm2 = fma(a - t, jm1, -jk);
        s += jm2;
        jp = jm1;
        jk = jm2;
        kd -= 2.0;
    }
    const double j1u = fma(2.0 * t, jk, -jp);
    const double j0u = fma(t, j1u, -jk);
    const double inv = PIC_RCP(fma(2.0, s, j0u));
    j0 = j0u * inv;
    j1 = j1u * inv;
}

// locate (include/solver_pic.h:245-249) with the reference's division, so that the cell index
// and the linear weight are the reference's bit for bit (a reciprocal multiply was measured:
// it moves the weight by ulp(u) ~ 1e-12 on an 8192-cell mesh, visible against the oracle, and
// the kernel is HBM-bound anyway); idx == nf (eta == +L after rounding) wraps to cell 0.
PIC_HD void pic_locate(const PicConst& d, double eta, int& idx, double& wt) {
    const double u = PIC_DIV(PIC_ADD(eta, d.L), d.cw);
    const long long i = (long long)u;
    wt = u - (double)i;
    idx = (i >= d.nf || i < 0) ? 0 : (int)i;
}

// gather phi, dphi (include/solver_pic.h:91-100) and evaluate vs = A phi + B dphi
PIC_HD d2 pic_velocity(const PicConst& d, const d2* fld, double eta, d2 A, d2 B) {
    int idx;
    double wt;
    pic_locate(d, eta, idx, wt);
    const int nf = d.nf;
    const int ip1 = (idx + 1 == nf) ? 0 : idx + 1;
    const int ip2 = (ip1 + 1 == nf) ? 0 : ip1 + 1;
    const int im1 = (idx == 0) ? nf - 1 : idx - 1;
    const d2 f0 = fld[idx], f1 = fld[ip1], f2 = fld[ip2], fm = fld[im1];
    const double w0 = 1.0 - wt;
    const d2 phi = mk2(fma(w0, f0.x, wt * f1.x), fma(w0, f0.y, wt * f1.y));
    const d2 dphi = mk2((w0 * (f1.x - fm.x) + wt * (f2.x - f0.x)) * d.inv_2cw,
                        (w0 * (f1.y - fm.y) + wt * (f2.y - f0.y)) * d.inv_2cw);
    d2 vs = cmul2(A, phi);
    const d2 bd = cmul2(B, dphi);
    vs.x += bd.x;
    vs.y += bd.y;
    return vs;
}

// eta <- bound(eta + v_para h/(qR)) in the reference's operation order
// (include/solver_pic.h:143,401-404)
PIC_HD double pic_push(const PicConst& d, double eta, double vpar, double h) {
    double e = PIC_ADD(PIC_ADD(eta, PIC_DIV(PIC_MUL(vpar, h), d.qR)), d.L);
    // fmod(e, 2L): for -2L < e < 4L it is e or e - 2L, both exact; anything else (a marker
    // crossing more than the whole domain in one stage) takes the library call
    const double twoL = 2.0 * d.L;
    if (e >= twoL) e = (e < 2.0 * twoL) ? PIC_ADD(e, -twoL) : fmod(e, twoL);
    else if (e <= -twoL) e = fmod(e, twoL);
    return e < 0 ? PIC_ADD(e, d.L) : PIC_ADD(e, -d.L);
}

// Everything of a marker that depends on its position: the density it deposits
// (include/solver_pic.h:262-275) and the velocity coefficients of the next stage.
template <bool SWITCH>
PIC_HD void pic_marker_at(const PicConst& d, double eta, double vpar, double vperp, double pw, d2 w,
                          d2& den, d2& A, d2& B, double& c) {
    double se, ce;
    sincos(eta, &se, &ce);
    const double xperp = vperp * d.inv_vt;
    const double sh_eta = d.shat * eta;
    const double sb2 = d.b_theta * fma(sh_eta, sh_eta, 1.0);
    const double rsb = PIC_RSQRT(sb2);   // 1/sb
    double j0, j1;
    bessel_j01(xperp * (sb2 * rsb), j0, j1);
    const double dj0 = -d.b_theta * d.shat * d.shat * xperp * eta * j1 * rsb;
    const double v2 = vpar * vpar, p2 = vperp * vperp;
    const double odv = (v2 + 0.5 * p2) * d.inv_2vt2;
    const double ost = d.omega_s_i * (1.0 + d.eta_i * ((v2 + p2) * d.inv_2vt2 - 1.5));
    const double od = d.omega_d_bar * fma(sh_eta, se, ce);
    const double cpar = vpar * d.inv_qR;
    // coefficient of phi: i (omega_st - omega_d omega_dv) j0 - v_para/(qR) dj0; of dphi: -v_para/(qR) j0
    const d2 a = mk2(-cpar * dj0, (ost - od * odv) * j0);
    const double b = -cpar * j0;
    if (SWITCH) {
        const double odi = (d.qR * PIC_RCP(vpar)) * d.omega_d_bar * (se * (1.0 + d.shat) - sh_eta * ce);
        double sp, cp;
        sincos(-odi * odv, &sp, &cp);
        den = cmul2(mk2(j0 * w.x, j0 * w.y), mk2(cp, sp));
        const d2 g = mk2(pw * cp, -pw * sp);  // p_weight conj(dc_pb)
        A = cmul2(g, a);
        B = mk2(g.x * b, g.y * b);
        c = 0.0;
    } else {
        den = mk2(j0 * w.x, j0 * w.y);
        A = mk2(pw * a.x, pw * a.y);
        B = mk2(pw * b, 0.0);
        c = od * odv;
    }
}

// omega_d(eta) omega_dv of the loaded markers: the drift term of the first velocity when the
// pull-back transformation is off (include/solver_pic.h:112-114)
PIC_HD double pic_initial_c(const PicConst& d, double eta, double vpar, double vperp) {
    const double od = d.omega_d_bar * (cos(eta) + d.shat * eta * sin(eta));
    return od * ((vpar * vpar + 0.5 * vperp * vperp) * d.inv_2vt2);
}

// Integrator::coef (include/solver_pic.h:466-470)
#define EMME_PIC_RK_COEF                                                       \
    {{1, 0.62653829327080},                                                    \
     {0, 1, -0.55111240553326},                                                \
     {0, 1.5220585509963, -0.52205855099628, 0.92457411226246},                \
     {1., 0.13686116839369, -1.1368611683937}}

}  // namespace emme
