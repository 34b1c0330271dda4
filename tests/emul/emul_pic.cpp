// tests/emul/emul_pic.cpp -- TEST INFRASTRUCTURE: CPU replay of pic.cu's stage kernel.
//
// Compiles the SAME per-marker arithmetic the CUDA kernel uses (emme_b200/csrc/pic_eval.cuh)
// with g++ and runs, marker by marker, what pic_stage_kernel does in one Runge-Kutta stage
// (gather + factored velocity, stage combination with only k1 stored, push, Miller J0/J1,
// pull-back phase, deposit, next-stage coefficients A/B, field = density * table).  It lets the
// CPU test suite check that restructuring against the reference's per-step field dumps without
// a GPU.  glibc's libm instead of CUDA's and a serial deposit order: it says nothing about
// device rounding or atomics; the -m gpu tests do that.  Never linked into libemme_b200.so.
#define _GNU_SOURCE 1
#include <cmath>
#include <cstring>
#include <vector>

#include "../../emme_b200/csrc/pic_eval.cuh"
#include "../../include/emme_b200.h"

using namespace emme;

static const double RK[4][4] = EMME_PIC_RK_COEF;

template <bool SWITCH>
static void run(const PicConst& k, long n, std::vector<double>& eta, const double* vpar,
                const double* vperp, std::vector<d2>& w, const double* pw, const double* coef,
                double dt, int nsteps, double* fields_out) {
    const int nf = k.nf;
    std::vector<d2> A(n, mk2(0, 0)), B(n, mk2(0, 0)), k1(n, mk2(0, 0)), field(nf, mk2(0, 0)), dens(nf);
    std::vector<double> c(n, 0.0);
    if (!SWITCH)
        for (long i = 0; i < n; ++i) c[i] = pic_initial_c(k, eta[i], vpar[i], vperp[i]);
    for (int step = 0; step < nsteps; ++step) {
        for (int stage = 0; stage < 3; ++stage) {
            const double h = RK[stage][stage + 1] * dt, c1 = RK[2][1], c2 = RK[2][2];
            for (int i = 0; i < nf; ++i) dens[i] = mk2(0, 0);
            for (long i = 0; i < n; ++i) {
                d2 vs = pic_velocity(k, field.data(), eta[i], A[i], B[i]);
                if (!SWITCH) {
                    vs.x += c[i] * w[i].y;
                    vs.y -= c[i] * w[i].x;
                }
                d2 v = vs;
                if (stage == 1) k1[i] = vs;
                else if (stage == 2) v = mk2(c1 * k1[i].x + c2 * vs.x, c1 * k1[i].y + c2 * vs.y);
                eta[i] = pic_push(k, eta[i], vpar[i], h);
                w[i].x = fma(v.x, h, w[i].x);
                w[i].y = fma(v.y, h, w[i].y);
                d2 den;
                pic_marker_at<SWITCH>(k, eta[i], vpar[i], vperp[i], pw[i], w[i], den, A[i], B[i], c[i]);
                int idx;
                double wt;
                pic_locate(k, eta[i], idx, wt);
                const int i1 = (idx + 1 == nf) ? 0 : idx + 1;
                dens[idx].x += den.x * (1.0 - wt);
                dens[idx].y += den.y * (1.0 - wt);
                dens[i1].x += den.x * wt;
                dens[i1].y += den.y * wt;
            }
            for (int i = 0; i < nf; ++i) field[i] = mk2(dens[i].x * coef[i], dens[i].y * coef[i]);
        }
        std::memcpy(fields_out + (size_t)step * nf * 2, field.data(), sizeof(d2) * nf);
    }
}

extern "C" int emul_pic_run(const emme_pic_params* p, long n, const double* eta0, const double* vpar,
                            const double* vperp, const double* w0, const double* pw, const double* coef,
                            double dt, int nsteps, double* fields_out, double* eta_out, double* w_out) {
    PicConst k;
    const double cell_width = 2 * p->length / p->npoints;
    k.nf = p->npoints;
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
    std::vector<double> eta(eta0, eta0 + n);
    std::vector<d2> w(n);
    std::memcpy(w.data(), w0, sizeof(d2) * n);
    if (p->drift_center_transformation_switch) run<true>(k, n, eta, vpar, vperp, w, pw, coef, dt, nsteps, fields_out);
    else run<false>(k, n, eta, vpar, vperp, w, pw, coef, dt, nsteps, fields_out);
    std::memcpy(eta_out, eta.data(), sizeof(double) * n);
    std::memcpy(w_out, w.data(), sizeof(d2) * n);
    return 0;
}

extern "C" void emul_bessel_j01(double x, double* j0, double* j1) { bessel_j01(x, *j0, *j1); }
