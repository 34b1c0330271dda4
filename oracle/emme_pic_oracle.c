/* oracle/emme_pic_oracle.c -- plain-C restatement of the reference's PIC method, in the
 * reference's own operation order (TEST INFRASTRUCTURE, see emme_pic_oracle.h).
 *
 * Follows include/solver_pic.h of ssskkkky/EMME:
 *   put_velocity            :76-135   (marker velocity d(weight)/dt from the gathered field)
 *   update                  :137-151  (push eta, advance weight, solve the field)
 *   initialize_marker_extras:207-238
 *   locate                  :245-249
 *   solve_field             :251-354  (256 sequential batches, summed in batch order)
 *   omega_d, omega_d_integral, cal_quasi_neutrality_coef, bound :361-404
 *   Integrator::step / coef :423-434,466-470
 *   util::calculate_omega   :475-529
 * std::complex arithmetic is restated with C99 double _Complex (same libgcc multiply), the
 * complex exponential with glibc's cexp (what libstdc++'s std::exp forwards to), and the two
 * special functions std::cyl_bessel_j / std::cyl_bessel_i come from libstdc++ through the
 * two-line C++ shim oracle/bessel_shim.cpp, because the reference calls exactly those.
 */
#include "emme_pic_oracle.h"

#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef double _Complex cplx;

double emme_shim_cyl_bessel_j(double nu, double x);
double emme_shim_cyl_bessel_i(double nu, double x);

#define BATCH_COUNT 256 /* include/solver_pic.h:252 */

struct emme_pic_oracle {
    emme_pic_oracle_params p;
    long n;
    int nf;
    double cell_width;
    double *eta, *v_para, *v_perp;
    cplx* weight;
    double *omega_dv, *omega_st, *p_weight, *j0;
    cplx* dc_pb;
    double* coef;
    cplx* field;
    cplx* k[3];
    cplx* buffer;
};

static const double COEF[4][4] = {{1, 0.62653829327080},
                                  {0, 1, -0.55111240553326},
                                  {0, 1.5220585509963, -0.52205855099628, 0.92457411226246},
                                  {1., 0.13686116839369, -1.1368611683937}};

static double omega_d(const emme_pic_oracle_params* p, double eta) {
    return p->omega_d_bar * (cos(eta) + p->shat * eta * sin(eta));
}

static double omega_d_integral(const emme_pic_oracle_params* p, double eta, double v_para) {
    return (p->q * p->R / v_para) * p->omega_d_bar *
           (sin(eta) * (1. + p->shat) - p->shat * eta * cos(eta));
}

static double bound(const emme_pic_oracle_params* p, double eta) {
    eta = fmod(eta + p->length, 2 * p->length);
    return eta < 0 ? eta + p->length : eta - p->length;
}

static void locate(const emme_pic_oracle* s, double eta, size_t* idx, double* w) {
    *idx = (size_t)((eta + s->p.length) / s->cell_width);
    *w = (eta + s->p.length) / s->cell_width - *idx;
}

static void solve_field(emme_pic_oracle* s) {
    const emme_pic_oracle_params* p = &s->p;
    const size_t nf = (size_t)s->nf;
    const long batch_size = s->n / BATCH_COUNT, remain = s->n % BATCH_COUNT;
    for (long b = 0; b < BATCH_COUNT; ++b) {
        const long begin = b * batch_size + (b < remain ? b : remain);
        const long end = (b + 1) * batch_size + (b < remain ? b + 1 : remain);
        cplx* buf = s->buffer + (size_t)b * nf;
        for (size_t i = 0; i < nf; ++i) buf[i] = 0;
        for (long i = begin; i < end; ++i) {
            const double eta = s->eta[i];
            const double x_perp = s->v_perp[i] / p->vt;
            const double sb = sqrt(p->b_theta * (1. + pow(p->shat * eta, 2)));
            s->j0[i] = emme_shim_cyl_bessel_j(0, x_perp * sb);
            s->dc_pb[i] = cexp(CMPLX(0., -omega_d_integral(p, eta, s->v_para[i]) * s->omega_dv[i]));
            const cplx den = p->drift_center_transformation_switch ? s->j0[i] * s->weight[i] * s->dc_pb[i]
                                                                   : s->j0[i] * s->weight[i];
            size_t idx;
            double w;
            locate(s, eta, &idx, &w);
            buf[idx] += den * (1. - w);
            buf[(idx + 1) % nf] += den * w;
        }
    }
    for (size_t i = 0; i < nf; ++i) s->field[i] = 0;
    for (long b = 0; b < BATCH_COUNT; ++b)
        for (size_t i = 0; i < nf; ++i) s->field[i] += s->buffer[(size_t)b * nf + i];
    for (size_t i = 0; i < nf; ++i) s->field[i] *= s->coef[i];
}

static void put_velocity(emme_pic_oracle* s, cplx* vs) {
    const emme_pic_oracle_params* p = &s->p;
    const size_t nf = (size_t)s->nf;
    const cplx* field = s->field;
    for (long i = 0; i < s->n; ++i) {
        const double eta = s->eta[i], v_para = s->v_para[i];
        const double x_perp = s->v_perp[i] / p->vt;
        const double sb = sqrt(p->b_theta * (1. + pow(p->shat * eta, 2)));
        const double dj0 =
            -p->b_theta * p->shat * p->shat * x_perp * eta * emme_shim_cyl_bessel_j(1, x_perp * sb) / sb;
        size_t c;
        double w;
        locate(s, eta, &c, &w);
        const cplx phi = (1. - w) * field[c] + w * field[(c + 1) % nf];
        const cplx dphi = ((1. - w) * (field[(c + 1) % nf] - field[(c + nf - 1) % nf]) +
                           w * (field[(c + 2) % nf] - field[c])) /
                          (2. * s->cell_width);
        const double omega_dv = s->omega_dv[i], omega_st = s->omega_st[i], p_weight = s->p_weight[i],
                     j0 = s->j0[i];
        const cplx dc_pb = s->dc_pb[i];
        const cplx iu = CMPLX(0., 1.);
        if (p->drift_center_transformation_switch) {
            vs[i] = p_weight * conj(dc_pb) *
                    (iu * ((omega_st - omega_d(p, eta) * omega_dv) * j0 * phi) -
                     v_para / (p->q * p->R) * (j0 * dphi + dj0 * phi));
        } else {
            vs[i] = -s->weight[i] * omega_d(p, eta) * omega_dv * iu +
                    p_weight * (iu * ((omega_st - omega_d(p, eta) * omega_dv) * j0 * phi) -
                                v_para / (p->q * p->R) * (j0 * dphi + dj0 * phi));
        }
    }
}

emme_pic_oracle* emme_pic_oracle_create(const emme_pic_oracle_params* p, long n, const double* eta,
                                        const double* v_para, const double* v_perp,
                                        const double* weight) {
    emme_pic_oracle* s = (emme_pic_oracle*)calloc(1, sizeof(*s));
    s->p = *p;
    s->n = n;
    s->nf = p->npoints;
    s->cell_width = 2 * p->length / p->npoints;
    const size_t nf = (size_t)s->nf;
    s->eta = (double*)malloc(sizeof(double) * n);
    s->v_para = (double*)malloc(sizeof(double) * n);
    s->v_perp = (double*)malloc(sizeof(double) * n);
    s->weight = (cplx*)malloc(sizeof(cplx) * n);
    s->omega_dv = (double*)malloc(sizeof(double) * n);
    s->omega_st = (double*)malloc(sizeof(double) * n);
    s->p_weight = (double*)malloc(sizeof(double) * n);
    s->j0 = (double*)calloc(n, sizeof(double));
    s->dc_pb = (cplx*)calloc(n, sizeof(cplx));
    s->coef = (double*)malloc(sizeof(double) * nf);
    s->field = (cplx*)calloc(nf, sizeof(cplx));
    s->buffer = (cplx*)calloc(BATCH_COUNT * nf, sizeof(cplx));
    for (int k = 0; k < 3; ++k) s->k[k] = (cplx*)calloc(n, sizeof(cplx));
    memcpy(s->eta, eta, sizeof(double) * n);
    memcpy(s->v_para, v_para, sizeof(double) * n);
    memcpy(s->v_perp, v_perp, sizeof(double) * n);
    memcpy(s->weight, weight, sizeof(cplx) * n);
    /* initialize_marker_extras */
    for (long i = 0; i < n; ++i) {
        const double vp = v_para[i], vq = v_perp[i];
        s->omega_dv[i] = (vp * vp + .5 * vq * vq) / (2. * p->vt * p->vt);
        s->omega_st[i] = p->omega_s_i * (1. + p->eta_i * ((vp * vp + vq * vq) / (2. * p->vt * p->vt) - 1.5));
        s->p_weight[i] = vq * exp(-(vp * vp * (1 - p->water_bag_weight_vpara) +
                                    vq * vq * (1 - p->water_bag_weight_vperp)) /
                                  (2 * p->vt * p->vt));
    }
    double sum = 0;
    for (long i = 0; i < n; ++i) sum += s->p_weight[i];
    const double inn = 2 * p->length / (sum);
    for (long i = 0; i < n; ++i) s->p_weight[i] = s->p_weight[i] * inn;
    /* cal_quasi_neutrality_coef */
    for (size_t idx = 0; idx < nf; ++idx) {
        const double b = p->b_theta * (1. + pow(p->shat * (idx * s->cell_width - p->length), 2));
        double g0 = emme_shim_cyl_bessel_i(0, b) * exp(-b);
        s->coef[idx] = 1. / ((1. + 1. / p->tau - g0) * s->cell_width);
    }
    return s;
}

void emme_pic_oracle_destroy(emme_pic_oracle* s) {
    if (!s) return;
    free(s->eta); free(s->v_para); free(s->v_perp); free(s->weight);
    free(s->omega_dv); free(s->omega_st); free(s->p_weight); free(s->j0); free(s->dc_pb);
    free(s->coef); free(s->field); free(s->buffer);
    for (int k = 0; k < 3; ++k) free(s->k[k]);
    free(s);
}

void emme_pic_oracle_step(emme_pic_oracle* s, double dt) {
    const emme_pic_oracle_params* p = &s->p;
    for (int st = 0; st < 3; ++st) {
        put_velocity(s, s->k[st]);
        const double h = COEF[st][st + 1] * dt;
        for (long i = 0; i < s->n; ++i) {
            /* the fold (... + coef[p][k] * intermediates[k]) of Integrator::step, left to right */
            cplx v = COEF[st][0] * s->k[0][i];
            for (int k = 1; k <= st; ++k) v = v + COEF[st][k] * s->k[k][i];
            s->eta[i] = bound(p, s->eta[i] + s->v_para[i] * h / (p->q * p->R));
            s->weight[i] += v * h;
        }
        solve_field(s);
    }
}

void emme_pic_oracle_field(const emme_pic_oracle* s, double* field) {
    memcpy(field, s->field, sizeof(cplx) * (size_t)s->nf);
}

void emme_pic_oracle_markers(const emme_pic_oracle* s, double* eta, double* weight) {
    memcpy(eta, s->eta, sizeof(double) * s->n);
    memcpy(weight, s->weight, sizeof(cplx) * s->n);
}

void emme_pic_oracle_extras(const emme_pic_oracle* s, double* omega_dv, double* omega_st,
                            double* p_weight, double* coef) {
    memcpy(omega_dv, s->omega_dv, sizeof(double) * s->n);
    memcpy(omega_st, s->omega_st, sizeof(double) * s->n);
    memcpy(p_weight, s->p_weight, sizeof(double) * s->n);
    memcpy(coef, s->coef, sizeof(double) * (size_t)s->nf);
}

void emme_pic_oracle_calculate_omega(const double* stats, long size, double dt, double* omega_re,
                                     double* omega_im) {
    const size_t n = (size_t)size / 2;
    double t = 0, weighted_sum = 0, sum = 0;
    for (size_t i = n; i < (size_t)size; ++i) {
        const double val = log(stats[3 * i + 2]);
        weighted_sum += val * t;
        sum += val;
        t += dt;
    }
    const double gamma = 6 * (2 * weighted_sum - dt * sum * (n + 1)) / (dt * dt * n * (n * n - 1));
    const size_t m = (size_t)size - n;
    double* rl = (double*)malloc(sizeof(double) * (m ? m : 1));
    for (size_t i = n; i < (size_t)size; ++i) rl[i - n] = log(fabs(stats[3 * i]));
    size_t count = 0;
    double first = 0, last = 0;
    for (size_t i = 1; i + 1 < m; ++i) {
        if (rl[i] > rl[i - 1] && rl[i] > rl[i + 1]) {
            if (!count) first = i * dt;
            last = i * dt;
            ++count;
        }
    }
    free(rl);
    double omega = 0;
    if (count > 1) omega = M_PI * (count - 1) / (last - first);
    *omega_re = omega;
    *omega_im = gamma;
}
