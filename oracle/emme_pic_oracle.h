/* oracle/emme_pic_oracle.h -- CPU restatement of the reference's PIC method (row N4).
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  It is never linked into libemme_b200.so and the
 * product has no CPU fallback.
 *
 * Parity status: PINNED -- checked against per-step field dumps of the unmodified reference
 * (include/solver_pic.h compiled by oracle/Makefile into oracle/_ref/pic_driver with a fixed
 * seed; fixtures tests/golden/pic_*.npz; tests/test_oracle.py::test_pic_*).  The reference's own
 * tests hold no vectors for this path (test/test_integrator.cpp exercises the RK3 Integrator on
 * a harmonic oscillator and no longer compiles).
 */
#ifndef EMME_PIC_ORACLE_H
#define EMME_PIC_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

/* The members of Parameters the PIC state reads (include/solver_pic.h:62-67 and below). */
typedef struct emme_pic_oracle_params {
    double q, R, vt, tau, shat, b_theta, length;
    double eta_i, omega_s_i, omega_d_bar;
    double water_bag_weight_vpara, water_bag_weight_vperp;
    int npoints;
    int drift_center_transformation_switch;
} emme_pic_oracle_params;

typedef struct emme_pic_oracle emme_pic_oracle;

/* PIC_State constructor (include/solver_pic.h:62-67) from given markers: eta, v_para, v_perp are
 * n doubles, weight n complex128 (re, im).  Extras (initialize_marker_extras :207-238), the
 * quasi-neutrality table (:381-399) and the zero field (:239-243) are derived here. */
emme_pic_oracle* emme_pic_oracle_create(const emme_pic_oracle_params* p, long n, const double* eta,
                                        const double* v_para, const double* v_perp,
                                        const double* weight);
void emme_pic_oracle_destroy(emme_pic_oracle* s);
/* Integrator::step (include/solver_pic.h:423-434): three stages of put_velocity + update. */
void emme_pic_oracle_step(emme_pic_oracle* s, double dt);
/* nf complex128 */
void emme_pic_oracle_field(const emme_pic_oracle* s, double* field);
void emme_pic_oracle_markers(const emme_pic_oracle* s, double* eta, double* weight);
void emme_pic_oracle_extras(const emme_pic_oracle* s, double* omega_dv, double* omega_st,
                            double* p_weight, double* coef);
/* util::calculate_omega (include/solver_pic.h:475-529), stats = n x {mean re, mean im, rms}. */
void emme_pic_oracle_calculate_omega(const double* stats, long n, double dt, double* omega_re,
                                     double* omega_im);

#ifdef __cplusplus
}
#endif
#endif
