"""ctypes loader for the CPU oracle (oracle/emme_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package (emme_b200) never does.
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "oracle" / "_ref" / "libemme_oracle.so"
REF_DRIVER = ROOT / "oracle" / "_ref" / "ref_driver"


class OracleParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("q", "R", "vt", "tau", "beta_e", "eta_i", "eta_e", "omega_s_i", "omega_s_e",
                 "omega_d_bar", "arc_coeff", "tol", "prec")] + [("maxdepth", C.c_int), ("order", C.c_int)]


def build():
    if not LIB.exists() or LIB.stat().st_mtime < (ROOT / "oracle" / "emme_oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "port"], check=True,
                       stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB))
        dp = C.POINTER(C.c_double)
        lp = C.POINTER(C.c_long)
        L.emme_oracle_assemble.argtypes = [C.POINTER(OracleParams), C.c_int, dp, dp, dp, C.c_double,
                                           C.c_double, C.c_double, dp, C.c_int, C.c_int, C.c_int, lp]
        L.emme_oracle_assemble.restype = None
        L.emme_oracle_kappa.argtypes = [C.POINTER(OracleParams), C.c_uint] + [C.c_double] * 8 + [dp, dp, lp]
        L.emme_oracle_kappa.restype = None
        L.emme_oracle_kappa_e.argtypes = [C.POINTER(OracleParams), C.c_uint] + [C.c_double] * 6 + [dp, dp]
        L.emme_oracle_kappa_e.restype = None
        L.emme_oracle_bessel_i_alter.argtypes = [C.c_double, C.c_double, dp, C.POINTER(C.c_int)]
        L.emme_oracle_bessel_i_alter.restype = None
        L.emme_oracle_weight.argtypes = [C.c_int, C.c_int, C.c_int]
        L.emme_oracle_weight.restype = C.c_double
        L.emme_oracle_grid.argtypes = [C.c_double, C.c_int, dp]
        L.emme_oracle_grid.restype = C.c_double
        L.emme_oracle_trace_step.argtypes = [C.c_int, dp, dp, dp, dp]
        L.emme_oracle_trace_step.restype = C.c_int
        L.emme_oracle_secant.argtypes = [C.c_long, dp, dp, C.c_double, C.c_double, dp]
        L.emme_oracle_secant.restype = None
        FN = C.CFUNCTYPE(None, C.c_double, C.c_void_p, dp, dp)
        L.emme_oracle_integrate.argtypes = [FN, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int,
                                            dp, dp, lp]
        L.emme_oracle_integrate.restype = C.c_int
        L.FN = FN
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def make_params(d):
    """d: dict with the OracleParams field names."""
    p = OracleParams()
    for n, _t in OracleParams._fields_:
        setattr(p, n, d[n])
    return p


def assemble(pd, eta, g, bi, dx, omega, rows=None, nthreads=0):
    """A(omega) as a (dim, dim) complex128 array + stats dict (oracle restatement of
    EigenSolver::matrixAssembler, include/solver.h:417-515)."""
    L = lib()
    p = make_params(pd)
    N = len(eta)
    dim = N if pd["beta_e"] == 0.0 else 2 * N
    out = np.zeros((dim, dim), dtype=np.complex128)
    stats = (C.c_long * 4)()
    r0, r1 = rows if rows else (0, N)
    eta = np.ascontiguousarray(eta, dtype=np.float64)
    g = np.ascontiguousarray(g, dtype=np.float64)
    bi = np.ascontiguousarray(bi, dtype=np.float64)
    L.emme_oracle_assemble(C.byref(p), N, _dp(eta), _dp(g), _dp(bi), dx, omega.real, omega.imag,
                           _dp(out.view(np.float64)), r0, r1, nthreads, stats)
    return out, dict(integrals=stats[0], evals=stats[1], fwd=stats[2], bwd=stats[3])


def trace_step(A, Ad):
    """delta = -1/trace(A^-1 Ad); inputs are copied (include/solver.h:129-140)."""
    L = lib()
    A = np.array(A, dtype=np.complex128, order="C")
    Ad = np.array(Ad, dtype=np.complex128, order="C")
    dr, di = C.c_double(), C.c_double()
    info = L.emme_oracle_trace_step(A.shape[0], _dp(A.view(np.float64)), _dp(Ad.view(np.float64)),
                                    C.byref(dr), C.byref(di))
    return complex(dr.value, di.value), info


def secant(A, Aold, delta):
    L = lib()
    A = np.ascontiguousarray(A, dtype=np.complex128)
    Aold = np.ascontiguousarray(Aold, dtype=np.complex128)
    out = np.empty_like(A)
    L.emme_oracle_secant(A.size, _dp(A.view(np.float64)), _dp(Aold.view(np.float64)), delta.real,
                         delta.imag, _dp(out.view(np.float64)))
    return out


def bessel_i_alter(z):
    L = lib()
    out = (C.c_double * 8)()
    trips = (C.c_int * 2)()
    L.emme_oracle_bessel_i_alter(z.real, z.imag, out, trips)
    return [complex(out[2 * k], out[2 * k + 1]) for k in range(4)], (trips[0], trips[1])


def integrate(fn, tol, prec, maxdepth, order):
    """util::integrate semi-infinite front end applied to a Python callable x -> complex."""
    L = lib()

    def tramp(x, _ctx, re, im):
        v = complex(fn(x))
        re[0] = v.real
        im[0] = v.imag

    cb = L.FN(tramp)
    re, im, ev = C.c_double(), C.c_double(), C.c_long()
    rc = L.emme_oracle_integrate(cb, None, tol, prec, maxdepth, order, C.byref(re), C.byref(im),
                                 C.byref(ev))
    if rc:
        raise RuntimeError("integration_start_points should be 15 or 31")
    return complex(re.value, im.value), ev.value


def have_ref_driver():
    return REF_DRIVER.exists() and os.access(REF_DRIVER, os.X_OK)


# ---- PIC method (row N4): oracle/emme_pic_oracle.c ------------------------------------------
PIC_LIB = ROOT / "oracle" / "_ref" / "libemme_pic_oracle.so"
PIC_DRIVER = ROOT / "oracle" / "_ref" / "pic_driver"


class PicOracleParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("q", "R", "vt", "tau", "shat", "b_theta", "length", "eta_i", "omega_s_i", "omega_d_bar",
                 "water_bag_weight_vpara", "water_bag_weight_vperp")] + [
        ("npoints", C.c_int), ("drift_center_transformation_switch", C.c_int)]


_pic_lib = None


def pic_lib():
    global _pic_lib
    if _pic_lib is None:
        src = ROOT / "oracle" / "emme_pic_oracle.c"
        if not PIC_LIB.exists() or PIC_LIB.stat().st_mtime < src.stat().st_mtime:
            subprocess.run(["make", "-C", str(ROOT / "oracle"), "port"], check=True, stdout=subprocess.DEVNULL)
        L = C.CDLL(str(PIC_LIB))
        dp = C.POINTER(C.c_double)
        L.emme_pic_oracle_create.restype = C.c_void_p
        L.emme_pic_oracle_create.argtypes = [C.POINTER(PicOracleParams), C.c_long, dp, dp, dp, dp]
        L.emme_pic_oracle_destroy.argtypes = [C.c_void_p]
        L.emme_pic_oracle_step.argtypes = [C.c_void_p, C.c_double]
        L.emme_pic_oracle_field.argtypes = [C.c_void_p, dp]
        L.emme_pic_oracle_markers.argtypes = [C.c_void_p, dp, dp]
        L.emme_pic_oracle_extras.argtypes = [C.c_void_p, dp, dp, dp, dp]
        L.emme_pic_oracle_calculate_omega.argtypes = [dp, C.c_long, C.c_double, dp, dp]
        L.emme_shim_cyl_bessel_j.restype = C.c_double
        L.emme_shim_cyl_bessel_j.argtypes = [C.c_double, C.c_double]
        _pic_lib = L
    return _pic_lib


class PicOracle:
    """The plain-C restatement of PIC_State + Integrator (include/solver_pic.h)."""

    def __init__(self, params, eta, v_para, v_perp, weight):
        self.L = pic_lib()
        self.p = PicOracleParams(**{k: params[k] for k, _ in PicOracleParams._fields_})
        self.n = len(eta)
        self.nf = self.p.npoints
        eta, v_para, v_perp = (np.ascontiguousarray(a, dtype=np.float64) for a in (eta, v_para, v_perp))
        weight = np.ascontiguousarray(weight, dtype=np.complex128)
        self.h = self.L.emme_pic_oracle_create(C.byref(self.p), self.n, _dp(eta), _dp(v_para), _dp(v_perp),
                                               _dp(weight.view(np.float64)))

    def step(self, dt):
        self.L.emme_pic_oracle_step(self.h, dt)

    def field(self):
        f = np.empty(self.nf, dtype=np.complex128)
        self.L.emme_pic_oracle_field(self.h, _dp(f.view(np.float64)))
        return f

    def markers(self):
        eta = np.empty(self.n)
        w = np.empty(self.n, dtype=np.complex128)
        self.L.emme_pic_oracle_markers(self.h, _dp(eta), _dp(w.view(np.float64)))
        return eta, w

    def extras(self):
        a, b, c = (np.empty(self.n) for _ in range(3))
        coef = np.empty(self.nf)
        self.L.emme_pic_oracle_extras(self.h, _dp(a), _dp(b), _dp(c), _dp(coef))
        return a, b, c, coef

    def __del__(self):
        if getattr(self, "h", None):
            self.L.emme_pic_oracle_destroy(self.h)
            self.h = None


def pic_calculate_omega(stats, dt):
    stats = np.ascontiguousarray(stats, dtype=np.float64)
    re, im = C.c_double(), C.c_double()
    pic_lib().emme_pic_oracle_calculate_omega(_dp(stats), stats.shape[0], dt, C.byref(re), C.byref(im))
    return complex(re.value, im.value)


def pic_field_stats(fields):
    """Per-step diagnostics of src/main.cpp:110-117 in the reference's accumulation order."""
    out = np.empty((fields.shape[0], 3))
    for k, f in enumerate(fields):
        re = im = nrm = 0.0
        for v in f:
            re += v.real
            im += v.imag
            nrm += v.real * v.real + v.imag * v.imag
        out[k] = (re / f.size, im / f.size, np.sqrt(nrm / f.size))
    return out
