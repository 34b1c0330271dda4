"""CPU emulation of kernel 1's group logic (tests/emul) against the reference goldens.

This checks the CUDA kernel's ALGEBRA (hoisted pair constants, 2*lambda and conj(e)/t identities,
reciprocal-based complex divisions, diagonal-major item decoding, reference-order panel sums)
with glibc's libm; device rounding is covered by the -m gpu tests.
"""
import ctypes as C
import subprocess

import numpy as np
import pytest

import cases
import parity
from emme_b200 import Input, capi

EMUL_DIR = cases.ROOT / "tests" / "emul"
EMUL_LIB = EMUL_DIR / "_build" / "libemul.so"


@pytest.fixture(scope="module")
def emul():
    src = EMUL_DIR / "emul_assembly.cpp"
    deps = [src] + list((cases.ROOT / "emme_b200" / "csrc").glob("*.h")) + \
        list((cases.ROOT / "emme_b200" / "csrc").glob("*.cuh"))
    if not EMUL_LIB.exists() or any(d.stat().st_mtime > EMUL_LIB.stat().st_mtime for d in deps):
        EMUL_LIB.parent.mkdir(exist_ok=True)
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fopenmp",
                        "-ffp-contract=off", "-o", str(EMUL_LIB), str(src)], check=True)
    L = C.CDLL(str(EMUL_LIB))
    dp = C.POINTER(C.c_double)
    L.emul_assemble.argtypes = [C.POINTER(capi.EmmeParams), C.c_int, dp, dp, dp, C.c_double, C.c_double,
                                dp, C.POINTER(C.c_ulonglong)]
    L.emul_assemble.restype = C.c_int
    return L


@pytest.mark.parametrize("case", ["c1_n32", "c1_n64", "c1_gk31_n128", "c1_em_n64", "c1_pos_n64",
                                  "c1_cyl_n64", "c1_tmd_n64", "c1_cylold_n64", "c3_n32", "c3_n64"])
def test_kernel_algebra_matches_reference(case, emul, golden, native_lib):
    inp = Input(cases.input_path(case))
    p, n = inp.params()
    eta, g, bi = inp.tables()
    w = complex(*golden["assemble"][case]["omega"])
    em = p.beta_e != 0
    dim = 2 * n if em else n
    out = np.zeros((dim, dim), dtype=np.complex128)
    st = (C.c_ulonglong * 8)()
    dp = C.POINTER(C.c_double)
    rc = emul.emul_assemble(C.byref(p), n, eta.ctypes.data_as(dp), g.ctypes.data_as(dp),
                            bi.ctypes.data_as(dp), w.real, w.imag,
                            out.view(np.float64).ctypes.data_as(dp), st)
    assert rc == 0
    ref = cases.ref_matrix(case)
    c = parity.assert_parity(out, ref, em=em, label=case)
    assert np.array_equal(out, out.T) or em      # ES matrix is exactly symmetric
    assert c["median_rel"] < 1e-14


def test_lean_cexp(emul):
    """cexp_lean (the straight-line exp/sincos of the integrand's exponential factor) against
    numpy in extended precision: <= 3 ulp on each component (CUDA documents 1 ulp for exp and 2 ulp for sincos) over the argument range the kernel
    sees (Re in [-40, 60], |Im| up to 5e5 including values next to multiples of pi/2)."""
    rng = np.random.default_rng(7)
    n = 400000
    a = rng.uniform(-40.0, 60.0, n)
    b = np.concatenate([rng.uniform(-50.0, 50.0, n // 2), rng.uniform(-5e5, 5e5, n // 4),
                        (rng.integers(-2000, 2000, n // 4) * (np.pi / 2)) * (1 + rng.uniform(-1e-9, 1e-9, n // 4))])
    out = np.zeros(2 * n)
    dp = C.POINTER(C.c_double)
    emul.emul_cexp.argtypes = [dp, dp, C.c_int, dp]
    emul.emul_cexp(a.ctypes.data_as(dp), b.ctypes.data_as(dp), n, out.ctypes.data_as(dp))
    la, lb = a.astype(np.longdouble), b.astype(np.longdouble)
    er = np.exp(la)
    ref_re, ref_im = er * np.cos(lb), er * np.sin(lb)
    for got, ref in ((out[0::2], ref_re), (out[1::2], ref_im)):
        ulp = np.spacing(np.abs(ref.astype(np.float64)))
        err = np.abs(got.astype(np.longdouble) - ref) / ulp
        assert float(err.max()) <= 3.0, float(err.max())


def test_reference_nan_corner_is_reproduced(emul, native_lib):
    """C5 point 63 (k_rho = 0.68) at the iterate where the reference fails: for the pairs (132, 842)
    and (181, 891) of the N=1024 mesh one quadrature node has lambda ~ 0, |z| ~ 3500, the
    reference's Bessel recurrence overflows and its safe_exp zero times inf is NaN (the entry is NaN,
    zsysv reports a singular D, the scan records "NaN").  The kernel's arithmetic -- which skips the
    recurrence behind the underflow guard -- must give NaN for the same pairs and finite values for
    their neighbours, like the oracle."""
    import oracle_lib as O
    from emme_b200 import workloads
    _, _, txt = workloads.c5_point(63)
    inp = Input(text=txt)
    p, n = inp.params()
    eta, g, bi = inp.tables()
    w = complex(-1.9458107725918097, -0.571341647480763)
    dp = C.POINTER(C.c_double)
    for (i, j), want_nan in (((132, 842), True), ((181, 891), True), ((131, 841), False), ((133, 843), False)):
        e2, g2, b2 = (np.ascontiguousarray(t[[i, j]]) for t in (eta, g, bi))
        out = np.zeros((2, 2), dtype=np.complex128)
        st = (C.c_ulonglong * 8)()
        rc = emul.emul_assemble(C.byref(p), 2, e2.ctypes.data_as(dp), g2.ctypes.data_as(dp), b2.ctypes.data_as(dp),
                                w.real, w.imag, out.view(np.float64).ctypes.data_as(dp), st)
        assert rc == 0
        ref, _ = O.assemble(cases.oracle_params(p), e2, g2, b2, p.dx, w)
        assert bool(np.isnan(ref[0, 1])) == want_nan, (i, j, ref)
        assert bool(np.isnan(out[0, 1])) == want_nan, (i, j, out)
        if not want_nan:
            assert abs(out[0, 1] - ref[0, 1]) <= 1e-10 * abs(ref[0, 1]), (out, ref)
