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
