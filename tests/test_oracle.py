"""Pin the CPU oracle (oracle/emme_oracle.c) against the UNMODIFIED reference.

The reference's own tests hold no vectors for the eigen path, so the
pins are fixtures dumped from the compiled reference (tests/golden/make_goldens.py, ref_driver).
The C restatement sequences its arithmetic like the reference, so with the same compiler and
libm (this image) agreement is BIT-FOR-BIT; a 4-ulp tolerance is the fallback if libm differs.
"""
import math

import numpy as np
import pytest

import cases
import oracle_lib as O
from emme_b200 import Input

SMALL = ["c1_n32", "c1_n64", "c1_gk31_n128", "c1_em_n64", "c1_pos_n64", "c1_cyl_n64", "c1_tmd_n64",
         "c1_cylold_n64", "c3_n32", "c3_n64"]


def close_ulps(a, b, ulps=4):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= ulps * np.spacing(np.maximum(np.abs(a), np.abs(b))))


def test_kat_gauss_kronrod(golden):
    kat = golden["kat"]
    for order in (15, 31):
        v, _ = O.integrate(lambda x: math.exp(-x), 1e-6, 1e-6, 100, order)
        assert [v.real, v.imag] == kat[f"gk{order}_exp"]
        v, _ = O.integrate(lambda x: complex(math.exp(-x * x), math.exp(-x) * math.sin(3 * x)),
                           1e-8, 1e-12, 100, order)
        assert [v.real, v.imag] == kat[f"gk{order}_gauss_osc"]
        assert abs(v.real - math.sqrt(math.pi) / 2) < 1e-9 and abs(v.imag - 0.3) < 2e-10
        # complex exponential: python's cmath.exp is not glibc's cexp -> rounding tolerance
        import cmath
        f3 = lambda x: cmath.exp(complex(-0.3, 2.0) * x) / (1.0 + x)
        v, _ = O.integrate(f3, 1e-5, 1e-2, 20, order)
        assert abs(v - complex(*kat[f"gk{order}_cexp"])) < 1e-13
        v, _ = O.integrate(f3, 1e-10, 1e-14, 30, order)
        assert abs(v - complex(*kat[f"gk{order}_cexp_tight"])) < 1e-13
    with pytest.raises(RuntimeError, match="should be 15 or 31"):
        O.integrate(lambda x: 0.0, 1e-6, 1e-6, 10, 21)


def test_kat_bessel(golden):
    kat = golden["kat"]
    k = 0
    while f"bessel_{k}_z" in kat:
        z = complex(*kat[f"bessel_{k}_z"])
        out, trips = O.bessel_i_alter(z)
        for c in range(4):
            ref = complex(*kat[f"bessel_{k}_{c}"])
            assert out[c] == ref, (k, c, out[c], ref)
        # scaled I0, I1 against scipy: y0/mu * exp(...) identity  I0(z) e^{-|Re z| sign} = y0/mu
        from scipy.special import ive
        zz = z if z.real >= 0 else -z
        # sanity only: the reference's Miller start/threshold rule is accurate to ~1e-9
        # the helper scales by exp(-zz) (complex), scipy's ive by exp(-|Re zz|)
        ph = np.exp(-1j * zz.imag)
        assert abs(out[0] / out[2] - ive(0, zz) * ph) < 1e-8 * abs(ive(0, zz))
        assert abs(out[1] / out[2] - (1 if z.real >= 0 else -1) * ive(1, zz) * ph) < 1e-8 * abs(ive(1, zz))
        k += 1
    assert k == 9


def test_kat_weights_and_grid(golden):
    kat = golden["kat"]
    L = O.lib()
    assert [L.emme_oracle_weight(12, 3, j) for j in range(12)] == kat["sh12_row3"]
    assert [L.emme_oracle_weight(12, 0, j) for j in range(12)] == kat["sh12_row0"]
    eta = np.empty(16)
    dx = L.emme_oracle_grid(20.0, 16, eta.ctypes.data_as(O.C.POINTER(O.C.c_double)))
    assert dx == kat["grid_20_16_dx"]
    assert eta.tolist() == kat["grid_20_16"]


@pytest.mark.parametrize("case", SMALL)
def test_assembly_matches_reference_bitwise(case, golden):
    inp = Input(cases.input_path(case))
    p, n = inp.params()
    eta, g, bi = cases.ref_tables(case)          # tables dumped from the reference itself
    w = complex(*golden["assemble"][case]["omega"])
    A, st = O.assemble(cases.oracle_params(p), eta, g, bi, p.dx, w)
    ref = cases.ref_matrix(case)
    assert A.shape == ref.shape
    if not np.array_equal(A.view(np.float64), ref.view(np.float64)):
        assert close_ulps(A.view(np.float64), ref.view(np.float64)), "oracle deviates from reference"
    nm = 3 if p.beta_e != 0 else 1
    assert st["integrals"] == n * (n - 1) // 2 * nm


def test_full_size_rows_match_reference(golden):
    """C1 at its real size (N=1024): the first rows of A(omega0) against the reference dump."""
    z = np.load(cases.GOLD / "rows_c1.npz")
    inp = Input(cases.input_path("c1"))
    p, n = inp.params()
    eta, g, bi = cases.ref_tables("c1")
    w = complex(*golden["assemble"]["c1"]["omega"])
    A, _ = O.assemble(cases.oracle_params(p), eta, g, bi, p.dx, w, rows=(0, 2))
    rows = z["rows"].tolist()
    for r in (0, 1):
        ref = z["data"][rows.index(r)]
        assert np.array_equal(A[r, r:].view(np.float64), ref[r:].view(np.float64))


@pytest.mark.parametrize("case", ["c1_n64", "c1_em_n64"])
def test_newton_iterates_match_reference(case, golden):
    """Seed + iterate rule of EigenSolver (include/solver.h:396-415,113-160) restated with the
    oracle pieces; the dense step is an LU instead of zsysv, so iterates agree to rounding."""
    rec = golden["newton"][case]
    inp = Input(cases.input_path(case))
    p, n = inp.params()
    pd = cases.oracle_params(p)
    eta, g, bi = cases.ref_tables(case)
    w0 = inp.initial_guess()
    w = 0.99 * w0
    dw = 0.01 * w0
    Aold, _ = O.assemble(pd, eta, g, bi, p.dx, w)
    w += dw
    assert [w.real, w.imag, dw.real, dw.imag] == rec["seed"]
    A, _ = O.assemble(pd, eta, g, bi, p.dx, w)
    Ad = O.secant(A, Aold, dw)
    for it in rec["iterates"][:4]:
        dw, info = O.trace_step(A, Ad)
        assert info == 0
        w += dw
        ref_w, ref_d = complex(it[0], it[1]), complex(it[2], it[3])
        assert abs(w - ref_w) <= 1e-12 * abs(ref_w)
        assert abs(dw - ref_d) <= 1e-9 * abs(ref_d) + 1e-15
        Aold = A
        A, _ = O.assemble(pd, eta, g, bi, p.dx, w)
        Ad = O.secant(A, Aold, dw)
