"""-m gpu: kernel 1 (CUDA assembly through the C ABI) against the reference goldens and the oracle."""
import numpy as np
import pytest

import cases
import parity
from emme_b200 import EigenSolver, Input

pytestmark = pytest.mark.gpu

SMALL = ["c1_n32", "c1_n64", "c1_n128", "c1_gk31_n128", "c1_em_n64", "c1_pos_n64", "c1_cyl_n64",
         "c1_tmd_n64", "c1_cylold_n64", "c3_n32", "c3_n64", "c3_n128"]


def report(label, c, st=None):
    print(f"\n[parity] {label}: entries={c['entries']} strict_frac={c['strict_frac']:.6f} "
          f"max_rel={c['max_rel']:.3e} median_rel={c['median_rel']:.3e} max_abs={c['max_abs']:.3e} "
          f"max_abs/max|A|={c['max_abs_over_max']:.3e} n_rel>1e-9={c['n_rel_gt_1e9']}"
          + (f" stats={st}" if st else ""))


@pytest.mark.parametrize("case", SMALL)
def test_assembly_matches_reference_golden(case, golden, floor, native_lib):
    inp = Input(cases.input_path(case))
    s = EigenSolver.from_input(inp)
    w = complex(*golden["assemble"][case]["omega"])
    A = s.matrixAssembler(w)
    ref = cases.ref_matrix(case)
    em = s.dim != s.npoints
    c = parity.assert_parity(A, ref, em=em, label=case)
    report(case, c, s.stats())
    # not worse than a small multiple of the reference's own FMA/no-FMA self-deviation
    assert c["max_rel"] <= max(1e-10, 8 * floor[case]["max_rel"])
    assert c["median_rel"] < 2e-14
    st = s.stats()
    nm = 3 if em else 1
    assert st["integrals"] == s.npoints * (s.npoints - 1) // 2 * nm
    if not em:
        assert np.array_equal(A, A.T)
        assert np.all(np.diag(A) == 1.0 + 1.0 / s.params.tau)


@pytest.mark.parametrize("case,omega", [("c1_n64", -0.3 + 0.9j), ("c1_n64", 1.7 - 0.05j),
                                        ("c3_n32", -0.2 + 0.4j), ("c1_em_n64", 0.5 + 0.3j)])
def test_assembly_matches_oracle_at_other_omega(case, omega, native_lib):
    import oracle_lib as O
    inp = Input(cases.input_path(case))
    p, n = inp.params()
    eta, g, bi = inp.tables()
    s = EigenSolver.from_input(inp)
    A = s.matrixAssembler(omega)
    ref, ost = O.assemble(cases.oracle_params(p), eta, g, bi, p.dx, omega)
    c = parity.assert_parity(A, ref, em=p.beta_e != 0, label=f"{case}@{omega}")
    report(f"{case}@{omega}", c)
    # identical adaptive trees: same number of integrand evaluations as the oracle
    assert s.stats()["evals"] == ost["evals"]


@pytest.mark.parametrize("case", ["c1", "c3"])
def test_full_size_rows_and_properties(case, golden, native_lib):
    """BASELINE configs at their real size (N=1024): selected rows against the reference dump,
    plus size-independent properties (symmetry structure, diagonal, column sums, norm)."""
    z = np.load(cases.GOLD / f"rows_{case}.npz")
    inp = Input(cases.input_path(case))
    s = EigenSolver.from_input(inp)
    w = complex(*golden["assemble"][case]["omega"])
    A = s.matrixAssembler(w)
    st = s.stats()
    n, dim = s.npoints, s.dim
    em = dim != n
    rows = z["rows"]
    sub = A[rows]
    ref = z["data"]
    d = np.abs(sub - ref)
    mag = np.abs(ref)
    scale = float(z["maxabs"])
    rel = np.where(mag > 0, d / np.where(mag > 0, mag, 1), 0)
    print(f"\n[parity] {case} rows: max_rel={rel.max():.3e} median_rel={np.median(rel[mag > 0]):.3e} "
          f"strict_frac={(d <= 1e-10 * mag).mean():.6f} max_abs/max|A|={d.max() / scale:.3e} "
          f"assemble_ms={st['assemble_ms']:.2f} evals={st['evals']} stats={st}")
    assert (d <= 1e-10 * mag + parity.EPS * scale).all()
    assert (d <= 1e-10 * mag).mean() >= 0.99
    assert (d <= 1e-8 * mag + 64 * parity.EPS * scale).all()
    assert np.allclose(np.diag(A), z["diag"], rtol=0, atol=0)
    # BLAS dot-product summation order differs between hosts: ~sqrt(n)*eps on the norm itself
    assert abs(np.linalg.norm(A) - float(z["fro"])) <= 1e-11 * float(z["fro"])
    assert np.abs(A.sum(axis=0) - z["colsum"]).max() <= 1e-11 * scale
    if not em:
        assert np.array_equal(A, A.T)
    else:
        # block structure of include/solver.h:476-504
        a00, a01, a10, a11 = A[:n, :n], A[:n, n:], A[n:, :n], A[n:, n:]
        assert np.array_equal(a00, a00.T) and np.array_equal(a11, a11.T)
        assert np.array_equal(a01, -a01.T) and np.array_equal(a10, a01.T)


def test_sharded_assembly_sums_to_full(native_lib):
    """Multi-GPU building block: P shards of the work items, each into a zeroed buffer, sum to
    the single-launch matrix bit-for-bit (the shards are disjoint entries)."""
    import torch
    inp = Input(cases.input_path("c1_em_n64"))
    s = EigenSolver.from_input(inp)
    w = -0.8 + 0.25j
    full = s.matrixAssembler(w)
    acc = np.zeros_like(full)
    P = 3
    for r in range(P):
        buf = torch.zeros((s.dim, s.dim), dtype=torch.complex128, device="cuda")
        torch.cuda.synchronize()       # the handle launches on its own non-blocking stream
        s.assemble_device(w, buf.data_ptr(), r, P)
        acc += buf.cpu().numpy()
    assert np.array_equal(acc, full)


@pytest.mark.parametrize("case", ["c1_n64", "c1_em_n64", "c3_n64"])
def test_entries_do_not_depend_on_the_item_order(case, native_lib, monkeypatch):
    """Kernel 1's item order (far diagonals first where the host finds them costly,
    capi.cu::choose_item_order) only schedules: every order gives the same matrix bit for bit."""
    inp = Input(cases.input_path(case))
    _, n = inp.params()
    w = inp.initial_guess()
    mats = []
    for split in (None, "0", "0.5", "0.9", "0.04"):
        if split is None:
            monkeypatch.delenv("EMME_ASM_FAR_SPLIT", raising=False)
        else:
            monkeypatch.setenv("EMME_ASM_FAR_SPLIT", split)
        s = EigenSolver.from_input(inp)
        mats.append(s.matrixAssembler(w))
        assert s.stats()["integrals"] == n * (n - 1) // 2 * (3 if s.dim == 2 * n else 1)
        s.close()
    for m in mats[1:]:
        assert np.array_equal(m, mats[0])


def test_deep_stack_spills_to_global(native_lib):
    """Tight tolerances force deep bisection: the interval stack must follow the reference's
    contract (depth <= integration_iteration_limit), past the shared-memory slots."""
    import oracle_lib as O
    txt = cases.input_path("c1_n32").read_text()
    txt = txt.replace('"integration_precision": 1.0e-6', '"integration_precision": 1.0e-13')
    txt = txt.replace('"integration_accuracy": 1.0e-6', '"integration_accuracy": 1.0e-15')
    inp = Input(text=txt)
    p, n = inp.params()
    eta, g, bi = inp.tables()
    s = EigenSolver.from_input(inp)
    A = s.matrixAssembler(-0.8 + 0.25j)
    ref, ost = O.assemble(cases.oracle_params(p), eta, g, bi, p.dx, -0.8 + 0.25j)
    st = s.stats()
    print("\n[deep]", st, ost)
    c = parity.compare(A, ref)
    assert c["max_abs_over_max"] < 1e-13
    assert st["max_stack"] >= 8


def _variant(case, **subs):
    txt = cases.input_path(case).read_text()
    for old, new in subs.values():
        assert old in txt, old
        txt = txt.replace(old, new)
    return Input(text=txt)


@pytest.mark.parametrize("label,inp_args,omega", [
    ("n2", dict(n=('"npoints": 32', '"npoints": 2')), -0.8 + 0.25j),
    ("n3", dict(n=('"npoints": 32', '"npoints": 3')), -0.8 + 0.25j),
    ("n33_ragged", dict(n=('"npoints": 32', '"npoints": 33')), -0.8 + 0.25j),
    ("n17_em", dict(n=('"npoints": 32', '"npoints": 17'), b=('"beta_e": 0.00', '"beta_e": 0.01')), 0.4 + 0.2j),
    ("re_omega_zero", dict(), 0.0 + 0.3j),
    ("prec_zero", dict(p=('"integration_accuracy": 1.0e-6', '"integration_accuracy": 0.0')), -0.8 + 0.25j),
    ("maxdepth_2", dict(d=('"integration_iteration_limit": 100', '"integration_iteration_limit": 2')), -0.8 + 0.25j),
    ("maxdepth_0", dict(d=('"integration_iteration_limit": 100', '"integration_iteration_limit": 0')), -0.8 + 0.25j),
    ("gk31_loose", dict(o=('"integration_start_points": 15', '"integration_start_points": 31'),
                        t=('"integration_precision": 1.0e-6', '"integration_precision": 1.0e-2')), -0.8 + 0.25j),
])
def test_edge_cases_against_oracle(label, inp_args, omega, native_lib):
    """Smallest and ragged grids, Re(omega) = 0, zero absolute accuracy, depth limits 0 and 2: the CUDA
    path and the oracle must take the same adaptive decisions and agree to rounding."""
    import oracle_lib as O
    inp = _variant("c1_n32", **inp_args)
    p, n = inp.params()
    eta, g, bi = inp.tables()
    s = EigenSolver.from_input(inp)
    A = s.matrixAssembler(omega)
    ref, ost = O.assemble(cases.oracle_params(p), eta, g, bi, p.dx, omega)
    c = parity.assert_parity(A, ref, em=p.beta_e != 0, label=label)
    report(label, c)
    assert s.stats()["evals"] == ost["evals"], (s.stats(), ost)


def test_bad_quadrature_order_is_rejected(native_lib):
    from emme_b200 import EmmeError, capi
    inp = _variant("c1_n32", o=('"integration_start_points": 15', '"integration_start_points": 21'))
    with pytest.raises(EmmeError, match="integration_start_points should be 15 or 31") as ei:
        EigenSolver.from_input(inp)
    assert ei.value.code == capi.E_BAD_ORDER


def test_bench_workload_full_size(native_lib, monkeypatch):
    """BASELINE configs[3] at its largest point -- the bench workload, C1 physics with npoints=8192:
    three rows of A(omega) against the oracle (near, middle and far end of the mesh), the
    size-independent properties (exact symmetry, diagonal 1 + 1/tau, every entry written), and one
    Newton/secant iterate through BOTH paths of kernel 2 (symmetric L D L^T and LU) giving the
    same omega."""
    import oracle_lib as O
    inp = Input(cases.input_path("c1"))
    inp.set_number("npoints", 8192)
    p, n = inp.params()
    assert n == 8192
    eta, g, bi = inp.tables()
    w = -0.8 + 0.25j
    s = EigenSolver.from_input(inp)
    A = s.matrixAssembler(w)
    st = s.stats()
    assert st["integrals"] == n * (n - 1) // 2
    assert np.array_equal(A, A.T)
    assert np.all(np.diag(A) == 1.0 + 1.0 / p.tau)
    assert np.isfinite(A.view(np.float64)).all() and np.count_nonzero(A) == A.size
    scale = float(np.abs(A - np.diag(np.diag(A))).max())
    for r0 in (0, 4000, 8100):
        ref, _ = O.assemble(cases.oracle_params(p), eta, g, bi, p.dx, w, rows=(r0, r0 + 1))
        got, want = A[r0, r0 + 1:], ref[r0, r0 + 1:]
        d, mag = np.abs(got - want), np.abs(want)
        print(f"\n[parity] n=8192 row {r0}: max_rel={np.max(d / mag):.3e} strict_frac={(d <= 1e-10 * mag).mean():.6f}")
        assert (d <= 1e-10 * mag + parity.EPS * scale).all()
        assert (d <= 1e-10 * mag).mean() >= 0.99
        del ref
    del A
    s.seed(w)
    s.newtonTraceSecantIteration()
    assert s.stats()["sym_steps"] == 1 and s.stats()["pivot_fallbacks"] == 0
    w_sym = s.eigen_value
    s.close()
    monkeypatch.setenv("EMME_DENSE_SYM", "0")
    s2 = EigenSolver.from_input(inp)
    s2.seed(w)
    s2.newtonTraceSecantIteration()
    assert s2.stats()["sym_steps"] == 0 and s2.stats()["pivot_fallbacks"] == 0
    print(f"[newton] n=8192 first iterate: symmetric path {w_sym!r}, LU path {s2.eigen_value!r}")
    assert abs(w_sym - s2.eigen_value) <= 1e-11 * abs(w_sym)
    s2.close()
