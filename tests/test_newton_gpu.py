"""-m gpu: kernel 2 (dense step) and the Newton/secant iterate sequence against the reference."""
import numpy as np
import pytest

import cases
from emme_b200 import EigenSolver, Input, solve_once_eigen

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim", [7, 32, 33, 100, 256, 257, 640])
def test_trace_delta_matches_numpy(dim, native_lib):
    """delta = -1/trace(A^-1 A') on random, badly row-scaled complex matrices that need pivoting."""
    rng = np.random.default_rng(dim)
    A = rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))
    A[::3] *= 1e-3                     # forces row interchanges
    A[0, 0] = 0.0                      # zero leading pivot
    B = rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))
    inp = Input(cases.input_path("c1_n32"))
    p, _ = inp.params()
    n = dim
    s = EigenSolver(p, n, np.linspace(-1, 1, n), np.zeros(n), np.ones(n))
    d = s.trace_delta(A, B)
    ref = -1.0 / np.trace(np.linalg.solve(A, B))
    assert abs(d - ref) <= 1e-10 * abs(ref), (d, ref)


def _sym_case(dim, seed=0, general_rhs=False):
    """Complex SYMMETRIC, diagonally strong A (the shape of EMME's matrices) and a right-hand side."""
    rng = np.random.default_rng(1000 * seed + dim)
    S = (rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))) * (0.4 / np.sqrt(dim))
    A = S + S.T + np.diag(2.0 + 0.3 * rng.standard_normal(dim) + 0.2j * rng.standard_normal(dim))
    B = rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))
    if not general_rhs:
        B = B + B.T
    return A, B


def _trace_solver(dim):
    inp = Input(cases.input_path("c1_n32"))
    p, _ = inp.params()
    return EigenSolver(p, dim, np.linspace(-1, 1, dim), np.zeros(dim), np.ones(dim))


@pytest.mark.parametrize("dim", [2, 5, 31, 32, 33, 64, 65, 100, 191, 256, 257, 640, 1000])
def test_symmetric_path_matches_numpy(dim, native_lib):
    """Symmetric path of kernel 2 (A = L D L^T, trace from the explicit inverse): taken for a
    symmetric, diagonally strong matrix, and delta agrees with numpy's LU solve.  Ragged sizes
    exercise every tile edge of the masked GEMMs."""
    A, B = _sym_case(dim)
    s = _trace_solver(dim)
    d = s.trace_delta(A, B)
    ref = -1.0 / np.trace(np.linalg.solve(A, B))
    st = s.stats()
    assert st["sym_steps"] == 1 and st["pivot_fallbacks"] == 0, st
    assert abs(d - ref) <= 1e-12 * abs(ref), (d, ref)
    # a second call replays the captured CUDA graph
    d2 = s.trace_delta(A, B)
    assert d2 == d and s.stats()["sym_steps"] == 2


@pytest.mark.parametrize("dim", [2100, 2304])
def test_symmetric_path_two_level_blocking(dim, native_lib):
    """dim > 2048 switches kernel 2 to 128-wide outer blocks (ragged and aligned sizes)."""
    A, B = _sym_case(dim, seed=1)
    s = _trace_solver(dim)
    d = s.trace_delta(A, B)
    ref = -1.0 / np.trace(np.linalg.solve(A, B))
    assert s.stats()["sym_steps"] == 1
    assert abs(d - ref) <= 1e-12 * abs(ref), (d, ref)


def test_symmetric_path_general_rhs(native_lib):
    """The contraction sum_ij (A^-1)_ij B_ji does not need a symmetric right-hand side."""
    A, B = _sym_case(300, seed=2, general_rhs=True)
    s = _trace_solver(300)
    d = s.trace_delta(A, B)
    ref = -1.0 / np.trace(np.linalg.solve(A, B))
    assert s.stats()["sym_steps"] == 1
    assert abs(d - ref) <= 1e-12 * abs(ref), (d, ref)


def test_symmetric_path_is_verified(native_lib):
    """The symmetric path checks its own preconditions on the device: one asymmetric entry (last
    bit) sends the step to the LU path, a symmetric matrix that needs interchanges to the pivoting
    LU; both still give the right delta."""
    A, B = _sym_case(200, seed=3)
    s = _trace_solver(200)
    A1 = A.copy()
    A1[150, 3] = np.nextafter(A1[150, 3].real, 10.0) + 1j * A1[150, 3].imag
    d = s.trace_delta(A1, B)
    st = s.stats()
    assert st["sym_steps"] == 0 and st["pivot_fallbacks"] == 0, st
    assert abs(d + 1.0 / np.trace(np.linalg.solve(A1, B))) <= 1e-12 * abs(d)
    A2 = A.copy()
    A2[0, 0] = 1e-9                      # symmetric, but partial pivoting must interchange row 0 (first pivot: nothing has been added to it yet)
    d = s.trace_delta(A2, B)
    st = s.stats()
    assert st["sym_steps"] == 0 and st["pivot_fallbacks"] == 1, st
    assert abs(d + 1.0 / np.trace(np.linalg.solve(A2, B))) <= 1e-10 * abs(d)
    d = s.trace_delta(A, B)
    assert s.stats()["sym_steps"] == 1
    assert abs(d + 1.0 / np.trace(np.linalg.solve(A, B))) <= 1e-12 * abs(d)


def test_symmetric_and_lu_paths_agree_on_real_matrices(golden, native_lib, monkeypatch):
    """Same EMME matrices through both paths of kernel 2 (EMME_DENSE_SYM=0 disables the symmetric one)."""
    inp = Input(cases.input_path("c1_em_n64"))
    s = EigenSolver.from_input(inp)
    A = s.matrixAssembler(-0.8 + 0.25j).copy()
    A2 = s.matrixAssembler(-0.79 + 0.251j).copy()
    Ad = (A - A2) / (0.01 - 0.001j)
    d_sym = s.trace_delta(A, Ad)
    assert s.stats()["sym_steps"] == 1
    monkeypatch.setenv("EMME_DENSE_SYM", "0")
    s2 = EigenSolver.from_input(inp)
    d_lu = s2.trace_delta(A, Ad)
    assert s2.stats()["sym_steps"] == 0
    ref = -1.0 / np.trace(np.linalg.solve(A, Ad))
    assert abs(d_sym - ref) <= 1e-12 * abs(ref) and abs(d_lu - ref) <= 1e-12 * abs(ref), (d_sym, d_lu, ref)


def test_trace_delta_matches_oracle_on_real_matrices(golden, native_lib):
    import oracle_lib as O
    inp = Input(cases.input_path("c1_n128"))
    s = EigenSolver.from_input(inp)
    A = cases.ref_matrix("c1_n128")
    A2 = s.matrixAssembler(-0.79 + 0.251j)
    Ad = (A - A2) / (0.01 - 0.001j)
    d = s.trace_delta(A, Ad)
    ref, info = O.trace_step(A, Ad)
    assert info == 0
    assert abs(d - ref) <= 1e-12 * abs(ref), (d, ref)


def test_singular_matrix_reports_info(native_lib):
    from emme_b200 import EmmeError
    inp = Input(cases.input_path("c1_n32"))
    p, _ = inp.params()
    n = 64
    s = EigenSolver(p, n, np.linspace(-1, 1, n), np.zeros(n), np.ones(n))
    A = np.eye(n, dtype=np.complex128)
    A[10] = 0
    with pytest.raises(EmmeError, match="Linear solve failed") as ei:
        s.trace_delta(A, np.eye(n, dtype=np.complex128))
    assert ei.value.code == 11


@pytest.mark.parametrize("case", ["c1_n64", "c1_n128", "c1_gk31_n128", "c1_em_n64", "c1_pos_n64", "c1", "c3"])
def test_newton_iterates_match_reference(case, golden, native_lib):
    """Same seeds, same step formula, same stop rule (include/solver.h:396-415,113-160;
    src/main.cpp:43-57): every iterate and the converged omega within 1e-8 relative."""
    rec = golden["newton"][case]
    inp = Input(cases.input_path(case))
    w0 = inp.initial_guess()
    w, iters, s = solve_once_eigen(inp, w0)
    print(f"\n[newton] {case}: {len(iters)} iterates, omega={w!r}, ref={rec['final']}, stats={s.stats()}")
    assert len(iters) == len(rec["iterates"]) == rec["final"][2]
    worst = 0.0
    for (wi, di), r in zip(iters, rec["iterates"]):
        rw = complex(r[0], r[1])
        worst = max(worst, abs(wi - rw) / abs(rw))
    print(f"[newton] {case}: worst iterate rel err {worst:.3e}")
    assert worst <= 1e-8
    rf = complex(rec["final"][0], rec["final"][1])
    assert abs(w - rf) <= 1e-8 * abs(rf)


def _qr_delta_lapack(A, Ad):
    """newtonQRSecantIteration's formula (include/solver.h:246-370) with LAPACK itself: scipy's
    qr(pivoting=True) is zgeqp3."""
    import scipy.linalg as sl
    n = A.shape[0]
    Q, R, piv = sl.qr(A, pivoting=True)
    x = sl.solve_triangular(R[:n - 1, :n - 1], R[:n - 1, n - 1])
    v = np.zeros(n, dtype=np.complex128)
    v[piv[:n - 1]] = -x
    v[piv[n - 1]] = 1.0
    t = Q.conj().T @ (Ad @ v)
    return -R[n - 1, n - 1] / t[n - 1], int(piv[n - 1])


@pytest.mark.parametrize("dim", [2, 3, 17, 64, 100, 129, 300])
def test_qr_delta_matches_lapack(dim, native_lib):
    """The QR-secant dense step on random complex matrices with well separated column norms
    (so that the pivot order is not decided by rounding) against LAPACK's zgeqp3 path."""
    rng = np.random.default_rng(40 + dim)
    A = rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))
    A *= rng.uniform(0.2, 5.0, dim)[None, :]
    B = rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))
    s = _trace_solver(dim)
    d = s.qr_delta(A, B)
    ref, _ = _qr_delta_lapack(A, B)
    assert abs(d - ref) <= 1e-10 * abs(ref), (d, ref)


def test_qr_delta_on_real_matrices(golden, native_lib):
    """Same, on an EMME matrix pair (all column norms within a few per cent of each other)."""
    inp = Input(cases.input_path("c1_n128"))
    s = EigenSolver.from_input(inp)
    A = cases.ref_matrix("c1_n128")
    A2 = s.matrixAssembler(-0.79 + 0.251j).copy()
    Ad = (A - A2) / (0.01 - 0.001j)
    d = s.qr_delta(A, Ad)
    ref, last = _qr_delta_lapack(A, Ad)
    print(f"\n[qr] delta={d!r} lapack={ref!r} last pivot column {last}")
    assert abs(d - ref) <= 1e-9 * abs(ref), (d, ref)


@pytest.mark.parametrize("case", ["c1_n64", "c1_n128", "c1_gk31_n128", "c1_em_n64", "c1_pos_n64"])
def test_qr_newton_iterates_match_reference(case, golden, native_lib):
    """iteration_method != "TraceSecant": EigenSolver::newtonQRSecantIteration
    (include/solver.h:210-383) against the iterate list of the compiled reference."""
    rec = golden["newton_qr"][case]
    inp = Input(text=cases.input_path(case).read_text().replace('"TraceSecant"', '"QRSecant"'))
    w, iters, s = solve_once_eigen(inp, inp.initial_guess())
    print(f"\n[newton-qr] {case}: {len(iters)} iterates, omega={w!r}, ref={rec['final']}, dense_ms={s.stats()['dense_ms']:.2f}")
    assert len(iters) == len(rec["iterates"]) == rec["final"][2]
    worst = 0.0
    for (wi, di), r in zip(iters, rec["iterates"]):
        rw = complex(r[0], r[1])
        worst = max(worst, abs(wi - rw) / abs(rw))
    print(f"[newton-qr] {case}: worst iterate rel err {worst:.3e}")
    # The bar is the converged eigenvalue (1e-8 relative).  Intermediate iterates of this method can
    # wander far from the root (c1_pos_n64: 12 iterates, |delta| ~ 1) and the secant quotient
    # (A - A_old)/delta amplifies rounding-level differences of the assembled matrices by ~1e3 per
    # iterate until the iteration contracts again: measured 1e-14, 4e-14, 5e-11, 3e-10, 1e-8,
    # 6e-8 (max), then back down to 2e-11 at convergence.  So: first iterate tight, every iterate
    # on the reference's trajectory, converged value to the bar.
    first = complex(*rec["iterates"][0][:2])
    assert abs(iters[0][0] - first) <= 1e-10 * abs(first)
    assert worst <= 1e-6
    rf = complex(rec["final"][0], rec["final"][1])
    assert abs(w - rf) <= 1e-8 * abs(rf)


def test_step_before_seed_is_an_error(native_lib):
    from emme_b200 import EmmeError, capi
    inp = Input(cases.input_path("c1_n32"))
    s = EigenSolver.from_input(inp)
    with pytest.raises(EmmeError) as ei:
        s.newtonTraceSecantIteration()
    assert ei.value.code == capi.E_STATE


def test_null_space_matches_svd(native_lib):
    """nullSpace (include/solver.h:58-112): right singular vector of the smallest singular value,
    compared with numpy's SVD of the same matrix up to the arbitrary complex phase."""
    inp = Input(cases.input_path("c1_n128"))
    w, iters, s = solve_once_eigen(inp, inp.initial_guess())
    A = s.eigen_matrix
    v = s.nullSpace()
    _, sv, vh = np.linalg.svd(A)
    ref = np.conj(vh[-1])
    assert abs(np.linalg.norm(v) - 1) < 1e-12
    overlap = abs(np.vdot(ref, v))
    print(f"\n[nullspace] sigma_min={sv[-1]:.3e} sigma_2={sv[-2]:.3e} |<ref,v>|={overlap:.15f}")
    assert overlap > 1 - 1e-9
    assert np.linalg.norm(A @ v) <= 1.0000001 * sv[-1] + 1e-14
    k = np.argmax(np.abs(v))
    assert abs(v[k].imag) < 1e-14 and v[k].real > 0


def test_host_program_end_to_end(tmp_path, golden, native_lib):
    """The C++ `emme` program on input.json -> output.json + eigenMatrics/eigenMatrix.bin."""
    import re
    import subprocess
    from emme_b200 import build
    build.build_all()
    (tmp_path / "input.json").write_text(cases.input_path("c1_n64").read_text())
    (tmp_path / "eigenMatrics").mkdir()
    r = subprocess.run([str(build.EXE)], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    rec = golden["newton"]["c1_n64"]
    assert r.stdout.count("        (") == rec["final"][2]                 # one line per iterate
    out = (tmp_path / "output.json").read_text()
    m = re.search(r'"eigenvalue": \[\s*([-0-9.e]+),\s*([-0-9.e]+)', out)
    assert abs(float(m.group(1)) - rec["final"][0]) < 1e-5 and abs(float(m.group(2)) - rec["final"][1]) < 1e-5
    assert '"scan_key": "(None)"' in out and '"eigenvector"' in out
    A = np.fromfile(tmp_path / "eigenMatrics" / "eigenMatrix.bin", dtype=np.complex128).reshape(64, 64)
    assert np.array_equal(A, A.T) and np.all(np.diag(A) == 2.0)


def test_pivoting_fallback_is_taken_and_correct(native_lib):
    """A matrix that needs row interchanges: the optimistic factorisation must raise its flag and
    the step must be repeated with partial pivoting (stats.pivot_fallbacks counts it)."""
    rng = np.random.default_rng(5)
    n = 200
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    B = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    inp = Input(cases.input_path("c1_n32"))
    p, _ = inp.params()
    s = EigenSolver(p, n, np.linspace(-1, 1, n), np.zeros(n), np.ones(n))
    d = s.trace_delta(A, B)
    ref = -1.0 / np.trace(np.linalg.solve(A, B))
    assert abs(d - ref) <= 1e-10 * abs(ref)
    assert s.stats()["pivot_fallbacks"] == 1
    D = A * 0.01 + 3 * np.eye(n)
    d = s.trace_delta(D, B)
    assert abs(d + 1.0 / np.trace(np.linalg.solve(D, B))) <= 1e-12 * abs(d)
    assert s.stats()["pivot_fallbacks"] == 1          # diagonally dominant: no fallback


def test_host_program_scan_matches_reference_program(tmp_path, native_lib):
    """A 4-point two-sided scan through the C++ `emme` program against output.json of the UNMODIFIED
    reference program (tests/golden/scan_c1_n32.json: continuation of omega between points, the
    turning point, eigenMatrics file names; the reference prints 6 significant digits)."""
    import json
    import re
    import subprocess
    from emme_b200 import build
    build.build_all()
    gold = json.loads((cases.GOLD / "scan_c1_n32.json").read_text())
    (tmp_path / "input.json").write_text(cases.input_path("c1_scan_n32").read_text())
    (tmp_path / "eigenMatrics").mkdir()
    r = subprocess.run([str(build.EXE)], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    out = (tmp_path / "output.json").read_text()
    vals = [float(v) for v in re.findall(r'"scan_value": ([-0-9.e+]+)', out)]
    evs = [(float(a), float(b)) for a, b in re.findall(r'"eigenvalue": \[\s*([-0-9.e+]+),\s*([-0-9.e+]+)', out)]
    assert vals == gold["scan_values"]
    assert len(evs) == len(gold["eigenvalues"])
    for (a, b), (ra, rb) in zip(evs, gold["eigenvalues"]):
        assert abs(a - ra) <= 2e-6 * max(1, abs(ra)) and abs(b - rb) <= 2e-6 * max(1, abs(rb)), (evs, gold)
    assert sorted(p.name for p in (tmp_path / "eigenMatrics").iterdir()) == gold["files"]


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_trace_delta_tiny(dim, native_lib):
    rng = np.random.default_rng(100 + dim)
    A = rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim)) + 3 * np.eye(dim)
    B = rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))
    inp = Input(cases.input_path("c1_n32"))
    p, _ = inp.params()
    n = max(dim, 2)
    s = EigenSolver(p, n, np.linspace(-1, 1, n), np.zeros(n), np.ones(n))
    if dim == 1:
        pytest.skip("npoints >= 2 is the smallest mesh (Grid divides by npoints-1)")
    d = s.trace_delta(A, B)
    ref = -1.0 / np.trace(np.linalg.solve(A, B))
    assert abs(d - ref) <= 1e-12 * abs(ref)


def test_scan_parallel_single_rank_matches_reference_scan(native_lib):
    """solve_scan_parallel on one rank: independent points with explicit starts reproduce the
    eigenvalues of the reference program's continuation scan (same roots, 6 printed digits)."""
    import json
    from emme_b200 import parallel
    gold = json.loads((cases.GOLD / "scan_c1_n32.json").read_text())
    base = cases.input_path("c1_n32").read_text()
    starts = [complex(*e) * (1 + 1e-3) for e in gold["eigenvalues"]]      # near each root
    recs = parallel.solve_scan_parallel(base, "omega_d_coeff", gold["scan_values"], starts)
    assert [r["scan_value"] for r in recs] == gold["scan_values"]
    for r, (ra, rb) in zip(recs, gold["eigenvalues"]):
        assert r["converged"], r
        assert abs(r["eigenvalue"][0] - ra) < 2e-6 and abs(r["eigenvalue"][1] - rb) < 2e-6, (r, ra, rb)
    bad = parallel.solve_scan_parallel(base.replace('"tokamak"', '"torus"'), "omega_d_coeff", [1.0], -0.8 + 0.25j)
    assert bad[0]["eigenvalue"] == "NaN" and "not supported" in bad[0]["reason"]


def test_async_matrix_download_overlaps_and_matches(native_lib):
    """emme_copy_matrix_async: the download of eigen_matrix into pinned memory is ordered after the
    iterate that produced it and may overlap the next iterate; the bytes equal the blocking copy."""
    import torch
    from emme_b200 import capi
    inp = Input(cases.input_path("c1_n128"))
    s = EigenSolver.from_input(inp)
    s.seed(inp.initial_guess())
    s.newtonTraceSecantIteration()
    want = s.eigen_matrix.copy()
    pin = torch.empty((s.dim, s.dim, 2), dtype=torch.float64).pin_memory()
    capi.check(s._lib.emme_copy_matrix_async(s._h, 0, pin.data_ptr()))
    s.newtonTraceSecantIteration()          # overlaps the copy; must not disturb it
    s.newtonTraceSecantIteration()          # overwrites the copied buffer: waits for the copy
    capi.check(s._lib.emme_copy_wait(s._h))
    got = pin.numpy().view(np.complex128).reshape(s.dim, s.dim)
    assert np.array_equal(got, want)


def test_iterating_past_convergence_is_an_error_not_a_fault(native_lib):
    """30 iterates on one point go far past convergence: A == A_old, A' = 0, the
    trace vanishes and -1/trace is not finite.  The reference divides unchecked
    (include/solver.h:139) and fails later in LAPACK; here the step that produces the non-finite
    delta returns EMME_E_NONFINITE ("Linear solve failed ..."), nothing is assembled at a
    non-finite omega, no CUDA error occurs and the handle stays usable."""
    from emme_b200 import EmmeError, capi
    inp = Input(cases.input_path("c1_n64"))
    s = EigenSolver.from_input(inp)
    s.seed(inp.initial_guess())
    tol = inp.number("iteration_precision")
    converged_at, failed_at = None, None
    for k in range(30):
        try:
            s.newtonTraceSecantIteration()
        except EmmeError as e:
            assert e.code == capi.E_NONFINITE, (e.code, str(e))
            assert "Linear solve failed" in str(e)
            failed_at = k
            break
        if converged_at is None and abs(s.d_eigen_value) < abs(tol * s.eigen_value):
            converged_at = k
    assert converged_at is not None and converged_at <= 8
    assert failed_at is not None and failed_at > converged_at, (converged_at, failed_at)
    assert not np.isfinite(abs(s.eigen_value))           # omega was updated like the reference does
    with pytest.raises(EmmeError):                        # and nothing assembles at that omega
        s.newtonTraceSecantIteration()
    # the device is healthy: a fresh seed on the same handle reproduces the reference iterates
    w, its, _ = solve_once_eigen(inp, inp.initial_guess(), solver=s)
    assert abs(its[-1][1]) < abs(tol * w)


@pytest.mark.parametrize("dim,path", [(96, "sym"), (96, "lu"), (300, "lu")])
def test_non_finite_matrix_reports_an_error(dim, path, native_lib):
    """A NaN anywhere in A must come back as EMME_E_NONFINITE from every path of kernel 2: the
    pivot search of the pivoting LU orders NaN candidates (it used to leave the search without a
    winner and use 0x7fffffff as a row index)."""
    from emme_b200 import EmmeError, capi
    A, B = _sym_case(dim, seed=5)
    if path == "lu":
        A = A + np.triu(np.ones((dim, dim)), 1) * 1e-3          # not symmetric: LU paths
    A[dim // 2, :] = np.nan
    A[:, dim // 2] = np.nan
    s = _trace_solver(dim)
    with pytest.raises(EmmeError) as ei:
        s.trace_delta(A, B)
    assert ei.value.code == capi.E_NONFINITE, (ei.value.code, str(ei.value))
    A2, B2 = _sym_case(dim, seed=6)                               # same handle, healthy afterwards
    d = s.trace_delta(A2, B2)
    assert abs(d + 1.0 / np.trace(np.linalg.solve(A2, B2))) <= 1e-11 * abs(d)


def test_zero_rhs_reports_an_error(native_lib):
    """A' = 0 (two identical assemblies): trace = 0 exactly, -1/0 is reported, not returned."""
    from emme_b200 import EmmeError, capi
    A, _ = _sym_case(128, seed=7)
    s = _trace_solver(128)
    with pytest.raises(EmmeError) as ei:
        s.trace_delta(A, np.zeros_like(A))
    assert ei.value.code == capi.E_NONFINITE


@pytest.mark.parametrize("dim", [2304, 3000])
def test_blocked_symmetric_path_independent_of_outer_block(dim, native_lib, monkeypatch):
    """The blocked (outer block > 32) symmetric path forms M = L^-1 per outer block through
    M_KK; different outer widths are different summation orders of the same quantity."""
    A, B = _sym_case(dim, seed=8)
    ref = -1.0 / np.trace(np.linalg.solve(A, B))
    got = []
    for nbo in ("64", "128", "256"):
        monkeypatch.setenv("EMME_DENSE_NBO", nbo)
        s = _trace_solver(dim)
        d = s.trace_delta(A, B)
        assert s.stats()["sym_steps"] == 1
        assert abs(d - ref) <= 1e-12 * abs(ref), (nbo, d, ref)
        got.append(d)
        s.close()
    monkeypatch.setenv("EMME_DENSE_NBO", "0")
    _trace_solver(4).close()                 # restore the default outer block (process-wide setting)


@pytest.mark.parametrize("k", [0, 9, 27, 54, 63])
def test_c5_points_match_reference(k, native_lib):
    """BASELINE configs[4] (64 independent wavenumbers, N=1024): the iterate lists of the unmodified
    reference for five of the points (tests/golden/c5.json from make_c5_goldens.py) -- a point that
    uses all 21 iterates without converging (k=0), a 10-iterate wandering trajectory (k=9), the
    4-iterate neighbour of C1 (k=27), and two points on which the reference itself fails at iterate 3
    with "Linear solve failed" (k=54, 63), which the scan records as NaN (src/main.cpp:311-318)."""
    import json
    from emme_b200 import EmmeError, workloads
    gold = json.loads((cases.GOLD / "c5.json").read_text())["points"][str(k)]
    k_rho, w0, txt = workloads.c5_point(k)
    assert k_rho == gold["k_rho"] and [w0.real, w0.imag] == gold["omega0"]
    inp = Input(text=txt)
    s = EigenSolver.from_input(inp)
    if "final" in gold:
        w, its, _ = solve_once_eigen(inp, w0, solver=s)
        assert len(its) == gold["final"][2]
        for (wi, di), g in zip(its, gold["iterates"]):
            wg = complex(g[0], g[1])
            # a wandering trajectory amplifies rounding-level matrix differences through the secant
            # quotient; the converged value is what the 1e-8 bar is about
            assert abs(wi - wg) <= 1e-6 * abs(wg), (k, wi, wg)
        wf = complex(gold["final"][0], gold["final"][1])
        assert abs(w - wf) <= 1e-8 * abs(wf), (w, wf)
    else:
        s.seed(w0)
        for n_ok, g in enumerate(gold["iterates"]):
            s.newtonTraceSecantIteration()
            wg = complex(g[0], g[1])
            assert abs(s.eigen_value - wg) <= 1e-7 * abs(wg), (k, n_ok, s.eigen_value, wg)
        # the next dense step is the one the reference's zsysv refuses: the matrix assembled at the
        # last iterate holds NaN entries (overflow of the reference's Bessel recurrence, reproduced)
        A = s.eigen_matrix
        nan_at = sorted(map(tuple, np.argwhere(np.isnan(A)).tolist()))
        assert len(nan_at) >= 2 and all((j, i) in nan_at for i, j in nan_at), nan_at
        if k == 63:      # the pairs found by assembling with the reference itself
            assert nan_at == [(132, 842), (181, 891), (842, 132), (891, 181)], nan_at
        with pytest.raises(EmmeError) as ei:
            s.newtonTraceSecantIteration()
        assert "Linear solve failed" in str(ei.value)
        recs = parallel_scan_one(txt, w0)
        assert recs[0]["eigenvalue"] == "NaN"
    s.close()


def parallel_scan_one(txt, w0):
    from emme_b200 import parallel
    return parallel.solve_scan_texts([txt], [w0])


@pytest.mark.parametrize("dim", [2304, 3000])
def test_chain_stream_is_bitwise_neutral(dim, native_lib, monkeypatch):
    """Blocked symmetric path with the panel chain on its own high-priority stream (default) against
    the same launches on one stream: identical delta, graph replay included."""
    A, B = _sym_case(dim, seed=21)
    got = {}
    for la in ("0", "1"):
        monkeypatch.setenv("EMME_DENSE_LOOKAHEAD", la)
        s = _trace_solver(dim)
        got[la] = (s.trace_delta(A, B), s.trace_delta(A, B))
        assert s.stats()["sym_steps"] == 2
        s.close()
    assert got["0"][0] == got["0"][1] == got["1"][0] == got["1"][1], got
    ref = -1.0 / np.trace(np.linalg.solve(A, B))
    assert abs(got["1"][0] - ref) <= 1e-12 * abs(ref)
