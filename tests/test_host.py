"""Host layer: JSON dialect, Parameters/tables bit-identity, scan generator, C-ABI surface."""
import ctypes as C
import re

import numpy as np
import pytest

import cases
from emme_b200 import EigenSolver, EmmeError, Input, capi, scan_values

ALL_TABLE_CASES = ["c1", "c3", "c1_n32", "c1_n64", "c1_n128", "c1_n256", "c1_n512", "c1_gk31_n128",
                   "c1_em_n64", "c1_em_n128", "c1_pos_n64", "c1_cyl_n64", "c1_tmd_n64",
                   "c1_cylold_n64", "c3_n32", "c3_n64", "c3_n128", "c3_n256"]


def test_every_declared_symbol_is_exported(native_lib):
    hdr = (cases.ROOT / "include" / "emme_b200.h").read_text()
    declared = set(re.findall(r"\b(emme_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(capi.PROTOTYPES), declared ^ set(capi.PROTOTYPES)
    for name in declared:
        assert hasattr(native_lib, name)
    assert b"sm_100a" in native_lib.emme_version()


@pytest.mark.parametrize("case", ALL_TABLE_CASES)
def test_tables_bit_identical_to_reference(case, golden, native_lib):
    """eta_i, g_integration_f(eta_i), bi(eta_i) and the derived constants equal what the
    compiled reference computes (fixtures from ref_driver tables), for all five geometries."""
    inp = Input(cases.input_path(case))
    p, n = inp.params()
    eta, g, bi = inp.tables()
    r_eta, r_g, r_bi = cases.ref_tables(case)
    assert n == r_eta.size
    assert np.array_equal(eta, r_eta)
    assert np.array_equal(bi, r_bi)
    assert np.array_equal(g, r_g), float(np.abs(g - r_g).max())
    info = golden["tables"][case]
    assert p.dx == info["dx"]
    assert p.omega_s_i == info["omega_s_i"]
    assert p.omega_s_e == info["omega_s_e"]
    assert p.omega_d_bar == info["omega_d_bar"]


def test_json_dialect_quirks(native_lib):
    # test/test.json of the reference and test_json.cpp's expectations (a=42.42, bs[0]=1)
    txt = '''{ "a":42.42, "bs": [1,2,{},4], "obj": {"one":1,"two":2,"empty_array":[]},
               "float": -3.25e-9, "answer" :"forty two", "primitives": [true, false, null] }'''
    inp = Input(text=txt)
    assert inp.number("a") == 42.42
    assert inp.number("bs[0]") == 1
    assert inp.number("float") == -3.25e-9
    assert inp.string("answer") == "forty two"
    # a number token without '.' is an INTEGER read with atoi: 1e-6 -> 1 (JsonParser.cpp:436-444)
    assert Input(text='{"x": 1e-6, "y": 1.0e-6, "z": 10}').number("x") == 1
    assert Input(text='{"x": 1e-6, "y": 1.0e-6, "z": 10}').number("y") == 1.0e-6
    with pytest.raises(EmmeError, match="Failed to accessing key: nope"):
        inp.number("nope")
    with pytest.raises(EmmeError, match=r"Incorrect JSON type, requires one of: ValueCategory::NumberFloat, "
                                         r"ValueCategory::NumberInt, actually: ValueCategory::String"):
        inp.number("answer")
    with pytest.raises(EmmeError, match="Incorrect JSON type, requires: ValueCategory::String"):
        inp.string("a")
    with pytest.raises(EmmeError, match="error: unrecognized token"):
        Input(text='{"a": @}')
    with pytest.raises(EmmeError, match="error: unexpected content"):
        Input(text='{"a": 1 "b": 2}')
    with pytest.raises(EmmeError, match="File /nonexistent/input.json not found"):
        Input("/nonexistent/input.json")


def test_missing_keys_and_bad_conf(native_lib):
    # the shipped stellarator example lacks keys the constructor needs
    txt = cases.input_path("c3").read_text().replace('"epsilon_r":0.0,', "")
    with pytest.raises(EmmeError, match="Failed to accessing key: epsilon_r"):
        Input(text=txt).params()
    txt = cases.input_path("c1_n32").read_text().replace('"tokamak"', '"torus"')
    with pytest.raises(EmmeError, match="Input configuration not supported yet."):
        Input(text=txt).params()


def test_scan_object_collapses_to_head(native_lib):
    inp = Input(cases.input_path("c3_n32"))       # beta_e is a scan object with head 0.02
    p, n = inp.params()
    assert p.beta_e == 0.02 and n == 32
    inp = Input(cases.input_path("c1_scan"))      # omega_d_coeff scan, head 1.01
    assert inp.number("omega_d_coeff") == 1.01
    inp.set_number("omega_d_coeff", 0.5)
    assert inp.number("omega_d_coeff") == 0.5


def test_scan_generator_sequences():
    # input-example.json: 11 continuation-chained values 1.01, 0.91 ... 0.01, no right branch
    vals = scan_values(1.01, 0.1, [0.01, 1.01])
    assert len(vals) == 11 and not any(t for _, t in vals)
    assert abs(vals[0][0] - 1.01) < 1e-15 and abs(vals[-1][0] - 0.01) < 1e-12
    # input-stellarator-example.json: the single value 0.02
    assert scan_values(0.02, -0.001, [0.02, 0.02]) == [(0.02, False)]
    # two-sided scan with a turning point
    vals = scan_values(1.0, 0.25, [0.5, 1.5])
    assert [round(v, 12) for v, _ in vals] == [1.0, 0.75, 0.5, 1.25, 1.5]
    assert [t for _, t in vals] == [False, False, False, True, False]
    # scalar tail: the other tail is head + 0.5*copysign(step, head - tail) -> nothing on that side
    vals = scan_values(1.0, 0.25, 0.5)
    assert [round(v, 12) for v, _ in vals] == [1.0, 0.75, 0.5]


def test_argument_errors_are_lapack_style(native_lib):
    p = capi.EmmeParams()
    h = C.c_void_p()
    dp = C.POINTER(C.c_double)
    assert native_lib.emme_create(None, 8, None, None, None, 0, C.byref(h)) == -1
    arr = np.zeros(8)
    a = arr.ctypes.data_as(dp)
    assert native_lib.emme_create(C.byref(p), 1, a, a, a, 0, C.byref(h)) == -2
    assert native_lib.emme_create(C.byref(p), 8, None, a, a, 0, C.byref(h)) == -3
    p.integration_start_points = 21
    assert native_lib.emme_create(C.byref(p), 8, a, a, a, 0, C.byref(h)) == capi.E_BAD_ORDER
    assert b"should be 15 or 31" in native_lib.emme_last_error()


def test_no_cpu_fallback(native_lib):
    """Without a CUDA device the compute path must fail loudly, not fall back."""
    if native_lib.emme_device_count() > 0:
        pytest.skip("a CUDA device is present")
    inp = Input(cases.input_path("c1_n32"))
    with pytest.raises(EmmeError, match="no CPU fallback") as ei:
        EigenSolver.from_input(inp)
    assert ei.value.code == capi.E_NO_DEVICE


def _run_emme(tmp_path, input_text):
    import subprocess
    from emme_b200 import build
    build.build_all()
    (tmp_path / "input.json").write_text(input_text)
    (tmp_path / "eigenMatrics").mkdir(exist_ok=True)
    return subprocess.run([str(build.EXE)], cwd=tmp_path, capture_output=True, text=True, timeout=600)


def test_host_program_fails_loudly_without_gpu(tmp_path, native_lib):
    """The C++ host program (mirror of the reference's main) must not fall back to the CPU:
    a single solve aborts, a scan records every point as NaN with the reason (src/main.cpp:300-318)."""
    if native_lib.emme_device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = _run_emme(tmp_path, cases.input_path("c1_n32").read_text())
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
    scan = cases.input_path("c1_n32").read_text().replace(
        '"omega_d_coeff": 1.0', '"omega_d_coeff": {"head": 1.0, "step": 0.25, "tail": 0.5}')
    r = _run_emme(tmp_path, scan)
    assert r.returncode == 0, r.stderr
    out = Input(tmp_path / "output.json")          # our own reader parses what the writer wrote
    txt = (tmp_path / "output.json").read_text()
    assert txt.count('"eigenvalue": "NaN"') == 3 and "no CPU fallback" in txt
    assert '"scan_key": "omega_d_coeff"' in txt


def test_host_program_rejects_unknown_method(tmp_path, native_lib):
    r = _run_emme(tmp_path, cases.input_path("c1_n32").read_text().replace('"method": "eigen"', '"method": "fluid"'))
    assert r.returncode != 0 and "Method 'fluid' is not supported, yet." in r.stderr


def test_host_program_pic_fails_loudly_without_gpu(tmp_path, native_lib):
    """`"method": "PIC"` (row N4) goes to the device path; without a device it aborts."""
    if native_lib.emme_device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = _run_emme(tmp_path, cases.input_path("pic_n32").read_text())
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_workload_texts_parse_and_match_goldens(native_lib):
    """emme_b200/workloads.py: the input texts of the BASELINE configs parse with the host layer, the C4
    sweep keeps C1's physics, and the C5 points are the ones the goldens were generated from."""
    import json
    from emme_b200 import Input, workloads
    c1 = Input(workloads.C1_PATH)
    p1, n1 = c1.params()
    for n in workloads.C4_SIZES:
        inp = Input(text=workloads.c4_text(n))
        p, nn = inp.params()
        assert nn == n and p.q == p1.q and p.tau == p1.tau and p.omega_s_i == p1.omega_s_i
        assert abs(p.dx * (n - 1) - p1.dx * (n1 - 1)) <= 1e-12          # same domain, Grid::dx = 2L/(n-1)
    pts = workloads.c5_points()
    assert len(pts) == 64 and pts[0][0] == 0.05 and pts[63][0] == 0.68
    assert len({round(k, 2) for k, _, _ in pts}) == 64
    gold = json.loads((cases.GOLD / "c5.json").read_text())["points"]
    for k, g in gold.items():
        k_rho, w0, txt = workloads.c5_point(int(k))
        assert k_rho == g["k_rho"] and [w0.real, w0.imag] == g["omega0"]
        inp = Input(text=txt)
        assert inp.number("k_rho") == k_rho and inp.initial_guess() == w0
    assert workloads.workload_name(8192).startswith("C4 sweep point")
