"""The BASELINE.json configurations as input.json texts.

    C1  input-example.json, method=eigen, omega_d_coeff=1.0            tests/golden/inputs/c1.json
    C3  input-stellarator-example.json + the seven missing keys        tests/golden/inputs/c3.json
    C4  grid-size sweep: C1 physics with npoints in 512 .. 8192        c4_text(npoints)
    C5  parameter scan: 64 independent wavenumbers k_rho = 0.05 + 0.01 k, k = 0 .. 63, each with
        its own explicit start (no continuation chain, the reference's scan is sequential:
        src/main.cpp:78,263,302)                                        c5_points()

Shared by bench.py, the tests and tests/golden/make_c5_goldens.py so that every arm (B200,
reference, goldens) sees byte-identical inputs.  Number spelling follows the input files
(`1.0e-6`, decimal points kept) because of the reference's lexer rule (src/JsonParser.cpp:436-443).
"""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
C1_PATH = ROOT / "tests" / "golden" / "inputs" / "c1.json"
C3_PATH = ROOT / "tests" / "golden" / "inputs" / "c3.json"
PIC_PATH = ROOT / "tests" / "golden" / "inputs" / "pic.json"

C4_SIZES = (512, 1024, 2048, 4096, 8192)
C5_NPOINTS = 64
C1_K_RHO = 0.3182
C1_START = (-0.8, 0.25)


def _sub(txt, pattern, repl):
    txt, n = re.subn(pattern, repl, txt)
    assert n == 1, pattern
    return txt


def c4_text(npoints, k_rho=None):
    """C1 physics on an `npoints` mesh (the sweep of BASELINE configs[3])."""
    txt = _sub(C1_PATH.read_text(), r'"npoints": 1024', f'"npoints": {int(npoints)}')
    if k_rho is not None:
        txt = _sub(txt, r'"k_rho": 0.3182', f'"k_rho": {float(k_rho)!r}')
    return txt


def workload_name(npoints):
    """config.workload of the bench line -- the same string in the B200 and the reference arm."""
    return (f"C4 sweep point: C1 physics (input-example.json, method=eigen, omega_d_coeff=1.0), "
            f"npoints={npoints}, dim={npoints}; step = one Newton/secant iterate of "
            f"newtonTraceSecantIteration (dense step + assembly + secant)")


def c5_point(k, npoints=1024):
    """(k_rho, omega0, input text) of scan point k of config C5.  The start scales with the
    wavenumber like the diamagnetic frequency does: omega0 = (-0.8, 0.25) * k_rho / 0.3182."""
    k_rho = round(0.05 + 0.01 * k, 2)
    s = k_rho / C1_K_RHO
    w0 = (round(C1_START[0] * s, 6), round(C1_START[1] * s, 6))
    txt = c4_text(npoints, k_rho)
    txt = _sub(txt, r'"initial_guess": \[-0.8, 0.25\]', f'"initial_guess": [{w0[0]!r}, {w0[1]!r}]')
    return k_rho, complex(*w0), txt


def c5_points(npoints=1024, count=C5_NPOINTS):
    return [c5_point(k, npoints) for k in range(count)]
