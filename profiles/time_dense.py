"""Quick timing of kernel 2 alone (dense step on host-provided matrices) and a numpy check."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from emme_b200 import EigenSolver, Input  # noqa: E402
inp = Input(ROOT / "tests" / "golden" / "inputs" / "c1_n32.json")
p, _ = inp.params()
kind = "dom"
args = []
for a in sys.argv[1:]:
    if a in ("dom", "piv", "sym", "qr"):
        kind = a
    else:
        args.append(int(a))
for n in args or [1024]:
    rng = np.random.default_rng(n)
    A = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) * (0.5 / np.sqrt(n)) + 2 * np.eye(n)
    if kind == "piv":
        A[::7] *= 0.01          # badly scaled rows: partial pivoting must interchange
    if kind in ("sym", "qr"):
        A = (A + A.T) / 2       # complex symmetric like EMME's matrices: symmetric path of kernel 2
    B = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    s = EigenSolver(p, n, np.linspace(-1, 1, n), np.zeros(n), np.ones(n))
    ms = []
    for _ in range(4):
        d = s.qr_delta(A, B) if kind == "qr" else s.trace_delta(A, B)
        ms.append(s.stats()["dense_ms"])
    err = None
    if n <= 2048 and kind != "qr":
        ref = -1.0 / np.trace(np.linalg.solve(A, B))
        err = abs(d - ref) / abs(ref)
    fl = s.stats()["dense_flops"]
    print(f"n={n}: dense_ms min {min(ms):.3f}  {fl / min(ms) / 1e9:.2f} TFLOP/s  rel.err vs numpy {err}  launches/step {s.stats()['launches'] // 4} fallbacks {s.stats()['pivot_fallbacks']} sym_steps {s.stats()['sym_steps']} ({kind})")
    s.close()
