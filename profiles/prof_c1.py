"""Profiling driver: C1 (input-example.json, method=eigen, N=1024): seed + 2 Newton iterates.
Used under ncu (launch list / --set full); never a bench number."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from emme_b200 import EigenSolver, Input  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "c1"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
inp = Input(ROOT / "tests" / "golden" / "inputs" / f"{case}.json")
s = EigenSolver.from_input(inp)
s.seed(inp.initial_guess())
for _ in range(iters):
    s.newtonTraceSecantIteration()
print(case, s.eigen_value, s.stats())
