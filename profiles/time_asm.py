"""Quick timing of kernel 1 alone (CUDA-event ms reported by the handle) for a golden case."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from emme_b200 import EigenSolver, Input  # noqa: E402
case = sys.argv[1] if len(sys.argv) > 1 else "c1"
inp = Input(ROOT / "tests" / "golden" / "inputs" / f"{case}.json")
if len(sys.argv) > 2:
    inp.set_number("npoints", float(sys.argv[2]))      # synthetic sweep point (BASELINE configs[3])
s = EigenSolver.from_input(inp)
w = inp.initial_guess()
ms = []
for _ in range(5):
    s.matrixAssembler(w)
    ms.append(s.stats()["assemble_ms"])
st = s.stats()
fl = st["evals"] * 354 + 14 * (st["fwd_trips"] + st["bwd_trips"])
print(f"{case} npoints={s.npoints}: assemble_ms min {min(ms):.3f} med {sorted(ms)[2]:.3f}  alg TFLOP/s {fl / min(ms) / 1e9:.2f}  {st}")
