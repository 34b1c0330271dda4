"""Timing of the PIC method (row N4): input-example.json as shipped (method PIC, 1024 cells x 1024
markers per cell), `steps` Integrator::step calls per timed call.  Usage: time_pic.py [mpc] [steps] [npoints]"""
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from emme_b200 import Input, pic  # noqa: E402
inp = Input(ROOT / "tests" / "golden" / "inputs" / "pic.json")
if len(sys.argv) > 3:
    inp.set_number("npoints", float(sys.argv[3]))
p, mpc, nt, dt = pic.pic_params(inp)
mpc = int(sys.argv[1]) if len(sys.argv) > 1 else mpc
steps = int(sys.argv[2]) if len(sys.argv) > 2 else nt
t0 = time.perf_counter()
markers = pic.load_markers(p, mpc * p.npoints, seed=1)
t1 = time.perf_counter()
s = pic.PIC_State.from_markers(p, *markers)
t2 = time.perf_counter()
s.step(dt, 3)   # warm-up (graph capture unless EMME_PIC_PERSISTENT=1)
ms = []
for _ in range(3):
    s.step(dt, steps)
    ms.append(s.timing()[0])
n = s.marker_num()
best = min(ms)
stats = s.field_stats()
print(f"pic npoints={p.npoints} markers={n} steps={steps}: load {t1 - t0:.3f}s create {t2 - t1:.3f}s; "
      f"{best:.3f} ms per call, {best / steps * 1e3:.1f} us/step, {best / steps / 3 * 1e3:.1f} us/stage, "
      f"{3.0 * n * steps / best / 1e6:.1f} G marker-stages/s; all {['%.3f' % m for m in ms]}; "
      f"omega {pic.calculate_omega(stats, dt)}")
