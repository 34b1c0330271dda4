"""Kernel 2 (dense step of one Newton iterate) at C1, C3 and the sweep sizes: minimum of four iterates."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emme_b200 import EigenSolver, Input, workloads
out = {}
for name, txt in (("c1", workloads.C1_PATH.read_text()), ("c3", workloads.C3_PATH.read_text()),
                  ("n512", workloads.c4_text(512)), ("n2048", workloads.c4_text(2048)), ("n4096", workloads.c4_text(4096)), ("n8192", workloads.c4_text(8192))):
    inp = Input(text=txt)
    s = EigenSolver.from_input(inp)
    s.seed(inp.initial_guess())
    dns = []
    for _ in range(4):
        s.newtonTraceSecantIteration()
        st = s.stats()
        dns.append(st["dense_ms"])
        if abs(s.d_eigen_value) < 1e-6 * abs(s.eigen_value):
            s.seed(inp.initial_guess())
    out[name] = {"dense_ms": round(min(dns), 4), "sym_steps": st["sym_steps"], "omega": [s.eigen_value.real, s.eigen_value.imag]}
    s.close()
print(json.dumps(out))
