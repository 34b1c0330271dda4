"""GPU diag: which dense steps of c1_n256 leave the symmetric path, per outer block width."""
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import cases
from emme_b200 import EigenSolver, Input
for nbo in ("0", "64", "128"):
    os.environ["EMME_DENSE_NBO"] = nbo
    for case in ("c1_n256", "c1_n128", "c1_n512"):
        inp = Input(cases.input_path(case))
        s = EigenSolver.from_input(inp)
        s.seed(inp.initial_guess())
        prev = s.stats()
        log = []
        for k in range(4):
            s.newtonTraceSecantIteration()
            st = s.stats()
            A = s.eigen_matrix_old
            log.append((st["sym_steps"] - prev["sym_steps"], st["pivot_fallbacks"] - prev["pivot_fallbacks"],
                        bool(np.array_equal(A, A.T)), float(np.abs(A - np.diag(np.diag(A))).max())))
            prev = st
        print(f"nbo {nbo} {case}: per step (sym, pivot_fallback, A_old symmetric, max offdiag) {log}", flush=True)
        s.close()
