"""EM (beta_e != 0) sanity at sizes above the golden cases: symmetric path taken, omega stable in N."""
import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import cases
from emme_b200 import Input, solve_once_eigen
for case, n in (("c3", 1024), ("c3", 2048), ("c1_em_n64", 2048)):
    inp = Input(cases.input_path(case))
    inp.set_number("npoints", float(n))
    t0 = time.perf_counter()
    w, iters, s = solve_once_eigen(inp, inp.initial_guess())
    t1 = time.perf_counter()
    st = s.stats()
    print(f"{case} npoints={n} dim={s.dim}: {len(iters)} iterates omega={w!r} wall={t1 - t0:.3f}s assemble_ms={st['assemble_ms']:.2f} "
          f"dense_ms={st['dense_ms']:.2f} sym_steps={st['sym_steps']} pivot_fallbacks={st['pivot_fallbacks']}")
    s.close()
