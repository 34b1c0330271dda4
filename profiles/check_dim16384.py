"""Largest configuration that the bench sizes imply: C1 physics made electromagnetic (beta_e = 0.02)
at npoints = 8192 -> dim 16384 (4 GiB per matrix, 24 GiB per handle): seed + 2 iterates on the
symmetric path against the same on the LU paths (EMME_DENSE_SYM=0), one GPU."""
import json, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from emme_b200 import EigenSolver, Input, workloads
    import re
    txt = workloads.c4_text(8192)
    txt, n = re.subn(r'"beta_e": 0.00', '"beta_e": 0.02', txt)
    assert n == 1
    inp = Input(text=txt)
    s = EigenSolver.from_input(inp)
    t0 = time.perf_counter()
    s.seed(inp.initial_guess())
    its = []
    for _ in range(2):
        s.newtonTraceSecantIteration()
        st = s.stats()
        its.append({"omega": [s.eigen_value.real, s.eigen_value.imag], "assemble_ms": st["assemble_ms"],
                    "dense_ms": st["dense_ms"], "dense_tflops": st["dense_flops"] / st["dense_ms"] / 1e9})
    print(json.dumps({"dim": s.dim, "seconds": time.perf_counter() - t0, "sym_steps": st["sym_steps"],
                      "pivot_fallbacks": st["pivot_fallbacks"], "iterates": its}))
else:
    out = {}
    for name, env in (("symmetric", {}), ("lu", {"EMME_DENSE_SYM": "0"})):
        r = subprocess.run([sys.executable, __file__, "child"], env={**os.environ, **env}, capture_output=True, text=True)
        out[name] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"error": r.stderr[-400:]}
    a, b = out["symmetric"], out["lu"]
    if "iterates" in a and "iterates" in b:
        wa, wb = complex(*a["iterates"][-1]["omega"]), complex(*b["iterates"][-1]["omega"])
        out["rel_diff_omega"] = abs(wa - wb) / abs(wb)
    print(json.dumps(out))
