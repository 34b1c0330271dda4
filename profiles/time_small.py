"""Timing of the small configurations (BASELINE configs C1, C3, sweep points <= 2048): kernel 1 with
3/4/5 resident CTAs per SM and kernel 2 with and without the look-ahead split."""
import os, sys, json, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from emme_b200 import EigenSolver, Input, workloads
    out = {}
    for name, txt in (("c1", workloads.C1_PATH.read_text()), ("c3", workloads.C3_PATH.read_text()),
                      ("n512", workloads.c4_text(512)), ("n2048", workloads.c4_text(2048))):
        inp = Input(text=txt)
        s = EigenSolver.from_input(inp)
        s.seed(inp.initial_guess())
        asm, dns = [], []
        for _ in range(4):
            s.newtonTraceSecantIteration()
            st = s.stats()
            asm.append(st["assemble_ms"]); dns.append(st["dense_ms"])
            if abs(s.d_eigen_value) < 1e-6 * abs(s.eigen_value):
                s.seed(inp.initial_guess())
        out[name] = {"assemble_ms": min(asm), "dense_ms": min(dns), "sym_steps": st["sym_steps"]}
        s.close()
    print(json.dumps(out))
else:
    for env in ({}, {"EMME_ASM_BLOCKS_PER_SM": "4"}, {"EMME_ASM_BLOCKS_PER_SM": "3"}, {"EMME_DENSE_LOOKAHEAD": "0"}):
        r = subprocess.run([sys.executable, __file__, "child"], env={**os.environ, **env}, capture_output=True, text=True)
        print(env, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:], flush=True)
