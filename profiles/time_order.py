"""Kernel 1 against the item order: EMME_ASM_FAR_SPLIT=f runs the diagonals d >= f*N first (far pairs,
longest Miller recurrences), then d = 1, 2, ...; unset = plain diagonal-major order.  Assembly time
(minimum over one seed + 4 iterates) and omega after the iterates (must not depend on the order)."""
import os, sys, json, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from emme_b200 import EigenSolver, Input, workloads
    out = {}
    for name, txt in (("c1", workloads.C1_PATH.read_text()), ("c3", workloads.C3_PATH.read_text()),
                      ("n512", workloads.c4_text(512)), ("n2048", workloads.c4_text(2048)),
                      ("n4096", workloads.c4_text(4096)), ("n8192", workloads.c4_text(8192))):
        inp = Input(text=txt)
        s = EigenSolver.from_input(inp)
        s.seed(inp.initial_guess())
        asm = []
        for _ in range(2 if name == "n8192" else 4):
            s.newtonTraceSecantIteration()
            asm.append(s.stats()["assemble_ms"])
        out[name] = [round(min(asm), 4), repr(s.eigen_value)]
        s.close()
    print(json.dumps(out))
else:
    envs = [{}] + [{"EMME_ASM_FAR_SPLIT": f} for f in ("0.5", "0.6", "0.75")] + [{}]
    for env in envs:
        r = subprocess.run([sys.executable, __file__, "child"], env={**os.environ, **env}, capture_output=True, text=True)
        print(json.dumps(env), r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:], flush=True)
