"""Kernel 1 on small grids: refill granularity (EMME_REFILL_MIN) x resident CTAs per SM
(EMME_ASM_BLOCKS_PER_SM) on C1, C3 and the sweep points 512 / 2048 (assembly time only, minimum of
the assemblies of one seed + 4 iterates).  Results are bitwise independent of both knobs."""
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
        asm = []
        for _ in range(4):
            s.newtonTraceSecantIteration()
            asm.append(s.stats()["assemble_ms"])
            if abs(s.d_eigen_value) < 1e-6 * abs(s.eigen_value):
                s.seed(inp.initial_guess())
        out[name] = round(min(asm), 4)
        s.close()
    print(json.dumps(out))
else:
    envs = [{}] + [{"EMME_REFILL_MIN": str(r)} for r in (1, 4, 8, 16, 24)] + \
           [{"EMME_REFILL_MIN": "16", "EMME_ASM_BLOCKS_PER_SM": "4"}, {"EMME_ASM_BLOCKS_PER_SM": "4"}, {}]
    for env in envs:
        r = subprocess.run([sys.executable, __file__, "child"], env={**os.environ, **env}, capture_output=True, text=True)
        print(json.dumps(env), r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:], flush=True)
