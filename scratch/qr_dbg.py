import sys, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import cases
from emme_b200 import Input, solve_once_eigen
g = json.load(open("/root/repo/tests/golden/golden.json"))
for case in ["c1_pos_n64", "c1_em_n64"]:
    rec = g["newton_qr"][case]
    inp = Input(text=cases.input_path(case).read_text().replace('"TraceSecant"', '"QRSecant"'))
    w, iters, s = solve_once_eigen(inp, inp.initial_guess())
    for k, ((wi, di), r) in enumerate(zip(iters, rec["iterates"])):
        rw = complex(r[0], r[1]); rd = complex(r[2], r[3])
        print(case, k, f"omega err {abs(wi-rw)/abs(rw):.2e}  delta err {abs(di-rd)/abs(rd):.2e} |delta| {abs(rd):.2e}")
