"""Driver of scratch/sim_warp.cpp: lane efficiency and tail of kernel 1's schedule on C1 (N=1024) for
different item orders and lane-private queue depths K.  Analysis tool."""
import ctypes as C, os, sys, subprocess, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
from emme_b200 import Input, workloads, capi
LIB = "/tmp/libsimwarp.so"
if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(ROOT, "scratch/sim_warp.cpp")):
    subprocess.run(["g++", "-O2", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB,
                    os.path.join(ROOT, "scratch/sim_warp.cpp")], check=True)
L = C.CDLL(LIB)
dp = C.POINTER(C.c_double)
L.sim_schedule.argtypes = [C.POINTER(capi.EmmeParams), C.c_int, dp, dp, dp, C.c_double, C.c_double,
                           C.POINTER(C.c_int), C.c_long, C.c_int, C.c_int, dp, dp]

def order(N, kind, band=1, nb=32):
    """kind: 'nbr' neighbours, 'mirror' (shipped); band > 1: `band` adjacent diagonals interleaved."""
    out = []
    d = 1
    while d < N:
        ds = [dd for dd in range(d, min(d + band, N))]
        Lmax = N - ds[0]
        for t in range(Lmax):
            pos = t if kind == "nbr" else ((t >> 1) if not (t & 1) else None)
            for dd in ds:
                Ld = N - dd
                if kind == "nbr":
                    i = t
                else:
                    i = (t >> 1) if not (t & 1) else Ld - 1 - (t >> 1)
                    # each diagonal must be covered exactly once: positions t < Ld in ITS OWN alternation
                if t < Ld: out.append((i, i + dd))
        d += band
    return np.array(out, dtype=np.int32)

def run(inp_text, tables, items, K, omega=(-0.8, 0.25), warps=2960):
    inp = Input(text=inp_text)
    p, _n = inp.params()
    eta, g, bi = tables
    res = (C.c_double * 5)()
    items = np.asarray(items, dtype=np.int32)
    if items.shape[1] == 2: items = np.concatenate([items, np.zeros((len(items), 1), np.int32)], axis=1)
    items = np.ascontiguousarray(items)
    L.sim_schedule(C.byref(p), len(eta), eta.ctypes.data_as(dp), g.ctypes.data_as(dp), bi.ctypes.data_as(dp),
                   omega[0], omega[1], items.ctypes.data_as(C.POINTER(C.c_int)), len(items), K, warps, res, None)
    use, tot32, mk, bal, ev = res
    return dict(lanes=32 * use / tot32, makespan=mk, balanced=bal, idle_pct=100 * (1 - bal / mk), evals=int(ev))

if __name__ == "__main__":
    N = 1024
    tables = cases.ref_tables("c1")
    txt = workloads.C1_PATH.read_text()
    base = None
    for name, kind, band in (("neighbours", "nbr", 1), ("mirror pairs (shipped)", "mirror", 1),
                             ("mirror, 2 diagonals", "mirror", 2), ("mirror, 4 diagonals", "mirror", 4)):
        it = order(N, kind, band)
        assert len(it) == N * (N - 1) // 2 and len({(a, b) for a, b in it.tolist()}) == len(it)
        for K in (1, 2):
            r = run(txt, tables, it, K)
            if base is None: base = r["makespan"]
            print(f"{name:24s} K={K}: lanes {r['lanes']:5.2f}/32, idle at the end {r['idle_pct']:4.1f} %, "
                  f"makespan {r['makespan']/base:6.3f} of the first line, evals {r['evals']}", flush=True)
