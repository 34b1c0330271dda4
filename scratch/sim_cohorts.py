"""CPU model of kernel 1's warp schedule on C1 (N=1024): per-item cost from the oracle (evaluations and
Miller trips per quadrature), then lane efficiency and tail of (a) the shipped order (mirror pairs,
whole-warp cohorts), (b) neighbours only, (c) lane-private queues of K cohorts, (d) cost-sorted cohorts
(upper bound of any reordering).  Test infrastructure only (uses oracle/)."""
import ctypes as C, os, sys, numpy as np
from multiprocessing import Pool
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
CACHE = "/tmp/c1_item_costs.npy"

def rows_job(rows):
    import oracle_lib, cases
    from emme_b200 import Input, workloads
    inp = Input(text=workloads.C1_PATH.read_text())
    pd = cases.oracle_params(inp.params()[0] if isinstance(inp.params(), tuple) else inp.params())
    eta, g, bi = cases.ref_tables("c1")
    L = oracle_lib.lib(); p = oracle_lib.make_params(pd)
    N = len(eta); out = []
    re = C.c_double(); im = C.c_double(); st = (C.c_long * 4)()
    for i in rows:
        for j in range(i + 1, N):
            for k in range(4): st[k] = 0
            L.emme_oracle_kappa(C.byref(p), 0, eta[i], eta[j], g[i], g[j], bi[i], bi[j], -0.8, 0.25,
                                C.byref(re), C.byref(im), st)
            out.append((i, j, st[0], st[1], st[2]))
    return out

def costs():
    if os.path.exists(CACHE): return np.load(CACHE)
    N = 1024
    jobs = [list(range(r, N - 1, 64)) for r in range(64)]
    with Pool(8) as pool: res = pool.map(rows_job, jobs)
    a = np.array([t for r in res for t in r], dtype=np.int64)
    np.save(CACHE, a); return a

def order_shipped(N, mirror=True):
    idx = []
    for d in range(1, N):
        L = N - d
        for t in range(L):
            i = (L - 1 - (t >> 1)) if (mirror and (t & 1)) else ((t >> 1) if mirror else t)
            idx.append((i, i + d))
    return idx

def simulate(panels, percost, K=1, warps=2960):
    """panels[n], percost[n] in schedule order.  A warp takes blocks of 32*K items; lane l runs items
    l, l+32, ... of the block back to back; every round each busy lane does one panel and the round costs
    the largest per-panel cost among the busy lanes.  Returns (useful / (32*busy), makespan, ideal)."""
    n = len(panels); B = 32 * K
    nb = (n + B - 1) // B
    blk_time = np.zeros(nb); blk_useful = np.zeros(nb)
    for b in range(nb):
        p = panels[b * B:(b + 1) * B]; c = percost[b * B:(b + 1) * B]
        lanes_p = [p[l::32] for l in range(32)]; lanes_c = [c[l::32] for l in range(32)]
        # expand each lane into its sequence of per-round costs
        seqs = [np.repeat(lc, lp) for lp, lc in zip(lanes_p, lanes_c)]
        T = max(len(s) for s in seqs)
        M = np.zeros((32, T))
        for l, s in enumerate(seqs): M[l, :len(s)] = s
        blk_time[b] = M.max(axis=0).sum(); blk_useful[b] = M.sum()
    # greedy dynamic assignment of blocks, in order, to the earliest free warp
    import heapq
    h = [0.0] * warps; heapq.heapify(h)
    for b in range(nb):
        t = heapq.heappop(h); heapq.heappush(h, t + blk_time[b])
    makespan = max(h)
    return blk_useful.sum() / (32 * blk_time.sum()), makespan, blk_time.sum() / warps

if __name__ == "__main__":
    a = costs(); N = 1024
    key = {(int(i), int(j)): k for k, (i, j) in enumerate(a[:, :2])}
    ev, fw, bw = a[:, 2], a[:, 3], a[:, 4]
    pan = ev // 15
    cost = ev * 354 + 14 * (fw + bw)          # SURVEY 8d flops per item
    per = cost / np.maximum(pan, 1)
    print("items", len(a), "evals", ev.sum(), "mean panels", pan.mean())
    for name, mirror in (("neighbours", False), ("mirror pairs (shipped)", True)):
        o = np.array([key[t] for t in order_shipped(N, mirror)])
        for K in (1, 2, 3):
            eff, mk, ideal = simulate(pan[o], per[o], K)
            print(f"{name:24s} K={K}: lanes {32*eff:5.2f}/32, tail {100*(mk/ideal-1):5.1f} % of the balanced time, total {mk/1e6:8.2f}")
    # cost-sorted inside each diagonal (oracle knowledge: upper bound)
    o = []
    pos = 0
    for d in range(1, N):
        L = N - d
        ids = np.array([key[(i, i + d)] for i in range(L)])
        o.extend(ids[np.argsort(-cost[ids], kind="stable")])
    o = np.array(o)
    for K in (1, 2):
        eff, mk, ideal = simulate(pan[o], per[o], K)
        print(f"{'cost-sorted per diagonal':24s} K={K}: lanes {32*eff:5.2f}/32, tail {100*(mk/ideal-1):5.1f} %, total {mk/1e6:8.2f}")
