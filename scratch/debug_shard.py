"""GPU debug: where does the column-sharded dense step first differ from the single-handle one?"""
import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np
import cases
from emme_b200 import EigenSolver, Input, capi, parallel
from test_newton_gpu import _sym_case

def run(dim, world, nbo):
    os.environ["EMME_DENSE_NBO"] = str(nbo)
    lib = capi.load()
    lib.emme_peer_set_timeout(5.0)
    A, B = _sym_case(dim, seed=11)
    inp = Input(cases.input_path("c1_n32"))
    p, _ = inp.params()
    mk = lambda: EigenSolver(p, dim, np.linspace(-1, 1, dim), np.zeros(dim), np.ones(dim))
    one = mk()
    d1 = one.trace_delta(A, B)
    W1, Y1 = one._matrix(3), one._matrix(4)
    g = parallel.LocalShardedGroup(p, dim, np.linspace(-1, 1, dim), np.zeros(dim), np.ones(dim), devices=[0] * world)
    out = [None] * world
    g._all(lambda s: out.__setitem__(g.ranks.index(s), s.trace_delta(A, B)))
    print(f"dim {dim} world {world} nbo {nbo}: single {d1} sharded {out[0]} equal {all(o == d1 for o in out)}", flush=True)
    nbk = (dim + nbo - 1) // nbo
    for r, s in enumerate(g.ranks):
        Wr, Yr = s._matrix(3), s._matrix(4)
        for name, M1, Mr in (("W", W1, Wr), ("Y", Y1, Yr)):
            bad = []
            for bi in range(nbk):
                for bj in range(nbk):
                    a = M1[bi*nbo:(bi+1)*nbo, bj*nbo:(bj+1)*nbo]; b = Mr[bi*nbo:(bi+1)*nbo, bj*nbo:(bj+1)*nbo]
                    # W: compare only what every rank must hold (L panels, U12, diagonal blocks) or owns
                    if name == "W" and bj > bi and False:
                        continue
                    if not np.array_equal(a, b):
                        bad.append((bi, bj, float(np.abs(a - b).max())))
            print(f"  rank {r} {name}: {len(bad)} differing blocks", bad[:12], flush=True)
    g.close(); one.close()

for cfg in [(2304, 4, 256), (2304, 2, 256), (2304, 4, 128), (1024, 4, 256), (1536, 2, 256), (768, 2, 256)]:
    try:
        run(*cfg)
    except Exception as e:
        print(cfg, "ERROR", e, flush=True)
