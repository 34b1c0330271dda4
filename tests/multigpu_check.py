"""Run under torchrun on >= 2 GPUs: the pair-sharded solver (peer stores and all-reduce variants)
must reproduce the single-GPU iterates and matrix bit for bit.  Driven by tests/test_multigpu.py."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from emme_b200 import EigenSolver, Input, parallel  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    case = sys.argv[1] if len(sys.argv) > 1 else "c1_n128"
    inp = Input(ROOT / "tests" / "golden" / "inputs" / f"{case}.json")
    p, n = inp.params()
    w0 = inp.initial_guess()
    single = EigenSolver.from_input(inp, w0, device=local)
    for _ in range(3):
        single.newtonTraceSecantIteration()
    A1 = single.eigen_matrix
    ok = True
    for exchange in ("p2p", "allreduce"):
        s = parallel.ShardedEigenSolver(p, n, *inp.tables(), device=local, exchange=exchange)
        s.seed(w0)
        for _ in range(3):
            s.newtonTraceSecantIteration()
        same_w = s.eigen_value == single.eigen_value and s.d_eigen_value == single.d_eigen_value
        same_A = np.array_equal(s.eigen_matrix, A1)
        print(f"[rank {rank}] {case} {exchange}: omega {s.eigen_value!r} same_omega={same_w} same_matrix={same_A}",
              flush=True)
        ok = ok and same_w and same_A
        s.close()
        dist.barrier()
    # scan-parallel: 6 independent k_rho points dealt to the ranks, gathered in scan order
    base = (ROOT / "tests" / "golden" / "inputs" / "c1_n64.json").read_text()
    ks = [0.25 + 0.02 * k for k in range(6)]
    recs = parallel.solve_scan_parallel(base, "k_rho", ks, w0, device=local)
    ok = ok and [r["scan_value"] for r in recs] == ks and all(r.get("converged") for r in recs)
    ok = ok and sorted({r["rank"] for r in recs}) == list(range(min(world, 6)))
    if rank == 0:
        print("[scan]", [(r["scan_value"], r["rank"], r["iterations"], r["eigenvalue"]) for r in recs], flush=True)
    t = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
