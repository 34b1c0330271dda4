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
from emme_b200 import EigenSolver, Input, parallel, pic  # noqa: E402


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
    # p2p + column-sharded dense step (EMME_DENSE_NBO=64 from the test gives these small matrices
    # several column blocks per rank), p2p with a replicated dense step, NCCL all-reduce baseline
    for exchange, shard_dense in (("p2p", True), ("p2p", False), ("allreduce", False)):
        s = parallel.ShardedEigenSolver(p, n, *inp.tables(), device=local, exchange=exchange,
                                        shard_dense=shard_dense)
        s.seed(w0)
        for k in range(3):
            if rank == k % world:
                import time
                time.sleep(0.3)                  # ranks out of step: the device barriers must hold
            s.newtonTraceSecantIteration()
        same_w = s.eigen_value == single.eigen_value and s.d_eigen_value == single.d_eigen_value
        same_A = np.array_equal(s.eigen_matrix, A1) and np.array_equal(s.eigen_matrix_old, single.eigen_matrix_old)
        st = s.stats()
        print(f"[rank {rank}] {case} {exchange} shard_dense={shard_dense}: omega {s.eigen_value!r} "
              f"same_omega={same_w} same_matrix={same_A} sym_steps={st['sym_steps']}", flush=True)
        # the path decisions (symmetric / pivoting fallback near convergence) are the single-GPU ones:
        # the flags are OR-ed over the ranks
        st1 = single.stats()
        ok = ok and same_w and same_A and (st["sym_steps"], st["pivot_fallbacks"]) == (st1["sym_steps"], st1["pivot_fallbacks"])
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
    # PIC method: markers in contiguous blocks, one density all-reduce per stage; the fields
    # agree with a single-GPU run to rounding (the deposit order differs), positions bit for bit
    pin = Input(ROOT / "tests" / "golden" / "inputs" / "pic_n64_wb.json")
    pp, mpc, _, pdt = pic.pic_params(pin)
    markers = pic.load_markers(pp, 4096 * pp.npoints // 64 + 3, seed=17)      # ragged split
    one = pic.PIC_State.from_markers(pp, *markers, device=local)
    one.step(pdt, 5)
    f1 = one.field_history()
    e1, w1 = one.markers()
    first, count = parallel.marker_shard(markers[0].shape[0], rank, world)
    for exchange in ("p2p", "nccl"):     # fused peer-store exchange inside the field kernel / NCCL baseline
        sh = parallel.ShardedPIC(pp, markers, device=local, exchange=exchange)
        sh.step(pdt, 2)
        if rank == 1:
            import time
            time.sleep(0.2)              # ranks out of step between calls
        sh.step(pdt, 3)
        sh.synchronize()
        fs = sh.field_history()
        ferr = float(np.abs(f1 - fs).max() / np.abs(f1).max())
        es, ws = sh.markers()
        same_eta = es.shape[0] == count and np.array_equal(es, e1[first:first + count])
        werr = float(np.abs(ws - w1[first:first + count]).max() / np.abs(w1).max())
        # every rank holds the same history: bit for bit with the fixed-order peer sum
        hist = [None] * world
        dist.all_gather_object(hist, fs.tobytes())
        same_hist = all(h == hist[0] for h in hist) if exchange == "p2p" else True
        pic_ok = ferr <= 1e-12 and same_eta and werr <= 1e-12 and same_hist
        print(f"[rank {rank}] pic sharded ({exchange}): field deviation {ferr:.2e} weights {werr:.2e} "
              f"same_eta={same_eta} same_history_on_all_ranks={same_hist} same_pic={pic_ok}", flush=True)
        ok = ok and pic_ok
        sh.close()
    one.close()
    # markers drawn per rank (from_seed): the blocks tile the range and the normalisation is global
    sf = parallel.ShardedPIC.from_seed(pp, 64 * pp.npoints + 5, seed=3, device=local)
    sf.step(pdt, 2)
    pw = sf.state.extras()[2]
    t = torch.tensor([float(pw.sum()), float(sf.state.marker_num())], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t)
    seed_ok = abs(float(t[0]) - 2 * pp.length) <= 1e-9 * pp.length and int(t[1]) == 64 * pp.npoints + 5 \
        and np.isfinite(sf.current_field()).all()
    print(f"[rank {rank}] pic from_seed: sum p_weight {float(t[0])!r} (2L = {2 * pp.length}) markers {int(t[1])} "
          f"same_pic={seed_ok}", flush=True)
    ok = ok and seed_ok
    sf.close()
    t = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
