"""torchrun helper: time the column-sharded dense step alone (CUDA events inside the library,
emme_stats.dense_ms) at the bench size for a few outer block widths.
    torchrun --nproc-per-node N profiles/time_dense_sharded.py [npoints] [nbo ...]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from emme_b200 import Input, parallel, workloads

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
npoints = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
nbos = [int(a) for a in sys.argv[2:]] or [256]
out = {}
for nbo in nbos:
    os.environ["EMME_DENSE_NBO"] = str(nbo)
    inp = Input(text=workloads.c4_text(npoints))
    p, n = inp.params()
    s = parallel.ShardedEigenSolver(p, n, *inp.tables(), device=local, shard_dense=True)
    s.seed(inp.initial_guess())
    ms, asm = [], []
    for k in range(5):
        s.newtonTraceSecantIteration()
        st = s.stats()
        ms.append(st["dense_ms"]); asm.append(st["assemble_ms"])
        if abs(s.d_eigen_value) < 1e-6 * abs(s.eigen_value):
            s.seed(inp.initial_guess() * (1.0 + 0.01 * (k + 1)))
    t = torch.tensor([min(ms[1:]), sorted(ms[1:])[len(ms[1:]) // 2], min(asm[1:])], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[nbo] = {"dense_ms_min": float(t[0]), "dense_ms_median": float(t[1]), "assemble_ms": float(t[2]),
                "sym_steps": st["sym_steps"], "omega": [s.eigen_value.real, s.eigen_value.imag]}
    s.close()
    dist.barrier()
if rank == 0:
    print(json.dumps({"n_gpus": world, "npoints": npoints, "dense": out}))
dist.destroy_process_group()
