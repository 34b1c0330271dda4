"""Host-side multi-GPU logic on CPU: shard arithmetic and the world_size-2 gloo gather."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
from emme_b200 import parallel


def test_shard_items_partition_is_exact():
    for n, em in ((32, False), (64, True), (1024, False), (1000, True)):
        total = parallel.n_work_items(n, em)
        for world in (1, 2, 3, 4, 8):
            parts = [parallel.shard_items(total, r, world) for r in range(world)]
            assert sum(parts) == total
            assert max(parts) - min(parts) <= 32
            # the kernel's rule: local item k -> global (k//32)*32*world + rank*32 + k%32, all < total
            seen = set()
            for r in range(world):
                for k in range(parts[r]):
                    gk = (k // 32) * 32 * world + r * 32 + k % 32
                    assert gk < total
                    seen.add(gk)
            assert len(seen) == total


def test_scan_partition_round_robin():
    pts = [0.05 + 0.01 * k for k in range(64)]
    seen = []
    for r in range(8):
        mine = parallel.scan_partition(pts, r, 8)
        assert len(mine) == 8 and all(k % 8 == r for k, _ in mine)
        seen += [k for k, _ in mine]
    assert sorted(seen) == list(range(64))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, str(cases.ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pts = [0.05 + 0.01 * k for k in range(7)]            # ragged: 7 points on 2 ranks
        mine = parallel.scan_partition(pts, rank, world)
        local = [(k, {"k_rho": p, "omega": complex(-p, p * p), "iters": k + 1, "rank": rank})
                 for k, p in mine]
        merged = parallel.gather_results(local)
        # sharded-assembly emulation: each rank owns items k*world+rank of a zeroed buffer;
        # the all-reduce(sum) of disjoint shares reproduces the full array exactly
        total = parallel.n_work_items(16, False)
        full = np.arange(1, total + 1, dtype=np.float64) * 1.25
        share = np.zeros(total)
        for k in range(parallel.shard_items(total, rank, world)):
            gk = (k // 32) * 32 * world + rank * 32 + k % 32
            share[gk] = full[gk]
        t = torch.from_numpy(share)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ok = bool(np.array_equal(t.numpy(), full))
        # PIC exchange emulation: every rank deposits its contiguous block of markers with the
        # linear weights of solve_field; the all-reduce(sum) of the densities is the full deposit
        rng = np.random.default_rng(5)
        nmk, nf = 1001, 16
        u = rng.uniform(0, nf, nmk)
        wgt = rng.normal(size=nmk)
        cell = u.astype(np.int64)
        frac = u - cell
        full_d = np.bincount(cell, wgt * (1 - frac), nf) + np.bincount((cell + 1) % nf, wgt * frac, nf)
        first, count = parallel.marker_shard(nmk, rank, world)
        sl = slice(first, first + count)
        mine_d = torch.from_numpy(np.bincount(cell[sl], (wgt * (1 - frac))[sl], nf) +
                                  np.bincount((cell[sl] + 1) % nf, (wgt * frac)[sl], nf))
        dist.all_reduce(mine_d, op=dist.ReduceOp.SUM)
        ok = ok and bool(np.allclose(mine_d.numpy(), full_d, rtol=1e-13, atol=1e-13))
        # max-over-ranks timing reduction used by bench.py
        ms = torch.tensor([10.0 + rank])
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            torch.save({"merged": merged, "ok": ok, "ms": float(ms)}, out)
    finally:
        dist.destroy_process_group()


def test_gloo_world2_gather_and_allreduce(tmp_path):
    out = tmp_path / "r0.pt"
    mp.spawn(_worker, args=(2, _free_port(), str(out)), nprocs=2, join=True)
    res = torch.load(out, weights_only=False)
    assert res["ok"] and res["ms"] == 11.0
    merged = res["merged"]
    assert [m["iters"] for m in merged] == [1, 2, 3, 4, 5, 6, 7]
    assert [m["rank"] for m in merged] == [0, 1, 0, 1, 0, 1, 0]


def test_marker_shard_partition_is_exact():
    for n in (1, 7, 1024, 1048576 + 5):
        for world in (1, 2, 3, 8):
            blocks = [parallel.marker_shard(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
            assert all(blocks[r][0] + blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def test_gather_without_process_group_is_identity():
    assert parallel.gather_results([(1, "b"), (0, "a")]) == ["a", "b"]
