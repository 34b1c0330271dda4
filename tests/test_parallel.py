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


def _pw_worker(rank, world, port, out):
    """from_seed's host logic over gloo: every rank draws ITS block from its own mt19937 stream, the
    un-normalised p_weight sums are added with one all-reduce, and the blocks tile the marker range."""
    sys.path.insert(0, str(cases.ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from emme_b200 import Input, pic
        p, _, _, _ = pic.pic_params(Input(cases.input_path("pic_n64_wb")))     # water-bag weights != 1
        n_total = 10007
        first, count = parallel.marker_shard(n_total, rank, world)
        block = pic.load_markers(p, count, seed=3 + rank)
        part = pic.pweight_sum(p, block[1], block[2])
        t = torch.tensor([part, float(count), float(first)], dtype=torch.float64)
        parts = [torch.zeros(3, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, t)
        tot = torch.tensor([part], dtype=torch.float64)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        if rank == 0:
            torch.save({"parts": [x.tolist() for x in parts], "total": float(tot), "n_total": n_total}, out)
    finally:
        dist.destroy_process_group()


def test_gloo_world2_block_markers_and_pweight_sum(tmp_path, native_lib):
    out = tmp_path / "pw.pt"
    mp.spawn(_pw_worker, args=(2, _free_port(), str(out)), nprocs=2, join=True)
    res = torch.load(out, weights_only=False)
    parts = res["parts"]
    assert sum(int(x[1]) for x in parts) == res["n_total"]
    assert int(parts[0][2]) == 0 and int(parts[1][2]) == int(parts[0][1])      # contiguous blocks
    assert abs(sum(x[0] for x in parts) - res["total"]) <= 1e-12 * res["total"]
    assert all(x[0] > 0 for x in parts)


def test_pweight_sum_is_additive_over_blocks(native_lib):
    """emme_pic_pweight_sum over the blocks of a marker set adds up to the sum over the whole set (to
    rounding: the reference sums sequentially), for unit and non-unit water-bag weights."""
    from emme_b200 import Input, pic
    for case in ("pic_n32", "pic_n64_wb"):
        p, _, _, _ = pic.pic_params(Input(cases.input_path(case)))
        m = pic.load_markers(p, 5000, seed=9)
        whole = pic.pweight_sum(p, m[1], m[2])
        parts = 0.0
        for r in range(3):
            first, count = parallel.marker_shard(5000, r, 3)
            parts += pic.pweight_sum(p, m[1][first:first + count], m[2][first:first + count])
        assert abs(parts - whole) <= 1e-13 * whole
        if case == "pic_n32":                    # unit weights: p_weight is v_perp itself
            assert whole == float(np.add.reduce(m[2], dtype=np.float64)) or abs(whole - m[2].sum()) <= 1e-12 * whole


def test_column_block_ownership_covers_every_block():
    """The column-block-cyclic ownership of the sharded dense step (csrc/dense.cu: block K belongs to
    rank K mod P; first_own(from) = smallest own block >= from) restated: every block has exactly one
    owner and the look-ahead split (block K+1 alone, then K+1+P, ...) visits each own block once."""
    def first_own(frm, me, P):
        return frm + ((me - frm) % P + P) % P
    for P in (1, 2, 3, 4, 8):
        for nb in (1, 2, 5, 9, 32, 33):
            for K in range(nb):
                seen = []
                for me in range(P):
                    ahead = K + 1 < nb and (K + 1) % P == me
                    blocks = list(range(first_own(K + 1, me, P), nb, P))
                    if ahead:
                        assert blocks[0] == K + 1
                        assert list(range(K + 1 + P, nb, P)) == blocks[1:]
                    assert all(b % P == me and b > K for b in blocks)
                    seen += blocks
                assert sorted(seen) == list(range(K + 1, nb))


def test_item_order_visits_every_pair_once():
    """Kernel 1's item order restated (csrc/assembly.cu::decode_item): pairs diagonal-major (d = j - i),
    inside a diagonal alternating between its two ends (0, L-1, 1, L-2, ...) so that a cohort of 32
    items holds 16 neighbours and their 16 mirror images.  d_split = 0: d ascending; d_split > 0: the far
    diagonals first (d = N-1 down to d_split), then d = 1 .. d_split-1.  Every pair i < j exactly once."""
    def decode(p, N, d_split=0):
        # closed forms for the diagonal with integer fix-ups, as on the device
        n_far = (N - d_split) * (N - d_split + 1) // 2 if d_split > 0 else 0
        if p < n_far:
            L = max(int(np.floor((1.0 + np.sqrt(1.0 + 8.0 * p)) * 0.5)), 1)
            while L > 1 and L * (L - 1) // 2 > p:
                L -= 1
            while L * (L + 1) // 2 <= p:
                L += 1
            d, base = N - L, L * (L - 1) // 2
        else:
            p -= n_far
            tn = 2.0 * N + 1.0
            disc = max(tn * tn - 8.0 * (N + p), 0.0)
            d = int(np.floor((tn - np.sqrt(disc)) * 0.5))
            d = min(max(d, 1), N - 1)
            while d > 1 and (d - 1) * (2 * N - d) // 2 > p:
                d -= 1
            while d < N - 1 and d * (2 * N - d - 1) // 2 <= p:
                d += 1
            base = (d - 1) * (2 * N - d) // 2
        L, t = N - d, p - base
        i = (L - 1 - (t >> 1)) if (t & 1) else (t >> 1)
        return i, i + d
    for N in (2, 3, 7, 32, 33, 100):
        pairs = [decode(p, N) for p in range(N * (N - 1) // 2)]
        assert len(set(pairs)) == len(pairs) == N * (N - 1) // 2
        assert all(0 <= i < j < N for i, j in pairs)
        ds = [j - i for i, j in pairs]
        assert ds == sorted(ds)                                   # diagonal-major
    for N, d_split in ((7, 3), (32, 16), (33, 16), (33, 2), (33, 32), (100, 50), (1024, 512)):
        n = N * (N - 1) // 2
        pairs = [decode(p, N, d_split) for p in range(n)]
        assert len(set(pairs)) == n and all(0 <= i < j < N for i, j in pairs)
        ds = [j - i for i, j in pairs]
        n_far = (N - d_split) * (N - d_split + 1) // 2
        assert ds[:n_far] == sorted(ds[:n_far], reverse=True) and min(ds[:n_far]) == d_split
        assert ds[n_far:] == sorted(ds[n_far:]) and (n_far == n or max(ds[n_far:]) == d_split - 1)
    # mirror pairing: consecutive even/odd positions of a diagonal are mirror images of each other
    N = 64
    first = [decode(p, N) for p in range(8)]                      # diagonal d = 1, L = 63
    assert first[:4] == [(0, 1), (62, 63), (1, 2), (61, 62)]
    assert all(first[2 * k][0] + first[2 * k + 1][1] == N - 1 for k in range(4))
    far = [decode(p, N, 32) for p in range(6)]                    # d = 63 (1 pair), 62 (2), 61 (3)
    assert far == [(0, 63), (0, 62), (1, 63), (0, 61), (2, 63), (1, 62)]

    # electromagnetic runs: the three integrals of a pair are three items; the modes sit inside the
    # diagonal (all pairs of d for m = 0, then m = 1, then m = 2)
    def decode_item(k, N, nm, d_split=0):
        i, j = decode(k // nm, N, d_split)            # some pair on the item's diagonal
        d, L = j - i, N - (j - i)
        n_far = (N - d_split) * (N - d_split + 1) // 2 if d_split > 0 else 0
        base = (N - d) * (N - d - 1) // 2 if (d_split > 0 and d >= d_split) else n_far + (d - 1) * (2 * N - d) // 2
        r = k - nm * base
        m, t = divmod(r, L)
        i = (L - 1 - (t >> 1)) if (t & 1) else (t >> 1)
        return i, i + d, m
    for N, d_split in ((2, 0), (7, 0), (33, 0), (33, 16), (64, 32), (100, 50)):
        n = 3 * N * (N - 1) // 2
        items = [decode_item(k, N, 3, d_split) for k in range(n)]
        assert len(set(items)) == n and all(0 <= i < j < N and 0 <= m < 3 for i, j, m in items)
        # a diagonal's items are contiguous, mode by mode
        key = [(j - i, m) for i, j, m in items]
        assert all(key[k] == key[k + 1] or key[k] not in key[k + 1:] for k in range(0, n - 1, max(1, n // 97)))
    assert [decode_item(k, 5, 3) for k in range(12)] == [(0, 1, 0), (3, 4, 0), (1, 2, 0), (2, 3, 0),
                                                          (0, 1, 1), (3, 4, 1), (1, 2, 1), (2, 3, 1),
                                                          (0, 1, 2), (3, 4, 2), (1, 2, 2), (2, 3, 2)]
