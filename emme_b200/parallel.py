"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink) for the
exchange; the compute stays in libemme_b200.so.

Two ways the path shards:

* scan-parallel (weak scaling, no data-path collective): independent scan points / omega
  starts are dealt round-robin to ranks; only (omega, iterations, status) is gathered at the end.
  `scan_partition`, `gather_results`.
* one problem over N GPUs (strong scaling), `ShardedEigenSolver`:
  - assembly: the work items of ONE assembly (pairs i<j in diagonal-major order, times 3 modes when
    electromagnetic) are dealt to ranks in chunks of 32 (`shard_items`).  Default
    (`exchange="p2p"`): the ranks map each other's buffers with CUDA IPC and the assembly kernel
    stores every entry it computes straight into the matrix of EVERY GPU over NVLink while it
    computes; two stream-ordered device barriers (release/acquire flags in peer memory, csrc/peer.cu)
    bracket the launch: nobody overwrites a buffer a peer may still read as eigen_matrix_old, nobody
    reads before every store has landed.  No NCCL, no host round trip.  Baseline to compare against
    (`exchange="allreduce"`): each rank fills its entries of a zeroed buffer and one NCCL all-reduce
    (sum) completes the matrix.  Either way the result is bit-identical to a single-GPU assembly.
  - dense step (`shard_dense`, default from dim 4096): the symmetric L D L^T path is column-block-
    cyclic over the ranks; the owner of a 256-column panel factors it and stores it into every
    peer's work matrix, announces it with a flag, and everybody updates its own columns; the
    partial traces are all-gathered by peer stores and reduced in one fixed order, so delta is
    BITWISE the single-GPU delta and every rank holds the same omega without a broadcast.
  `LocalShardedGroup` drives the same protocol from ONE process (one host thread per rank, several
  ranks per device allowed): the CPU-less way to test it on a single GPU.
* PIC method (row N4): the markers are dealt to ranks in contiguous blocks (`marker_shard`); every
  Runge-Kutta stage deposits the rank's markers and ends with ONE all-reduce (sum) of the density,
  2 * npoints doubles (16 KB at 1024 cells) -- the path's only real exchange step -- after which
  every rank applies the quasi-neutrality table and holds the same field.  `ShardedPIC`.
"""
import ctypes as C

import numpy as np

from . import capi
from .solver import EigenSolver


def shard_items(n_items, rank, world, chunk=32):
    """Number of work items rank `rank` owns: chunks of `chunk` consecutive items are dealt
    round-robin (launch_assembly in emme_b200/csrc/assembly.cu uses chunk = 32, one warp cohort)."""
    n_chunks = (n_items + chunk - 1) // chunk
    if n_chunks <= rank:
        return 0
    mine = (n_chunks - rank + world - 1) // world
    n = mine * chunk
    if (mine - 1) * world + rank == n_chunks - 1:      # I own the (possibly partial) last chunk
        n -= n_chunks * chunk - n_items
    return n


def n_work_items(npoints, electromagnetic):
    return npoints * (npoints - 1) // 2 * (3 if electromagnetic else 1)


def scan_partition(points, rank, world):
    """Round-robin deal of independent scan points to ranks: [(global_index, point), ...]."""
    return [(k, p) for k, p in enumerate(points) if k % world == rank]


def gather_results(local, world=None):
    """All-gather small per-point records (python objects) and return them ordered by their
    global index.  `local` is a list of (global_index, record).  Works with any initialised
    torch.distributed backend (NCCL on GPU boxes, gloo in the CPU tests)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [r for _, r in sorted(local, key=lambda t: t[0])]
    bucket = [None] * dist.get_world_size()
    dist.all_gather_object(bucket, local)
    merged = [item for part in bucket for item in part]
    return [r for _, r in sorted(merged, key=lambda t: t[0])]


def solve_scan_parallel(base_text, key, values, omega0, device=0, group=None):
    """Scan-parallel driver (BASELINE config C5): `values` of the input key `key` are independent
    scan points (each with the explicit start `omega0`, a complex or one per point); they are dealt
    round-robin to the ranks of the process group, every rank solves its points with
    solve_once_eigen on its own GPU, and the records {value, omega, iterations, status} are
    gathered on every rank in scan order.  No collective touches the data path.

    A failing point is recorded like the reference's scan does (src/main.cpp:300-318):
    {"eigenvalue": "NaN", "reason": <message>}."""
    import re

    import torch.distributed as dist

    from .solver import Input, solve_once_eigen
    rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    starts = list(omega0) if isinstance(omega0, (list, tuple)) else [omega0] * len(values)
    local = []
    solver = None
    for k, v in scan_partition(list(values), rank, world):
        rec = {"scan_value": v, "rank": rank}
        try:
            inp = Input(text=base_text)
            inp.set_number(key, v)
            p, n = inp.params()
            if solver is not None and (n != solver.npoints or solver.dim != (n if p.beta_e == 0.0 else 2 * n)):
                solver.close()                      # the scanned key changed the mesh: new buffers
                solver = None
            if solver is None:
                solver = EigenSolver.from_input(inp, device=device)
            else:                                   # same mesh: reuse the handle and its buffers
                capi.check(solver._lib.emme_set_params(solver._h, p))
                tabs = [np.ascontiguousarray(t) for t in inp.tables()]
                capi.check(solver._lib.emme_set_tables(solver._h, *[t.ctypes.data for t in tabs]))
                solver.synchronize()
            w, iters, _ = solve_once_eigen(inp, starts[k], solver=solver)
            rec.update(eigenvalue=[w.real, w.imag], iterations=len(iters),
                       converged=abs(iters[-1][1]) < abs(inp.number("iteration_precision") * w))
        except Exception as e:                       # noqa: BLE001 - recorded, scan continues
            rec.update(eigenvalue="NaN", reason=str(e))
        local.append((k, rec))
    if solver is not None:
        solver.close()
    return gather_results(local)


def solve_scan_texts(texts, starts, device=0, group=None):
    """Scan-parallel driver over complete input.json texts (BASELINE config C5: one file per
    point, each with its own start): point k goes to rank k % world, a rank reuses one handle for
    all its points of the same mesh, records are gathered on every rank in scan order."""
    import torch.distributed as dist

    from .solver import Input, solve_once_eigen
    on = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if on else 0
    world = dist.get_world_size(group) if on else 1
    local = []
    solver = None
    for k, txt in scan_partition(list(texts), rank, world):
        rec = {"point": k, "rank": rank}
        try:
            inp = Input(text=txt)
            p, n = inp.params()
            if solver is not None and (n != solver.npoints or solver.dim != (n if p.beta_e == 0.0 else 2 * n)):
                solver.close()                      # another mesh: the handle's buffers do not fit
                solver = None
            if solver is None:
                solver = EigenSolver.from_input(inp, device=device)
            else:                                   # same mesh: reuse the handle and its buffers
                capi.check(solver._lib.emme_set_params(solver._h, p))
                tabs = [np.ascontiguousarray(t) for t in inp.tables()]
                capi.check(solver._lib.emme_set_tables(solver._h, *[t.ctypes.data for t in tabs]))
                solver.synchronize()
            w, iters, _ = solve_once_eigen(inp, starts[k], solver=solver)
            rec.update(eigenvalue=[w.real, w.imag], iterations=len(iters),
                       converged=bool(abs(iters[-1][1]) < abs(inp.number("iteration_precision") * w)))
        except Exception as e:                       # noqa: BLE001 - recorded, scan continues
            rec.update(eigenvalue="NaN", reason=str(e))
            # a numerical failure ("Linear solve failed", a singular pivot) leaves the handle healthy
            # (tests/test_newton_gpu.py); only a device or peer error retires it
            if solver is not None and getattr(e, "code", None) in (capi.E_CUDA, capi.E_PEER, capi.E_NO_DEVICE):
                solver.close()
                solver = None
        local.append((k, rec))
    if solver is not None:
        solver.close()
    return gather_results(local)


def marker_shard(n_markers, rank, world):
    """(first, count) of the contiguous block of markers rank `rank` keeps (the same split as
    emme_pic_create_shard in emme_b200/csrc/pic.cu)."""
    per, rem = divmod(n_markers, world)
    return rank * per + min(rank, rem), per + (1 if rank < rem else 0)


class _DevVector:
    """Zero-copy torch view of `count` doubles of handle-owned device memory."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {
            "shape": (count,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class ShardedPIC:
    """PIC_State + Integrator with the markers split over the ranks of a process group.

    exchange="p2p" (default): the ranks map each other's exchange buffers with CUDA IPC once; from
    then on step() is the plain emme_pic_step -- the field kernel of every stage stores the rank's
    density into every peer's buffer over NVLink, signals per 32-cell block with release stores and
    adds the contributions in rank order (csrc/pic.cu, pic_field_kernel mode 3).  No collective call,
    no host synchronisation inside a step, every rank ends with the same field history bit for bit.
    exchange="nccl" is the baseline it replaced: one all-reduce of 2*npoints doubles per stage.

    ShardedPIC(params, markers): every rank passes ALL markers (the p_weight normalisation then runs
    over all of them in the reference's order) and keeps its block.  ShardedPIC.from_seed draws only
    the rank's own block and reduces the normalisation with one 8-byte all-reduce."""

    def __init__(self, params, markers=None, device=0, group=None, exchange="p2p", _state=None):
        import torch
        import torch.distributed as dist

        from .pic import PIC_State
        self._group = group
        self._device = device
        on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if on else 0
        self.world = dist.get_world_size(group) if on else 1
        self.exchange = exchange if self.world > 1 else "none"
        self.state = _state or PIC_State(params, markers=markers, device=device, shard=(self.rank, self.world))
        self.nf = self.state.nf
        if self.exchange == "p2p":
            everyone = [None] * self.world
            dist.all_gather_object(everyone, self.state.ipc_export(), group=group)
            for r, h in enumerate(everyone):
                self.state.ipc_import(r, self.world, h)
            dist.barrier(group=group)                # nobody stores before everyone has mapped
        elif self.exchange == "nccl":
            self._stream = torch.cuda.ExternalStream(self.state.stream(), device=device)
            self._dens = torch.as_tensor(_DevVector(self.state.density_ptr(), 2 * self.nf),
                                         device=f"cuda:{device}")

    @classmethod
    def from_seed(cls, params, n_markers, seed=1, device=0, group=None, exchange="p2p"):
        """Rank r draws ITS block of the n_markers markers from its own std::mt19937 stream (seed + r,
        the reference's distributions and draw order) -- no rank generates or holds another rank's
        markers -- and the p_weight normalisation (a sum over all markers) is completed with one
        8-byte all-reduce.  The marker set is therefore not the one a single stream would give; the
        reference seeds from std::random_device (include/solver_pic.h:356-359), so its own runs
        differ from each other in the same way."""
        import torch
        import torch.distributed as dist

        from .pic import PIC_State, load_markers, pweight_sum
        on = dist.is_available() and dist.is_initialized()
        rank = dist.get_rank(group) if on else 0
        world = dist.get_world_size(group) if on else 1
        first, count = marker_shard(n_markers, rank, world)
        block = load_markers(params, count, seed=seed + rank)
        total = pweight_sum(params, block[1], block[2])
        if world > 1:
            dev = f"cuda:{device}" if dist.get_backend(group) == "nccl" else "cpu"
            t = torch.tensor([total], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            total = float(t.item())
        state = PIC_State.from_block(params, n_markers, first, block, total, (rank, world), device=device)
        return cls(params, device=device, group=group, exchange=exchange, _state=state)

    def exchange_description(self):
        if self.exchange == "nccl":
            return "one NCCL all-reduce (sum) of 2*npoints doubles per stage on the handle's stream"
        return ("fused into the field kernel: per-stage density stored into every peer's buffer over NVLink "
                "(CUDA IPC), release/acquire flags per 32-cell block, fixed-order sum; no collective call")

    def step(self, dt, nsteps=1):
        import torch
        import torch.distributed as dist
        if self.exchange != "nccl":
            self.state.step(dt, nsteps)
            return
        with torch.cuda.stream(self._stream):      # collectives are ordered after the stage kernels
            for _ in range(nsteps):
                for stage in range(3):
                    self.state.stage_begin(dt, stage)
                    dist.all_reduce(self._dens, op=dist.ReduceOp.SUM, group=self._group)
                    self.state.stage_finish(stage)

    def synchronize(self):
        import torch
        torch.cuda.synchronize(self._device)

    def current_field(self):
        return self.state.current_field()

    def field_history(self, first=0, count=None):
        return self.state.field_history(first, count)

    def markers(self):
        """(eta, weight) of this rank's block, in the caller's marker order."""
        return self.state.markers()

    def close(self):
        import torch.distributed as dist
        if self.exchange == "p2p" and dist.is_initialized() and getattr(self.state, "_h", None):
            self.synchronize()
            dist.barrier(group=self._group)          # peers may still be storing into my buffer
        self.state.close()


class LocalShardedPIC:
    """ShardedPIC's p2p protocol inside ONE process: `world` states (rank r on devices[r], several
    ranks per device allowed) attached with emme_pic_peer_attach, one host thread per rank."""

    def __init__(self, params, markers, devices):
        from .pic import PIC_State
        self.world = len(devices)
        self.ranks = [PIC_State(params, markers=markers, device=d, shard=(r, self.world))
                      for r, d in enumerate(devices)]
        for s in self.ranks:
            for q, peer in enumerate(self.ranks):
                s.peer_attach(q, self.world, peer)

    def step(self, dt, nsteps=1):
        import threading
        errs = []

        def run(s):
            try:
                s.step(dt, nsteps)
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        th = [threading.Thread(target=run, args=(s,)) for s in self.ranks]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errs:
            raise errs[0]

    def close(self):
        for s in self.ranks:
            s.close()


class _DevMatrix:
    """Zero-copy torch view of a handle-owned device matrix (for NCCL collectives)."""

    def __init__(self, ptr, dim, device):
        self.__cuda_array_interface__ = {
            "shape": (dim, dim, 2), "typestr": "<f8", "data": (int(ptr), False), "version": 2}
        self.device = device


def device_view(solver, which=0, device=0):
    import torch
    ptr = solver.matrix_device_ptr(which)
    return torch.as_tensor(_DevMatrix(ptr, solver.dim, device), device=f"cuda:{device}")


class ShardedEigenSolver(EigenSolver):
    """EigenSolver whose assemblies -- and, from dim 4096 up, dense steps -- are split over the
    ranks of a process group (one process per GPU).

    Same public surface (seed, newtonTraceSecantIteration, eigen_value ...); every rank ends each
    call with the full matrices and the same eigen_value."""

    def __init__(self, params, npoints, eta, g, bi, device=0, group=None, exchange="p2p", shard_dense=None):
        import torch.distributed as dist
        super().__init__(params, npoints, eta, g, bi, device=device)
        self._group = group
        self._device = device
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.shard_config(self.rank, self.world)
        self.exchange = exchange if self.world > 1 else "none"
        self.dense_sharded = False
        if self.exchange == "p2p":
            self._map_peers()
            if shard_dense is None:
                shard_dense = self.dim >= 4096
            if shard_dense:
                capi.check(self._lib.emme_shard_dense(self._h, 1))
                self.dense_sharded = True

    def _map_peers(self):
        """Exchange CUDA IPC handles of the peer-visible buffers; every rank maps every peer."""
        import torch.distributed as dist
        mine = []
        for which in range(capi.PEER_BUFS):
            buf = C.create_string_buffer(64)
            capi.check(self._lib.emme_ipc_export(self._h, which, buf))
            mine.append(buf.raw)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self._group)
        for r, handles in enumerate(everyone):
            for which in range(capi.PEER_BUFS):
                capi.check(self._lib.emme_ipc_import(self._h, r, self.world, which,
                                                     C.create_string_buffer(handles[which], 64)))
        dist.barrier(group=self._group)          # nobody starts storing before everyone has mapped

    def close(self):
        import torch.distributed as dist
        if getattr(self, "_h", None) and self.exchange == "p2p" and dist.is_initialized():
            self.synchronize()
            dist.barrier(group=self._group)      # peers may still be storing into my buffers
        super().close()

    __del__ = EigenSolver.close

    # ---- exchange="allreduce" (NCCL baseline): begin / complete / finish ----
    def _complete(self):
        import torch
        import torch.distributed as dist
        self.synchronize()
        t = device_view(self, 0, self._device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self._group)
        torch.cuda.current_stream().synchronize()

    def seed(self, omega0):
        if self.exchange != "allreduce":
            return super().seed(omega0)          # p2p: the exchange is inside the library call
        w = complex(omega0)
        capi.check(self._lib.emme_seed_begin(self._h, w.real, w.imag))
        self._complete()
        capi.check(self._lib.emme_seed_middle(self._h))
        self._complete()
        capi.check(self._lib.emme_seed_finish(self._h))
        self._pull()

    def _step(self, begin):
        rc = begin(self._h)
        self._pull()
        capi.check(rc)
        self._complete()
        v = [C.c_double() for _ in range(4)]
        capi.check(self._lib.emme_step_finish(self._h, *[C.byref(x) for x in v]))
        self._pull()

    def newtonTraceSecantIteration(self):
        if self.exchange != "allreduce":
            return super().newtonTraceSecantIteration()
        self._step(self._lib.emme_step_begin)

    def newtonQRSecantIteration(self):
        if self.exchange != "allreduce":
            return super().newtonQRSecantIteration()
        self._step(self._lib.emme_qr_step_begin)


class LocalShardedGroup:
    """The ShardedEigenSolver protocol driven from ONE process: `world` handles (rank r on
    devices[r]; several ranks may share a device) attached to each other with emme_peer_attach, one
    host thread per rank for the calls that synchronise (seed, iterates).  Every rank ends with the
    same eigen_value; `ranks[r]` are plain EigenSolver objects for inspection."""

    def __init__(self, params, npoints, eta, g, bi, devices, shard_dense=True):
        self.world = len(devices)
        self.ranks = [EigenSolver(params, npoints, eta, g, bi, device=d) for d in devices]
        lib = self.ranks[0]._lib
        for r, s in enumerate(self.ranks):
            s.shard_config(r, self.world)
        for r, s in enumerate(self.ranks):
            for q, peer in enumerate(self.ranks):
                capi.check(lib.emme_peer_attach(s._h, q, self.world, peer._h))
        self.dense_sharded = False
        if shard_dense:
            for s in self.ranks:
                capi.check(lib.emme_shard_dense(s._h, 1))
            self.dense_sharded = True

    def _all(self, fn):
        import threading
        errs = [None] * self.world

        def run(r):
            try:
                fn(self.ranks[r])
            except Exception as e:  # noqa: BLE001 - re-raised below on the calling thread
                errs[r] = e
        th = [threading.Thread(target=run, args=(r,)) for r in range(self.world)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        bad = [(r, e) for r, e in enumerate(errs) if e is not None]
        if bad:
            # a rank that failed for its own reason makes its peers time out: show every rank's error
            msg = "; ".join(f"rank {r}: [{getattr(e, 'code', type(e).__name__)}] {e}" for r, e in bad)
            codes = {getattr(e, "code", None) for _, e in bad}
            own = [c for c in codes if c not in (None, capi.E_PEER)]     # prefer a rank's own failure
            raise capi.EmmeError(own[0] if own else bad[0][1].code if hasattr(bad[0][1], "code") else -1, msg)

    def seed(self, omega0):
        self._all(lambda s: s.seed(omega0))

    def newtonTraceSecantIteration(self):
        self._all(lambda s: s.newtonTraceSecantIteration())

    def newtonQRSecantIteration(self):
        self._all(lambda s: s.newtonQRSecantIteration())

    @property
    def eigen_value(self):
        return self.ranks[0].eigen_value

    @property
    def d_eigen_value(self):
        return self.ranks[0].d_eigen_value

    def close(self):
        for s in self.ranks:
            s.synchronize()
        for s in self.ranks:
            s.close()
