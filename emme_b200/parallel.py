"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink) for the
exchange; the compute stays in libemme_b200.so.

Two ways the path shards:

* scan-parallel (weak scaling, no data-path collective): independent scan points / omega
  starts are dealt round-robin to ranks; only (omega, iterations, status) is gathered at the end.
  `scan_partition`, `gather_results`.
* row/pair-sharded assembly (strong scaling): the work items of ONE assembly (pairs i<j in
  diagonal-major order, times 3 modes when electromagnetic) are dealt round-robin to ranks
  (`shard_items`); each rank fills its entries of a zeroed dim x dim buffer and one all-reduce
  (sum) over NVLink completes the matrix on every rank -- the shares are disjoint, so the sum is
  bit-identical to a single-GPU assembly; the dense step is then replicated on every rank
  (identical inputs give identical delta, no broadcast needed).  `ShardedEigenSolver`.
"""
import ctypes as C

import numpy as np

from . import capi
from .solver import EigenSolver


def shard_items(n_items, rank, world):
    """Number of work items rank `rank` owns under the kernel's round-robin rule
    (global item = k*world + rank, emme_b200/csrc/assembly.cu)."""
    return (n_items - rank + world - 1) // world if n_items > rank else 0


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
    """EigenSolver whose assemblies are split over the ranks of a process group.

    Same public surface (seed, newtonTraceSecantIteration, eigen_value ...); every rank ends each
    call with the full matrices and the same eigen_value."""

    def __init__(self, params, npoints, eta, g, bi, device=0, group=None):
        import torch.distributed as dist
        super().__init__(params, npoints, eta, g, bi, device=device)
        self._group = group
        self._device = device
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.shard_config(self.rank, self.world)

    def _complete(self):
        """Sum the ranks' disjoint shares of eigen_matrix (all-reduce over NVLink)."""
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return
        self.synchronize()                       # shard kernel finished on the handle's stream
        t = device_view(self, 0, self._device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self._group)
        torch.cuda.current_stream().synchronize()

    def seed(self, omega0):
        w = complex(omega0)
        capi.check(self._lib.emme_seed_begin(self._h, w.real, w.imag))
        self._complete()
        capi.check(self._lib.emme_seed_middle(self._h))
        self._complete()
        capi.check(self._lib.emme_seed_finish(self._h))
        self._pull()

    def newtonTraceSecantIteration(self):
        rc = self._lib.emme_step_begin(self._h)
        self._pull()
        capi.check(rc)
        self._complete()
        v = [C.c_double() for _ in range(4)]
        capi.check(self._lib.emme_step_finish(self._h, *[C.byref(x) for x in v]))
        self._pull()
