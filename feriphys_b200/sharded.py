"""One flock over several GPUs: one process per GPU, ``torch.distributed`` for the
rendezvous only (it broadcasts the 128-byte ``ncclUniqueId``); the data path is the
library's own NCCL calls over NVLink (``fp_shard.cu``).

* all-pairs: ranks own contiguous boid-index ranges, one ``ncclAllGather`` of
  positions/velocities per step -- bit-identical to a single GPU;
* grid: ranks own x-slabs of whole cell layers, one halo + migration exchange with
  each neighbour per step.

Every rank must make the same calls in the same order (SPMD).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, f32c, ptr
from .flocking import Simulation


def shard_range(n: int, rank: int, world: int):
    """Index range ``(first, count)`` rank ``rank`` holds at creation: ``per = ceil(n/world)``
    rows each, the last ranks possibly fewer (must match ``fp_flock_create_sharded``)."""
    per = (n + world - 1) // world
    first = min(rank * per, n)
    return first, min(per, n - first)


def broadcast_unique_id(dist, device=None) -> np.ndarray:
    """rank 0 makes the NCCL unique id; everyone else receives it through ``dist``."""
    import torch
    lib = _lib.load()
    uid = np.zeros(128, np.uint8)
    if dist.get_rank() == 0:
        check(lib.fp_nccl_unique_id(ptr(uid)))
    t = torch.from_numpy(uid)
    if dist.get_backend() == "nccl":
        t = t.cuda(device)
    dist.broadcast(t, src=0)
    return t.cpu().numpy().copy()


class ShardedSimulation(Simulation):
    """``Simulation`` whose flock is sharded over ``dist.get_world_size()`` GPUs."""

    def __init__(self, local_state, n_global: int, first: int, dist, **kw):
        self._dist = dist
        self._n_global = int(n_global)
        self._first = int(first)
        self._uid = broadcast_unique_id(dist, kw.get("device", 0))
        super().__init__(None, 0, kw.pop("bounding_box", None), kw.pop("lead_boids", None),
                         kw.pop("obstacles", None), kw.pop("attractors", None), _state=local_state, **kw)

    @classmethod
    def from_global_slice(cls, local_state, n_global, first, dist, **kw):
        return cls(local_state, n_global, first, dist, **kw)

    @classmethod
    def from_global_state(cls, global_state, dist, **kw):
        """Every rank passes the same full ``[n, 6]`` state and keeps its own range."""
        st = f32c(global_state, (-1, 6))
        first, count = shard_range(len(st), dist.get_rank(), dist.get_world_size())
        return cls(st[first:first + count], len(st), first, dist, **kw)

    def _create(self, state: np.ndarray, device: int):
        self._n = self._n_global  # per-boid outputs are global, identical on every rank
        h = C.c_void_p()
        cfg = self.config.to_c()
        check(self._lib.fp_flock_create_sharded(
            C.byref(h), C.byref(cfg), self._n_global, self._first, len(state), ptr(state), device,
            self._dist.get_rank(), self._dist.get_world_size(), ptr(self._uid)))
        return h

    def write_state(self, state) -> None:
        raise NotImplementedError("write_state is not supported on a sharded flock")
