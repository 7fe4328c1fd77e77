"""Minimal host-side analogue of the PartitionedArrays.jl pieces the assembly path touches.

The reference delegates all index-set bookkeeping and neighbour exchanges to
PartitionedArrays.jl v0.3 (not vendored under /root/reference).
What the hot path needs from it is small and is restated here:

* index sets ``LocalIndices`` (arbitrary own/ghost interleaving: FE-space dofs,
  reference FESpaces.jl:249-255) and ``OwnAndGhostIndices`` (own first, then ghosts:
  rows/cols of the linear system, reference Algebra.jl:1245-1252),
* ``PRange`` = one index set per part,
* two execution backends mirroring ``with_debug`` / ``with_mpi`` (reference README.md:21-37,
  test/sequential/*.jl vs test/mpi/*.jl): ``DebugBackend`` keeps every part in this process,
  ``DistBackend`` keeps ONE part per ``torch.distributed`` rank (one rank per GPU),
* sparse neighbour ``exchange`` of variable-length segments, ``scan`` and ``reduction``.

All ids are the reference's: parts, local ids and global ids are 1-based; ``0`` means "absent".
Everything here is control plane (setup, once).  The data plane (ghost-row values, halo of x)
runs inside libgraft.so over NCCL.
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------------------------
# Backends
# ----------------------------------------------------------------------------------------------


class DebugBackend:
    """All parts live in this process (PartitionedArrays' DebugArray, reference README.md:23)."""

    name = "debug"

    def __init__(self, nparts: int):
        self.nparts = int(nparts)
        self.parts_here = list(range(1, self.nparts + 1))

    def exchange(self, snd):
        """snd[k] = {dest_part: ndarray} for local part k.  Returns rcv[k] = {src_part: ndarray}."""
        rcv = [dict() for _ in self.parts_here]
        for k, p in enumerate(self.parts_here):
            for q, data in snd[k].items():
                rcv[q - 1][p] = np.array(data, copy=True)
        return rcv

    def scan_exclusive(self, vals, init=0):
        out, acc = [], init
        for v in vals:
            out.append(acc)
            acc = acc + v
        return out

    def reduction(self, vals):
        s = sum(vals)
        return [s for _ in vals]

    def allgather(self, vals):
        return [list(vals) for _ in vals]

    def barrier(self):
        pass


class DistBackend:
    """One part per torch.distributed rank (the analogue of ``with_mpi``; one rank per GPU).

    Host-side control-plane exchanges go through a gloo group so that they also work in the
    CPU tests (world_size 2, gloo) and do not occupy the NCCL data-plane communicator.
    """

    name = "dist"

    def __init__(self, group=None):
        import torch.distributed as dist

        self.dist = dist
        self.rank = dist.get_rank()
        self.nparts = dist.get_world_size()
        self.parts_here = [self.rank + 1]
        if group is None:
            if dist.get_backend() == "gloo":
                group = dist.group.WORLD
            else:
                group = dist.new_group(backend="gloo")
        self.group = group

    def allgather(self, vals):
        out = [None] * self.nparts
        self.dist.all_gather_object(out, vals[0], group=self.group)
        return [out]

    def scan_exclusive(self, vals, init=0):
        allv = self.allgather(vals)[0]
        acc = init
        for v in allv[: self.rank]:
            acc = acc + v
        return [acc]

    def reduction(self, vals):
        return [sum(self.allgather(vals)[0])]

    def exchange(self, snd):
        import torch

        dist = self.dist
        me = self.rank + 1
        mine = snd[0]
        meta = {q: (str(np.asarray(a).dtype), int(np.asarray(a).size)) for q, a in mine.items()}
        allmeta = self.allgather([meta])[0]
        reqs, keep, rcv = [], [], {}
        for src0, m in enumerate(allmeta):
            src = src0 + 1
            if src != me and me in m:
                dt, n = m[me]
                buf = torch.empty(n, dtype=getattr(torch, _TORCH_DT[dt]))
                rcv[src] = buf
                if n:
                    reqs.append(dist.irecv(buf, src=src0, group=self.group))
        for q, a in mine.items():
            if q == me:
                rcv[me] = torch.from_numpy(np.array(a, copy=True).ravel())
                continue
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(a).ravel()))
            keep.append(t)
            if t.numel():
                reqs.append(dist.isend(t, dst=q - 1, group=self.group))
        for r in reqs:
            r.wait()
        return [{s: b.numpy() for s, b in rcv.items()}]

    def barrier(self):
        self.dist.barrier(group=self.group)


_TORCH_DT = {"int32": "int32", "int64": "int64", "float64": "float64", "bool": "bool", "uint8": "uint8"}


def with_debug(fn, nparts=None):
    """``with_debug() do distribute`` (reference test/sequential/PoissonTests.jl:5-7)."""

    def distribute(n):
        return DebugBackend(int(np.prod(n)))

    return fn(distribute)


def with_dist(fn):
    """``with_mpi() do distribute`` (reference test/mpi/runtests_np4.jl:42-52), over torch.distributed."""

    def distribute(n):
        b = DistBackend()
        assert int(np.prod(n)) == b.nparts, "number of parts must equal the world size"
        return b

    return fn(distribute)


# ----------------------------------------------------------------------------------------------
# Index sets
# ----------------------------------------------------------------------------------------------


class _G2L:
    """global->local lookup without a dense n_global table (133 M gids at 256^3)."""

    def __init__(self, l2g):
        self.order = np.argsort(l2g, kind="stable")
        self.sorted = l2g[self.order]

    def __call__(self, gids):
        gids = np.asarray(gids, dtype=np.int64)
        pos = np.searchsorted(self.sorted, gids)
        pos = np.minimum(pos, len(self.sorted) - 1) if len(self.sorted) else np.zeros_like(pos)
        if len(self.sorted) == 0:
            return np.zeros(gids.shape, dtype=np.int64)
        hit = self.sorted[pos] == gids
        return np.where(hit, self.order[pos] + 1, 0).astype(np.int64)


class LocalIndices:
    """``LocalIndices(n_global, me, local_to_global, local_to_owner)`` (reference FESpaces.jl:254)."""

    def __init__(self, n_global, part, l2g, l2o):
        self.n_global = int(n_global)
        self.part = int(part)
        self.l2g = np.ascontiguousarray(l2g, dtype=np.int64)
        self.l2o = np.ascontiguousarray(l2o, dtype=np.int32)
        own = self.l2o == self.part
        self.own_to_local = (np.flatnonzero(own) + 1).astype(np.int32)
        self.ghost_to_local = (np.flatnonzero(~own) + 1).astype(np.int32)
        self.local_to_own = np.zeros(len(self.l2g), dtype=np.int32)
        self.local_to_own[self.own_to_local - 1] = np.arange(1, len(self.own_to_local) + 1)
        self.local_to_ghost = np.zeros(len(self.l2g), dtype=np.int32)
        self.local_to_ghost[self.ghost_to_local - 1] = np.arange(1, len(self.ghost_to_local) + 1)
        self._g2l = None

    # PartitionedArrays accessors ---------------------------------------------------------
    @property
    def local_length(self):
        return len(self.l2g)

    @property
    def own_length(self):
        return len(self.own_to_local)

    @property
    def ghost_length(self):
        return len(self.ghost_to_local)

    @property
    def own_to_global(self):
        return self.l2g[self.own_to_local - 1]

    @property
    def ghost_to_global(self):
        return self.l2g[self.ghost_to_local - 1]

    @property
    def ghost_to_owner(self):
        return self.l2o[self.ghost_to_local - 1]

    def global_to_local(self, gids):
        if self._g2l is None:
            self._g2l = _G2L(self.l2g)
        return self._g2l(gids)


class OwnAndGhostIndices(LocalIndices):
    """``OwnAndGhostIndices(OwnIndices(n,me,own_to_global), GhostIndices(n,g2g,g2o))``.

    Local numbering: own ids first (in ``own_to_global`` order), then ghosts (SURVEY A.8;
    built at reference Algebra.jl:1178-1180 and :1249-1251).
    """

    def __init__(self, n_global, part, own_to_global, ghost_to_global, ghost_to_owner):
        o2g = np.asarray(own_to_global, dtype=np.int64)
        g2g = np.asarray(ghost_to_global, dtype=np.int64)
        g2o = np.asarray(ghost_to_owner, dtype=np.int32)
        l2g = np.concatenate([o2g, g2g])
        l2o = np.concatenate([np.full(len(o2g), part, dtype=np.int32), g2o])
        super().__init__(n_global, part, l2g, l2o)


class PRange:
    """One index set per part held by this process (``PRange(indices)``)."""

    def __init__(self, backend, indices):
        self.backend = backend
        self.indices = list(indices)

    def partition(self):
        return self.indices

    def __len__(self):
        return self.indices[0].n_global if self.indices else 0


def assembly_neighbors(prange: PRange):
    """``assembly_neighbors`` [ext]: parts_snd = sorted owners of my ghosts; parts_rcv = the parts
    that hold ghosts owned by me, in ascending part order (assumption recorded in DESIGN.md).
    """
    b = prange.backend
    snd = [sorted(set(int(o) for o in np.unique(ids.ghost_to_owner))) for ids in prange.indices]
    allsnd = b.allgather(snd)
    rcv = []
    for k, p in enumerate(b.parts_here):
        lists = allsnd[k]
        rcv.append([q + 1 for q, lst in enumerate(lists) if p in lst])
    return snd, rcv


def uniform_local_range(p, np_, n, ghost=False, periodic=False):
    """``_local_range(p,np,n,ghost,periodic)`` (reference PArraysExtras.jl:72-84); 1-based inclusive.

    The remainder cells go to the LAST parts.
    """
    l, rem = divmod(n, np_)
    offset = l * (p - 1)
    if rem >= (np_ - p + 1):
        l += 1
        offset += p - (np_ - rem) - 1
    g = 1 if ghost else 0
    start = 1 + offset - g
    stop = l + offset + g
    if periodic:
        return start, stop
    return max(1, start), min(n, stop)
