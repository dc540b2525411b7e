"""Peer-memory exchange of the cell-sharded pass: host side of ``wsage_peer_reduce`` (include/wsage.h).

One ``PeerGroup`` per process holds this rank's peer allocation and the mapped allocations of the other ranks of the
node.  ``reduce`` replaces  sum of the split-K slabs -> all-reduce over the ranks -> scale / self-loop epilogue  of the
gene destinations (/root/reference/models/gnn.py:65 has a single process and no counterpart) by one kernel that reads
the other GPUs' partial sums over NVLink.  The handles travel through ``torch.distributed`` (plumbing only).
"""
import ctypes
import os
from ctypes import c_int32, c_void_p
from typing import Optional

import torch

from . import _lib


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class PeerGroup:
    def __init__(self, rank: int, world: int, max_elems: int, bases, own_base, opened, timeout_s: float = 60.0, grid: int = 0):
        self.rank, self.world, self.max_elems, self.timeout_s, self.grid = rank, world, int(max_elems), float(timeout_s), int(grid)
        self._bases = (c_void_p * world)(*bases)
        self._own, self._opened = own_base, opened
        self._epoch = 1
        self.calls = 0

    # ---- construction ----
    @staticmethod
    def _alloc(max_elems):
        lib = _lib.load()
        base = c_void_p()
        handle = ctypes.create_string_buffer(64)
        _lib.check(lib.wsage_peer_alloc(int(max_elems), ctypes.byref(base), handle), "wsage_peer_alloc")
        return base.value, handle.raw

    @classmethod
    def create(cls, max_elems: int, group=None, timeout_s: float = 60.0) -> "PeerGroup":
        """Collective over ``group`` (default: the world), all ranks on one node with peer access between their GPUs."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        if world > _lib.PEER_MAX:
            raise RuntimeError(f"peer exchange spans at most {_lib.PEER_MAX} ranks of one node, got {world}")
        lib = _lib.load()
        own = handle = err = None
        try:
            own, handle = cls._alloc(max_elems)
        except RuntimeError as e:
            err = str(e)
        handles = [None] * world
        dist.all_gather_object(handles, handle, group=group)
        bases, opened = [], []
        for r in range(world):
            if r == rank or own is None or handles[r] is None:
                bases.append(own if r == rank else None)
                continue
            b = c_void_p()
            if lib.wsage_peer_open(handles[r], ctypes.byref(b)) != _lib.OK:
                err = lib.wsage_last_error().decode("utf-8", "replace")
                bases.append(None)
            else:
                bases.append(b.value)
                opened.append(b.value)
        # every rank learns whether every rank could map every peer: either all use the kernel or none does
        errs = [None] * world
        dist.all_gather_object(errs, err, group=group)
        pg = cls(rank, world, max_elems, bases, own, opened, timeout_s)
        if any(e is not None for e in errs):
            pg.close()
            raise RuntimeError("peer memory is not available on every rank: " + "; ".join(f"rank {r}: {e}" for r, e in enumerate(errs) if e))
        return pg

    @classmethod
    def local_ranks(cls, world: int, max_elems: int, timeout_s: float = 60.0):
        """``world`` groups inside ONE process on one device (the kernels of the ranks then run side by side on separate
        streams): exercises the exchange without a second GPU.  All the ranks' CTAs must be resident together (two
        per SM fit), so each rank gets its share of the device."""
        allocs = [cls._alloc(max_elems)[0] for _ in range(world)]
        grid = 0 if world == 1 else 296 // world
        return [cls(r, world, max_elems, allocs, allocs[r], [], timeout_s, grid) for r in range(world)]

    # ---- the exchange ----
    def reduce(self, slabs: torch.Tensor, rows: int, *, slot_of_row: Optional[torch.Tensor] = None, dscale=None, selfcoef=None,
               hself=None, out=None, raw=None, stream=None):
        """raw[r] = Σ_ranks Σ_slabs slabs[k, slot_of_row[r]];  out[r] = dscale[r]·raw[r] + selfcoef[r]·hself[r]."""
        if self._own is None:
            raise RuntimeError("PeerGroup is closed")
        assert slabs.dim() == 3 and slabs.is_contiguous() and slabs.dtype == torch.float32
        a = _lib.PeerReduceArgs()
        a.rank, a.world, a.bases, a.max_elems, a.epoch = self.rank, self.world, self._bases, self.max_elems, self._epoch
        a.slabs, a.n_slabs, a.slab_rows, a.slot_of_row = _ptr(slabs), slabs.shape[0], slabs.shape[1], _ptr(slot_of_row)
        a.rows, a.dim = int(rows), slabs.shape[2]
        a.dscale, a.selfcoef = _ptr(dscale), _ptr(selfcoef)
        if selfcoef is not None:
            a.hself, a.ld_hself = _ptr(hself), hself.stride(0)
        if out is not None:
            a.out, a.ld_out = _ptr(out), out.stride(0)
        if raw is not None:
            a.raw, a.ld_raw = _ptr(raw), raw.stride(0)
        a.timeout_s, a.grid = self.timeout_s, self.grid
        st = stream if stream is not None else torch.cuda.current_stream()
        _lib.check(_lib.load().wsage_peer_reduce(ctypes.byref(a), c_void_p(st.cuda_stream)), "wsage_peer_reduce")
        self._epoch += 2
        self.calls += 1

    def check(self):
        """Synchronises; raises if a barrier of an earlier ``reduce`` timed out (a rank did not take part)."""
        if self._own is None:
            return
        status = c_int32(0)
        _lib.check(_lib.load().wsage_peer_status(c_void_p(self._own), ctypes.byref(status)), "wsage_peer_status")
        if status.value != 0:
            raise RuntimeError(f"wsage_peer_reduce: a cross-GPU barrier timed out on rank {self.rank} after {self.timeout_s} s; "
                               "the exchange buffers are unusable")

    def close(self, free=True):
        lib = _lib.load()
        for b in self._opened:
            lib.wsage_peer_close(c_void_p(b))
        self._opened = []
        if free and self._own is not None:
            lib.wsage_peer_free(c_void_p(self._own))
        self._own = None


_active: Optional[PeerGroup] = None


def active() -> Optional[PeerGroup]:
    return _active


def enable(max_elems: int, group=None) -> Optional[PeerGroup]:
    """Collective.  Sets up the process-wide group the sharded forward uses; returns None (and leaves the NCCL all-reduce
    in place) when WSAGE_PEER=0 or when the GPUs cannot map each other's memory."""
    global _active
    import torch.distributed as dist
    if os.environ.get("WSAGE_PEER", "1") == "0" or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) < 2:
        return None
    if _active is not None and _active.max_elems >= max_elems and _active._own is not None:
        return _active
    disable()
    try:
        _active = PeerGroup.create(max_elems, group)
    except RuntimeError as e:
        import warnings
        warnings.warn(f"peer-memory exchange unavailable, using the NCCL all-reduce: {e}")
        _active = None
    return _active


def disable():
    global _active
    if _active is not None:
        import torch.distributed as dist
        torch.cuda.synchronize()
        if dist.is_available() and dist.is_initialized():
            dist.barrier()                  # nobody still reads this rank's buffers
        _active.close()
        _active = None
