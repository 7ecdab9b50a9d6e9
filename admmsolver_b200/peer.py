"""Peer mailboxes of the sharded batch-wide criterion (include/admm_b200.h ``admm_peer_comm``).

The path shards over the batch index with no data-path collective.  The one exchange -- the all-reduce of
the ten squared-norm sums behind ``residual()`` / ``check_convergence()`` / ``update_mu()`` of the packed
batch (reference optimizer.py:232-299) -- is done by the kernels themselves: the reduction kernel pushes its
sums into every rank's mailbox over NVLink (peer-mapped ``cudaIpc`` memory) and the decision kernel polls its
own mailbox.  ``torch.distributed`` is only the bootstrap that carries the 64-byte IPC handles (any backend:
NCCL on the GPU box, gloo in the CPU tests of this host logic).
"""
from __future__ import annotations

import ctypes as C
import os
import socket
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import MAILBOX_BYTES, MAX_PEERS, PeerComm

__all__ = ["exchange_handles", "check_one_box", "PeerMailbox"]


def exchange_handles(group, payload: Tuple) -> List[Tuple]:
    """All-gather one small picklable record per rank over ``group`` (rank order).  Pure host logic."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out: List[Optional[Tuple]] = [None] * world
    dist.all_gather_object(out, payload, group=group)
    return out  # type: ignore[return-value]


def check_one_box(records: Sequence[Tuple]) -> None:
    """The mailboxes are CUDA-IPC mappings: every rank has to live on the same host, and no two ranks may be
    the same process (a process cannot open its own handle)."""
    hosts = {r[0] for r in records}
    if len(hosts) != 1:
        raise NotImplementedError("peer mailboxes need all ranks on one box, got hosts %s" % sorted(hosts))
    pids = [r[1] for r in records]
    if len(set(pids)) != len(pids):
        raise NotImplementedError("peer mailboxes need one process per rank")
    if len(records) > MAX_PEERS:
        raise NotImplementedError("at most %d ranks (one box), got %d" % (MAX_PEERS, len(records)))


class PeerMailbox:
    """This rank's mailbox plus the mapped mailboxes of all peers, as the ``admm_peer_comm`` the kernels take."""

    def __init__(self, group):
        import torch.distributed as dist
        dev = _lib.require_cuda()
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        own = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        _lib.call("admm_peer_alloc", MAILBOX_BYTES, C.byref(own), handle)
        self._own = own
        self._opened: List[C.c_void_p] = []
        records = exchange_handles(group, (socket.gethostname(), os.getpid(), bytes(handle)))
        check_one_box(records)
        self.ctrl = torch.zeros(4, dtype=torch.int32, device=dev)
        self.comm = PeerComm()
        self.comm.rank, self.comm.world = self.rank, self.world
        self.comm.ctrl = self.ctrl.data_ptr()
        for r, (_, _, h) in enumerate(records):
            if r == self.rank:
                self.comm.mbox[r] = own.value
                continue
            p = C.c_void_p()
            buf = (C.c_ubyte * 64).from_buffer_copy(h)
            _lib.call("admm_peer_open", buf, C.byref(p))
            self._opened.append(p)
            self.comm.mbox[r] = p.value
        torch.cuda.synchronize()
        dist.barrier(group=group)          # nobody posts before every mailbox is mapped everywhere

    def close(self) -> None:
        if self._own is None:
            return
        torch.cuda.synchronize()
        for p in self._opened:
            _lib.call("admm_peer_close", p)
        self._opened = []
        try:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.barrier(group=self.group)     # peers have unmapped before the owner frees
        except Exception:
            pass
        _lib.call("admm_peer_free", self._own)
        self._own = None

    def __del__(self):
        # best effort without collectives (interpreter shutdown order is arbitrary)
        try:
            if self._own is not None:
                for p in self._opened:
                    _lib.lib.admm_peer_close(p)
                _lib.lib.admm_peer_free(self._own)
                self._own = None
        except Exception:
            pass
