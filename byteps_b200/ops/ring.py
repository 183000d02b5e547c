"""Python face of the descriptor-ring exchange kernel (csrc/kernels/pushpull_ring.cu).

A :class:`RingTable` is a device-resident table of bucket / partition descriptors that one
``pushpull_ring`` launch consumes.  Every rank builds the same table (same windows, slots and
priorities); only the optimizer-state pointers are rank local.  This is the replacement for the
reference's per-partition NCCL groups driven by ``BytePSScheduledQueue``
(/root/reference/byteps/common/scheduled_queue.cc:82-163, core_loops.cc:271-360).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from .. import _native

_FMT = "<4Q4Qf3iI3I"      # csrc/kernels/pushpull_ring.cuh::RingDesc (96 bytes)


@dataclass
class RingEntry:
    grad_off: int
    numel: int                 # padded to a multiple of 8
    wire: int                  # WIRE_F32 / WIRE_BF16 / WIRE_F16
    slot: int
    kind: int = 0              # RING_ALLREDUCE / RING_SGD / RING_ADAM
    param_off: int = 0
    scale: float = 1.0
    priority: int = 0
    master: int = 0
    state0: int = 0
    state1: int = 0
    hp: int = 0

    @property
    def nbytes(self) -> int:
        return self.numel * (4 if self.wire == 0 else 2)

    def pack(self) -> bytes:
        if self.numel % 8:
            raise ValueError("ring descriptors cover multiples of 8 elements")
        return struct.pack(_FMT, self.grad_off, self.param_off, self.numel // 8, self.nbytes, self.master,
                           self.state0, self.state1, self.hp, self.scale, self.kind, self.wire, self.priority,
                           self.slot, 0, 0, 0)


class RingTable:
    """Descriptors of ONE class (wire dtype, kind) in consumption order, resident on the device."""

    def __init__(self, entries: Sequence[RingEntry], device):
        cu = _native.cuda()
        if not entries:
            raise ValueError("empty ring table")
        if len(entries) > cu.RING_SLOTS:
            raise ValueError("a ring launch takes at most %d descriptors" % cu.RING_SLOTS)
        cls = {(e.wire, e.kind) for e in entries}
        if len(cls) != 1:
            raise ValueError("all descriptors of a ring launch must share wire dtype and kind")
        if len({e.slot for e in entries}) != len(entries):
            raise ValueError("slots must be unique inside a launch")
        self.cu = cu
        cu.ring_preload()       # no first-time (lazy) kernel load may happen while a ring kernel is spinning
        self.entries: List[RingEntry] = list(entries)
        self.wire, self.kind = next(iter(cls))
        blob = b"".join(e.pack() for e in entries)
        assert len(blob) == cu.RING_DESC_BYTES * len(entries)
        host = torch.frombuffer(bytearray(blob), dtype=torch.uint8)
        self.dev = host.to(device)
        self.total_bytes = sum(e.nbytes for e in entries)

    def __len__(self):
        return len(self.entries)

    def ptr(self, first: int = 0) -> int:
        return self.dev.data_ptr() + first * self.cu.RING_DESC_BYTES

    def launch(self, view, blocks: int, stream: int, *, nvls: bool = False, sched: bool = False,
               self_mark: bool = True, credit_bytes: int = 0, first: int = 0, count: Optional[int] = None,
               solo: bool = False):
        n = len(self.entries) - first if count is None else count
        self.cu.pushpull_ring(view, self.wire, self.kind, self.ptr(first), n, blocks, nvls, sched, self_mark,
                              int(credit_bytes), stream, solo)

    def slots(self, first: int = 0, count: Optional[int] = None) -> List[int]:
        n = len(self.entries) - first if count is None else count
        return [e.slot for e in self.entries[first:first + n]]
