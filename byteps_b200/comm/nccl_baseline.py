"""The reference's single-box data path, reproduced with NCCL as the BASELINE.

For every <= BYTEPS_PARTITION_BYTES partition, in FIFO order, in groups of
BYTEPS_NCCL_GROUP_SIZE (4) reduce + 4 broadcast tasks on one dedicated
high-priority stream: ``ncclReduceScatter(sum)`` (+ ``ncclReduce`` of the
``len % nGPU`` tail to the root) into the output, then ``ncclAllGather``
(+ ``ncclBroadcast`` of the tail) in place; after all partitions of a tensor are
done, ``output.div_(size)`` (/root/reference/byteps/common/core_loops.cc:190-376,
nccl_manager.cc:74-165, torch/ops.cc:78-91).  This is what BASELINE.md calls the
"reference-style NCCL path"; nothing of it is used by the product path.
"""
from __future__ import annotations

import os
from contextlib import contextmanager, nullcontext

import torch
import torch.distributed as dist


class NcclReferencePath:
    def __init__(self, group=None, partition_bytes=None, group_size=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.partition_bytes = partition_bytes or int(os.environ.get("BYTEPS_PARTITION_BYTES", 4096000))
        self.group_size = group_size or int(os.environ.get("BYTEPS_NCCL_GROUP_SIZE", 4))
        self.stream = torch.cuda.Stream(priority=-1)
        self.root = self.world - 1            # the reference's root is the highest local rank

    @contextmanager
    def _grouped(self, device):
        cm = getattr(dist, "_coalescing_manager", None)
        if cm is None:
            yield
            return
        try:
            with cm(group=self.group, device=device, async_ops=False):
                yield
        except Exception:  # noqa: BLE001
            yield

    def _partitions(self, t: torch.Tensor):
        es = t.element_size()
        page = 4096 * max(self.world, 1)
        bound = (self.partition_bytes + page - 1) // page * page
        per = max(bound // es, 1)
        flat = t.view(-1)
        return [flat[i:i + per] for i in range(0, flat.numel(), per)]

    def push_pull_(self, tensors, average=True):
        """In-place push_pull of a list of tensors, reference style.  Returns an
        event recorded on the comm stream when everything (incl. div_) is done."""
        dev = tensors[0].device
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev))
        self.stream.wait_event(ready)
        parts = [p for t in tensors for p in self._partitions(t)]
        N = self.world
        with torch.cuda.stream(self.stream):
            for g0 in range(0, len(parts), self.group_size):
                grp = parts[g0:g0 + self.group_size]
                with (self._grouped(dev) if N > 1 else nullcontext()):     # 4 REDUCE tasks
                    for p in grp:
                        if N == 1:
                            continue
                        n = p.numel() // N
                        if n:
                            out = p[self.rank * n:(self.rank + 1) * n]
                            dist.reduce_scatter_tensor(out, p[:n * N], op=dist.ReduceOp.SUM, group=self.group)
                        if p.numel() % N:
                            dist.reduce(p[n * N:], dst=self.root, op=dist.ReduceOp.SUM, group=self.group)
                with (self._grouped(dev) if N > 1 else nullcontext()):     # 4 BROADCAST tasks
                    for p in grp:
                        if N == 1:
                            continue
                        n = p.numel() // N
                        if n:
                            dist.all_gather_into_tensor(p[:n * N], p[self.rank * n:(self.rank + 1) * n],
                                                        group=self.group)
                        if p.numel() % N:
                            dist.broadcast(p[n * N:], src=self.root, group=self.group)
            if average:
                for t in tensors:
                    t.div_(N)          # host-callback div_ in the reference
            done = torch.cuda.Event()
            done.record(self.stream)
        return done


class NativeNcclReferencePath:
    """Same path, issued by the native NcclManager (csrc/comm/nccl_manager.cc): raw
    ncclReduceScatter/AllGather (+Reduce/Broadcast tails) in ncclGroupStart/End batches on
    the manager's own highest-priority stream - no torch.distributed dispatch in between."""

    def __init__(self, group_ops, device, partition_bytes=None, group_size=None, num_rings=None):
        from .. import _native
        from .engine import core_dtype

        cu = _native.cuda()
        self._dt = core_dtype
        self.world, self.rank = group_ops.world, group_ops.rank
        self.partition_bytes = partition_bytes or int(os.environ.get("BYTEPS_PARTITION_BYTES", 4096000))
        gs = group_size or int(os.environ.get("BYTEPS_NCCL_GROUP_SIZE", 4))
        rings = num_rings or int(os.environ.get("BYTEPS_NCCL_NUM_RINGS", 1))
        dev = torch.device(device)
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        self.mgr = cu.NcclManager(self.rank, self.world, idx, rings, gs)
        ids = [cu.NcclManager.make_unique_id() for _ in range(rings)] if self.rank == 0 else None
        ids = group_ops.broadcast_object(ids, 0)
        self.mgr.init(ids)
        self.stream = torch.cuda.ExternalStream(self.mgr.stream(0), device=dev)
        self._key = 0

    def push_pull_(self, tensors, average=True):
        dev = tensors[0].device
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev))
        for t in tensors:
            flat = t if t.is_contiguous() else torch.as_strided(t, (t.numel(),), (1,), t.storage_offset())
            self.mgr.push_pull(flat.data_ptr(), flat.numel() * flat.element_size(), self._dt(t.dtype), self._key,
                               self.partition_bytes, ready.cuda_event, 0)
            self._key += 1
        with torch.cuda.stream(self.stream):
            if average:
                for t in tensors:
                    t.div_(self.world)
            done = torch.cuda.Event()
            done.record(self.stream)
        return done
