"""Framework-neutral plugin: push_pull for anything that speaks DLPack or the
buffer protocol (numpy arrays, CuPy, JAX, TensorFlow and MXNet tensors all do).

The reference ships one plugin per framework (tensorflow / keras / mxnet, SURVEY
P8-P13), each wrapping the same C core.  Those frameworks are not installable in
this image, so instead of three untestable copies this module exposes the
engine through the exchange format they all support: a tensor is imported
zero-copy with ``torch.from_dlpack`` (CPU or CUDA), pushed through the same
engine, and the result is visible in the caller's own tensor.

    import numpy as np, byteps_b200.dlpack as bps
    bps.init()
    g = np.ones(1000, dtype=np.float32)
    bps.push_pull_inplace(g, name="g")          # g now holds the average over all ranks
"""
from __future__ import annotations

import numpy as np
import torch

from ..torch import ops as _ops

init = _ops.init
shutdown = _ops.shutdown
suspend = _ops.suspend
resume = _ops.resume
rank = _ops.rank
size = _ops.size
local_rank = _ops.local_rank
local_size = _ops.local_size
poll = _ops.poll
declare = _ops.declare


def _as_torch(x) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("Tensor is required to be contiguous.")
        return torch.from_numpy(x)
    if hasattr(x, "__dlpack__"):
        return torch.from_dlpack(x)
    raise TypeError("push_pull needs a numpy array or an object with __dlpack__, got %s" % type(x))


_keep = {}


def push_pull_async_inplace(tensor, average=True, name=None, version=0, priority=0):
    """In-place push_pull of a DLPack/numpy tensor; returns a handle."""
    t = _as_torch(tensor)
    h = _ops.push_pull_async_inplace(t, average, name, version, priority)
    _keep[h] = tensor
    return h


def synchronize(handle):
    _ops.synchronize(handle)
    return _keep.pop(handle, None)


def push_pull_inplace(tensor, average=True, name=None, version=0, priority=0):
    return synchronize(push_pull_async_inplace(tensor, average, name, version, priority))


def push_pull(tensor, average=True, name=None, version=0, priority=0):
    """Out-of-place: returns a new numpy array (CPU inputs) or torch tensor (device inputs)."""
    t = _as_torch(tensor)
    out = _ops.synchronize(_ops.push_pull_async(t, average, name, version, priority))
    return out.numpy() if isinstance(tensor, np.ndarray) else out


def broadcast(tensor, root_rank=0, name=None):
    """Root's values everywhere (zero elsewhere, then sum - like the reference's broadcast)."""
    t = _as_torch(tensor)
    if rank() != root_rank:
        t.zero_()
    return push_pull_inplace(tensor, average=False, name=name)
