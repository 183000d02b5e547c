"""MXNet tensor entry points (parity: /root/reference/byteps/mxnet/ops.py:48-123 and the C
side /root/reference/byteps/mxnet/ops.cc:73-158).

The reference pushes an async function onto the MXNet engine that hands the NDArray's raw
pointer to the C core.  Here an NDArray reaches the SAME engine the torch front end uses,
zero-copy through DLPack when the array exports it (GPU arrays then ride the NVLink kernels),
through a numpy round trip otherwise.  The call is ordered against MXNet's engine with
``wait_to_read()`` before and a host-blocking synchronise after, so MXNet operators queued
later see the reduced values.
"""
from __future__ import annotations

import numpy as np
import torch

from ..torch import ops as _ops

init = _ops.init
shutdown = _ops.shutdown
suspend = _ops.suspend
resume = _ops.resume
size = _ops.size
rank = _ops.rank
local_size = _ops.local_size
local_rank = _ops.local_rank


def _torch_view(tensor):
    """(torch tensor aliasing `tensor`, write_back or None)."""
    to_dl = getattr(tensor, "to_dlpack_for_write", None)
    if to_dl is not None:
        try:
            return torch.utils.dlpack.from_dlpack(to_dl()), None
        except Exception:  # noqa: BLE001 - fall back to the copy path
            pass
    host = np.ascontiguousarray(tensor.asnumpy())
    t = torch.from_numpy(host)

    def write_back():
        tensor[:] = host
    return t, write_back


def byteps_push_pull(tensor, version=0, priority=0, name=None, is_average=True):
    """In-place push_pull of an NDArray: afterwards it holds the sum (or the average) over all
    workers.  ``priority`` orders transmissions (higher first), ``name`` must be identical on
    every worker for the same logical tensor."""
    if name is None:
        raise AssertionError("byteps_push_pull needs a name")
    wait = getattr(tensor, "wait_to_read", None)
    if wait is not None:
        wait()
    t, write_back = _torch_view(tensor)
    h = _ops.push_pull_async_inplace(t, average=is_average, name=name, version=version, priority=priority)
    _ops.synchronize(h)
    if t.is_cuda:
        torch.cuda.current_stream(t.device).synchronize()
    if write_back is not None:
        write_back()
    return tensor


def byteps_declare_tensor(name, **kwargs):
    """Declare a tensor (fixing its key) with optional ``byteps_*`` compressor attributes,
    e.g. ``byteps_compressor_type="topk", byteps_compressor_k=0.01`` (docs/gradient-compression.md)."""
    return _ops.declare(name, **{k: str(v) for k, v in kwargs.items()})


def set_learning_rate(lr):
    _ops.set_learning_rate(lr)
