"""push_pull primitives for PyTorch tensors.

Parity: /root/reference/byteps/torch/ops.py:38-236 and the native adapter behind
it (/root/reference/byteps/torch/ops.cc:54-206): same functions, same argument
order, same name-keyed semantics, handle-based completion.  Differences:

* completion is a CUDA event, so ``synchronize`` makes the *current stream*
  wait (no host polling with 1 ms sleeps, ops.cc:129-135); pass
  ``BYTEPS_SYNC_HOST=1`` to also block the host,
* bf16 tensors are accepted,
* with one process the out-of-place variant returns a real copy (the reference
  returns an uninitialised buffer).
"""
import os

import torch

from ..common import BytePSBasics as _BytePSBasics
from ..common import engine as _engine
from ..common import remember_declared as _remember
from .compression import Compression

_basics = _BytePSBasics()

init = _basics.init
shutdown = _basics.shutdown
suspend = _basics.suspend
resume = _basics.resume
size = _basics.size
local_size = _basics.local_size
rank = _basics.rank
local_rank = _basics.local_rank
get_pushpull_speed = _basics.get_pushpull_speed

_SYNC_HOST = os.environ.get("BYTEPS_SYNC_HOST", "0") not in ("0", "", "false")

# handle -> (input, output): keeps tensors alive until the operation finished
_handle_map = {}
_num_grads = 0
_grad_count = 0


def _check(tensor):
    if not tensor.is_contiguous():
        raise ValueError('Tensor is required to be contiguous.')


def _do_push_pull_async(tensor, output, average, name, version=0, priority=0):
    _check(tensor)
    eng = _engine()
    if name is not None:
        _remember("byteps." + name)
    handle = eng.push_pull_async(tensor, output, average, name, version, priority)
    _handle_map[handle] = (tensor, output)
    return handle


def _do_push_pull_group_sync(tensor, output, average, name, version=0, priority=0):
    """DDP helper: like the async call, but also counts gradients so the caller
    knows when the last one of the iteration has been issued
    (/root/reference/byteps/torch/ops.cc:137-166)."""
    global _grad_count
    handle = _do_push_pull_async(tensor, output, average, name, version, priority)
    _grad_count += 1
    curr = _grad_count
    if _num_grads and _grad_count >= _num_grads:
        _grad_count = 0
    return handle, curr


def push_pull_async(tensor, average=True, name=None, version=0, priority=0):
    """Asynchronously average (or sum) ``tensor`` over all processes; the input
    is not modified.  Returns a handle for ``poll()`` / ``synchronize()``."""
    output = tensor.new(tensor.shape)
    return _do_push_pull_async(tensor, output, average, name, version, priority)


class BytePSPushPull(torch.autograd.Function):
    """An autograd function that performs push_pull on a tensor."""

    @staticmethod
    def forward(ctx, tensor, average, name, version, priority):
        ctx.average = average
        ctx.name = name
        ctx.version = version
        ctx.priority = priority
        handle = push_pull_async(tensor, average, name, version, priority)
        return synchronize(handle)

    @staticmethod
    def backward(ctx, grad_output):
        return push_pull(grad_output, ctx.average, ctx.name, ctx.version, ctx.priority), None, None, None, None


def push_pull(tensor, average=True, name=None, version=0, priority=0, compression=Compression.none):
    """Average (or sum) ``tensor`` over all processes.  Differentiable: the
    backward pass push_pulls the incoming gradient under the same name."""
    if name is None:
        raise AssertionError("To manually call push_pull, you must specify a name by name=...")
    tensor_compressed, ctx = compression.compress(tensor)
    summed_tensor_compressed = BytePSPushPull.apply(tensor_compressed, average, name, version, priority)
    return compression.decompress(summed_tensor_compressed, ctx)


def push_pull_async_inplace(tensor, average=True, name=None, version=0, priority=0):
    """Asynchronous in-place push_pull."""
    return _do_push_pull_async(tensor, tensor, average, name, version, priority)


def push_pull_group_sync_inplace(tensor, average=True, name=None, version=0, priority=0):
    return _do_push_pull_group_sync(tensor, tensor, average, name, version, priority)


def push_pull_inplace(tensor, average=True, name=None, version=0, priority=0):
    """In-place push_pull; returns the tensor once it holds the result."""
    handle = push_pull_async_inplace(tensor, average, name, version, priority)
    return synchronize(handle)


def poll(handle):
    """True once the operation behind ``handle`` has completed (then
    ``synchronize`` returns without waiting)."""
    return _engine().poll(handle)


def declare(name, **kwargs):
    """Declare a tensor name (fixes its key order).  Optional kwargs configure gradient
    compression for that tensor: ``compressor_type`` (onebit|topk|randomk|dithering),
    ``compressor_k``, ``compressor_onebit_scaling``, ``ef_type`` (vanilla), ``momentum_type``
    (nesterov), ``momentum_mu``, ``seed``, ``dithering_partition``, ``dithering_normalize``.
    A leading ``byteps_`` on a key is accepted (the attribute spelling of the MXNet trainer)."""
    _remember("byteps." + name)
    _engine().declare("byteps." + name)
    if kwargs:
        kw = {(k[7:] if k.startswith("byteps_") else k): v for k, v in kwargs.items()}
        _engine().set_compression(name, kw)
    return 0


def set_learning_rate(lr):
    """Tell error-feedback compressors the current learning rate."""
    _engine().set_learning_rate(lr)


def byteps_torch_set_num_grads(num_grads_):
    global _num_grads, _grad_count
    _num_grads = int(num_grads_)
    _grad_count = 0
    return 0


def synchronize(handle):
    """Wait for an asynchronous push_pull and return its output tensor."""
    if handle not in _handle_map:
        return
    _engine().synchronize(handle, block_host=_SYNC_HOST)
    _, output = _handle_map.pop(handle)
    return output
