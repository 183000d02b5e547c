"""TensorFlow tensor bridge (parity: /root/reference/byteps/tensorflow/ops.py:102-205 and the
AsyncOpKernel in /root/reference/byteps/tensorflow/ops.cc:136-231).

The reference compiles a custom op against TF's headers.  Here a tf.Tensor reaches the shared
engine zero-copy through DLPack (``tf.experimental.dlpack``), so GPU tensors use the NVLink
kernels; CPU tensors and TF builds without DLPack go through numpy.  Inside ``tf.function``
graphs the exchange is a ``tf.py_function`` node.  Gradients are registered with
``tf.custom_gradient``: d(push_pull)/dx is a push_pull of the upstream gradient, d(broadcast)
is a push_pull that only the root keeps.
"""
from __future__ import annotations

import re
import warnings
from enum import Enum

import tensorflow as tf
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
get_pushpull_speed = _ops.get_pushpull_speed


class ReduceOps(Enum):
    """Reduction named by the `op` argument of push_pull / DistributedOptimizer
    (/root/reference/byteps/tensorflow/ops.py:92-97).  Averaging is done by the framework-side code on top of a
    sum; Adasum is part of the vocabulary the reference inherited from Horovod and is rejected where it is used."""
    Average = "Average"
    Sum = "Sum"
    Adasum = "Adasum"


def handle_average_backwards_compatibility(op, average):
    """`average=` is the deprecated spelling of `op=`: old call sites keep their behaviour, mixing both is an
    error, neither means Average."""
    if op is not None:
        if average is not None:
            raise ValueError("The op parameter supersedes average. Please provide only one of them.")
        return op
    if average is not None:
        warnings.warn("Parameter `average` has been replaced with `op` and will be removed", DeprecationWarning)
        return ReduceOps.Average if average else ReduceOps.Sum
    return ReduceOps.Average


def _normalize_name(name):
    """Normalizes operation name to TensorFlow rules."""
    return re.sub("[^a-zA-Z0-9_]", "_", name)


def _stable_name(tensor, scope, name, kind):
    if name is None:
        tname = getattr(tensor, "name", None)
        if not isinstance(tname, str) or not tname:
            # eager tensors are anonymous: dtype+shape is the only identity that is stable across steps
            tname = "anon_%s_%s" % (getattr(tensor.dtype, "name", tensor.dtype),
                                    "x".join(str(int(d)) for d in tensor.shape))
        name = "%s_%s" % (kind, _normalize_name(tname))
    return (scope or "") + name


def _exchange(value, full_name, root_rank=None):
    """Sum `value` (anything with .numpy(), or a DLPack exporter) over all workers; returns a torch
    tensor on the same device.  With root_rank set, non-root contributions are zeroed (broadcast)."""
    t = None
    try:
        t = torch.utils.dlpack.from_dlpack(tf.experimental.dlpack.to_dlpack(value))
    except Exception:  # noqa: BLE001 - no DLPack in this TF build / unsupported dtype
        t = torch.from_numpy(value.numpy().copy())
    src = t.contiguous()
    if root_rank is not None and rank() != root_rank:
        src = torch.zeros_like(src)
    h = _ops.push_pull_async(src, average=False, name=full_name)
    out = _ops.synchronize(h)
    if out.is_cuda:
        torch.cuda.current_stream(out.device).synchronize()
    return out


def _to_tf(out, like):
    try:
        return tf.experimental.dlpack.from_dlpack(torch.utils.dlpack.to_dlpack(out))
    except Exception:  # noqa: BLE001
        return tf.convert_to_tensor(out.cpu().numpy(), dtype=like.dtype)


def _run(tensor, full_name, root_rank=None):
    if tf.executing_eagerly():
        return _to_tf(_exchange(tensor, full_name, root_rank), tensor)
    out = tf.py_function(lambda x: _to_tf(_exchange(x, full_name, root_rank), x), [tensor], tensor.dtype)
    out.set_shape(tensor.shape)
    return out


def _push_pull(tensor, scope="", name=None):
    """Sum of `tensor` over all workers (differentiable)."""
    full = _stable_name(tensor, scope, name, "BytePSPushPull")
    _ops.declare(full)

    @tf.custom_gradient
    def op(x):
        def grad(dy):
            return _push_pull(dy, scope, name=full + "_grad")
        return _run(x, full), grad
    return op(tensor)


def broadcast(tensor, root_rank, scope="", name=None, is_variable=True):
    """Root's value on every worker (differentiable: the gradient is summed onto the root)."""
    full = _stable_name(tensor, scope, name, "BytePSBroadcast")
    _ops.declare(full)

    @tf.custom_gradient
    def op(x):
        def grad(dy):
            g = _push_pull(dy, scope, name=full + "_grad")
            return g if rank() == root_rank else g * 0
        return _run(x, full, root_rank=root_rank), grad
    value = tensor.value() if (is_variable and hasattr(tensor, "value")) else tensor
    return op(value)
