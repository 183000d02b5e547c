"""MXNet front end: ``DistributedOptimizer``, gluon ``DistributedTrainer``, ``broadcast_parameters``
(parity: /root/reference/byteps/mxnet/__init__.py:35-360).

MXNet is not part of this image, so the module imports lazily against whatever ``mxnet``
package is importable; everything below the tensor bridge (``ops.py``: DLPack/numpy into the
shared engine) is plain python and is exercised by tests/test_plugins.py with a minimal stand-in.
Differences from the reference, on purpose:

* the learning rate reaches error-feedback compressors through ``set_learning_rate`` instead of
  the mmap'd ``lr.s`` file;
* ``compression_params`` become declare-time kwargs through the same translation every front
  end uses (common/compression_params.py).
"""
from __future__ import annotations

import copy
import os
import warnings

try:
    import mxnet as mx
except ImportError as e:  # pragma: no cover - exercised only without mxnet
    raise ImportError("byteps_b200.mxnet needs the `mxnet` package (not shipped in this image); "
                      "the torch front end `byteps_b200.torch` and the framework-neutral "
                      "`byteps_b200.dlpack` are always available") from e

from ..common.compression_params import translate as _translate
from .compression import Compression
from .ops import (byteps_declare_tensor, byteps_push_pull, init, local_rank, local_size, rank, resume,
                  set_learning_rate, shutdown, size, suspend)

parameter_index = 0


class DistributedOptimizer(mx.optimizer.Optimizer):
    """Wraps an MXNet optimizer: every ``update`` first sums-and-averages the gradient over all
    workers (priority = -index, so parameters the next forward needs first travel first).  In
    async mode (``BYTEPS_ENABLE_ASYNC=1``) the local update runs first and the weight DELTA is
    pushed; the pulled value is the server's current weight."""

    def __init__(self, optimizer):
        self._optimizer = optimizer
        self._enable_async = int(os.getenv("BYTEPS_ENABLE_ASYNC", 0)) != 0
        self._seen = set()

    def __getattr__(self, item):
        return getattr(self._optimizer, item)

    def create_state_multi_precision(self, index, weight):
        return self._optimizer.create_state_multi_precision(index, weight)

    def _declare_once(self, name):
        if name not in self._seen:
            byteps_declare_tensor(name)
            self._seen.add(name)

    def _do_push_pull(self, index, grad):
        if isinstance(index, (tuple, list)):
            for i in range(len(index)):
                self._declare_once("gradient_" + str(index[i]))
                byteps_push_pull(grad[i], version=0, priority=-index[i], name="gradient_" + str(index[i]),
                                 is_average=True)
        else:
            self._declare_once("gradient_" + str(index))
            byteps_push_pull(grad, version=0, priority=-index, name="gradient_" + str(index), is_average=True)

    def _do_push_pull_param(self, index, delta_weight):
        if isinstance(index, (tuple, list)):
            for i in range(len(index)):
                self._declare_once("weight_" + str(index[i]))
                byteps_push_pull(delta_weight[i], version=0, priority=-index[i], name="weight_" + str(index[i]),
                                 is_average=False)
        else:
            self._declare_once("weight_" + str(index))
            byteps_push_pull(delta_weight, version=0, priority=-index, name="weight_" + str(index),
                             is_average=False)

    def _update(self, fn, index, weight, grad, state):
        if self._enable_async:
            # local step, then exchange weight deltas through the server
            many = isinstance(index, (tuple, list))
            ws = list(weight) if many else [weight]
            before = [w.copy() for w in ws]
            fn(index, weight, grad, state)
            for w, b in zip(ws, before):
                w -= b                     # w now holds the delta
            self._do_push_pull_param(index, weight)
        else:
            self._do_push_pull(index, grad)
            fn(index, weight, grad, state)

    def update(self, index, weight, grad, state):
        self._update(self._optimizer.update, index, weight, grad, state)

    def update_multi_precision(self, index, weight, grad, state):
        self._update(self._optimizer.update_multi_precision, index, weight, grad, state)

    def set_learning_rate(self, lr):
        self._optimizer.set_learning_rate(lr)

    def set_lr_mult(self, args_lr_mult):
        self._optimizer.set_lr_mult(args_lr_mult)

    def set_wd_mult(self, args_wd_mult):
        self._optimizer.set_wd_mult(args_wd_mult)


def broadcast_parameters(params, root_rank=0):
    """Make ``params`` (dict name->NDArray, or a gluon ParameterDict) equal to root's copy:
    zero everywhere else, then a sum push_pull - declaration order is fixed by sorted name."""
    global parameter_index
    tensors = []
    if isinstance(params, dict):
        tensors = [p for _, p in sorted(params.items())]
    elif isinstance(params, mx.gluon.parameter.ParameterDict):
        for _, p in sorted(params.items()):
            try:
                tensors.append(p.data())
            except mx.gluon.parameter.DeferredInitializationError:
                # shapes unknown until the first forward: broadcast right after initialisation
                def _hook(init_impl, p=p):
                    def wrapped(self, *args, **kwargs):
                        init_impl(*args, **kwargs)
                        broadcast_parameters({self.name: self.data()}, root_rank=root_rank)
                    return wrapped
                p._init_impl = _hook(p._init_impl).__get__(p, type(p))
    else:
        raise ValueError("Invalid params of type: %s" % type(params))
    for t in tensors:
        name = "parameter_" + str(parameter_index)
        byteps_declare_tensor(name)
        if rank() != root_rank:
            t *= 0
        byteps_push_pull(t, version=0, priority=0, name=name, is_average=False)
        parameter_index += 1
    for t in tensors:
        wait = getattr(t, "wait_to_read", None)
        if wait is not None:
            wait()


class DistributedTrainer(mx.gluon.Trainer):
    """gluon Trainer whose gradient aggregation is a BytePS push_pull (sum, pre-scaled by
    1/(batch*workers)) instead of a kvstore, with optional gradient compression:

        trainer = DistributedTrainer(net.collect_params(), "sgd", {"learning_rate": .1, "momentum": .9},
                                     compression_params={"compressor": "onebit", "ef": "vanilla",
                                                         "momentum": "nesterov", "scaling": True})
    """

    def __init__(self, params, optimizer, optimizer_params=None, root_rank=0, compression_params=None):
        if isinstance(optimizer, DistributedOptimizer):
            optimizer = optimizer._optimizer
            warnings.warn("DistributedTrainer does not take DistributedOptimizer as its optimizer. "
                          "We have unwrapped it for you.")
        param_list = []
        if isinstance(params, mx.gluon.ParameterDict):
            for key in sorted(list(params.keys())):
                param_list.append(params[key])
        else:
            param_list = list(params)
        optimizer_params = dict(optimizer_params or {})
        self._compress_kwargs, self._intra_compressor = self._register_compressor(
            optimizer_params, compression_params)
        super(DistributedTrainer, self).__init__(param_list, optimizer, optimizer_params=optimizer_params,
                                                 kvstore=None)
        self._bps_size = size()
        self.root_rank = root_rank
        self._intra_compressors = {}
        for i, param in enumerate(self._params):
            byteps_declare_tensor("parameter_" + str(i))
            self._intra_compressors[param.name] = copy.deepcopy(self._intra_compressor)
            if param.grad_req != "null":
                byteps_declare_tensor("gradient_" + str(i), **self._compress_kwargs)

    @staticmethod
    def _register_compressor(optimizer_params, compression_params):
        """(declare kwargs, intra-node compressor).  When the compressor chain takes over the
        momentum (and, for 1-bit, the weight decay) those are REMOVED from ``optimizer_params``
        so they are not applied twice."""
        intra = Compression.none
        if not compression_params:
            return {}, intra
        if compression_params.get("fp16"):
            intra = Compression.fp16
        if "compressor" not in compression_params:
            warnings.warn("Compressor is not defined")
            return {}, intra
        kwargs = {"byteps_" + k: v for k, v in _translate(compression_params, optimizer_params).items()}
        if compression_params.get("momentum"):
            threshold = int(os.environ.get("BYTEPS_MIN_COMPRESS_BYTES", 65536))
            mu = optimizer_params["momentum"]
            if compression_params["compressor"] == "onebit" and "wd" in optimizer_params:
                intra = Compression.wdmom(intra, mu, optimizer_params.pop("wd"), threshold)
            intra = Compression.nag(intra, mu, threshold)
            del optimizer_params["momentum"]
        return kwargs, intra

    def step(self, batch_size, ignore_stale_grad=False):
        # gradients are normalised by batch_size in _allreduce_grads; stop Trainer.step doing it again
        self._scale = batch_size
        super(DistributedTrainer, self).step(batch_size, ignore_stale_grad)

    def _allreduce_grads(self):
        set_learning_rate(self.learning_rate)
        for i, param in enumerate(self._params):
            if param.grad_req == "null":
                continue
            g = param._grad[0]
            g *= 1.0 / self._scale / self._bps_size
            comp = self._intra_compressors[param.name]
            compressed, ctx = comp.compress(g)
            byteps_push_pull(compressed, is_average=False, name="gradient_" + str(i), priority=-i)
            param._grad[0][:] = comp.decompress(compressed, ctx, x=param._data[0])

    def _init_params(self):
        later = []
        for param in self._params_to_init:
            if param._deferred_init:
                later.append(param)
                continue
            arrays = param._check_and_get(param._data, list)
            idx = self._param2idx[param.name]
            if rank() != self.root_rank:
                arrays[0] *= 0
            byteps_push_pull(arrays[0], version=0, priority=0, name="parameter_" + str(idx), is_average=False)
        self._params_to_init = later


__all__ = ["init", "shutdown", "suspend", "resume", "size", "rank", "local_size", "local_rank",
           "byteps_push_pull", "byteps_declare_tensor", "DistributedOptimizer", "DistributedTrainer",
           "broadcast_parameters", "Compression", "set_learning_rate"]
