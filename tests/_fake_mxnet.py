"""A numpy-backed stand-in for the slice of MXNet the byteps_b200.mxnet front end touches
(NDArray, optimizer.Optimizer/SGD, gluon.Parameter/ParameterDict/Trainer).  Only for tests:
MXNet itself is not installable in this image."""
import sys
import types

import numpy as np


class NDArray:
    def __init__(self, a):
        self.a = np.array(a)

    shape = property(lambda self: self.a.shape)
    dtype = property(lambda self: self.a.dtype)

    def asnumpy(self):
        return self.a.copy()

    def copy(self):
        return NDArray(self.a.copy())

    def astype(self, dtype, copy=True):
        if not copy and np.dtype(dtype) == self.a.dtype:
            return self
        return NDArray(self.a.astype(dtype))

    def wait_to_read(self):
        return None

    def __setitem__(self, k, v):
        self.a[k] = v.a if isinstance(v, NDArray) else v

    def __getitem__(self, k):
        return NDArray(self.a[k])

    def _v(self, o):
        return o.a if isinstance(o, NDArray) else o

    def __iadd__(self, o):
        self.a += self._v(o)
        return self

    def __isub__(self, o):
        self.a -= self._v(o)
        return self

    def __imul__(self, o):
        self.a *= self._v(o)
        return self

    def __mul__(self, o):
        return NDArray(self.a * self._v(o))

    __rmul__ = __mul__

    def __add__(self, o):
        return NDArray(self.a + self._v(o))

    def __sub__(self, o):
        return NDArray(self.a - self._v(o))


class Optimizer:
    def __init__(self, learning_rate=0.01, wd=0.0, rescale_grad=1.0, **kw):
        self.lr, self.wd, self.rescale_grad = learning_rate, wd, rescale_grad

    def set_learning_rate(self, lr):
        self.lr = lr

    def set_lr_mult(self, m):
        self.lr_mult = m

    def set_wd_mult(self, m):
        self.wd_mult = m

    def create_state(self, index, weight):
        return None

    def create_state_multi_precision(self, index, weight):
        return self.create_state(index, weight)

    def update_multi_precision(self, index, weight, grad, state):
        self.update(index, weight, grad, state)


class SGD(Optimizer):
    def __init__(self, momentum=0.0, **kw):
        super().__init__(**kw)
        self.momentum = momentum

    def create_state(self, index, weight):
        return NDArray(np.zeros_like(weight.a)) if self.momentum else None

    def update(self, index, weight, grad, state):
        g = grad.a * self.rescale_grad + self.wd * weight.a
        if state is not None:
            state.a[...] = self.momentum * state.a - self.lr * g
            weight.a += state.a
        else:
            weight.a -= self.lr * g


def _create(name, **kw):
    return {"sgd": SGD}[name](**kw)


class DeferredInitializationError(Exception):
    pass


class Parameter:
    def __init__(self, name, value, grad_req="write"):
        self.name, self.grad_req = name, grad_req
        self._data = [NDArray(value)]
        self._grad = [NDArray(np.zeros_like(np.asarray(value)))]
        self._deferred_init = ()

    def data(self):
        return self._data[0]

    def _check_and_get(self, arr_list, ctx):
        return arr_list


class ParameterDict(dict):
    pass


class Trainer:
    """The control flow of gluon.Trainer.step that DistributedTrainer hooks into."""

    def __init__(self, params, optimizer, optimizer_params=None, kvstore="device"):
        self._params = list(params)
        self._param2idx = {p.name: i for i, p in enumerate(self._params)}
        self._params_to_init = list(self._params)
        self._optimizer = _create(optimizer, **(optimizer_params or {})) if isinstance(optimizer, str) else optimizer
        self._states = [self._optimizer.create_state(i, p.data()) for i, p in enumerate(self._params)]
        self._scale = 1.0

    @property
    def learning_rate(self):
        return self._optimizer.lr

    def set_learning_rate(self, lr):
        self._optimizer.set_learning_rate(lr)

    def step(self, batch_size, ignore_stale_grad=False):
        rescale = self._scale / batch_size
        self._optimizer.rescale_grad = rescale
        if self._params_to_init:
            self._init_params()
        self._allreduce_grads()
        for i, p in enumerate(self._params):
            if p.grad_req != "null":
                self._optimizer.update(i, p._data[0], p._grad[0], self._states[i])


def install():
    mx = types.ModuleType("mxnet")
    nd = types.ModuleType("mxnet.nd")
    nd.NDArray, nd.array = NDArray, NDArray
    nd.zeros_like = lambda x: NDArray(np.zeros_like(x.a))
    opt = types.ModuleType("mxnet.optimizer")
    opt.Optimizer, opt.SGD, opt.create = Optimizer, SGD, _create
    gluon = types.ModuleType("mxnet.gluon")
    parameter = types.ModuleType("mxnet.gluon.parameter")
    parameter.ParameterDict, parameter.Parameter = ParameterDict, Parameter
    parameter.DeferredInitializationError = DeferredInitializationError
    gluon.parameter, gluon.ParameterDict, gluon.Parameter, gluon.Trainer = parameter, ParameterDict, Parameter, Trainer
    mx.nd, mx.optimizer, mx.gluon = nd, opt, gluon
    for name, mod in (("mxnet", mx), ("mxnet.nd", nd), ("mxnet.optimizer", opt), ("mxnet.gluon", gluon),
                      ("mxnet.gluon.parameter", parameter)):
        sys.modules[name] = mod
    return mx
