"""A numpy-backed stand-in for the slice of TensorFlow/tf.keras the byteps_b200.tensorflow and
byteps_b200.tensorflow.keras front ends touch (eager mode only).  TensorFlow is not installable
in this image; this exists so the front ends' logic runs in CI."""
import contextlib
import sys
import types

import numpy as np


class DType:
    def __init__(self, name):
        self.name = name
        self.np = np.dtype(name)
        self.is_floating = self.np.kind == "f"

    def __eq__(self, o):
        return isinstance(o, DType) and o.name == self.name

    def __hash__(self):
        return hash(self.name)

    def __repr__(self):
        return "tf." + self.name


_DT = {n: DType(n) for n in ("float16", "float32", "float64", "int32", "int64")}


def _dt(npdtype):
    return _DT[np.dtype(npdtype).name]


def _raw(x):
    if isinstance(x, (T, Variable)):
        return x.a
    if isinstance(x, IndexedSlices):
        dense = np.zeros(x.dense_shape, dtype=_raw(x.values).dtype)
        np.add.at(dense, _raw(x.indices), _raw(x.values))
        return dense
    return np.asarray(x)


class T:
    def __init__(self, a, name=None):
        self.a = np.array(a)
        if name:
            self.name = name

    dtype = property(lambda self: _dt(self.a.dtype))
    shape = property(lambda self: self.a.shape)

    def numpy(self):
        return self.a

    def set_shape(self, s):
        pass

    def __truediv__(self, o):
        return T(self.a / _raw(o))

    def __mul__(self, o):
        return T(self.a * _raw(o))

    __rmul__ = __mul__

    def __add__(self, o):
        return T(self.a + _raw(o))

    __radd__ = __add__

    def __sub__(self, o):
        return T(self.a - _raw(o))

    def __getitem__(self, k):
        return T(self.a[k])


class Variable(T):
    def __init__(self, a, name="var"):
        super().__init__(np.array(a, dtype=np.float32 if np.asarray(a).dtype.kind == "f" else None))
        self.name = name

    def assign(self, v):
        self.a[...] = _raw(v)
        return self

    def value(self):
        return T(self.a.copy(), name=self.name)


class IndexedSlices:
    def __init__(self, values, indices, dense_shape):
        self.values, self.indices, self.dense_shape = values, indices, tuple(dense_shape)
        self.dtype = _dt(_raw(values).dtype)


class GradientTape:
    """`preset` maps id(source) -> gradient; enough to exercise DistributedGradientTape."""

    def __init__(self, preset=None):
        self.preset = preset or {}

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def watch(self, x):
        self.watched = x

    def gradient(self, target, sources, output_gradients=None):
        if isinstance(sources, (list, tuple)):
            return [self.preset.get(id(s)) for s in sources]
        return self.preset.get(id(sources))


class Optimizer:
    def __init__(self, learning_rate=0.1, momentum=0.0, **kw):
        self.lr = types.SimpleNamespace(value=learning_rate)
        self.momentum = types.SimpleNamespace(value=momentum)

    def get_config(self):
        return {"learning_rate": self.lr.value, "momentum": self.momentum.value}

    @classmethod
    def from_config(cls, cfg):
        return cls(**cfg)

    def apply_gradients(self, grads_and_vars, **kw):
        for g, v in grads_and_vars:
            if g is not None:
                v.assign(v.a - self.lr.value * _raw(g))

    def variables(self):
        return []


class SGD(Optimizer):
    pass


class Callback:
    def __init__(self):
        self.model, self.params = None, {}

    def set_model(self, m):
        self.model = m

    def set_params(self, p):
        self.params = p


class _Distributed:
    """PerReplica / Mirrored: a tuple of per-device tensors."""

    def __init__(self, values):
        self.values = tuple(values)


class PerReplica(_Distributed):
    pass


class Mirrored(_Distributed):
    pass


class ReduceOp:
    class _Op:
        def __init__(self, name):
            self.name = name

    SUM, MEAN = _Op("SUM"), _Op("MEAN")


class CrossDeviceOps:
    def __init__(self):
        pass

    def reduce(self, reduce_op, per_replica_value, destinations, options=None):
        return self.reduce_implementation(reduce_op, per_replica_value, destinations, options)

    def batch_reduce(self, reduce_op, value_destination_pairs, options=None):
        return self.batch_reduce_implementation(reduce_op, value_destination_pairs, options)

    def _gather(self, per_replica_value, destinations, axis, options=None):
        return self._gather_implementation(per_replica_value, destinations, axis, options)


class ReductionToOneDevice(CrossDeviceOps):
    def reduce_implementation(self, reduce_op, per_replica_value, destinations, options=None):
        vals = per_replica_value.values
        tot = sum(_raw(v) for v in vals)
        if reduce_op is ReduceOp.MEAN:
            tot = tot / len(vals)
        return Mirrored([T(tot) for _ in vals])

    def broadcast_implementation(self, tensor, destinations):
        return Mirrored([T(_raw(tensor).copy())])

    def _gather_implementation(self, per_replica_value, destinations, axis, options=None):
        vals = per_replica_value.values
        cat = np.concatenate([_raw(v) for v in vals], axis=axis)
        return Mirrored([T(cat) for _ in vals])


class BaseMirroredStrategy:
    """One replica per entry of `devices`; `run` calls fn once per replica with that replica's args."""

    def __init__(self, devices=None, cross_device_ops=None):
        self.devices = list(devices or ["/cpu:0"])
        self.cross_device_ops = cross_device_ops or ReductionToOneDevice()
        self.num_replicas_in_sync = len(self.devices)

    def scope(self):
        return contextlib.nullcontext()

    def run(self, fn, args=()):
        outs = []
        for r in range(len(self.devices)):
            outs.append(fn(*[a.values[r] if isinstance(a, _Distributed) else a for a in args]))
        return PerReplica(outs)

    def reduce(self, reduce_op, value, axis=None):
        return self.cross_device_ops.reduce(reduce_op, value, None).values[0]

    def batch_reduce(self, reduce_op, values):
        return [m.values[0] for m in self.cross_device_ops.batch_reduce(reduce_op, [(v, None) for v in values])]

    def gather(self, value, axis):
        return self.cross_device_ops._gather(value, None, axis).values[0]


def install():
    tf = types.ModuleType("tensorflow")
    for n, d in _DT.items():
        setattr(tf, n, d)
    tf.Tensor, tf.Variable, tf.IndexedSlices, tf.GradientTape = T, Variable, IndexedSlices, GradientTape
    tf.constant = lambda v, dtype=None, name=None: T(np.asarray(v, dtype=dtype.np if dtype else None), name=name)
    tf.convert_to_tensor = lambda v, dtype=None: T(_raw(v).astype(dtype.np) if dtype else _raw(v))
    tf.cast = lambda v, dtype: T(_raw(v).astype(dtype.np))
    tf.identity = lambda v: T(_raw(v).copy())
    tf.executing_eagerly = lambda: True
    tf.device = lambda d: contextlib.nullcontext()
    tf.custom_gradient = lambda f: (lambda *a: f(*a)[0])
    tf.concat = lambda vs, axis=0: T(np.concatenate([_raw(v) for v in vs], axis=axis))
    tf.reshape = lambda v, shape: T(_raw(v).reshape(tuple(shape)))
    tf.zeros_like = lambda v: T(np.zeros_like(_raw(v)))
    tf.config = types.SimpleNamespace(list_logical_devices=lambda kind=None: [])
    tf.distribute = types.SimpleNamespace(CrossDeviceOps=CrossDeviceOps, ReductionToOneDevice=ReductionToOneDevice,
                                          MirroredStrategy=BaseMirroredStrategy, ReduceOp=ReduceOp,
                                          PerReplica=PerReplica, Mirrored=Mirrored)
    tf.experimental = types.SimpleNamespace()          # no dlpack: the numpy fallback is taken
    gv = []
    tf.compat = types.SimpleNamespace(v1=types.SimpleNamespace(global_variables=lambda: gv))
    tf._global_variables = gv
    keras = types.ModuleType("tensorflow.keras")
    backend = types.ModuleType("tensorflow.keras.backend")
    backend.get_value = lambda x: x.value
    backend.set_value = lambda x, v: setattr(x, "value", v)
    keras.backend = backend
    keras.callbacks = types.SimpleNamespace(Callback=Callback)
    keras.optimizers = types.SimpleNamespace(Optimizer=Optimizer, SGD=SGD)
    keras.models = types.SimpleNamespace(load_model=lambda path, custom_objects=None: custom_objects)
    tf.keras = keras
    sys.modules.update({"tensorflow": tf, "tensorflow.keras": keras, "tensorflow.keras.backend": backend})
    return tf
