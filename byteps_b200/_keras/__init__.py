"""Shared implementation behind ``byteps_b200.keras`` (standalone Keras) and
``byteps_b200.tensorflow.keras`` (tf.keras) - parity: /root/reference/byteps/_keras/__init__.py:20-121.
Everything is parameterised by the keras module / backend handed in by the thin wrappers."""
from __future__ import annotations

import byteps_b200.tensorflow as bps


def create_distributed_optimizer(keras, optimizer, name, device_dense, device_sparse, compression,
                                 sparse_as_dense):
    """Subclass the optimizer's own class so Keras (de)serialises it under the same name; gradient
    aggregation is intercepted at every entry point Keras generations have used
    (``get_gradients``, ``_aggregate_gradients``, ``_compute_gradients``)."""
    base = optimizer.__class__
    scope = (name or "Distributed%s" % base.__name__) + "."

    def reduce_list(grads, tag):
        if bps.size() <= 1:
            return grads
        out = []
        for i, g in enumerate(grads):
            if g is None:
                out.append(None)
                continue
            if sparse_as_dense and isinstance(g, bps.tf.IndexedSlices):
                g = bps.tf.convert_to_tensor(g)
            out.append(bps.push_pull(g, scope, device_dense=device_dense, device_sparse=device_sparse,
                                     compression=compression, name="%s_%d" % (tag, i)))
        return out

    class _DistributedOptimizer(base):
        _bps_aggregated = False

        def get_gradients(self, loss, params):                  # Keras 2.x
            return reduce_list(super().get_gradients(loss, params), "keras_grad")

        def _aggregate_gradients(self, grads_and_vars):         # TF 2.2-2.10
            gv = list(grads_and_vars)
            self._bps_aggregated = True
            return reduce_list([g for g, _ in gv], "keras_agg")

        def _compute_gradients(self, *args, **kwargs):          # OptimizerV2.minimize
            gv = list(super()._compute_gradients(*args, **kwargs))
            if self._bps_aggregated:
                return gv
            self._bps_aggregated = True
            return list(zip(reduce_list([g for g, _ in gv], "keras_cg"), [v for _, v in gv]))

        def apply_gradients(self, grads_and_vars, *args, **kwargs):
            gv = list(grads_and_vars)
            if not self._bps_aggregated:
                gv = list(zip(reduce_list([g for g, _ in gv], "keras_ag"), [v for _, v in gv]))
            self._bps_aggregated = False
            return super().apply_gradients(gv, *args, **kwargs)

    cls = _DistributedOptimizer
    cls.__name__ = cls.__qualname__ = base.__name__      # saved models record the optimizer's own name
    if hasattr(optimizer, "get_config") and hasattr(cls, "from_config"):
        try:
            return cls.from_config(optimizer.get_config())
        except Exception:  # noqa: BLE001 - fall through to state adoption
            pass
    obj = cls.__new__(cls)
    obj.__dict__.update(optimizer.__dict__)
    return obj


def _eval(backend, op_or_result):
    if bps.tf.executing_eagerly():
        return op_or_result
    return backend.get_session().run(op_or_result)


def broadcast_global_variables(backend, root_rank):
    return _eval(backend, bps.broadcast_global_variables(root_rank))


def push_pull(backend, value, name, average):
    return _eval(backend, bps.push_pull(bps.tf.constant(value, name=name), average=average, name=name))


def broadcast(backend, value, root_rank, name):
    return _eval(backend, bps.broadcast(bps.tf.constant(value, name=name), root_rank, name=name,
                                        is_variable=False))


def load_model(keras, wrap_optimizer, optimizer_modules, filepath, custom_optimizers, custom_objects):
    """``keras.models.load_model`` with every known optimizer class replaced by its distributed
    wrapper, so a restored model keeps averaging gradients."""
    from inspect import isclass

    bps_objects = {}
    for module in optimizer_modules:
        for name in dir(module):
            sub = getattr(module, name)
            if isclass(sub) and issubclass(sub, keras.optimizers.Optimizer) and sub is not keras.optimizers.Optimizer:
                bps_objects[sub.__name__] = wrap_optimizer(sub)
    for opt in custom_optimizers or []:
        bps_objects[opt.__name__] = wrap_optimizer(opt)
    if custom_objects is not None:
        bps_objects.update(custom_objects)
    return keras.models.load_model(filepath, custom_objects=bps_objects)
