"""tf.keras front end (parity: /root/reference/byteps/tensorflow/keras/__init__.py): DistributedOptimizer,
broadcast_global_variables, push_pull, broadcast, load_model - thin bindings of
byteps_b200._keras to this Keras flavour."""
from tensorflow import keras
from tensorflow.keras import backend as K

from byteps_b200 import _keras as _impl
from byteps_b200.tensorflow import Compression, init, local_rank, local_size, rank, shutdown, size  # noqa: F401

from . import callbacks  # noqa: F401,E402


def DistributedOptimizer(optimizer, name=None, device_dense="", device_sparse="", compression=Compression.none,
                         sparse_as_dense=False):
    """Wrap a Keras optimizer so gradients are averaged over all workers before they are applied."""
    return _impl.create_distributed_optimizer(keras, optimizer, name, device_dense, device_sparse, compression,
                                              sparse_as_dense)


def broadcast_global_variables(root_rank):
    return _impl.broadcast_global_variables(K, root_rank)


def push_pull(value, name=None, average=True):
    """Average (or sum) a python/numpy value over all workers and return the result."""
    return _impl.push_pull(K, value, name, average)


def broadcast(value, root_rank, name=None):
    return _impl.broadcast(K, value, root_rank, name)


def load_model(filepath, custom_optimizers=None, custom_objects=None, compression=Compression.none):
    """Load a saved model with its optimizer re-wrapped in DistributedOptimizer."""
    def wrap_optimizer(cls):
        return lambda **kwargs: DistributedOptimizer(cls(**kwargs), compression=compression)
    return _impl.load_model(keras, wrap_optimizer, [keras.optimizers], filepath, custom_optimizers, custom_objects)
