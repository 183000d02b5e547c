"""standalone Keras callbacks (parity: /root/reference/byteps/keras/callbacks.py)."""
import keras
from keras import backend as K

from byteps_b200._keras import callbacks as _impl


class BroadcastGlobalVariablesCallback(_impl.BroadcastGlobalVariablesCallbackImpl, keras.callbacks.Callback):
    """Broadcast all variables from root_rank after the first batch, so all workers continue
    from identical weights and optimizer state (random init or restored checkpoint)."""

    def __init__(self, root_rank, device=""):
        super().__init__(K, root_rank, device)


class MetricAverageCallback(_impl.MetricAverageCallbackImpl, keras.callbacks.Callback):
    """Average the epoch-end metrics over workers; list it before ReduceLROnPlateau, TensorBoard
    and other metric-driven callbacks."""

    def __init__(self, device=""):
        super().__init__(K, device)


class LearningRateScheduleCallback(_impl.LearningRateScheduleCallbackImpl, keras.callbacks.Callback):
    """lr = initial_lr * multiplier(epoch) between start_epoch and end_epoch."""

    def __init__(self, multiplier, start_epoch=0, end_epoch=None, staircase=True, momentum_correction=True,
                 steps_per_epoch=None):
        super().__init__(K, multiplier, start_epoch, end_epoch, staircase, momentum_correction, steps_per_epoch)


class LearningRateWarmupCallback(_impl.LearningRateWarmupCallbackImpl, keras.callbacks.Callback):
    """Gradual warm-up from lr/size to lr over the first epochs (large-minibatch SGD)."""

    def __init__(self, warmup_epochs=5, momentum_correction=True, steps_per_epoch=None, verbose=0):
        super().__init__(K, warmup_epochs, momentum_correction, steps_per_epoch, verbose)
