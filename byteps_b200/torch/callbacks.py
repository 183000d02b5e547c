"""Training-loop callbacks: the torch flavour of the reference's Keras callbacks
(/root/reference/byteps/_keras/callbacks.py:23-196): broadcast initial state,
average metrics over workers, scheduled / warm-up learning rates scaled by the
number of workers.  They are plain objects with ``on_train_begin``,
``on_epoch_begin``, ``on_batch_begin``, ``on_batch_end``, ``on_epoch_end``
hooks so any loop (or trainer framework) can drive them."""
from __future__ import annotations

import torch

from . import broadcast_optimizer_state, broadcast_parameters
from .ops import push_pull, size


class Callback:
    def on_train_begin(self, logs=None): ...
    def on_epoch_begin(self, epoch, logs=None): ...
    def on_batch_begin(self, batch, logs=None): ...
    def on_batch_end(self, batch, logs=None): ...
    def on_epoch_end(self, epoch, logs=None): ...


class BroadcastGlobalVariablesCallback(Callback):
    """Broadcast model (and optimizer) state from root_rank at the start of training, so
    all workers start identical (random init or restored checkpoint)."""

    def __init__(self, model, optimizer=None, root_rank=0):
        self.model, self.optimizer, self.root_rank = model, optimizer, root_rank
        self.broadcast_done = False

    def on_batch_end(self, batch, logs=None):   # like the reference: after the first batch created all state
        if self.broadcast_done:
            return
        broadcast_parameters(self.model.state_dict(), self.root_rank)
        if self.optimizer is not None:
            broadcast_optimizer_state(self.optimizer, self.root_rank)
        self.broadcast_done = True

    on_train_begin = lambda self, logs=None: self.on_batch_end(-1, logs)  # noqa: E731


class MetricAverageCallback(Callback):
    """Average the metrics in `logs` over all workers at the end of every epoch."""

    def on_epoch_end(self, epoch, logs=None):
        if logs is None:
            return
        for k in sorted(logs.keys()):
            v = logs[k]
            if isinstance(v, (int, float)) or (torch.is_tensor(v) and v.numel() == 1):
                t = torch.tensor(float(v))
                logs[k] = push_pull(t, average=True, name="metric.%s" % k).item()


class LearningRateScheduleCallback(Callback):
    """lr = initial_lr * multiplier(epoch) inside [start_epoch, end_epoch); `staircase`
    adjusts per epoch, otherwise per batch (fractional epochs)."""

    def __init__(self, optimizer, multiplier, start_epoch=0, end_epoch=None, staircase=True, steps_per_epoch=None,
                 momentum_correction=True):
        self.optimizer = optimizer
        self.start_epoch, self.end_epoch, self.staircase = start_epoch, end_epoch, staircase
        self.steps_per_epoch = steps_per_epoch
        self.momentum_correction = momentum_correction
        self.initial_lr = [g["lr"] for g in optimizer.param_groups]
        self.multiplier = multiplier if callable(multiplier) else (lambda epoch: multiplier)
        self.current_epoch = 0

    def _adjust(self, epoch):
        for g, lr0 in zip(self.optimizer.param_groups, self.initial_lr):
            old = g["lr"]
            g["lr"] = lr0 * self.multiplier(epoch)
            if self.momentum_correction and "momentum" in g and old > 0 and g["momentum"]:
                g["_momentum_correction"] = g["lr"] / old   # informational: SGD in torch is lr-inside-momentum

    def _in_range(self, epoch):
        return epoch >= self.start_epoch and (self.end_epoch is None or epoch < self.end_epoch)

    def on_epoch_begin(self, epoch, logs=None):
        self.current_epoch = epoch
        if self.staircase and self._in_range(epoch):
            self._adjust(epoch)

    def on_batch_begin(self, batch, logs=None):
        if self.staircase or not self.steps_per_epoch:
            return
        epoch = self.current_epoch + float(batch) / self.steps_per_epoch
        if self._in_range(epoch):
            self._adjust(epoch)


class LearningRateWarmupCallback(LearningRateScheduleCallback):
    """Gradual warm-up from lr/size to lr over `warmup_epochs` (Goyal et al.), the usual
    companion of scaling the learning rate by the number of workers."""

    def __init__(self, optimizer, warmup_epochs=5, steps_per_epoch=None, momentum_correction=True, verbose=0):
        def multiplier(epoch):
            epoch += 1.0 / max(self.steps_per_epoch or 1, 1)
            return 1.0 / size() * (epoch * (size() - 1) / warmup_epochs + 1)
        super().__init__(optimizer, multiplier, start_epoch=0, end_epoch=warmup_epochs, staircase=False,
                         steps_per_epoch=steps_per_epoch, momentum_correction=momentum_correction)
        self.verbose = verbose
