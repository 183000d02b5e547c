"""Callback logic shared by the Keras flavours (parity:
/root/reference/byteps/_keras/callbacks.py:23-196).  The classes here are mixed into
``keras.callbacks.Callback`` by the wrappers; ``backend`` is ``keras.backend``."""
from __future__ import annotations

import warnings

import byteps_b200.tensorflow as bps


class BroadcastGlobalVariablesCallbackImpl(object):
    def __init__(self, backend, root_rank, device="", *args):
        super().__init__(*args)
        self.backend, self.root_rank, self.device = backend, root_rank, device
        self.broadcast_done = False

    def on_batch_end(self, batch, logs=None):
        """After the FIRST batch: by then the optimizer has created its slot variables too."""
        if self.broadcast_done:
            return
        with bps.tf.device(self.device):
            if bps.tf.executing_eagerly() and hasattr(self.model, "variables"):
                bps.broadcast_variables(self.model.variables, root_rank=self.root_rank)
                opt = getattr(self.model, "optimizer", None)
                opt_vars = opt.variables() if callable(getattr(opt, "variables", None)) else getattr(opt, "variables", [])
                bps.broadcast_variables(opt_vars or [], root_rank=self.root_rank)
            else:
                self.backend.get_session().run(bps.broadcast_global_variables(self.root_rank))
        self.broadcast_done = True


class MetricAverageCallbackImpl(object):
    def __init__(self, backend, device="", *args):
        super().__init__(*args)
        self.backend, self.device = backend, device

    def _average_metrics_in_place(self, logs):
        logs = logs or {}
        # sorted: every worker reduces the metrics in the same order
        for metric, value in sorted(logs.items()):
            if isinstance(value, (int, float)) or getattr(value, "shape", None) == ():
                t = bps.tf.constant(float(value), dtype=bps.tf.float32)
                red = bps.push_pull(t, average=True, name="metric_" + str(metric))
                logs[metric] = float(red.numpy()) if hasattr(red, "numpy") else float(self.backend.get_value(red))

    def on_epoch_end(self, epoch, logs=None):
        self._average_metrics_in_place(logs)


class LearningRateScheduleCallbackImpl(object):
    """lr = initial_lr * multiplier(epoch) for start_epoch <= epoch < end_epoch; with
    ``staircase=False`` the multiplier is evaluated at fractional epochs every batch.  Momentum is
    corrected by lr_new/lr_old for the step where the rate changes (Goyal et al., 2017)."""

    def __init__(self, backend, multiplier, start_epoch=0, end_epoch=None, staircase=True,
                 momentum_correction=True, steps_per_epoch=None, *args):
        super().__init__(*args)
        self.backend = backend
        self.start_epoch, self.end_epoch, self.staircase = start_epoch, end_epoch, staircase
        self.momentum_correction, self.steps_per_epoch = momentum_correction, steps_per_epoch
        self.initial_lr = None
        self.restore_momentum = None
        self.current_epoch = None
        self.multiplier = multiplier if callable(multiplier) else (lambda epoch: multiplier)
        if not callable(multiplier):
            self.staircase = True

    def _autodetect_steps_per_epoch(self):
        p = getattr(self, "params", {}) or {}
        if p.get("steps"):
            return p["steps"]
        if p.get("samples") and p.get("batch_size"):
            return p["samples"] // p["batch_size"]
        raise ValueError("Could not autodetect the number of steps per epoch. Please specify the "
                         "steps_per_epoch parameter to the %s()." % self.__class__.__name__)

    def _adjust_learning_rate(self, epoch):
        old_lr = self.backend.get_value(self.model.optimizer.lr)
        new_lr = self.initial_lr * self.multiplier(epoch)
        self.backend.set_value(self.model.optimizer.lr, new_lr)
        if hasattr(self.model.optimizer, "momentum") and self.momentum_correction and old_lr > 0:
            self.restore_momentum = self.backend.get_value(self.model.optimizer.momentum)
            self.backend.set_value(self.model.optimizer.momentum, self.restore_momentum * new_lr / old_lr)

    def _restore_momentum_if_needed(self):
        if self.restore_momentum is not None:
            self.backend.set_value(self.model.optimizer.momentum, self.restore_momentum)
            self.restore_momentum = None

    def on_train_begin(self, logs=None):
        self.initial_lr = self.backend.get_value(self.model.optimizer.lr)
        if not self.staircase and not self.steps_per_epoch:
            self.steps_per_epoch = self._autodetect_steps_per_epoch()

    def on_epoch_begin(self, epoch, logs=None):
        self.current_epoch = epoch

    def on_batch_begin(self, batch, logs=None):
        e = self.current_epoch
        if e < self.start_epoch or (self.end_epoch is not None and e >= self.end_epoch):
            return
        if self.staircase and batch == 0:
            self._adjust_learning_rate(e)
        elif not self.staircase:
            self._adjust_learning_rate(e + float(batch) / self.steps_per_epoch)

    def on_batch_end(self, batch, logs=None):
        self._restore_momentum_if_needed()

    def on_epoch_end(self, epoch, logs=None):
        if logs is not None:
            logs["lr"] = self.backend.get_value(self.model.optimizer.lr)


class LearningRateWarmupCallbackImpl(LearningRateScheduleCallbackImpl):
    """Ramp from lr/size to lr over ``warmup_epochs`` (the user sets lr already scaled by size)."""

    def __init__(self, backend, warmup_epochs=5, momentum_correction=True, steps_per_epoch=None, verbose=0, *args):
        def multiplier(epoch):
            epoch += 1.0 / self.steps_per_epoch      # so the last warm-up batch lands exactly on lr
            return 1.0 / bps.size() * (epoch * (bps.size() - 1) / warmup_epochs + 1)
        super().__init__(backend, multiplier, start_epoch=0, end_epoch=warmup_epochs, staircase=False,
                         momentum_correction=momentum_correction, steps_per_epoch=steps_per_epoch, *args)
        self.verbose = verbose

    def on_epoch_end(self, epoch, logs=None):
        super().on_epoch_end(epoch, logs)
        if epoch == self.end_epoch - 1 and self.verbose > 0:
            print("\nEpoch %d: finished gradual learning rate warmup to %g." % (
                epoch + 1, self.backend.get_value(self.model.optimizer.lr)))
        if epoch == self.end_epoch - 1 and bps.size() > 1 and self.initial_lr is None:
            warnings.warn("warm-up finished before on_train_begin was called")
