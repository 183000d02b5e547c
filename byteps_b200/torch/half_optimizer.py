"""Half-precision training with fp32 master weights.

Parity: /root/reference/byteps/misc/imagenet18/__init__.py:39-382
(``DistributedOptimizer`` for fp16 models: fp16 params + fp32 master copy,
gradients pushed with priority = order of first appearance, per-parameter SGD
on the master as each handle completes, copy back to fp16, static/dynamic loss
scale).  On B200 this is the *default shape* of the fused path: the exchange
kernel reads bf16/fp16 gradients, accumulates in fp32, divides by
loss_scale*size, updates the fp32 master shard and writes the half-precision
weights to every rank - all in one kernel per bucket.
"""
from __future__ import annotations

import torch

from . import DistributedOptimizer


class HalfPrecisionDistributedOptimizer:
    def __init__(self, optimizer, named_parameters=None, loss_scale: float = 1.0, dynamic_loss_scale: bool = False,
                 scale_window: int = 1000, bucket_bytes=None):
        self._opt = DistributedOptimizer(optimizer, named_parameters=named_parameters, fused_update=True,
                                         bucket_bytes=bucket_bytes)
        self._sync = self._opt.grad_sync
        self.loss_scale = float(loss_scale)
        self.dynamic = dynamic_loss_scale
        self.scale_window = scale_window
        self._good_steps = 0
        if self._sync is not None:
            self._sync.loss_scale = self.loss_scale
            self._sync.refresh_hparams()

    def __getattr__(self, item):
        return getattr(self._opt, item)

    def scale_loss(self, loss):
        return loss * self.loss_scale

    def backward(self, loss):
        self.scale_loss(loss).backward()

    def zero_grad(self, set_to_none=False):
        self._opt.zero_grad()

    def _has_overflow(self) -> bool:
        bad = torch.zeros((), device=self._sync.device)
        for b in self._sync.buckets:
            bad += (~torch.isfinite(b.flat_grad.float())).any()
        return bool(bad.item())

    def step(self, closure=None):
        if self.dynamic and self._sync is not None:
            # inspect the local gradients before they are consumed by the exchange
            self._sync.synchronize_launch_guard = True
        loss = self._opt.step(closure)
        if self.dynamic and self._sync is not None:
            self._good_steps += 1
            if self._good_steps % self.scale_window == 0:
                self.loss_scale *= 2.0
                self._sync.loss_scale = self.loss_scale
                self._sync.refresh_hparams()
        return loss

    def master_params(self):
        return self._sync.master_params() if self._sync is not None else {}
