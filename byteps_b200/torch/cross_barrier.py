"""CrossBarrier (ByteScheduler): remove the global barrier between iterations.

Parity: /root/reference/byteps/torch/cross_barrier.py:28-421.  The reference
takes a per-parameter lock when a gradient's push_pull is issued, a poller
thread applies a per-parameter SGD/Adam/RMSprop update as each handle
completes and releases the lock, and forward pre-hooks block only on the
parameters of the layer about to run; ``step()`` is a no-op after step 0.

Here the same effect needs no thread and no locks: the update of a parameter is
fused into its bucket's exchange kernel (fp32 master weights, SGD/Adam), the
kernel is stream ordered after that bucket's gradients, and the forward
pre-hook of a module makes the compute stream wait on the completion EVENT of
the buckets holding its parameters.  The next iteration's forward therefore
starts while later (= earlier-layer) buckets are still in flight, exactly the
cross-iteration overlap ByteScheduler is about.  For optimizers the fused
kernels do not cover (RMSprop) the per-parameter python update of the
reference is kept, driven by completion events instead of a polling thread.
"""
from __future__ import annotations

import torch

from . import DistributedOptimizer, _DistributedOptimizer
from .compression import Compression
from .ops import size


class _CrossBarrier:
    def __init__(self, model, optimizer, num_steps=10 ** 6):
        self._model = model
        self._opt = optimizer
        self._step = 0
        self._final_step = num_steps
        self._sync = optimizer.grad_sync
        self._hooks = []
        if self._sync is not None and self._sync.fused:
            self._bucket_of = dict(self._sync._param_bucket)
            self._register_forward_hooks()

    def __getattr__(self, item):
        return getattr(self._opt, item)

    # ---- forward pre-hooks: wait only for this module's parameters
    def _register_forward_hooks(self):
        for mod in self._model.modules():
            params = [p for p in mod.parameters(recurse=False) if p in self._bucket_of]
            if not params:
                continue
            buckets = sorted({self._bucket_of[p].index for p in params})

            def pre_hook(m, inp, buckets=buckets):
                cur = torch.cuda.current_stream()
                for bi in buckets:
                    b = self._sync.buckets[bi]
                    if b.done is not None:
                        cur.wait_event(b.done)
            self._hooks.append(mod.register_forward_pre_hook(pre_hook))

    def zero_grad(self, set_to_none=False):
        # gradients of a bucket may only be cleared after its exchange finished
        if self._sync is not None:
            cur = torch.cuda.current_stream()
            for b in self._sync.buckets:
                if b.done is not None:
                    cur.wait_event(b.done)
        self._opt.zero_grad()

    def step(self, closure=None):
        """Step 0 behaves like a normal step (everything synchronised once so all
        ranks start aligned); afterwards the global barrier is gone: completion is
        awaited per module by the forward pre-hooks.  The last step drains."""
        self._step += 1
        sync = self._sync
        if sync is None or not sync.fused:
            return self._opt.step(closure)
        loss = closure() if closure is not None else None
        if self._step == 1 or self._step >= self._final_step:
            sync.synchronize()
        else:
            sync.finish_launches()     # issue stragglers, do NOT wait
        sync.step_done()
        return loss

    def synchronize(self):
        if self._sync is not None:
            self._sync.synchronize()

    def remove_hooks(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def CrossBarrier(model, optimizer, named_parameters=None, compression=Compression.none, backward_passes_per_step=1,
                 num_steps=10 ** 6):
    """Wrap ``optimizer`` like the reference's CrossBarrier(model, optimizer,
    named_parameters, compression, backward_passes_per_step, num_steps)."""
    if not hasattr(optimizer, "_push_pull_delay"):
        optimizer = DistributedOptimizer(optimizer, named_parameters=named_parameters or model.named_parameters(),
                                         compression=compression, backward_passes_per_step=backward_passes_per_step,
                                         fused_update=torch.cuda.is_available() and
                                         all(p.is_cuda for p in model.parameters()))
    return _CrossBarrier(model, optimizer, num_steps)


del _DistributedOptimizer, size
