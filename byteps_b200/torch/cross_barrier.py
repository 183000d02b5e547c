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
cross-iteration overlap ByteScheduler is about.

When the fused kernels do not apply (CPU/gloo or CPU-server transports, or an
optimizer other than SGD/Adam/AdamW such as RMSprop) the reference's scheme is
kept in a generic form: every parameter gets its own single-parameter instance of
the user's optimizer class (so ANY torch optimizer works, not only the three the
reference re-implements), ``step()`` returns without waiting, and the forward
pre-hook of a module waits for exactly its parameters' handles, applies their
updates and clears their gradients.
"""
from __future__ import annotations

import torch

from . import DistributedOptimizer, _DistributedOptimizer
from .compression import Compression
from .ops import size, synchronize


class _CrossBarrier:
    def __init__(self, model, optimizer, num_steps=10 ** 6):
        self._model = model
        self._opt = optimizer
        self._step = 0
        self._final_step = num_steps
        self._sync = optimizer.grad_sync
        self._hooks = []
        self._pending = {}          # generic path: param -> (handle, ctx) still in flight
        self._per_param = {}
        self._generic = False
        self._zero_stream = None
        self._zeroed = {}           # bucket index -> event: its gradient window has been cleared for the next pass
        if self._sync is not None and self._sync.fused:
            self._bucket_of = dict(self._sync._param_bucket)
            self._register_forward_hooks()
            self._register_backward_hooks()
        elif self._sync is None and size() > 1 and hasattr(optimizer, "_handles"):
            self._generic = True
            base_cls = type(optimizer).__mro__[1]          # the user's optimizer class
            self._group_of = {}
            for g in optimizer.param_groups:
                hyper = {k: v for k, v in g.items() if k != "params"}
                for q in g["params"]:
                    if q.requires_grad:
                        self._per_param[q] = base_cls([q], **hyper)
                        self._group_of[q] = g
            self._register_generic_hooks()

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

    # ---- backward pre-hooks: a module's backward waits only for ITS buckets to have been cleared
    def _register_backward_hooks(self):
        for mod in self._model.modules():
            params = [p for p in mod.parameters(recurse=False) if p in self._bucket_of]
            if not params:
                continue
            buckets = sorted({self._bucket_of[p].index for p in params})

            def bwd_pre_hook(m, grad_output, buckets=buckets):
                cur = torch.cuda.current_stream()
                for bi in buckets:
                    ev = self._zeroed.get(bi)
                    if ev is not None:
                        cur.wait_event(ev)
            self._hooks.append(mod.register_full_backward_pre_hook(bwd_pre_hook))

    # ---- generic path: per-parameter completion + update -----------------------------------------
    def _finish_param(self, p):
        """Wait for p's exchange, apply p's update with its own optimizer instance, clear its gradient."""
        entry = self._pending.pop(p, None)
        if entry is None:
            return
        handle, ctx, hyper = entry
        opt = self._opt
        output = synchronize(handle)
        if ctx is not None:
            cctx, _ = ctx
            tmp = opt._compression.decompress(output, cctx)
            if tmp.data_ptr() != p.grad.data_ptr():
                if tmp.shape != p.grad.shape:
                    torch.as_strided(p.grad, (p.grad.numel(),), (1,), p.grad.storage_offset()).copy_(tmp)
                else:
                    p.grad.copy_(tmp)
        po = self._per_param[p]
        po.param_groups[0].update(hyper)     # the hyper-parameters in force when step() was called
        po.step()
        p.grad.zero_()

    def _register_generic_hooks(self):
        for mod in self._model.modules():
            params = [p for p in mod.parameters(recurse=False) if p in self._per_param]
            if not params:
                continue

            def pre_hook(m, inp, params=params):
                for q in params:
                    self._finish_param(q)
            self._hooks.append(mod.register_forward_pre_hook(pre_hook))

    def _drain(self):
        for q in list(self._pending):
            self._finish_param(q)

    def zero_grad(self, set_to_none=False):
        if self._generic:
            # gradients still being exchanged are cleared by _finish_param after their update
            for q in self._per_param:
                if q not in self._pending and q.grad is not None:
                    q.grad.zero_()
            return
        if self._sync is not None and self._sync.fused:
            # Gradients of a bucket may only be cleared after ITS exchange has finished - but the compute
            # stream must not wait for that here: `zero_grad(); forward; backward; step()` would put the global
            # barrier back (round 1 did).  Each bucket is cleared on a side stream behind its own completion
            # event, and only the backward of the modules that write into it waits for the clearing.
            dev = self._sync.device
            if self._zero_stream is None:
                self._zero_stream = torch.cuda.Stream(device=dev)
            zs = self._zero_stream
            zs.wait_stream(torch.cuda.current_stream(dev))       # nothing of the previous backward is still writing
            for b in self._sync.buckets:
                if b.done is not None:
                    zs.wait_event(b.done)
                with torch.cuda.stream(zs):
                    b.flat_grad.zero_()
                    ev = torch.cuda.Event()
                    ev.record(zs)
                self._zeroed[b.index] = ev
            return
        self._opt.zero_grad()

    def step(self, closure=None):
        """Step 0 behaves like a normal step (everything synchronised once so all
        ranks start aligned); afterwards the global barrier is gone: completion is
        awaited per module by the forward pre-hooks.  The last step drains."""
        self._step += 1
        sync = self._sync
        if self._generic:
            return self._generic_step(closure)
        if sync is None or not sync.fused:
            return self._opt.step(closure)
        loss = closure() if closure is not None else None
        if self._step == 1 or self._step >= self._final_step:
            sync.synchronize()
        else:
            sync.finish_launches()     # issue stragglers, do NOT wait
        sync.step_done()
        return loss

    def _generic_step(self, closure):
        opt = self._opt
        loss = closure() if closure is not None else None
        # parameters whose hook did not fire (unused in this iteration) are exchanged too, in name order
        missing = opt._requires_update - set(opt._handles.keys())
        for q in sorted(missing, key=lambda t: opt._parameter_names.get(t)):
            opt._handles[q] = opt._push_pull_grad_async(q)
        for q, (h, ctx) in list(opt._handles.items()):
            if h is None:
                opt._handles[q] = opt._push_pull_grad_async(q)
        # lr schedules act on the wrapped optimizer's groups; the delayed update uses the values of THIS step
        snap = {id(g): {k: v for k, v in g.items() if k != "params"} for g in opt.param_groups}
        for q, (h, ctx) in opt._handles.items():
            self._pending[q] = (h, ctx, snap[id(self._group_of[q])])
        opt._handles.clear()
        for q in self._pending:
            opt._push_pull_delay[q] = opt.backward_passes_per_step
        if self._step == 1 or self._step >= self._final_step:
            self._drain()          # first step: everyone aligned; last step: nothing left in flight
        return loss

    def synchronize(self):
        if self._generic:
            self._drain()
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


del _DistributedOptimizer
