"""GPU gradient compression with error feedback, exchanged over NVLink.

Same algorithms and kwargs as the CPU compressors in csrc/compress (and the
reference's /root/reference/byteps/common/compressor): onebit (+scaling), topk,
randomk, dithering (linear|natural x max|l2), vanilla error feedback with the
lr ratio, nesterov momentum.  Pipeline for one tensor, all on one stream:

    [momentum] -> corrected = g + (lr_prev/lr)*e -> compress (+ e update)
      -> payload in the symmetric arena
      -> ONE kernel: flag barrier, read every peer's payload over NVLink,
         decompress and sum in fp32, flag barrier
      -> ["server" stage: compress the sum again, with its own error state]
      -> decompress / cast into the user's tensor (x 1/size when averaging)

The second stage reproduces the reference's server-side compression
(server.cc:92-118) so results follow the same two-stage contract its tests
check; every rank computes it redundantly from the same data, bit-identically.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from ..comm.symm import SymmContext, wire_code


def _k_of(kw: Dict[str, str], numel: int) -> int:
    f = float(kw["compressor_k"])
    if f <= 0:
        raise ValueError("compressor_k must be positive")
    if f < 1:
        return max(1, int(f * numel))
    return min(int(f), int(numel))        # an absolute k larger than a small tensor keeps everything


class GpuCompressor:
    @staticmethod
    def payload_bytes_for(kwargs: Dict[str, str], numel: int) -> int:
        """Bytes of symmetric memory one tensor's payload window needs."""
        kw = {str(k): str(v) for k, v in kwargs.items()}
        kind, n = kw.get("compressor_type"), int(numel)
        if kind == "onebit":
            b = (n + 31) // 32 * 4 + 4
        elif kind == "topk":
            b = _k_of(kw, n) * 8
        elif kind == "randomk":
            b = _k_of(kw, n) * 4
        elif kind == "dithering":
            b = (n + 15) // 16 * 16 + 16
        else:
            raise ValueError("unknown compressor_type %r" % kind)
        return (b + 255) // 256 * 256

    def __init__(self, ctx: SymmContext, kwargs: Dict[str, str], numel: int, dtype: torch.dtype, payload_off: int = 0,
                 two_stage: bool = True):
        self.ctx, self.cu = ctx, ctx.cu
        self.kw = {str(k): str(v) for k, v in kwargs.items()}
        self.kind = self.kw.get("compressor_type")
        if self.kind not in ("onebit", "topk", "randomk", "dithering"):
            raise ValueError("unknown compressor_type %r" % self.kind)
        self.n, self.dtype, self.off = int(numel), dtype, int(payload_off)
        self.code = wire_code(dtype)
        dev = ctx.device
        self.use_ef = self.kw.get("ef_type") == "vanilla"
        self.mu = float(self.kw["momentum_mu"]) if self.kw.get("momentum_type") == "nesterov" else None
        self.two_stage = two_stage
        f32 = dict(dtype=torch.float32, device=dev)
        n = self.n
        self.corrected = torch.empty(n, **f32)
        self.err = torch.zeros(n, **f32) if self.use_ef else None
        self.err2 = torch.zeros(n, **f32) if (self.use_ef and two_stage) else None
        self.mom = torch.zeros(n, **f32) if self.mu is not None else None
        self.sum = torch.empty(n, **f32)
        self.acc = torch.zeros(4 + 3 * 148 * 8, **f32)   # results + per-block partials (kEfAccFloats)
        self.lr_prev = self.lr_cur = 1.0
        self.step = 0
        # CTAs of the cross-rank exchange kernels (each CTA owns a flag-barrier slot): scale with the work
        self.blocks = max(1, min(256, (n + 8191) // 8192))
        if self.kind == "onebit":
            self.scaled = self.kw.get("compressor_onebit_scaling", "false").lower() in ("1", "true", "yes")
            self.payload_bytes = (n + 31) // 32 * 4 + 4
            self.local2 = torch.empty((n + 31) // 32 + 1, dtype=torch.int32, device=dev)
        elif self.kind == "topk":
            self.k = _k_of(self.kw, n)
            self.payload_bytes = self.k * 8
            self.scratch = torch.zeros(1024, dtype=torch.int32, device=dev)
            self.local2 = torch.empty(2 * self.k, dtype=torch.int32, device=dev)
        elif self.kind == "randomk":
            self.k = _k_of(self.kw, n)
            self.payload_bytes = self.k * 4
            seed = int(self.kw.get("seed", "0")) or 0x9E3779B97F4A7C15
            # xorshift128+ is serial: the index streams (worker stage and "server" stage, both seeded like
            # the CPU compressors) are drawn on the host into pinned buffers and copied in on the stream
            from .. import _native

            core = _native.core()
            self.rng, self.rng2 = core.XorShift128Plus(), core.XorShift128Plus()
            self.rng.set_seed(seed & 0xFFFFFFFFFFFFFFFF)
            self.rng2.set_seed(seed & 0xFFFFFFFFFFFFFFFF)
            self.idx = torch.empty(self.k, dtype=torch.int32, device=dev)
            self.idx2 = torch.empty(self.k, dtype=torch.int32, device=dev)
            self._idx_host = [torch.empty(self.k, dtype=torch.int32).pin_memory() for _ in range(2)]
            self._idx_ev = [None, None]
            self.blocks = max(1, min(256, (self.k + 8191) // 8192))
            self.vals = torch.empty(self.k, **f32)
            self.vals2 = torch.empty(self.k, **f32)
        else:
            self.s = int(float(self.kw["compressor_k"]))
            self.partition = int(self.kw.get("dithering_partition", "0"))
            self.normalize = int(self.kw.get("dithering_normalize", "0"))
            self.seed = int(self.kw.get("seed", "0")) or 12345
            self.lv_bytes = (n + 15) // 16 * 16
            self.payload_bytes = self.lv_bytes + 16
            self.levels2 = torch.empty(self.lv_bytes, dtype=torch.int8, device=dev)
            self.scale2 = torch.empty(4, **f32)
        self.payload_bytes = (self.payload_bytes + 255) // 256 * 256
        if self.off + self.payload_bytes > ctx.data_bytes:
            raise ValueError("payload window does not fit in the arena")
        self.payload = ctx.arena[self.off:self.off + self.payload_bytes]

    def set_lr(self, lr: float):
        self.lr_cur = float(lr)

    def _draw_indices(self, rng, slot: int, dst: torch.Tensor):
        """k draws of randint(0, n) -> dst (device), through pinned slot `slot`."""
        ev = self._idx_ev[slot]
        if ev is not None:
            ev.synchronize()                     # the previous copy out of this pinned buffer is done
        host = self._idx_host[slot]
        rng.fill_randint_u32(host.data_ptr(), self.k, self.n)
        dst.copy_(host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.ctx.device))
        self._idx_ev[slot] = ev

    # ------------------------------------------------------------------
    def push_pull(self, grad: torch.Tensor, out: Optional[torch.Tensor] = None, average: bool = True, stream=None):
        """grad (may be modified by momentum) -> out (defaults to grad, in place)."""
        st = stream or torch.cuda.current_stream(self.ctx.device)
        with torch.cuda.stream(st):
            for phase in self.phases(grad, out, average, st.cuda_stream):
                phase()
        return grad if out is None else out

    def phases(self, grad, out, average, s):
        """The pipeline as a list of closures.  Odd entries are the cross-rank kernels (they
        spin on peer flags); a single-process multi-rank harness must issue each phase for ALL
        virtual ranks before the next one, so no spinning kernel sits in front of a peer's work
        in a shared hardware queue."""
        cu, ctx, n = self.cu, self.ctx, self.n
        out = grad if out is None else out
        assert grad.numel() == n and grad.dtype == self.dtype and grad.is_contiguous() and out.is_contiguous()
        mult = (1.0 / ctx.world) if average else 1.0
        ratio = (self.lr_prev / self.lr_cur) if self.lr_cur > 0 else 1.0
        self.lr_prev = self.lr_cur
        err = self.err.data_ptr() if self.err is not None else 0
        e2 = self.err2.data_ptr() if self.err2 is not None else 0
        pay = self.payload.data_ptr()
        kind = self.kind
        cor, acc, sm = self.corrected.data_ptr(), self.acc.data_ptr(), self.sum.data_ptr()

        def correct():
            if self.mom is not None:
                cu.nesterov(grad.data_ptr(), self.code, self.mom.data_ptr(), self.mu, n, s)
            cu.ef_correct(grad.data_ptr(), self.code, err, ratio, cor, n, acc, s)

        def server_correct():
            cu.ef_correct(sm, 0, e2, 1.0, cor, n, acc, s)

        if kind == "onebit":
            def pre():
                correct()
                cu.onebit_pack(cor, n, acc, self.scaled, pay, err, s)

            def xchg():
                cu.onebit_exchange_sum(ctx.view, self.off, n, sm, self.blocks, 0, s)

            def post():
                if self.two_stage:
                    server_correct()
                    cu.onebit_pack(cor, n, acc, self.scaled, self.local2.data_ptr(), e2, s)
                    cu.onebit_unpack(self.local2.data_ptr(), n, out.data_ptr(), self.code, mult, s)
                else:
                    cu.cast_scale(sm, n, out.data_ptr(), self.code, mult, s)
            return [pre, xchg, post]
        if kind == "topk":
            def pre():
                correct()
                cu.topk_select(cor, n, self.k, pay, err, self.scratch.data_ptr(), s)
                self.sum.zero_()

            def bar():
                cu.barrier(ctx.view, 1, 0, s)

            def adds():   # fixed peer order, indices unique inside a payload: bit-reproducible sum
                for p in range(ctx.world):
                    cu.sparse_add(ctx.view.data_ptr(p) + self.off, self.k, n, sm, s)

            def post():
                if self.two_stage:
                    server_correct()
                    cu.topk_select(cor, n, self.k, self.local2.data_ptr(), e2, self.scratch.data_ptr(), s)
                    cu.sparse_scatter(self.local2.data_ptr(), self.k, n, out.data_ptr(), self.code, mult, s)
                else:
                    cu.cast_scale(sm, n, out.data_ptr(), self.code, mult, s)
            return [pre, bar, adds, bar, post]
        if kind == "randomk":
            def pre():
                correct()
                self._draw_indices(self.rng, 0, self.idx)
                cu.randomk_gather(cor, self.idx.data_ptr(), self.k, n, pay, err, s)

            def xchg():
                cu.dense_exchange_sum(ctx.view, self.off, self.k, self.vals.data_ptr(), self.blocks, 0, s)

            def post():
                if self.two_stage:
                    # server: D(worker payloads) summed = scatter(idx, vals); then its own random-k draw
                    cu.index_scatter(self.idx.data_ptr(), self.vals.data_ptr(), self.k, n, sm, 0, 1.0, s)
                    server_correct()
                    self._draw_indices(self.rng2, 1, self.idx2)
                    cu.randomk_gather(cor, self.idx2.data_ptr(), self.k, n, self.vals2.data_ptr(), e2, s)
                    cu.index_scatter(self.idx2.data_ptr(), self.vals2.data_ptr(), self.k, n, out.data_ptr(),
                                     self.code, mult, s)
                else:
                    cu.index_scatter(self.idx.data_ptr(), self.vals.data_ptr(), self.k, n, out.data_ptr(), self.code,
                                     mult, s)
            return [pre, xchg, post]

        self.step += 1
        step = self.step

        def pre():
            correct()
            cu.dither_quantize(cor, n, acc, self.s, self.partition, self.normalize, self.seed, step, pay,
                               pay + self.lv_bytes, err, s)

        def xchg():
            cu.dither_exchange_sum(ctx.view, self.off, n, self.s, self.partition, sm, self.blocks, 0, s)

        def post():
            if self.two_stage:
                server_correct()
                cu.dither_quantize(cor, n, acc, self.s, self.partition, self.normalize, self.seed ^ 0x5555, step,
                                   self.levels2.data_ptr(), self.scale2.data_ptr(), e2, s)
                cu.dither_unpack(self.levels2.data_ptr(), self.scale2.data_ptr(), n, self.s, self.partition,
                                 out.data_ptr(), self.code, mult, s)
            else:
                cu.cast_scale(sm, n, out.data_ptr(), self.code, mult, s)
        return [pre, xchg, post]
