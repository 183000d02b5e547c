"""Intra-node gradient "compressors" of the MXNet front end
(/root/reference/byteps/mxnet/compression.py:26-164): fp16 cast, and the two adapters the
trainer stacks on top when lossy inter-node compression takes the momentum out of the
optimizer - Nesterov momentum for tensors too small to be compressed, and the separate
weight-decay momentum used with 1-bit compression."""
from __future__ import annotations


def size(shape):
    n = 1
    for s in shape:
        n *= int(s)
    return n


def _zeros_like(x):
    try:
        import mxnet as mx

        return mx.nd.zeros_like(x)
    except Exception:  # noqa: BLE001
        import numpy as np

        return np.zeros_like(x)


def _is_wide_float(dtype):
    name = str(dtype)
    return "float" in name and "16" not in name


class Compressor(object):
    """compress(tensor) -> (payload, context); decompress(payload, context, **kw) -> tensor."""

    def compress(self, tensor, *args, **kwargs):
        raise NotImplementedError

    def decompress(self, tensor, ctx, *args, **kwargs):
        raise NotImplementedError


class NoneCompressor(Compressor):
    compress = lambda self, tensor, *a, **k: (tensor, None)        # noqa: E731
    decompress = lambda self, tensor, ctx, *a, **k: tensor          # noqa: E731


class _CastCompressor(Compressor):
    """fp32/fp64 gradients travel as `wire` and come back in their own dtype."""
    wire = "float16"

    def compress(self, tensor, *args, **kwargs):
        original = tensor.dtype
        return (tensor.astype(self.wire, copy=False) if _is_wide_float(original) else tensor), original

    def decompress(self, tensor, ctx, *args, **kwargs):
        wide = ctx is not None and _is_wide_float(ctx) and str(tensor.dtype) != str(ctx)
        return tensor.astype(ctx, copy=False) if wide else tensor


class FP16Compressor(_CastCompressor):
    wire = "float16"


class _MomentumOnTop(Compressor):
    """Shared machinery of the two adapters: wrap an inner compressor and keep ONE momentum buffer
    ``m <- mu * (m + delta)`` for the tensor, but only if the tensor's size is on the side of ``threshold`` the
    adapter cares about - decided the first time a tensor is seen (sizes never change)."""

    def __init__(self, compressor, mu, threshold, keep_if_small):
        self.compressor, self.mu, self.threshold = compressor, mu, threshold
        self._keep_if_small = keep_if_small
        self._active = None          # unknown until the first tensor
        self.mom = None

    def compress(self, tensor, *args, **kwargs):
        return self.compressor.compress(tensor)

    def _momentum(self, like, delta):
        """Advance the buffer by `delta` and return it, or None when this tensor carries no momentum here."""
        if self._active is None:
            small = size(like.shape) < self.threshold
            self._active = small if self._keep_if_small else not small
            if self._active:
                self.mom = _zeros_like(like)
        if not self._active:
            return None
        self.mom += delta
        self.mom *= self.mu
        return self.mom


class NagAdapter(_MomentumOnTop):
    """Nesterov momentum applied explicitly to gradients that are NOT compressed (fewer than ``threshold``
    elements); the compressed ones get it inside the compressor chain (momentum_type=nesterov)."""

    def __init__(self, compressor, mu, threshold, *args, **kwargs):
        super().__init__(compressor, mu, threshold, keep_if_small=True)

    def decompress(self, tensor, ctx, *args, **kwargs):
        tensor = self.compressor.decompress(tensor, ctx, *args, **kwargs)
        m = self._momentum(tensor, tensor)
        if m is not None:
            tensor += m
        return tensor


class WeightDecayMomentumAdapter(_MomentumOnTop):
    """1-bit compression keeps weight decay out of the compressed signal: with ``d = wd * x`` the gradient
    becomes ``g + mu*(m + d) + d`` for compressed-size tensors and ``g + d`` for the small ones."""

    def __init__(self, compressor, mu, wd, threshold, *args, **kwargs):
        super().__init__(compressor, mu, threshold, keep_if_small=False)
        self.wd = wd

    def decompress(self, tensor, ctx, *args, **kwargs):
        if "x" not in kwargs:
            raise ValueError("x is missing")
        decay = kwargs["x"].astype(tensor.dtype, copy=False) * self.wd
        m = self._momentum(tensor, decay)
        if m is not None:
            tensor += m
        tensor += decay
        return self.compressor.decompress(tensor, ctx, *args, **kwargs)


class Compression(object):
    none = NoneCompressor()
    fp16 = FP16Compressor()
    wdmom = WeightDecayMomentumAdapter
    nag = NagAdapter
