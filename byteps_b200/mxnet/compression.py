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


class NagAdapter(Compressor):
    """Nesterov momentum applied explicitly to gradients that are NOT compressed (smaller than
    ``threshold`` elements); compressed ones get it inside the compressor chain."""

    def __init__(self, compressor, mu, threshold, *args, **kwargs):
        self.compressor, self.mu, self.threshold = compressor, mu, threshold
        self.mom = None
        self.inited = False
        self.nag = False

    def compress(self, tensor, *args, **kwargs):
        return self.compressor.compress(tensor)

    def decompress(self, tensor, ctx, *args, **kwargs):
        tensor = self.compressor.decompress(tensor, ctx, *args, **kwargs)
        if not self.inited:
            if size(tensor.shape) < self.threshold:
                self.mom = _zeros_like(tensor)
                self.nag = True
            self.inited = True
        if self.nag:
            self.mom += tensor
            self.mom *= self.mu
            tensor += self.mom
        return tensor


class WeightDecayMomentumAdapter(Compressor):
    """1-bit compression keeps weight decay out of the compressed signal:
    ``m = mu*(m + wd*x)``; ``g += m + wd*x`` (momentum only for compressed-size tensors)."""

    def __init__(self, compressor, mu, wd, threshold, *args, **kwargs):
        self.compressor, self.mu, self.wd, self.threshold = compressor, mu, wd, threshold
        self.mom = None
        self.inited = False
        self.wdmom = False

    def compress(self, tensor, *args, **kwargs):
        return self.compressor.compress(tensor)

    def decompress(self, tensor, ctx, *args, **kwargs):
        if "x" not in kwargs:
            raise ValueError("x is missing")
        x = kwargs["x"].astype(tensor.dtype, copy=False)
        if not self.inited:
            if size(tensor.shape) >= self.threshold:
                self.mom = _zeros_like(tensor)
                self.wdmom = True
            self.inited = True
        decay = x * self.wd
        if self.wdmom:
            self.mom += decay
            self.mom *= self.mu
            tensor += self.mom
        tensor += decay
        return self.compressor.decompress(tensor, ctx, *args, **kwargs)


class Compression(object):
    none = NoneCompressor()
    fp16 = FP16Compressor()
    wdmom = WeightDecayMomentumAdapter
    nag = NagAdapter
