"""Wire-format selection for gradients (the reference calls it intra-box "compression").

API parity with /root/reference/byteps/torch/compression.py: ``Compression.none`` /
``Compression.fp16`` objects exposing ``compress(tensor) -> (tensor, ctx)`` and
``decompress(tensor, ctx)``; ``Compression.bf16`` is new.  On the symmetric-memory transport the
down-cast is the pack phase of the fused push-pull kernel (BucketedGradSync ``wire_dtype``), so these
objects mostly act as tags there; the explicit casts below serve the gloo / NCCL / CPU-server
transports.
"""
import torch


class Compressor(object):
    """compress() returns (payload, context); decompress(payload, context) undoes it."""

    @staticmethod
    def compress(tensor):
        raise NotImplementedError

    @staticmethod
    def decompress(tensor, ctx):
        raise NotImplementedError


def _cast_compressor(name: str, wire: torch.dtype, doc: str):
    """Build a Compressor class that ships floating-point tensors as `wire` and restores the dtype."""

    def compress(tensor):
        if tensor.dtype.is_floating_point and tensor.dtype != wire:
            return tensor.to(wire), tensor.dtype
        return tensor, tensor.dtype

    def decompress(tensor, ctx):
        if ctx is not None and ctx.is_floating_point and tensor.dtype != ctx:
            return tensor.to(ctx)
        return tensor

    return type(name, (Compressor,), {"wire": wire, "__doc__": doc, "compress": staticmethod(compress),
                                      "decompress": staticmethod(decompress)})


class NoneCompressor(Compressor):
    """Identity."""

    @staticmethod
    def compress(tensor):
        return tensor, None

    @staticmethod
    def decompress(tensor, ctx):
        return tensor


FP16Compressor = _cast_compressor("FP16Compressor", torch.float16, "IEEE half precision on the wire.")
BF16Compressor = _cast_compressor("BF16Compressor", torch.bfloat16, "bfloat16 on the wire (fp32 range, 8-bit mantissa).")


class Compression(object):
    """Namespace of the available wire formats: ``none`` (default), ``fp16``, ``bf16``."""
    none = NoneCompressor
    fp16 = FP16Compressor
    bf16 = BF16Compressor
