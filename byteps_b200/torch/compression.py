"""Intra-box gradient "compression" by down-casting on the wire.

Parity: /root/reference/byteps/torch/compression.py:21-75 (Compressor /
NoneCompressor / FP16Compressor / Compression).  bf16 is added.  When the
symmetric-memory transport is active the cast is not a separate ATen kernel: it
is the pack/unpack phase of the fused push-pull kernel (``wire_dtype``), so
``compress`` only tags the request; the explicit-cast path below is kept for
the gloo/NCCL/PS transports and for API compatibility.
"""
import torch


class Compressor(object):
    """Interface for compressing and decompressing a given tensor."""

    @staticmethod
    def compress(tensor):
        """Compresses a tensor and returns it with the context needed to decompress it."""
        raise NotImplementedError

    @staticmethod
    def decompress(tensor, ctx):
        """Decompress the tensor with the given context."""
        raise NotImplementedError


class NoneCompressor(Compressor):
    """Default no-op compression."""

    @staticmethod
    def compress(tensor):
        return tensor, None

    @staticmethod
    def decompress(tensor, ctx):
        return tensor


class _CastCompressor(Compressor):
    wire = torch.float16

    @classmethod
    def compress(cls, tensor):
        tensor_compressed = tensor
        if tensor.dtype.is_floating_point and tensor.dtype != cls.wire:
            tensor_compressed = tensor.to(cls.wire)
        return tensor_compressed, tensor.dtype

    @classmethod
    def decompress(cls, tensor, ctx):
        tensor_decompressed = tensor
        dtype = ctx
        if dtype is not None and dtype.is_floating_point and tensor.dtype != dtype:
            tensor_decompressed = tensor.to(dtype)
        return tensor_decompressed


class FP16Compressor(_CastCompressor):
    """Compress all floating point gradients to 16-bit (IEEE half)."""
    wire = torch.float16


class BF16Compressor(_CastCompressor):
    """Compress all floating point gradients to bfloat16 (new; the reference has no bf16)."""
    wire = torch.bfloat16


class Compression(object):
    """Optional gradient compression algorithm used during push_pull."""

    """Do not compress the gradients. This is the default."""
    none = NoneCompressor

    """Compress all floating point gradients to 16-bit."""
    fp16 = FP16Compressor

    """Compress all floating point gradients to bfloat16."""
    bf16 = BF16Compressor
