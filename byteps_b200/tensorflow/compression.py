"""Wire formats for the TensorFlow front end: ``Compression.none`` and ``Compression.fp16``
(API parity with /root/reference/byteps/tensorflow/compression.py)."""
import tensorflow as tf


class Compressor(object):
    """compress() -> (payload, context); decompress(payload, context) restores the original dtype."""

    @staticmethod
    def compress(tensor):
        raise NotImplementedError

    @staticmethod
    def decompress(tensor, ctx):
        raise NotImplementedError


class NoneCompressor(Compressor):
    @staticmethod
    def compress(tensor):
        return tensor, None

    @staticmethod
    def decompress(tensor, ctx):
        return tensor


def _is_float(dtype):
    return bool(getattr(dtype, "is_floating", False))


class FP16Compressor(Compressor):
    """Half precision on the wire for every floating-point tensor."""

    @staticmethod
    def compress(tensor):
        original = tensor.dtype
        payload = tf.cast(tensor, tf.float16) if _is_float(original) and original != tf.float16 else tensor
        return payload, original

    @staticmethod
    def decompress(tensor, ctx):
        return tf.cast(tensor, ctx) if ctx is not None and _is_float(ctx) and tensor.dtype != ctx else tensor


class Compression(object):
    none = NoneCompressor
    fp16 = FP16Compressor
