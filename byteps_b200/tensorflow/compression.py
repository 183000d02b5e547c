"""Wire formats for the TensorFlow front end: ``Compression.none``, ``Compression.fp16`` and (new)
``Compression.bf16`` - API parity with /root/reference/byteps/tensorflow/compression.py, built from one
cast-compressor factory."""
import tensorflow as tf


def _is_float(dtype):
    return bool(getattr(dtype, "is_floating", False))


class Compressor(object):
    """compress() -> (payload, context); decompress(payload, context) restores the original dtype."""

    @staticmethod
    def compress(tensor):
        raise NotImplementedError

    @staticmethod
    def decompress(tensor, ctx):
        raise NotImplementedError


class NoneCompressor(Compressor):
    compress = staticmethod(lambda tensor: (tensor, None))
    decompress = staticmethod(lambda tensor, ctx: tensor)


def _cast_compressor(name, wire_name, doc):
    """A Compressor class shipping every floating-point tensor as tf.<wire_name>."""

    def compress(tensor):
        wire = getattr(tf, wire_name)
        original = tensor.dtype
        return (tf.cast(tensor, wire) if _is_float(original) and original != wire else tensor), original

    def decompress(tensor, ctx):
        return tf.cast(tensor, ctx) if ctx is not None and _is_float(ctx) and tensor.dtype != ctx else tensor

    return type(name, (Compressor,), {"__doc__": doc, "compress": staticmethod(compress),
                                      "decompress": staticmethod(decompress)})


FP16Compressor = _cast_compressor("FP16Compressor", "float16", "Half precision on the wire.")
BF16Compressor = _cast_compressor("BF16Compressor", "bfloat16", "bfloat16 on the wire (no loss scaling needed).")


class Compression(object):
    none = NoneCompressor
    fp16 = FP16Compressor
    bf16 = BF16Compressor
