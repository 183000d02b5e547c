"""Intra-node gradient compression for the TensorFlow front end
(/root/reference/byteps/tensorflow/compression.py:21-75): none or fp16 on the wire."""
import tensorflow as tf


class Compressor(object):
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


class FP16Compressor(Compressor):
    @staticmethod
    def compress(tensor):
        dtype = tensor.dtype
        if getattr(dtype, "is_floating", False) and dtype != tf.float16:
            return tf.cast(tensor, tf.float16), dtype
        return tensor, dtype

    @staticmethod
    def decompress(tensor, ctx):
        if ctx is not None and getattr(ctx, "is_floating", False) and tensor.dtype != ctx:
            return tf.cast(tensor, ctx)
        return tensor


class Compression(object):
    none = NoneCompressor
    fp16 = FP16Compressor
