#!/usr/bin/env python
"""Throughput of the CPU compressors (csrc/compress) on one 4 MB partition - what a worker spends in its COMPRESS /
DECOMPRESS stages and a server per push in CPU-server mode.  Prints GB/s of gradient bytes processed.

    python benchmarks/cpu_compress_bench.py [--bytes 4096000] [--iters 30]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from byteps_b200 import _native  # noqa: E402

CONFIGS = [
    ("onebit+scaling", {"compressor_type": "onebit", "compressor_onebit_scaling": "true"}),
    ("onebit+scaling+ef", {"compressor_type": "onebit", "compressor_onebit_scaling": "true", "ef_type": "vanilla"}),
    ("topk 1%", {"compressor_type": "topk", "compressor_k": "0.01"}),
    ("topk 1%+ef+nesterov", {"compressor_type": "topk", "compressor_k": "0.01", "ef_type": "vanilla",
                            "momentum_type": "nesterov", "momentum_mu": "0.9"}),
    ("randomk 1%", {"compressor_type": "randomk", "compressor_k": "0.01", "seed": "7"}),
    ("dithering s=4 linear/max", {"compressor_type": "dithering", "compressor_k": "4", "seed": "7"}),
    ("dithering s=4 natural/l2", {"compressor_type": "dithering", "compressor_k": "4", "seed": "7",
                                 "dithering_partition": "1", "dithering_normalize": "1"}),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bytes", type=int, default=4096000)
    ap.add_argument("--iters", type=int, default=30)
    args = ap.parse_args()
    c = _native.core()
    n = args.bytes // 4
    g0 = np.random.RandomState(0).randn(n).astype(np.float32)
    red = c.CpuReducer(1)
    print("%-28s %10s %11s %14s %14s %10s" % ("compressor", "comp GB/s", "decomp GB/s", "decomp+sum GB/s", "decomp_add GB/s",
                                              "payload B"))
    for name, kw in CONFIGS:
        comp = c.Compressor(kw, n * 4, c.F32)
        buf = np.zeros(comp.max_compressed_bytes() + 64, dtype=np.uint8)
        out = np.zeros(n, dtype=np.float32)
        g = g0.copy()
        m = comp.compress(g.ctypes.data, buf.ctypes.data)
        comp.decompress(buf.ctypes.data, m, out.ctypes.data)
        tc = td = ts = ta = 0.0
        acc = np.zeros(n, dtype=np.float32)
        for _ in range(args.iters):
            g[:] = g0
            t0 = time.perf_counter()
            m = comp.compress(g.ctypes.data, buf.ctypes.data)
            t1 = time.perf_counter()
            comp.decompress(buf.ctypes.data, m, out.ctypes.data)
            t2 = time.perf_counter()
            red.sum(acc.ctypes.data, out.ctypes.data, n * 4, c.F32)      # what a server did per push: scratch + sum
            t3 = time.perf_counter()
            comp.decompress_add(buf.ctypes.data, m, acc.ctypes.data)     # what it does now
            t4 = time.perf_counter()
            tc += t1 - t0
            td += t2 - t1
            ts += t3 - t1
            ta += t4 - t3
        gbs = n * 4 * args.iters / 1e9
        print("%-28s %10.2f %11.2f %14.2f %14.2f %10d" % (name, gbs / tc, gbs / td, gbs / ts, gbs / ta, m))


if __name__ == "__main__":
    main()
