#!/usr/bin/env python
"""Host reducer bandwidth (the server's summation kernel): dst += src over 64 MB buffers for the dtypes
the server sees, at 1 and N OpenMP threads.  GB/s counts the three streams (read dst, read src, write dst)."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=64)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--threads", default="1,4,8")
    args = ap.parse_args()
    from byteps_b200 import _native

    c = _native.core()
    print("| dtype | threads | ms | GB/s (3 streams) | torch add_ GB/s |")
    print("|---|---|---|---|---|")
    for code, dt in (("F32", torch.float32), ("BF16", torch.bfloat16), ("F16", torch.float16), ("F64", torch.float64)):
        n = args.mb * (1 << 20) // torch.empty((), dtype=dt).element_size()
        a, b = torch.randn(n).to(dt), torch.randn(n).to(dt)
        nbytes = n * a.element_size()
        torch.set_num_threads(1)
        t0 = time.perf_counter()
        for _ in range(3):
            a.add_(b)
        ref = 3 * 3 * nbytes / (time.perf_counter() - t0) / 1e9
        for th in [int(x) for x in args.threads.split(",")]:
            red = c.CpuReducer(th)
            red.sum(a.data_ptr(), b.data_ptr(), nbytes, getattr(c, code))
            t0 = time.perf_counter()
            for _ in range(args.iters):
                red.sum(a.data_ptr(), b.data_ptr(), nbytes, getattr(c, code))
            ms = (time.perf_counter() - t0) / args.iters * 1e3
            print("| %s | %d | %.2f | %.1f | %.1f (1 thread) |" % (code, th, ms, 3 * nbytes / ms / 1e6, ref))


if __name__ == "__main__":
    main()
