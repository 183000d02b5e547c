#!/usr/bin/env python
"""CPU-server summation mode with pinned host buffers (BASELINE.json config 5): push_pull of a GPU
gradient through scheduler + server(s) + workers on ONE host, every worker on its own GPU.

    python -m byteps_b200.launcher.local_cluster -n 2 -s 1 python benchmarks/ps_bench.py --mb 100

Path per step: D2H copy of the gradient into pinned (shared-memory) staging on a side stream -> push over
the TCP van or, colocated, through POSIX shm (BYTEPS_ENABLE_IPC=1) -> AVX summation on the server ->
pull -> H2D.  Prints the goodput per worker (gradient bytes / wall time of one push_pull) and the
end-to-end rate.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=100)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu", action="store_true", help="host tensors (no GPU)")
    ap.add_argument("--out", default="")
    ap.add_argument("--compressor", default="", help="onebit | topk | randomk | dithering (CPU compressors, both sides)")
    ap.add_argument("--k", default="0.01")
    ap.add_argument("--ef", action="store_true", help="vanilla error feedback")
    args = ap.parse_args()
    wid = int(os.environ.get("DMLC_WORKER_ID", "0"))
    use_cuda = torch.cuda.is_available() and not args.cpu
    if use_cuda:
        torch.cuda.set_device(wid % torch.cuda.device_count())
    import byteps_b200.torch as bps

    bps.init()
    n = args.mb * 1000 * 1000 // 4
    g = torch.full((n,), float(bps.rank() + 1), device="cuda" if use_cuda else "cpu")
    expect = sum(range(1, bps.size() + 1)) / bps.size()

    if args.compressor:
        kw = {"byteps_compressor_type": args.compressor, "byteps_compressor_k": args.k}
        if args.compressor == "onebit":
            kw["byteps_compressor_onebit_scaling"] = "true"
        if args.ef:
            kw["byteps_ef_type"] = "vanilla"
        os.environ.setdefault("BYTEPS_MIN_COMPRESS_BYTES", "0")
        bps.declare("ps.g", **kw)

    def run():
        bps.push_pull_inplace(g, average=True, name="ps.g")
        if use_cuda:
            torch.cuda.synchronize()
    for _ in range(args.warmup):
        g.fill_(float(bps.rank() + 1))
        run()
        if not args.compressor:
            assert abs(g[0].item() - expect) < 1e-6 and abs(g[-1].item() - expect) < 1e-6
    ts = []
    for _ in range(args.iters):
        g.fill_(float(bps.rank() + 1))
        if use_cuda:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        run()
        ts.append(time.perf_counter() - t0)
    ms = 1e3 * sum(ts) / len(ts)
    row = {"workers": bps.size(), "servers": int(os.environ.get("DMLC_NUM_SERVER", "1")), "bytes": n * 4, "ms": ms,
           "gbs_per_worker": n * 4 / ms / 1e6, "device": "cuda" if use_cuda else "cpu",
           "ipc": os.environ.get("BYTEPS_ENABLE_IPC", "0"), "lanes": os.environ.get("DMLC_NUM_PORTS", "2"), "compressor": args.compressor + ("+ef" if args.ef else "")}
    if bps.rank() == 0:
        print("ps push_pull %d MB x %d workers: %.2f ms  (%.2f GB/s per worker; ipc=%s lanes=%s device=%s%s)" % (
            args.mb, bps.size(), ms, row["gbs_per_worker"], row["ipc"], row["lanes"], row["device"],
            " compressor=" + row["compressor"] if row["compressor"] else ""), flush=True)
        if args.out:
            json.dump(row, open(args.out, "w"))
    bps.shutdown()


if __name__ == "__main__":
    main()
