#!/usr/bin/env python
"""CPU tensors of ranks that share a host: shared-memory reduction (csrc/core/host_reduce.h) vs gloo.

    python benchmarks/host_reduce_bench.py                 # 2 and 4 ranks on one host, BYTEPS_HOST_SHM_REDUCE auto vs 0
    python benchmarks/host_reduce_bench.py --hosts 2       # 4 ranks as 2 "hosts" x 2 (the roots all-reduce over gloo)
    python benchmarks/host_reduce_bench.py --phases        # time of every phase of HostLocalReduce on the root
    python benchmarks/host_reduce_bench.py --many          # 162 tensors of 1 KB - 1 MB per step

Every rank calls `bps.push_pull_inplace` on a 100 MB fp32 CPU tensor (BASELINE.json config 1); host wall clock, mean
of the iterations after two warm-up rounds, printed by rank 0.  Results: profiles/cpu_round2.md.
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from _mp import run_workers  # noqa: E402  (spawns `world` processes with torchrun-style variables)


def _pushpull(rank, world, local, mb, iters, mode):
    os.environ.update({"LOCAL_RANK": str(rank % local), "LOCAL_WORLD_SIZE": str(local), "GROUP_RANK": str(rank // local),
                       "BYTEPS_HOST_SHM_REDUCE": mode})
    import torch

    import byteps_b200.torch as bps

    bps.init()
    n = mb * 1000 * 1000 // 4
    g = torch.full((n,), float(rank + 1))
    expect = sum(range(1, world + 1)) / world
    ts = []
    for _ in range(iters + 2):
        g.fill_(float(rank + 1))
        t0 = time.perf_counter()
        bps.push_pull_inplace(g, average=True, name="g")
        ts.append(time.perf_counter() - t0)
        assert abs(g[0].item() - expect) < 1e-6 and abs(g[-1].item() - expect) < 1e-6
    if rank == 0:
        print("%d ranks as %d host(s) x %d, BYTEPS_HOST_SHM_REDUCE=%-4s: %6.1f ms per %d MB push_pull (best %.1f)" % (
            world, world // local, local, mode, 1e3 * sum(ts[2:]) / len(ts[2:]), mb, 1e3 * min(ts)), flush=True)
    bps.shutdown()


def _many(rank, world, iters, mode):
    """162 tensors of 1 KB - 1 MB in flight at once (a CNN's gradient set): per-operation overhead, not bandwidth"""
    os.environ["BYTEPS_HOST_SHM_REDUCE"] = mode
    import torch

    import byteps_b200.torch as bps

    bps.init()
    sizes = [256, 1024, 4096, 16384, 65536, 262144] * 27
    ts = [torch.ones(n) for n in sizes]
    res = []
    for _ in range(iters + 2):
        t0 = time.perf_counter()
        hs = [bps.push_pull_async_inplace(t, average=True, name="t%d" % i) for i, t in enumerate(ts)]
        for h in hs:
            bps.synchronize(h)
        res.append(time.perf_counter() - t0)
    if rank == 0:
        print("%d ranks, BYTEPS_HOST_SHM_REDUCE=%-4s: %d tensors (%.1f MB) in %.1f ms per step" % (
            world, mode, len(sizes), sum(sizes) * 4 / 1e6, 1e3 * sum(res[2:]) / iters), flush=True)
    bps.shutdown()


def _phases(rank, world, mb, iters, tag):
    import numpy as np

    from byteps_b200 import _native

    c = _native.core()
    hr = c.HostLocalReduce(rank, world, tag, 0, "/tmp")
    n = mb * 1000 * 1000 // 4
    x = np.full(n, rank + 1, dtype=np.float32)
    out = np.zeros_like(x)
    time.sleep(0.5)           # every rank's socket is bound
    acc = [0.0, 0.0, 0.0]
    for it in range(iters + 2):
        t0 = time.perf_counter()
        hr.contribute(7, x.ctypes.data, x.nbytes, 20000)
        t1 = time.perf_counter()
        if hr.is_root():
            hr.reduce(7, x.nbytes, c.F32, 20000, 1.0 / world)
            t2 = time.perf_counter()
            hr.publish(7, out.ctypes.data, x.nbytes, 20000)
            t3 = time.perf_counter()
            if it >= 2:
                for i, v in enumerate((t1 - t0, t2 - t1, t3 - t2)):
                    acc[i] += v
        else:
            hr.collect(7, out.ctypes.data, x.nbytes, 20000, c.F32, 1.0 / world)
    if hr.is_root():
        print("%d ranks: contribute %.1f ms, reduce (wait + every rank sums and scales its share) %.1f ms, "
              "publish (copy out + acknowledgements) %.1f ms" % (world, *[1e3 * a / iters for a in acc]), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=100)
    ap.add_argument("--iters", type=int, default=6)
    ap.add_argument("--hosts", type=int, default=1, help="pretend the ranks are spread over this many hosts")
    ap.add_argument("--phases", action="store_true")
    ap.add_argument("--many", action="store_true", help="162 small tensors per step instead of one 100 MB tensor")
    args = ap.parse_args()
    if args.many:
        for mode in ("0", "auto"):
            run_workers(_many, world=2, args=(args.iters, mode), timeout=600)
        return
    if args.phases:
        for world in (2, 4):
            run_workers(_phases, world=world, args=(args.mb, args.iters, "bench%d" % os.getpid()), timeout=300)
        return
    worlds = (2, 4) if args.hosts == 1 else (2 * args.hosts,)
    for world in worlds:
        for mode in ("0", "auto"):
            run_workers(_pushpull, world=world, args=(world // args.hosts, args.mb, args.iters, mode), timeout=600)


if __name__ == "__main__":
    main()
