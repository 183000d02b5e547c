#!/usr/bin/env python
"""Transport benchmark: N workers x N servers on localhost, push / pull / push_pull of
fixed-size keys; prints goodput (Gbps) and ns per key.  Same protocol as ps-lite's
tests/test_benchmark.cc (the reference's CI runs it 1x1, 2x2, 4x4), on our TCP van
(optionally with the colocated shared-memory IPC path: --ipc).

    python benchmarks/kv_benchmark.py --workers 2 --servers 2 --len 4096000 --repeat 50
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=2)
    ap.add_argument("--servers", type=int, default=2)
    ap.add_argument("--len", type=int, default=4096000, help="bytes per key")
    ap.add_argument("--keys", type=int, default=8)
    ap.add_argument("--repeat", type=int, default=20)
    ap.add_argument("--ipc", action="store_true")
    ap.add_argument("--local", action="store_true", help="DMLC_LOCAL: Unix-domain sockets instead of TCP")
    ap.add_argument("--van", default="tcp", choices=["tcp", "shm"], help="DMLC_PS_VAN_TYPE")
    ap.add_argument("--lanes", type=int, default=2, help="DMLC_NUM_PORTS: connections per peer")
    args = ap.parse_args()
    from _cluster import Cluster

    from byteps_b200 import _native

    c = _native.core()
    cl = Cluster(args.workers, args.servers, extra={"enable_ipc": args.ipc, "local": args.local, "num_lanes": args.lanes, "van_type": args.van}).start()
    res = {}

    def work(rank, w, po):
        n = args.len // 4
        bufs = []
        for k in range(args.keys):
            key = c.make_key(k, 0)
            if args.ipc:
                import ctypes

                name = "bps_kvbench_%d_%d_%d" % (os.getpid(), rank, k)
                ptr = c.shm_create(name, n * 4)
                arr = np.frombuffer((ctypes.c_float * n).from_address(ptr), dtype=np.float32)
            else:
                arr = np.zeros(n, dtype=np.float32)
                ptr = arr.ctypes.data
            w.init_key(key, ptr, n * 4, c.F32)
            bufs.append((key, ptr, arr))
        po.barrier(0, c.GROUP_WORKER)
        t0 = time.time()
        for _ in range(args.repeat):
            hs = [w.push_pull("k%d" % i, ptr, c.F32, [(key, 0, n * 4)], 0, 0, 1.0) for i, (key, ptr, _) in enumerate(bufs)]
            for h in hs:
                w.wait(h)
        dt = time.time() - t0
        res[rank] = dt
    cl.run_workers(work)
    cl.stop()
    if args.ipc:
        for rank in range(args.workers):
            for k in range(args.keys):
                c.shm_release("bps_kvbench_%d_%d_%d" % (os.getpid(), rank, k))
    dt = max(res.values())
    total_bytes = 2.0 * args.len * args.keys * args.repeat * args.workers       # push + pull
    print("push_pull: %d workers x %d servers, %d keys x %d B x %d rounds in %.3f s -> goodput %.2f Gbps, "
          "%.0f ns per key" % (args.workers, args.servers, args.keys, args.len, args.repeat, dt,
                               total_bytes * 8 / dt / 1e9, dt / (args.keys * args.repeat) * 1e9))


if __name__ == "__main__":
    main()
