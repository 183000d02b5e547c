#!/usr/bin/env python
"""Compressed push-pull with error feedback on the NVLink path (BASELINE.json config 4).

torchrun, one process per GPU.  For a 100 MB fp32 gradient (25 M elements) it times, on the device with
CUDA events (max over ranks), one `push_pull` through the public API:

* uncompressed (fused pack/reduce/unpack kernel),
* onebit + scaling + vanilla error feedback,
* top-k (1 %) + vanilla error feedback,
* random-k (1 %) + error feedback, dithering (s = 4, linear, max-normalised),

each with the reference's two-stage contract (worker compression, "server" recompression of the sum).
Payload bytes that cross NVLink per rank are printed next to the time.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=100)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import byteps_b200.torch as bps
    from byteps_b200.ops.compress import GpuCompressor

    bps.init()
    world = bps.size()
    n = args.mb * 1000 * 1000 // 4
    torch.manual_seed(bps.rank())
    g = torch.randn(n, device=dev)
    cases = [
        ("uncompressed", None),
        ("onebit+scaling+ef", dict(compressor_type="onebit", compressor_onebit_scaling="true", ef_type="vanilla")),
        ("topk1%+ef", dict(compressor_type="topk", compressor_k=0.01, ef_type="vanilla")),
        ("randomk1%+ef", dict(compressor_type="randomk", compressor_k=0.01, ef_type="vanilla", seed=7)),
        ("dithering4", dict(compressor_type="dithering", compressor_k=4, dithering_partition=0,
                            dithering_normalize=0, seed=7)),
    ]
    rows = []
    for name, kw in cases:
        tname = "cb." + name
        if kw:
            bps.declare(tname, **kw)
        out = torch.empty_like(g)

        def run():
            h = bps.push_pull_async(g, average=True, name=tname)
            return bps.synchronize(h)
        for _ in range(args.warmup):
            run()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        tot = 0.0
        for _ in range(args.iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = run()
            e1.record()
            torch.cuda.synchronize(dev)
            tot += e0.elapsed_time(e1)
        ms = tot / args.iters
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        payload = GpuCompressor.slot_bytes_for(kw, n) if kw else n * 4
        rows.append({"case": name, "ms": ms, "payload_bytes_per_rank": payload, "gradient_bytes": n * 4,
                     "ratio": n * 4 / payload, "finite": bool(torch.isfinite(out).all().item())})
        if bps.rank() == 0:
            print("%-22s %8.3f ms   payload %10d B (%.0fx smaller)" % (name, ms, payload, n * 4 / payload), flush=True)
    if bps.rank() == 0 and args.out:
        json.dump({"n_gpus": world, "rows": rows}, open(args.out, "w"), indent=1)
    bps.shutdown()


if __name__ == "__main__":
    main()
