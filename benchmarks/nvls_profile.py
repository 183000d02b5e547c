#!/usr/bin/env python
"""ncu target for the multi-GPU exchange kernels: every rank maps the symmetric arena and the NVLS
multicast object, then ONLY rank 0 launches the kernels - without the cross-rank flag waits, because
ncu replays each kernel dozens of times while the peers are idle.  The data path is the real one:
multimem.ld_reduce pulls the shard out of all GPUs through the switch, multimem.st multicasts the
result into all of them.  Launch one process per GPU by hand (tools/round2/r2_8gpu.sh) with rank 0 under ncu.
"""
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    import byteps_b200.torch as bps
    from byteps_b200.common import engine
    from byteps_b200.comm.symm import SymmContext, wire_code
    from byteps_b200.ops.ring import RingEntry, RingTable

    bps.init()
    ctx = SymmContext(engine().group, dev, 512 << 20, "auto", "try")
    cu, view = ctx.cu, ctx.view
    nvls = bool(ctx.nvls)
    dt = torch.bfloat16
    n = 100 * 1000 * 1000 // 2 // 8 * 8                  # 100 MB of bf16
    ctx.tensor(0, n, dt).fill_(1.0)
    # ring: 4 buckets of 16 MB with the fused AdamW epilogue (what a BERT step launches)
    nb, bn = 4, (16 << 20) // 2
    goff = 128 << 20
    poff = 256 << 20
    hp = torch.zeros(64, dtype=torch.uint8, device=dev)
    cu.write_blob(hp.data_ptr(), struct.pack("<9f3if3i", 1e-4, 0.01, 0, 0, 0.9, 0.999, 1e-8, 0.1, 0.001, 0, 1, 1, 1.0,
                                             0, 0, 0), torch.cuda.current_stream().cuda_stream)
    entries, keep = [], []
    for i in range(nb):
        b, e = cu.shard_units(bn // 8, world, rank)
        m = [torch.zeros(max((e - b) * 8, 8), device=dev) for _ in range(3)]
        keep.append(m)
        ctx.tensor(goff + i * (16 << 20), bn, dt).fill_(0.5)
        entries.append(RingEntry(grad_off=goff + i * (16 << 20), param_off=poff + i * (16 << 20), numel=bn,
                                 wire=wire_code(dt), slot=i, kind=cu.RING_ADAM, scale=1.0 / world,
                                 master=m[0].data_ptr(), state0=m[1].data_ptr(), state1=m[2].data_ptr(),
                                 hp=hp.data_ptr()))
    table = RingTable(entries, dev)
    plain = RingTable([RingEntry(grad_off=goff + i * (16 << 20), numel=bn, wire=wire_code(dt), slot=8 + i,
                                 scale=1.0 / world) for i in range(nb)], dev)
    s = torch.cuda.current_stream().cuda_stream
    cu.ring_mark(view, list(range(nb)) + [8 + i for i in range(nb)], s)     # every rank: "my gradients are ready"
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        for _ in range(2):
            cu.pushpull_inplace(view, wire_code(dt), 0, n, 1.0 / world, 64, 512, -1, nvls, s)    # channel -1: no barriers
        table.launch(view, 32, s, nvls=nvls, self_mark=False, solo=True)
        plain.launch(view, 32, s, nvls=nvls, self_mark=False, solo=True)
        torch.cuda.synchronize()
        print("profiled on rank 0: nvls=%s world=%d" % (nvls, world), flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
