#!/usr/bin/env python
"""Synthetic benchmark with byteps_b200.torch.parallel.DistributedDataParallel and a
plain torch optimizer (reference: example/pytorch/benchmark_byteps_ddp.py)."""
import argparse
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import byteps_b200.torch as bps  # noqa: E402
from byteps_b200.models import get_model  # noqa: E402
from byteps_b200.torch.parallel import DistributedDataParallel as DDP  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--model", default="resnet50")
p.add_argument("--batch-size", type=int, default=32)
p.add_argument("--num-warmup-batches", type=int, default=10)
p.add_argument("--num-batches-per-iter", type=int, default=10)
p.add_argument("--num-iters", type=int, default=10)
p.add_argument("--image-size", type=int, default=224)
p.add_argument("--no-cuda", action="store_true")
args = p.parse_args()
bps.init()
cuda = torch.cuda.is_available() and not args.no_cuda
if cuda:
    torch.cuda.set_device(bps.local_rank())
dev = torch.device("cuda", bps.local_rank()) if cuda else torch.device("cpu")
total = args.num_warmup_batches + args.num_batches_per_iter * args.num_iters
model = DDP(get_model(args.model).to(dev), device_ids=[bps.local_rank()] if cuda else None)
opt = torch.optim.SGD(model.parameters(), lr=0.01)      # gradients are averaged by DDP during backward
x = torch.rand(args.batch_size, 3, args.image_size, args.image_size, device=dev)
y = torch.randint(0, 1000, (args.batch_size,), device=dev)


def sync():
    if cuda:
        torch.cuda.synchronize()


for i in range(total):
    if i == args.num_warmup_batches:
        sync()
        t0 = time.time()
    opt.zero_grad()
    F.cross_entropy(model(x), y).backward()
    opt.step()
sync()
if bps.rank() == 0:
    n = total - args.num_warmup_batches
    print("Total img/sec on %d %s(s): %.1f" % (bps.size(), "GPU" if cuda else "CPU worker",
                                               bps.size() * args.batch_size * n / (time.time() - t0)))
bps.shutdown()
