#!/usr/bin/env python
"""Synthetic benchmark with cross-iteration overlap (reference:
example/pytorch/benchmark_cross_barrier_byteps.py): step() does not wait for the exchange; the next
forward blocks per layer only on the parameters it is about to use."""
import argparse
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import byteps_b200.torch as bps  # noqa: E402
from byteps_b200.models import get_model  # noqa: E402
from byteps_b200.torch.cross_barrier import CrossBarrier  # noqa: E402

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
model = get_model(args.model).to(dev)
opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9)
opt = CrossBarrier(model, opt, named_parameters=model.named_parameters(), num_steps=total)   # drains at the last step
bps.broadcast_parameters(model.state_dict(), root_rank=0)
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
