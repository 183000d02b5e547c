#!/usr/bin/env python
"""Synthetic benchmark with CrossBarrier (reference:
example/pytorch/benchmark_cross_barrier_byteps.py): the optimizer update of each
parameter bucket is fused into its exchange kernel and the next forward only
waits for the buckets of the layer it is about to run."""
import argparse
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import byteps_b200.torch as bps  # noqa: E402
from byteps_b200.models import get_model  # noqa: E402
from byteps_b200.torch.cross_barrier import CrossBarrier  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--model", default="resnet50")
p.add_argument("--batch-size", type=int, default=32)
p.add_argument("--steps", type=int, default=50)
args = p.parse_args()
bps.init()
torch.cuda.set_device(bps.local_rank())
model = get_model(args.model).cuda()
opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9)
opt = CrossBarrier(model, opt, named_parameters=model.named_parameters(), num_steps=args.steps + 10)
bps.broadcast_parameters(model.state_dict(), root_rank=0)
x = torch.rand(args.batch_size, 3, 224, 224, device="cuda")
y = torch.randint(0, 1000, (args.batch_size,), device="cuda")
for i in range(args.steps + 10):
    if i == 10:
        torch.cuda.synchronize()
        t0 = time.time()
    opt.zero_grad()
    F.cross_entropy(model(x), y).backward()
    opt.step()
torch.cuda.synchronize()
if bps.rank() == 0:
    print("Total img/sec on %d GPU(s): %.1f" % (bps.size(), bps.size() * args.batch_size * args.steps / (time.time() - t0)))
bps.shutdown()
