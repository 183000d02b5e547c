#!/usr/bin/env python
"""Synthetic benchmark (same protocol as the reference's
example/pytorch/benchmark_byteps.py: 10 warm-up batches, then N iterations of 10
batches, img/sec mean +- 1.96 sigma, total = per-GPU x size).

    bpslaunch python examples/pytorch/benchmark_byteps.py --model resnet50
    torchrun --nproc-per-node 8 examples/pytorch/benchmark_byteps.py --fp16-pushpull
"""
import argparse
import os
import sys
import timeit

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import byteps_b200.torch as bps  # noqa: E402
from byteps_b200.models import get_model  # noqa: E402

p = argparse.ArgumentParser(description="PyTorch Synthetic Benchmark")
p.add_argument("--fp16-pushpull", action="store_true", help="use fp16 compression during push_pull")
p.add_argument("--bf16-pushpull", action="store_true", help="use bf16 on the wire")
p.add_argument("--model", default="resnet50")
p.add_argument("--batch-size", type=int, default=32)
p.add_argument("--num-warmup-batches", type=int, default=10)
p.add_argument("--num-batches-per-iter", type=int, default=10)
p.add_argument("--num-iters", type=int, default=10)
p.add_argument("--num-classes", type=int, default=1000)
p.add_argument("--no-cuda", action="store_true")
p.add_argument("--fused-update", action="store_true", help="fuse the SGD step into the exchange kernel")
args = p.parse_args()
args.cuda = not args.no_cuda and torch.cuda.is_available()

bps.init()
if args.cuda:
    torch.cuda.set_device(bps.local_rank())
torch.backends.cudnn.benchmark = True
model = get_model(args.model, num_classes=args.num_classes)
if args.cuda:
    model.cuda()
optimizer = torch.optim.SGD(model.parameters(), lr=0.01)
compression = bps.Compression.fp16 if args.fp16_pushpull else (
    bps.Compression.bf16 if args.bf16_pushpull else bps.Compression.none)
optimizer = bps.DistributedOptimizer(optimizer, named_parameters=model.named_parameters(), compression=compression,
                                     fused_update=args.fused_update)
bps.broadcast_parameters(model.state_dict(), root_rank=0)
bps.broadcast_optimizer_state(optimizer, root_rank=0)

data = torch.rand(args.batch_size, 3, 224, 224)
target = torch.randint(0, args.num_classes, (args.batch_size,))
if args.cuda:
    data, target = data.cuda(), target.cuda()


def benchmark_step():
    optimizer.zero_grad()
    loss = F.cross_entropy(model(data), target)
    loss.backward()
    optimizer.step()
    if args.cuda:
        torch.cuda.synchronize()


def log(s):
    if bps.local_rank() == 0:
        print(s, flush=True)


log("Model: %s\nBatch size: %d\nNumber of %ss: %d" % (args.model, args.batch_size, "GPU" if args.cuda else "CPU",
                                                      bps.size()))
timeit.timeit(benchmark_step, number=args.num_warmup_batches)
img_secs = []
for x in range(args.num_iters):
    t = timeit.timeit(benchmark_step, number=args.num_batches_per_iter)
    img_sec = args.batch_size * args.num_batches_per_iter / t
    log("Iter #%d: %.1f img/sec per device" % (x, img_sec))
    img_secs.append(img_sec)
m, c = np.mean(img_secs), 1.96 * np.std(img_secs)
log("Img/sec per device: %.1f +-%.1f" % (m, c))
log("Total img/sec on %d device(s): %.1f +-%.1f" % (bps.size(), bps.size() * m, bps.size() * c))
bps.shutdown()
