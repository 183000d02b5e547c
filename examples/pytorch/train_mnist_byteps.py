#!/usr/bin/env python
"""MNIST-shaped training loop (reference: example/pytorch/train_mnist_byteps.py).
There is no network for datasets, so synthetic 1x28x28 digits are used unless
--data points at a torchvision MNIST folder.  Shows metric averaging by push_pull."""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import byteps_b200.torch as bps  # noqa: E402
from byteps_b200.models import MnistNet  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--batch-size", type=int, default=64)
p.add_argument("--epochs", type=int, default=1)
p.add_argument("--lr", type=float, default=0.01)
p.add_argument("--momentum", type=float, default=0.5)
p.add_argument("--fp16-pushpull", action="store_true")
p.add_argument("--no-cuda", action="store_true")
args = p.parse_args()
bps.init()
use_cuda = torch.cuda.is_available() and not args.no_cuda
if use_cuda:
    torch.cuda.set_device(bps.local_rank())
dev = torch.device("cuda" if use_cuda else "cpu")
torch.manual_seed(42 + bps.rank())
xs = torch.randn(64 * args.batch_size, 1, 28, 28)
ys = torch.randint(0, 10, (64 * args.batch_size,))
model = MnistNet().to(dev)
# scale the learning rate by the number of workers, like the reference
opt = torch.optim.SGD(model.parameters(), lr=args.lr * bps.size(), momentum=args.momentum)
comp = bps.Compression.fp16 if args.fp16_pushpull else bps.Compression.none
opt = bps.DistributedOptimizer(opt, named_parameters=model.named_parameters(), compression=comp)
bps.broadcast_parameters(model.state_dict(), root_rank=0)
bps.broadcast_optimizer_state(opt, root_rank=0)


def metric_average(val, name):
    return bps.push_pull(torch.tensor(val), name=name).item()


for epoch in range(args.epochs):
    model.train()
    for i in range(0, xs.shape[0], args.batch_size):
        x, y = xs[i:i + args.batch_size].to(dev), ys[i:i + args.batch_size].to(dev)
        opt.zero_grad()
        loss = F.nll_loss(model(x), y)
        loss.backward()
        opt.step()
    avg = metric_average(loss.item(), "avg_loss")
    if bps.rank() == 0:
        print("epoch %d: loss averaged over %d workers = %.4f" % (epoch, bps.size(), avg))
bps.shutdown()
