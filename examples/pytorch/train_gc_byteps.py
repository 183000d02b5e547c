#!/usr/bin/env python
"""Gradient compression end to end (the torch counterpart of the reference's MXNet-only
example/mxnet/train_gluon_mnist_byteps_gc.py): per-tensor onebit / topk / randomk / dithering with
error feedback and Nesterov momentum.  On the NVLink backend the compressors are GPU kernels whose
payloads travel through symmetric memory; in CPU-server mode the worker and the server run the native
CPU compressors.

    torchrun --nproc-per-node 2 examples/pytorch/train_gc_byteps.py --compressor topk --k 0.01 --ef vanilla
"""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import byteps_b200.torch as bps  # noqa: E402
from byteps_b200.models import MnistNet  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--compressor", default="onebit", choices=["onebit", "topk", "randomk", "dithering"])
p.add_argument("--k", type=float, default=0.01, help="top-k/random-k count or fraction; dithering levels")
p.add_argument("--ef", default="vanilla", choices=["", "vanilla"])
p.add_argument("--compress-momentum", default="", choices=["", "nesterov"])
p.add_argument("--onebit-scaling", action="store_true")
p.add_argument("--partition", default="linear", choices=["linear", "natural"])
p.add_argument("--normalize", default="max", choices=["max", "l2"])
p.add_argument("--seed", type=int, default=2020)
p.add_argument("--lr", type=float, default=0.05)
p.add_argument("--momentum", type=float, default=0.9)
p.add_argument("--steps", type=int, default=50)
p.add_argument("--batch-size", type=int, default=64)
p.add_argument("--no-cuda", action="store_true")
args = p.parse_args()

bps.init()
cuda = torch.cuda.is_available() and not args.no_cuda
if cuda:
    torch.cuda.set_device(bps.local_rank())
device = torch.device("cuda", bps.local_rank()) if cuda else torch.device("cpu")
torch.manual_seed(args.seed)
model = MnistNet().to(device)

cp = {"compressor": args.compressor, "seed": args.seed}
if args.compressor == "onebit":
    cp["scaling"] = args.onebit_scaling
else:
    cp["k"] = args.k if args.compressor != "dithering" else int(max(args.k, 2))
if args.compressor == "dithering":
    cp.update(partition=args.partition, normalize=args.normalize)
if args.ef:
    cp["ef"] = args.ef
if args.compress_momentum:
    cp["momentum"] = args.compress_momentum

opt = torch.optim.SGD(model.parameters(), lr=args.lr * bps.size(), momentum=args.momentum)
opt = bps.DistributedOptimizer(opt, named_parameters=model.named_parameters(), compression_params=cp)
bps.broadcast_parameters(model.state_dict(), root_rank=0)
bps.broadcast_optimizer_state(opt, root_rank=0)

gen = torch.Generator().manual_seed(100 + bps.rank())
for step in range(args.steps):
    x = torch.rand(args.batch_size, 1, 28, 28, generator=gen).to(device)
    y = (x.flatten(1).sum(1) * 7).long().remainder(10)      # a learnable synthetic rule
    opt.zero_grad()
    loss = F.nll_loss(model(x), y)
    loss.backward()
    opt.step()
    if step % 10 == 0 and bps.rank() == 0:
        print("step %3d  loss %.4f  (%s%s)" % (step, loss.item(), args.compressor, "+ef" if args.ef else ""))
bps.shutdown()
