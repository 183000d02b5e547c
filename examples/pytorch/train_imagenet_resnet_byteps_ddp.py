#!/usr/bin/env python
"""ResNet training with byteps_b200's DistributedDataParallel and a DistributedSampler
(the DDP flavour of the reference's ImageNet example, example/pytorch/train_imagenet_resnet_byteps_ddp.py).
Without --train-dir it trains on a synthetic ImageNet-shaped dataset.

    torchrun --nproc-per-node 8 examples/pytorch/train_imagenet_resnet_byteps_ddp.py --epochs 1
"""
import argparse
import os
import sys

import torch
import torch.nn.functional as F
from torch.utils.data import DataLoader, TensorDataset
from torch.utils.data.distributed import DistributedSampler

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import byteps_b200.torch as bps  # noqa: E402
from byteps_b200.models import get_model  # noqa: E402
from byteps_b200.torch.parallel import DistributedDataParallel as DDP  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--train-dir", default="")
p.add_argument("--model", default="resnet50")
p.add_argument("--batch-size", type=int, default=32)
p.add_argument("--epochs", type=int, default=1)
p.add_argument("--base-lr", type=float, default=0.0125)
p.add_argument("--momentum", type=float, default=0.9)
p.add_argument("--wd", type=float, default=5e-5)
p.add_argument("--synthetic-samples", type=int, default=512)
p.add_argument("--image-size", type=int, default=224)
p.add_argument("--no-cuda", action="store_true")
args = p.parse_args()

bps.init()
cuda = torch.cuda.is_available() and not args.no_cuda
if cuda:
    torch.cuda.set_device(bps.local_rank())
device = torch.device("cuda", bps.local_rank()) if cuda else torch.device("cpu")

if args.train_dir:
    from torchvision import datasets, transforms   # only needed for real data

    ds = datasets.ImageFolder(args.train_dir, transforms.Compose([
        transforms.RandomResizedCrop(args.image_size), transforms.RandomHorizontalFlip(), transforms.ToTensor()]))
else:
    g = torch.Generator().manual_seed(0)
    ds = TensorDataset(torch.rand(args.synthetic_samples, 3, args.image_size, args.image_size, generator=g),
                       torch.randint(0, 1000, (args.synthetic_samples,), generator=g))
# every worker reads its own shard of every epoch
sampler = DistributedSampler(ds, num_replicas=bps.size(), rank=bps.rank())
loader = DataLoader(ds, batch_size=args.batch_size, sampler=sampler, num_workers=0, pin_memory=cuda)

model = get_model(args.model).to(device)
model = DDP(model, device_ids=[bps.local_rank()] if cuda else None)   # broadcasts the initial state
optimizer = torch.optim.SGD(model.parameters(), lr=args.base_lr * bps.size(), momentum=args.momentum,
                            weight_decay=args.wd)

for epoch in range(args.epochs):
    sampler.set_epoch(epoch)
    model.train()
    seen, loss_sum = 0, 0.0
    for x, y in loader:
        x, y = x.to(device, non_blocking=True), y.to(device, non_blocking=True)
        optimizer.zero_grad()
        loss = F.cross_entropy(model(x), y)
        loss.backward()            # gradient exchange overlaps with the rest of backward
        optimizer.step()           # DDP made the gradients global averages
        seen += x.size(0)
        loss_sum += loss.item() * x.size(0)
    # metric averaging across workers, like the reference's Metric helper
    avg = bps.push_pull(torch.tensor([loss_sum / max(seen, 1)]), average=True, name="epoch_loss").item()
    if bps.rank() == 0:
        print("epoch %d: mean loss %.4f over %d workers" % (epoch, avg, bps.size()))
bps.shutdown()
