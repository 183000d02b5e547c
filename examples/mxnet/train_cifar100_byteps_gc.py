#!/usr/bin/env python
"""Gluon CIFAR-100-style training with every gradient-compression knob on the command line (cf. the reference's
example/mxnet/train_cifar100_byteps_gc.py and train_gluon_imagenet_byteps_gc.py): compressor (onebit / topk /
randomk / dithering), error feedback, compressed-momentum, fp16 push_pull, learning-rate decay.
Needs `mxnet` (not part of this image); data are synthetic unless --data-dir holds CIFAR-100 records.

    bpslaunch python examples/mxnet/train_cifar100_byteps_gc.py --compressor topk --k 0.01 --ef vanilla
    bpslaunch python examples/mxnet/train_cifar100_byteps_gc.py --compressor dithering --k 4 \
        --partition natural --normalize l2 --compress-momentum nesterov
"""
import argparse
import os
import sys
import time

import mxnet as mx
from mxnet import autograd, gluon

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import byteps_b200.mxnet as bps  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--batch-size", type=int, default=32, help="per worker")
p.add_argument("--epochs", type=int, default=3)
p.add_argument("--batches-per-epoch", type=int, default=100)
p.add_argument("--lr", type=float, default=0.1)
p.add_argument("--lr-decay", type=float, default=0.1)
p.add_argument("--lr-decay-epoch", default="100,150")
p.add_argument("--momentum", type=float, default=0.9)
p.add_argument("--wd", type=float, default=5e-4)
p.add_argument("--data-dir", default="")
# gradient compression (docs/gradient-compression.md)
p.add_argument("--compressor", default="", choices=["", "onebit", "topk", "randomk", "dithering"])
p.add_argument("--ef", default="", choices=["", "vanilla"])
p.add_argument("--compress-momentum", default="", choices=["", "nesterov"])
p.add_argument("--onebit-scaling", action="store_true")
p.add_argument("--k", type=float, default=1, help="topk/randomk: count or fraction; dithering: levels")
p.add_argument("--partition", default="linear", choices=["linear", "natural"])
p.add_argument("--normalize", default="max", choices=["max", "l2"])
p.add_argument("--seed", type=int, default=2020)
p.add_argument("--fp16-pushpull", action="store_true")
args = p.parse_args()

bps.init()
ctx = mx.gpu(bps.local_rank()) if mx.context.num_gpus() else mx.cpu()
mx.random.seed(args.seed + bps.rank())


def block(ch, stride):
    b = gluon.nn.HybridSequential()
    b.add(gluon.nn.Conv2D(ch, 3, stride, 1, use_bias=False), gluon.nn.BatchNorm(), gluon.nn.Activation("relu"),
          gluon.nn.Conv2D(ch, 3, 1, 1, use_bias=False), gluon.nn.BatchNorm(), gluon.nn.Activation("relu"))
    return b


net = gluon.nn.HybridSequential()
net.add(gluon.nn.Conv2D(32, 3, 1, 1, use_bias=False), gluon.nn.BatchNorm(), gluon.nn.Activation("relu"),
        block(64, 2), block(128, 2), block(256, 2), gluon.nn.GlobalAvgPool2D(), gluon.nn.Dense(100))
net.initialize(mx.init.Xavier(), ctx=ctx)
net.hybridize()

optimizer_params = {"learning_rate": args.lr * bps.size(), "wd": args.wd, "momentum": args.momentum}
compression_params = {
    "compressor": args.compressor, "ef": args.ef, "momentum": args.compress_momentum,
    "scaling": args.onebit_scaling, "k": args.k, "partition": args.partition, "normalize": args.normalize,
    "seed": args.seed, "fp16": args.fp16_pushpull,
}
# with a compressed momentum the optimizer's own momentum moves into the compressor (DistributedTrainer does that)
trainer = bps.DistributedTrainer(net.collect_params(), "sgd", optimizer_params,
                                 compression_params=compression_params)
loss_fn = gluon.loss.SoftmaxCrossEntropyLoss()
decay_epochs = [int(e) for e in args.lr_decay_epoch.split(",") if e]


def batches():
    if args.data_dir:
        it = mx.io.ImageRecordIter(path_imgrec=os.path.join(args.data_dir, "train.rec"), data_shape=(3, 32, 32),
                                   batch_size=args.batch_size, rand_crop=True, rand_mirror=True, shuffle=True,
                                   num_parts=bps.size(), part_index=bps.rank())
        for b in it:
            yield b.data[0].as_in_context(ctx), b.label[0].as_in_context(ctx)
    else:
        for _ in range(args.batches_per_epoch):
            x = mx.nd.random.uniform(shape=(args.batch_size, 3, 32, 32), ctx=ctx)
            yield x, (x.reshape((args.batch_size, -1)).sum(axis=1) * 13).astype("int32") % 100


for epoch in range(args.epochs):
    if epoch in decay_epochs:
        trainer.set_learning_rate(trainer.learning_rate * args.lr_decay)     # also reaches the error feedback
    tic, seen, metric = time.time(), 0, mx.metric.Accuracy()
    for x, y in batches():
        with autograd.record():
            out = net(x)
            loss = loss_fn(out, y)
        loss.backward()
        trainer.step(args.batch_size)
        metric.update([y], [out])
        seen += args.batch_size
    mx.nd.waitall()
    if bps.rank() == 0:
        print("[Epoch %d] %.1f samples/sec per worker  loss %.4f  %s=%.4f" % (
            epoch, seen / (time.time() - tic), loss.mean().asscalar(), *metric.get()))
