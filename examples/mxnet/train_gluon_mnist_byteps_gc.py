#!/usr/bin/env python
"""Gluon training with DistributedTrainer and gradient compression (cf. the reference's
example/mxnet/train_gluon_mnist_byteps_gc.py).  Needs `mxnet` (not part of this image).

    bpslaunch python examples/mxnet/train_gluon_mnist_byteps_gc.py --compressor onebit --ef vanilla \
        --compress-momentum nesterov --onebit-scaling
"""
import argparse
import os
import sys

import mxnet as mx
from mxnet import autograd, gluon

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import byteps_b200.mxnet as bps  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--batch-size", type=int, default=64)
p.add_argument("--epochs", type=int, default=2)
p.add_argument("--lr", type=float, default=0.01)
p.add_argument("--momentum", type=float, default=0.9)
p.add_argument("--compressor", default="")
p.add_argument("--ef", default="")
p.add_argument("--compress-momentum", default="")
p.add_argument("--onebit-scaling", action="store_true")
p.add_argument("--k", type=float, default=1)
p.add_argument("--fp16-pushpull", action="store_true")
args = p.parse_args()

bps.init()
ctx = mx.gpu(bps.local_rank()) if mx.context.num_gpus() else mx.cpu()
net = gluon.nn.Sequential()
net.add(gluon.nn.Conv2D(20, 5, activation="relu"), gluon.nn.MaxPool2D(2, 2), gluon.nn.Flatten(),
        gluon.nn.Dense(128, activation="relu"), gluon.nn.Dense(10))
net.initialize(mx.init.Xavier(), ctx=ctx)
params = net.collect_params()

optimizer_params = {"momentum": args.momentum, "learning_rate": args.lr * bps.size()}
compression_params = {"compressor": args.compressor, "ef": args.ef, "momentum": args.compress_momentum,
                      "scaling": args.onebit_scaling, "k": args.k, "fp16": args.fp16_pushpull}
trainer = bps.DistributedTrainer(params, "sgd", optimizer_params, compression_params=compression_params)
loss_fn = gluon.loss.SoftmaxCrossEntropyLoss()

for epoch in range(args.epochs):
    for _ in range(100):
        x = mx.nd.random.uniform(shape=(args.batch_size, 1, 28, 28), ctx=ctx)
        y = (x.reshape((args.batch_size, -1)).sum(axis=1) * 7).astype("int32") % 10
        with autograd.record():
            loss = loss_fn(net(x), y)
        loss.backward()
        trainer.step(args.batch_size)
    if bps.rank() == 0:
        print("epoch %d loss %.4f" % (epoch, loss.mean().asscalar()))
