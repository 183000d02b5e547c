#!/usr/bin/env python
"""Symbolic (Module API) ImageNet-style training with `bps.DistributedOptimizer` (cf. the reference's
example/mxnet/train_imagenet_byteps.py + common/fit_byteps.py).  `--benchmark 1` trains on synthetic data and
prints img/sec the way the reference's script does.  Needs `mxnet` (not part of this image).

    bpslaunch python examples/mxnet/train_imagenet_byteps.py --benchmark 1 --batch-size 64 --network resnet --num-layers 50
"""
import argparse
import os
import sys
import time

import mxnet as mx

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import byteps_b200.mxnet as bps  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--network", default="resnet")
p.add_argument("--num-layers", type=int, default=50)
p.add_argument("--batch-size", type=int, default=64, help="per worker")
p.add_argument("--image-shape", default="3,224,224")
p.add_argument("--num-classes", type=int, default=1000)
p.add_argument("--num-epochs", type=int, default=1)
p.add_argument("--num-batches", type=int, default=100, help="batches per epoch in benchmark mode")
p.add_argument("--lr", type=float, default=0.1)
p.add_argument("--mom", type=float, default=0.9)
p.add_argument("--wd", type=float, default=1e-4)
p.add_argument("--benchmark", type=int, default=1)
p.add_argument("--data-train", default="", help="RecordIO file (each worker reads its own shard)")
p.add_argument("--disp-batches", type=int, default=20)
p.add_argument("--model-prefix", default="")
args = p.parse_args()

bps.init()
ctx = mx.gpu(bps.local_rank()) if mx.context.num_gpus() else mx.cpu()
shape = tuple(int(x) for x in args.image_shape.split(","))


def residual_unit(data, filters, stride, match, name):
    bn1 = mx.sym.BatchNorm(data, fix_gamma=False, name=name + "_bn1")
    act1 = mx.sym.Activation(bn1, act_type="relu")
    c1 = mx.sym.Convolution(act1, num_filter=filters // 4, kernel=(1, 1), no_bias=True, name=name + "_c1")
    a2 = mx.sym.Activation(mx.sym.BatchNorm(c1, fix_gamma=False, name=name + "_bn2"), act_type="relu")
    c2 = mx.sym.Convolution(a2, num_filter=filters // 4, kernel=(3, 3), stride=stride, pad=(1, 1), no_bias=True,
                            name=name + "_c2")
    a3 = mx.sym.Activation(mx.sym.BatchNorm(c2, fix_gamma=False, name=name + "_bn3"), act_type="relu")
    c3 = mx.sym.Convolution(a3, num_filter=filters, kernel=(1, 1), no_bias=True, name=name + "_c3")
    sc = data if match else mx.sym.Convolution(act1, num_filter=filters, kernel=(1, 1), stride=stride, no_bias=True,
                                               name=name + "_sc")
    return c3 + sc


def resnet(num_layers, num_classes):
    units = {50: [3, 4, 6, 3], 101: [3, 4, 23, 3], 152: [3, 8, 36, 3]}[num_layers]
    data = mx.sym.Variable("data")
    body = mx.sym.Convolution(data, num_filter=64, kernel=(7, 7), stride=(2, 2), pad=(3, 3), no_bias=True, name="conv0")
    body = mx.sym.Activation(mx.sym.BatchNorm(body, fix_gamma=False, name="bn0"), act_type="relu")
    body = mx.sym.Pooling(body, kernel=(3, 3), stride=(2, 2), pad=(1, 1), pool_type="max")
    for i, n in enumerate(units):
        filters = 256 * 2 ** i
        body = residual_unit(body, filters, (1 if i == 0 else 2,) * 2, False, "stage%d_unit1" % (i + 1))
        for j in range(n - 1):
            body = residual_unit(body, filters, (1, 1), True, "stage%d_unit%d" % (i + 1, j + 2))
    body = mx.sym.Activation(mx.sym.BatchNorm(body, fix_gamma=False, name="bn1"), act_type="relu")
    pool = mx.sym.Pooling(body, global_pool=True, kernel=(7, 7), pool_type="avg")
    fc = mx.sym.FullyConnected(mx.sym.Flatten(pool), num_hidden=num_classes, name="fc1")
    return mx.sym.SoftmaxOutput(fc, name="softmax")


class SyntheticIter(mx.io.DataIter):
    def __init__(self, batches):
        super().__init__(args.batch_size)
        self.batches, self.cur = batches, 0
        self.data = mx.nd.random.uniform(-1, 1, shape=(args.batch_size,) + shape, ctx=ctx)
        self.label = mx.nd.array([i % args.num_classes for i in range(args.batch_size)], ctx=ctx)
        self.provide_data = [mx.io.DataDesc("data", self.data.shape)]
        self.provide_label = [mx.io.DataDesc("softmax_label", (args.batch_size,))]

    def reset(self):
        self.cur = 0

    def next(self):
        if self.cur >= self.batches:
            raise StopIteration
        self.cur += 1
        return mx.io.DataBatch(data=(self.data,), label=(self.label,), pad=0)


if args.benchmark or not args.data_train:
    train = SyntheticIter(args.num_batches)
else:       # every worker reads its own part of the record file
    train = mx.io.ImageRecordIter(path_imgrec=args.data_train, data_shape=shape, batch_size=args.batch_size,
                                  rand_crop=True, rand_mirror=True, shuffle=True,
                                  num_parts=bps.size(), part_index=bps.rank())

model = mx.mod.Module(resnet(args.num_layers, args.num_classes), context=ctx)
model.bind(data_shapes=train.provide_data, label_shapes=train.provide_label)
model.init_params(mx.init.Xavier(rnd_type="gaussian", factor_type="in", magnitude=2))

# same start on every worker: rank 0's parameters are broadcast through push_pull
arg_params, aux_params = model.get_params()
bps.broadcast_parameters(arg_params, root_rank=0)
bps.broadcast_parameters(aux_params, root_rank=0)
model.set_params(arg_params=arg_params, aux_params=aux_params)

# gradients are summed over workers inside the optimizer's update(); rescale by the global batch
opt = mx.optimizer.create("sgd", learning_rate=args.lr * bps.size(), momentum=args.mom, wd=args.wd,
                          rescale_grad=1.0 / (args.batch_size * bps.size()))
model.init_optimizer(kvstore=None, optimizer=bps.DistributedOptimizer(opt))

metric = mx.metric.create(["accuracy"])
for epoch in range(args.num_epochs):
    train.reset()
    metric.reset()
    tic = time.time()
    for i, batch in enumerate(train, 1):
        model.forward_backward(batch)
        model.update()
        model.update_metric(metric, batch.label)
        if i % args.disp_batches == 0:
            mx.nd.waitall()
            speed = args.disp_batches * args.batch_size / (time.time() - tic)
            if bps.rank() == 0:
                print("Epoch[%d] Batch[%d]\tSpeed: %.1f samples/sec per worker (%.1f total)\t%s=%f" % (
                    epoch, i, speed, speed * bps.size(), *metric.get()))
            tic = time.time()
    if args.model_prefix and bps.rank() == 0:       # checkpoints from one worker only
        model.save_checkpoint(args.model_prefix, epoch + 1)
