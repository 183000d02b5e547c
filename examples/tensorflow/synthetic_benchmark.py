#!/usr/bin/env python
"""TF1-style (graph + session) synthetic benchmark: `DistributedOptimizer` wrapping `compute_gradients` and
`BroadcastGlobalVariablesHook` (cf. the reference's example/tensorflow/synthetic_benchmark.py).
Needs `tensorflow` (run through `tf.compat.v1` on TF2).

    bpslaunch python examples/tensorflow/synthetic_benchmark.py --model ResNet50 --batch-size 32
"""
import argparse
import os
import sys
import timeit

import numpy as np
import tensorflow as tf

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import byteps_b200.tensorflow as bps  # noqa: E402

tf1 = tf.compat.v1
p = argparse.ArgumentParser()
p.add_argument("--model", default="ResNet50")
p.add_argument("--batch-size", type=int, default=32)
p.add_argument("--fp16-pushpull", action="store_true")
p.add_argument("--num-warmup-batches", type=int, default=10)
p.add_argument("--num-batches-per-iter", type=int, default=10)
p.add_argument("--num-iters", type=int, default=10)
args = p.parse_args()

bps.init()
tf1.disable_eager_execution()
config = tf1.ConfigProto()
config.gpu_options.allow_growth = True
config.gpu_options.visible_device_list = str(bps.local_rank())

model = getattr(tf.keras.applications, args.model)(weights=None)
compression = bps.Compression.fp16 if args.fp16_pushpull else bps.Compression.none
opt = bps.DistributedOptimizer(tf1.train.GradientDescentOptimizer(0.01), compression=compression)

data = tf.random.uniform([args.batch_size, 224, 224, 3])
target = tf.random.uniform([args.batch_size, 1], minval=0, maxval=999, dtype=tf.int64)
loss = tf1.losses.sparse_softmax_cross_entropy(target, model(data, training=True))
train_op = opt.minimize(loss)                       # compute_gradients -> push_pull -> apply_gradients
hooks = [bps.BroadcastGlobalVariablesHook(0)]       # rank 0's initial values everywhere

with tf1.train.MonitoredTrainingSession(hooks=hooks, config=config) as session:
    timeit.timeit(lambda: session.run(train_op), number=args.num_warmup_batches)
    rates = []
    for i in range(args.num_iters):
        t = timeit.timeit(lambda: session.run(train_op), number=args.num_batches_per_iter)
        rates.append(args.batch_size * args.num_batches_per_iter / t)
        if bps.rank() == 0:
            print("Iter #%d: %.1f img/sec per GPU" % (i, rates[-1]))
if bps.rank() == 0:
    m, ci = np.mean(rates), 1.96 * np.std(rates)
    print("Img/sec per GPU: %.1f +-%.1f; total on %d GPU(s): %.1f +-%.1f" % (m, ci, bps.size(), bps.size() * m,
                                                                             bps.size() * ci))
