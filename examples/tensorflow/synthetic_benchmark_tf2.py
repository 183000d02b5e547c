#!/usr/bin/env python
"""TF2 synthetic benchmark (cf. the reference's example/tensorflow/synthetic_benchmark_tf2.py): a Keras
application model trained with a DistributedGradientTape.  Needs `tensorflow` (not part of this image).

    bpslaunch python examples/tensorflow/synthetic_benchmark_tf2.py --model ResNet50 --batch-size 32
"""
import argparse
import os
import sys
import timeit

import numpy as np
import tensorflow as tf

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import byteps_b200.tensorflow as bps  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--model", default="ResNet50")
p.add_argument("--batch-size", type=int, default=32)
p.add_argument("--fp16-pushpull", action="store_true")
p.add_argument("--num-warmup-batches", type=int, default=10)
p.add_argument("--num-batches-per-iter", type=int, default=10)
p.add_argument("--num-iters", type=int, default=10)
args = p.parse_args()

bps.init()
gpus = tf.config.experimental.list_physical_devices("GPU")
for g in gpus:
    tf.config.experimental.set_memory_growth(g, True)
if gpus:
    tf.config.experimental.set_visible_devices(gpus[bps.local_rank()], "GPU")

model = getattr(tf.keras.applications, args.model)(weights=None)
opt = tf.optimizers.SGD(0.01)
compression = bps.Compression.fp16 if args.fp16_pushpull else bps.Compression.none
data = tf.random.uniform([args.batch_size, 224, 224, 3])
target = tf.random.uniform([args.batch_size, 1], minval=0, maxval=999, dtype=tf.int64)


def benchmark_step(first_batch):
    with tf.GradientTape() as tape:
        loss = tf.losses.sparse_categorical_crossentropy(target, model(data, training=True))
    tape = bps.DistributedGradientTape(tape, compression=compression)     # averages over all workers
    grads = tape.gradient(loss, model.trainable_variables)
    opt.apply_gradients(zip(grads, model.trainable_variables))
    if first_batch:     # after the first step so the optimizer slots exist too
        bps.broadcast_variables(model.variables, root_rank=0)
        bps.broadcast_variables(opt.variables(), root_rank=0)


benchmark_step(True)
timeit.timeit(lambda: benchmark_step(False), number=args.num_warmup_batches)
rates = []
for _ in range(args.num_iters):
    t = timeit.timeit(lambda: benchmark_step(False), number=args.num_batches_per_iter)
    rates.append(args.batch_size * args.num_batches_per_iter / t)
if bps.rank() == 0:
    m, ci = np.mean(rates), 1.96 * np.std(rates)
    print("Img/sec per GPU: %.1f +-%.1f; total on %d GPU(s): %.1f +-%.1f" % (m, ci, bps.size(), bps.size() * m,
                                                                             bps.size() * ci))
