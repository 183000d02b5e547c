#!/usr/bin/env python
"""tf.keras `model.fit` synthetic benchmark (cf. the reference's example/keras/keras_synthetic_benchmark_tf2.py):
the keras `DistributedOptimizer` plus a timing callback that reports img/sec the way the other benchmarks do.
Needs `tensorflow`.

    bpslaunch python examples/keras/keras_synthetic_benchmark_tf2.py --model ResNet50 --batch-size 32
"""
import argparse
import os
import sys
from timeit import default_timer as timer

import numpy as np
import tensorflow as tf

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import byteps_b200.tensorflow.keras as bps  # noqa: E402
from byteps_b200.tensorflow.keras import callbacks as bcb  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--model", default="ResNet50")
p.add_argument("--batch-size", type=int, default=32)
p.add_argument("--fp16-pushpull", action="store_true")
p.add_argument("--num-warmup-batches", type=int, default=2)
p.add_argument("--num-batches-per-iter", type=int, default=10)
p.add_argument("--num-iters", type=int, default=10)
args = p.parse_args()

bps.init()
gpus = tf.config.experimental.list_physical_devices("GPU")
for g in gpus:
    tf.config.experimental.set_memory_growth(g, True)
if gpus:
    tf.config.experimental.set_visible_devices(gpus[bps.local_rank() % len(gpus)], "GPU")

model = getattr(tf.keras.applications, args.model)(weights=None)
compression = bps.Compression.fp16 if args.fp16_pushpull else bps.Compression.none
opt = bps.DistributedOptimizer(tf.keras.optimizers.SGD(0.01), compression=compression)
# the wrapped optimizer exchanges gradients inside get_gradients/_aggregate_gradients: keep keras from
# re-aggregating them itself
model.compile(loss="sparse_categorical_crossentropy", optimizer=opt, experimental_run_tf_function=False)

data = tf.random.uniform([args.batch_size, 224, 224, 3])
target = tf.random.uniform([args.batch_size, 1], minval=0, maxval=999, dtype=tf.int64)
dataset = tf.data.Dataset.from_tensor_slices((data, target)).cache().repeat().batch(args.batch_size)


class TimingCallback(tf.keras.callbacks.Callback):
    def on_train_begin(self, logs=None):
        self.rates = []

    def on_epoch_begin(self, epoch, logs=None):
        self.t0 = timer()

    def on_epoch_end(self, epoch, logs=None):
        rate = args.batch_size * args.num_batches_per_iter / (timer() - self.t0)
        self.rates.append(rate)
        if bps.rank() == 0:
            print("Iter #%d: %.1f img/sec per GPU" % (epoch, rate))


timing = TimingCallback()
model.fit(dataset, steps_per_epoch=args.num_warmup_batches, epochs=1, verbose=0,
          callbacks=[bcb.BroadcastGlobalVariablesCallback(0)])
model.fit(dataset, steps_per_epoch=args.num_batches_per_iter, epochs=args.num_iters, verbose=0, callbacks=[timing])
if bps.rank() == 0:
    m, ci = np.mean(timing.rates), 1.96 * np.std(timing.rates)
    print("Img/sec per GPU: %.1f +-%.1f; total on %d GPU(s): %.1f +-%.1f" % (m, ci, bps.size(), bps.size() * m,
                                                                             bps.size() * ci))
