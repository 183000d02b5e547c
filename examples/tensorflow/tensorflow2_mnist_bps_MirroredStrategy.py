#!/usr/bin/env python
"""MNIST-style training under `byteps_b200.tensorflow.distribute.MirroredStrategy` (cf. the reference's
example/tensorflow/tensorflow2_mnist_bps_MirroredStrategy.py).  One process per GPU: the strategy mirrors the
model on this process's GPU and its cross-replica reduction ends in a push_pull over all workers.
Needs `tensorflow`.

    bpslaunch python examples/tensorflow/tensorflow2_mnist_bps_MirroredStrategy.py
"""
import os
import sys

import numpy as np
import tensorflow as tf

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import byteps_b200.tensorflow as bps  # noqa: E402
from byteps_b200.tensorflow.distribute import MirroredStrategy  # noqa: E402

bps.init()
gpus = tf.config.experimental.list_physical_devices("GPU")
if gpus:
    tf.config.experimental.set_visible_devices(gpus[bps.local_rank() % len(gpus)], "GPU")

BATCH = 64
rng = np.random.RandomState(bps.rank())
x = rng.rand(8192, 28, 28, 1).astype("float32")
y = (x.reshape(len(x), -1).sum(1) * 7).astype("int64") % 10

strategy = MirroredStrategy()       # devices: this process's GPU; cross_device_ops: BytepsAllReduce()
with strategy.scope():
    model = tf.keras.Sequential([
        tf.keras.layers.Conv2D(32, 3, activation="relu", input_shape=(28, 28, 1)),
        tf.keras.layers.Conv2D(64, 3, activation="relu"), tf.keras.layers.MaxPooling2D(),
        tf.keras.layers.Flatten(), tf.keras.layers.Dense(128, activation="relu"), tf.keras.layers.Dense(10)])
    opt = tf.keras.optimizers.SGD(0.01 * bps.size(), momentum=0.9)
    loss_obj = tf.keras.losses.SparseCategoricalCrossentropy(from_logits=True,
                                                             reduction=tf.keras.losses.Reduction.NONE)
strategy.broadcast_variables(model.variables, root_rank=0)      # identical start on all workers

# each worker reads its own shard of the data
ds = tf.data.Dataset.from_tensor_slices((x, y)).shard(bps.size(), bps.rank()).shuffle(4096).batch(BATCH).repeat()
dist_ds = strategy.experimental_distribute_dataset(ds)


def replica_step(images, labels):
    with tf.GradientTape() as tape:
        per_example = loss_obj(labels, model(images, training=True))
        # scale by the GLOBAL batch: the strategy SUMs gradients over every replica of every worker
        loss = tf.nn.compute_average_loss(per_example, global_batch_size=BATCH * bps.size())
    grads = tape.gradient(loss, model.trainable_variables)
    opt.apply_gradients(zip(grads, model.trainable_variables))       # all-reduce happens in here
    return loss


@tf.function
def train_step(images, labels):
    per_replica = strategy.run(replica_step, args=(images, labels))
    return strategy.reduce(tf.distribute.ReduceOp.SUM, per_replica, axis=None)


for step, (images, labels) in enumerate(dist_ds):
    loss = train_step(images, labels)
    if step % 50 == 0 and bps.rank() == 0:
        print("step %d  loss %.4f" % (step, float(loss)))
    if step >= 500 // bps.size():
        break
if bps.rank() == 0:
    model.save_weights("./mirrored-mnist.ckpt")
