#!/usr/bin/env python
"""TF2 custom training loop with `DistributedGradientTape` (cf. the reference's
example/tensorflow/tensorflow2_mnist.py).  Needs `tensorflow`.

    bpslaunch python examples/tensorflow/tensorflow2_mnist.py
"""
import os
import sys

import numpy as np
import tensorflow as tf

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import byteps_b200.tensorflow as bps  # noqa: E402

bps.init()
gpus = tf.config.experimental.list_physical_devices("GPU")
for g in gpus:
    tf.config.experimental.set_memory_growth(g, True)
if gpus:
    tf.config.experimental.set_visible_devices(gpus[bps.local_rank() % len(gpus)], "GPU")

rng = np.random.RandomState(1234 + bps.rank())
x = rng.rand(8192, 28, 28, 1).astype("float32")
y = (x.reshape(len(x), -1).sum(1) * 7).astype("int64") % 10
dataset = tf.data.Dataset.from_tensor_slices((x, y)).repeat().shuffle(10000).batch(128)

model = tf.keras.Sequential([
    tf.keras.layers.Conv2D(32, [3, 3], activation="relu"), tf.keras.layers.Conv2D(64, [3, 3], activation="relu"),
    tf.keras.layers.MaxPooling2D(pool_size=(2, 2)), tf.keras.layers.Dropout(0.25), tf.keras.layers.Flatten(),
    tf.keras.layers.Dense(128, activation="relu"), tf.keras.layers.Dropout(0.5),
    tf.keras.layers.Dense(10, activation="softmax")])
loss_fn = tf.losses.SparseCategoricalCrossentropy()
opt = tf.optimizers.Adam(0.001 * bps.size())          # learning rate scaled by the number of workers
checkpoint = tf.train.Checkpoint(model=model, optimizer=opt)


@tf.function
def training_step(images, labels, first_batch):
    with tf.GradientTape() as tape:
        loss_value = loss_fn(labels, model(images, training=True))
    tape = bps.DistributedGradientTape(tape)           # gradients averaged over all workers
    grads = tape.gradient(loss_value, model.trainable_variables)
    opt.apply_gradients(zip(grads, model.trainable_variables))
    # broadcast AFTER the first step so that the optimizer's slot variables exist and are covered too
    if first_batch:
        bps.broadcast_variables(model.variables, root_rank=0)
        bps.broadcast_variables(opt.variables(), root_rank=0)
    return loss_value


for batch, (images, labels) in enumerate(dataset.take(2000 // bps.size())):
    loss_value = training_step(images, labels, batch == 0)
    if batch % 50 == 0 and bps.local_rank() == 0:
        print("Step #%d\tLoss: %.6f" % (batch, loss_value))
if bps.rank() == 0:     # only one worker writes checkpoints
    checkpoint.save("./checkpoints/ckpt")
