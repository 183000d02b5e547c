#!/usr/bin/env python
"""tf.keras MNIST-style training with the distributed optimizer and callbacks (cf. the reference's
example/keras/keras_mnist.py and example/tensorflow/tensorflow_keras_mnist.py).  Needs `tensorflow`.

    bpslaunch python examples/keras/keras_mnist.py
"""
import os
import sys

import numpy as np
import tensorflow as tf

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import byteps_b200.tensorflow.keras as bps  # noqa: E402
from byteps_b200.tensorflow.keras import callbacks as bcb  # noqa: E402

bps.init()
rng = np.random.RandomState(bps.rank())
x = rng.rand(4096, 28, 28, 1).astype("float32")
y = (x.reshape(len(x), -1).sum(1) * 7).astype("int64") % 10

model = tf.keras.Sequential([
    tf.keras.layers.Conv2D(32, 3, activation="relu", input_shape=(28, 28, 1)),
    tf.keras.layers.MaxPooling2D(), tf.keras.layers.Flatten(),
    tf.keras.layers.Dense(128, activation="relu"), tf.keras.layers.Dense(10, activation="softmax")])
# scale the learning rate by the number of workers; warm it up from lr/size over the first epochs
opt = bps.DistributedOptimizer(tf.keras.optimizers.SGD(0.01 * bps.size(), momentum=0.9))
model.compile(loss="sparse_categorical_crossentropy", optimizer=opt, metrics=["accuracy"],
              experimental_run_tf_function=False)
callbacks = [
    bcb.BroadcastGlobalVariablesCallback(0),        # identical start on all workers
    bcb.MetricAverageCallback(),                     # before any metric-driven callback
    bcb.LearningRateWarmupCallback(warmup_epochs=2, steps_per_epoch=len(x) // 64, verbose=1),
]
if bps.rank() == 0:
    callbacks.append(tf.keras.callbacks.ModelCheckpoint("./checkpoint-{epoch}.h5"))
model.fit(x, y, batch_size=64, epochs=3, callbacks=callbacks, verbose=1 if bps.rank() == 0 else 0)
