"""DLPack/numpy plugin and training callbacks."""
import numpy as np
import torch

from _mp import run_workers


def _dlpack(rank, world):
    import byteps_b200.dlpack as bps

    bps.init()
    g = np.full(1000, float(rank + 1), dtype=np.float32)
    out = bps.push_pull_inplace(g, average=True, name="np.g")
    assert out is g and np.allclose(g, sum(range(1, world + 1)) / world)
    t = torch.arange(10, dtype=torch.float64) * (rank + 1)          # any __dlpack__ provider
    r = bps.push_pull(t, average=False, name="dl.t")
    assert torch.equal(r, torch.arange(10, dtype=torch.float64) * sum(range(1, world + 1)))
    b = np.arange(5, dtype=np.int64) + 100 * rank
    bps.broadcast(b, root_rank=1, name="np.b")
    assert b.tolist() == (np.arange(5) + 100).tolist()
    bps.shutdown()


def test_dlpack_plugin_two_processes():
    run_workers(_dlpack, world=2)


def _callbacks(rank, world):
    import byteps_b200.torch as bps
    from byteps_b200.torch.callbacks import (BroadcastGlobalVariablesCallback, LearningRateScheduleCallback,
                                            LearningRateWarmupCallback, MetricAverageCallback)

    bps.init()
    torch.manual_seed(rank)
    m = torch.nn.Linear(3, 2)
    opt = torch.optim.SGD(m.parameters(), lr=0.4, momentum=0.9)
    BroadcastGlobalVariablesCallback(m, opt).on_train_begin()
    w = m.weight.detach().clone()
    ws = [torch.zeros_like(w) for _ in range(world)]
    import torch.distributed as dist

    dist.all_gather(ws, w)
    assert all(torch.equal(x, ws[0]) for x in ws)
    logs = {"loss": float(rank), "acc": 1.0}
    MetricAverageCallback().on_epoch_end(0, logs)
    assert abs(logs["loss"] - sum(range(world)) / world) < 1e-6 and logs["acc"] == 1.0
    sched = LearningRateScheduleCallback(opt, multiplier=lambda e: 0.1 ** (e // 30), start_epoch=2)
    warm = LearningRateWarmupCallback(opt, warmup_epochs=2, steps_per_epoch=10)
    warm.on_epoch_begin(0)
    warm.on_batch_begin(0)
    assert abs(opt.param_groups[0]["lr"] - 0.4 / world * (0.1 * (world - 1) / 2 + 1)) < 1e-9
    sched.on_epoch_begin(30)
    assert abs(opt.param_groups[0]["lr"] - 0.04) < 1e-9
    bps.shutdown()


def test_callbacks_two_processes():
    run_workers(_callbacks, world=2)


def _mxnet_front_end(rank, world):
    import _fake_mxnet

    mx = _fake_mxnet.install()
    import byteps_b200.mxnet as bps

    bps.init()
    tot = sum(range(1, world + 1))
    # tensor API
    x = mx.nd.array(np.arange(12, dtype=np.float32).reshape(3, 4) * (rank + 1))
    bps.byteps_declare_tensor("mx.x")
    bps.byteps_push_pull(x, name="mx.x", is_average=False)
    assert np.allclose(x.asnumpy(), np.arange(12).reshape(3, 4) * tot)
    # broadcast_parameters: dict and ParameterDict
    params = {"b": mx.nd.array(np.full(3, 10.0 + rank)), "a": mx.nd.array(np.full(2, 20.0 + rank))}
    bps.broadcast_parameters(params, root_rank=1)
    assert np.allclose(params["a"].asnumpy(), 21.0) and np.allclose(params["b"].asnumpy(), 11.0)
    # DistributedOptimizer: averaged gradient, then the wrapped update
    opt = bps.DistributedOptimizer(mx.optimizer.SGD(learning_rate=0.5))
    w = mx.nd.array(np.zeros(4, dtype=np.float32))
    g = mx.nd.array(np.full(4, float(rank + 1), dtype=np.float32))
    opt.update(3, w, g, None)
    assert np.allclose(w.asnumpy(), -0.5 * tot / world)
    opt.set_learning_rate(0.25)
    assert opt.lr == 0.25           # __getattr__ forwards to the wrapped optimizer
    # gluon DistributedTrainer: parameters start from root's values, gradients are pre-scaled by
    # 1/(batch*workers) and summed
    pd = mx.gluon.ParameterDict()
    pd["w1"] = mx.gluon.Parameter("w1", np.full(5, 1.0 + rank, dtype=np.float32))
    pd["w0"] = mx.gluon.Parameter("w0", np.full(3, 7.0 * (rank + 1), dtype=np.float32))
    pd["frozen"] = mx.gluon.Parameter("frozen", np.ones(2, dtype=np.float32), grad_req="null")
    tr = bps.DistributedTrainer(pd, "sgd", {"learning_rate": 1.0}, root_rank=0)
    assert [p.name for p in tr._params] == ["frozen", "w0", "w1"]     # sorted, identical on all workers
    batch = 4
    pd["w1"]._grad[0][:] = np.full(5, 8.0 * (rank + 1), dtype=np.float32)
    pd["w0"]._grad[0][:] = np.full(3, 4.0, dtype=np.float32)
    tr.step(batch)
    assert np.allclose(pd["w1"].data().asnumpy(), 1.0 - 8.0 * tot / world / batch), pd["w1"].data().asnumpy()
    assert np.allclose(pd["w0"].data().asnumpy(), 7.0 - 1.0)
    # compression_params: momentum/wd are moved out of the optimizer into the adapters
    op = {"learning_rate": 0.1, "momentum": 0.9, "wd": 1e-4}
    kwargs, intra = bps.DistributedTrainer._register_compressor(
        op, {"compressor": "onebit", "ef": "vanilla", "momentum": "nesterov", "scaling": True, "fp16": True})
    assert kwargs == {"byteps_compressor_type": "onebit", "byteps_ef_type": "vanilla",
                      "byteps_momentum_type": "nesterov", "byteps_compressor_onebit_scaling": "True",
                      "byteps_momentum_mu": "0.9"}
    assert "momentum" not in op and "wd" not in op
    from byteps_b200.mxnet.compression import NagAdapter, WeightDecayMomentumAdapter

    assert isinstance(intra, NagAdapter) and isinstance(intra.compressor, WeightDecayMomentumAdapter)
    # adapters: small tensor -> explicit nesterov, no wd momentum
    gsmall = mx.nd.array(np.ones(4, dtype=np.float32))
    c, ctx = intra.compress(gsmall)
    out = intra.decompress(c, ctx, x=mx.nd.array(np.full(4, 100.0, dtype=np.float32)))
    # fp16 cast, + wd*x = 0.01, then nag: mom = mu*(0 + g), g += mom
    assert np.allclose(out.asnumpy(), (1.0 + 0.01) * 1.9, rtol=2e-3)
    bps.shutdown()


def test_mxnet_front_end_two_processes():
    run_workers(_mxnet_front_end, world=2)


def _tensorflow_front_end(rank, world):
    import _fake_tensorflow

    tf = _fake_tensorflow.install()
    import byteps_b200.tensorflow as bps
    import byteps_b200.tensorflow.keras as bkeras
    from byteps_b200.tensorflow.keras import callbacks as bcb

    bps.init()
    tot = sum(range(1, world + 1))
    x = tf.constant(np.arange(6, dtype=np.float32) * (rank + 1))
    assert np.allclose(bps.push_pull(x, name="tf.x").numpy(), np.arange(6) * tot / world)         # average
    assert np.allclose(bps.push_pull(x, op=bps.Sum, name="tf.xs").numpy(), np.arange(6) * tot)
    # the op vocabulary of the reference (ops.py:74-99): Adasum is a known name that no reduction runs
    assert bps.ReduceOps.Average is bps.Average and bps.handle_average_backwards_compatibility(None, None) is bps.Average
    import warnings as _w
    with _w.catch_warnings(record=True) as seen:
        _w.simplefilter("always")
        assert bps.handle_average_backwards_compatibility(None, False) is bps.Sum
    assert any(issubclass(x.category, DeprecationWarning) for x in seen)
    for bad in (lambda: bps.push_pull(x, op=bps.Adasum, name="tf.ada"),
                lambda: bps.DistributedOptimizer(tf.keras.optimizers.SGD(learning_rate=0.1), op=bps.Adasum)):
        try:
            bad()
            raise AssertionError("Adasum must be rejected")
        except ValueError as e:
            assert "Adasum" in str(e)
    assert bps.get_pushpull_speed is not None and len(bps.get_pushpull_speed()) == 2
    assert np.allclose(bps.push_pull(x, average=False, compression=bps.Compression.fp16, name="tf.xh").numpy(),
                       np.arange(6) * tot)
    try:
        bps.push_pull(x, average=True, op=bps.Sum)
        raise SystemExit("op and average together must be rejected")
    except ValueError:
        pass
    sp = tf.IndexedSlices(tf.constant(np.ones((2, 3), dtype=np.float32)), tf.constant(np.array([0, 2])), (4, 3))
    dense = bps.push_pull(sp, op=bps.Sum, name="tf.sparse").numpy()
    assert dense.shape == (4, 3) and np.allclose(dense[[0, 2]], world) and np.allclose(dense[[1, 3]], 0)
    # broadcast_variables
    vs = [tf.Variable(np.full(3, 5.0 + rank), name="v0"), tf.Variable(np.full(2, -1.0 * rank), name="v1")]
    bps.broadcast_variables(vs, root_rank=1)
    assert np.allclose(vs[0].numpy(), 6.0) and np.allclose(vs[1].numpy(), -1.0)
    # DistributedGradientTape
    w = tf.Variable(np.zeros(4), name="w")
    tape = bps.DistributedGradientTape(tf.GradientTape({id(w): tf.constant(np.full(4, rank + 1.0, dtype=np.float32))}))
    with tape:
        pass
    (g,) = tape.gradient(None, [w])
    assert np.allclose(g.numpy(), tot / world)
    # DistributedOptimizer: apply_gradients called directly (custom training loop)
    opt = bps.DistributedOptimizer(tf.keras.optimizers.SGD(learning_rate=0.5))
    opt.apply_gradients([(tf.constant(np.full(4, rank + 1.0, dtype=np.float32)), w), (None, vs[0])])
    assert np.allclose(w.numpy(), -0.5 * tot / world)
    # keras flavour: optimizer wrapper keeps the class name (so saved models restore), callbacks
    kopt = bkeras.DistributedOptimizer(tf.keras.optimizers.SGD(learning_rate=0.25, momentum=0.9))
    assert type(kopt).__name__ == "SGD" and kopt.lr.value == 0.25
    w2 = tf.Variable(np.zeros(2), name="w2")
    kopt.apply_gradients([(tf.constant(np.full(2, 2.0 * (rank + 1), dtype=np.float32)), w2)])
    assert np.allclose(w2.numpy(), -0.25 * 2.0 * tot / world)
    wrapped = bkeras.load_model("unused.h5")
    assert isinstance(wrapped["SGD"](learning_rate=0.1), tf.keras.optimizers.SGD)
    import types

    model = types.SimpleNamespace(optimizer=tf.keras.optimizers.SGD(learning_rate=0.4 , momentum=0.9),
                                  variables=[tf.Variable(np.full(2, float(rank)), name="mv")])
    cb = bcb.BroadcastGlobalVariablesCallback(0)
    cb.set_model(model)
    cb.on_batch_end(0)
    assert np.allclose(model.variables[0].numpy(), 0.0) and cb.broadcast_done
    logs = {"loss": float(rank), "name": "x"}
    m = bcb.MetricAverageCallback()
    m.set_model(model)
    m.on_epoch_end(0, logs)
    assert abs(logs["loss"] - sum(range(world)) / world) < 1e-6 and logs["name"] == "x"
    warm = bcb.LearningRateWarmupCallback(warmup_epochs=2, steps_per_epoch=10)
    warm.set_model(model)
    warm.on_train_begin()
    warm.on_epoch_begin(0)
    warm.on_batch_begin(0)
    want = 0.4 / world * (0.1 * (world - 1) / 2 + 1)
    assert abs(model.optimizer.lr.value - want) < 1e-9
    assert abs(model.optimizer.momentum.value - 0.9 * want / 0.4) < 1e-9      # momentum correction ...
    warm.on_batch_end(0)
    assert model.optimizer.momentum.value == 0.9                               # ... for one step only
    sched = bcb.LearningRateScheduleCallback(multiplier=lambda e: 0.1 ** (e // 30), start_epoch=2,
                                             momentum_correction=False)
    sched.set_model(model)
    sched.initial_lr = 0.4
    sched.on_epoch_begin(30)
    sched.on_batch_begin(0)
    assert abs(model.optimizer.lr.value - 0.04) < 1e-9
    bps.shutdown()


def test_tensorflow_front_end_two_processes():
    run_workers(_tensorflow_front_end, world=2)


def _tf_mirrored_strategy(rank, world):
    import _fake_tensorflow

    tf = _fake_tensorflow.install()
    import byteps_b200.tensorflow as bps
    from byteps_b200.tensorflow.distribute import BytepsAllReduce, BytepsCrossDeviceOps, MirroredStrategy

    bps.init()
    tot = sum(range(1, world + 1))
    # default: one replica (this process's device), push_pull all-reduce
    s1 = MirroredStrategy()
    assert s1.num_replicas_in_sync == 1 and isinstance(s1.cross_device_ops, BytepsAllReduce)
    with s1.scope():
        w = tf.Variable(np.full(3, float(rank)), name="w")
    s1.broadcast_variables([w], root_rank=1)
    assert np.allclose(w.numpy(), 1.0)
    per = s1.run(lambda x: x * (rank + 1.0), args=(tf.constant(np.arange(4, dtype=np.float32)),))
    assert np.allclose(s1.reduce(tf.distribute.ReduceOp.SUM, per).numpy(), np.arange(4) * tot)
    assert np.allclose(s1.reduce(tf.distribute.ReduceOp.MEAN, per).numpy(), np.arange(4) * tot / world)
    # two local replicas per process: local reduction first, then the exchange; mean over ALL replicas
    s2 = MirroredStrategy(devices=["/cpu:0", "/cpu:1"], cross_device_ops=BytepsCrossDeviceOps(num_packs=2))
    pr = tf.distribute.PerReplica([tf.constant(np.full(5, rank + 1.0, dtype=np.float32)),
                                   tf.constant(np.full(5, 10.0 * (rank + 1), dtype=np.float32))])
    assert np.allclose(s2.reduce(tf.distribute.ReduceOp.SUM, pr).numpy(), 11.0 * tot)
    assert np.allclose(s2.reduce(tf.distribute.ReduceOp.MEAN, pr).numpy(), 11.0 * tot / (2 * world))
    # batch_reduce: mixed shapes and dtypes packed into num_packs exchanges per dtype, and the unpacked flavour
    grads = [tf.distribute.PerReplica([tf.constant(np.full(sh, (rank + 1.0) * (i + 1), dtype=dt))] * 2)
             for i, (sh, dt) in enumerate([((2, 3), np.float32), ((4,), np.float32), ((), np.float32),
                                           ((3,), np.float64), ((2, 2), np.float32)])]
    for ops in (BytepsCrossDeviceOps(num_packs=2), BytepsCrossDeviceOps(num_packs=0), BytepsAllReduce()):
        s3 = MirroredStrategy(devices=["/cpu:0", "/cpu:1"], cross_device_ops=ops)
        outs = s3.batch_reduce(tf.distribute.ReduceOp.SUM, grads)
        for i, (o, g) in enumerate(zip(outs, grads)):
            assert o.shape == g.values[0].shape and o.dtype == g.values[0].dtype
            assert np.allclose(o.numpy(), 2.0 * tot * (i + 1)), (i, o.numpy())
    # strategy.gather: local replicas are concatenated first, then the workers' blocks in rank order
    s4 = MirroredStrategy(devices=["/cpu:0", "/cpu:1"])
    pr = tf.distribute.PerReplica([tf.constant(np.full((2, 3), 10.0 * rank, dtype=np.float32)),
                                   tf.constant(np.full((2, 3), 10.0 * rank + 1, dtype=np.float32))])
    got = s4.gather(pr, axis=0).numpy()
    want = np.concatenate([np.full((2, 3), 10.0 * r + j, dtype=np.float32) for r in range(world) for j in range(2)])
    assert got.shape == (4 * world, 3) and np.array_equal(got, want)
    try:
        BytepsAllReduce(num_packs=-1)
        raise SystemExit("negative num_packs must be rejected")
    except ValueError:
        pass
    bps.shutdown()


def test_tensorflow_mirrored_strategy_two_processes():
    run_workers(_tf_mirrored_strategy, world=2)
