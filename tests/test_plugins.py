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
