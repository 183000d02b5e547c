"""init / shutdown / suspend / resume, checkpoint-style broadcast, error paths."""
import os

import pytest
import torch

from _mp import run_workers


def _suspend_resume(rank, world):
    import byteps_b200.torch as bps
    from byteps_b200.common import engine

    bps.init()
    for n in ("z", "a", "m"):
        bps.declare(n)
    x = torch.ones(10) * (rank + 1)
    assert bps.push_pull(x, name="a", average=False)[0].item() == sum(r + 1 for r in range(world))
    order0 = engine().registry.declared_names()
    bps.suspend()
    with pytest.raises(ValueError):
        bps.rank()
    bps.resume(int(os.environ.get("DMLC_NUM_WORKER", world)), 0)
    assert engine().registry.declared_names()[:len(order0)] == order0     # keys stay stable
    assert bps.push_pull(x, name="a", average=True)[0].item() == sum(r + 1 for r in range(world)) / world
    # checkpoint pattern: only rank 0 "loads", everyone ends up identical (model + optimizer state)
    torch.manual_seed(rank)
    model = torch.nn.Linear(5, 3)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    model(torch.randn(2, 5)).sum().backward()
    opt.step()
    bps.broadcast_parameters(model.state_dict(), root_rank=0)
    bps.broadcast_optimizer_state(opt, root_rank=0)
    flat = torch.cat([p.detach().flatten() for p in model.parameters()] +
                     [opt.state[p]["exp_avg"].flatten() for p in model.parameters()])
    ref = flat.clone()
    import torch.distributed as dist

    dist.broadcast(ref, src=0)
    assert torch.equal(flat, ref)
    bps.shutdown()
    bps.shutdown()       # idempotent


def test_suspend_resume_and_checkpoint_broadcast():
    run_workers(_suspend_resume, world=2)


def test_errors_before_init_and_bad_inputs():
    import byteps_b200.torch as bps

    with pytest.raises(ValueError):
        bps.size()
    bps.init()
    try:
        with pytest.raises(AssertionError):
            bps.push_pull(torch.ones(3))                       # name is mandatory for push_pull
        with pytest.raises(ValueError):
            bps.push_pull_inplace(torch.ones(4, 4).t(), name="nc")   # non-contiguous
        with pytest.raises(ValueError):
            bps.push_pull_inplace(torch.ones(3, dtype=torch.complex64), name="cplx")
        m = torch.nn.Linear(2, 2)
        with pytest.raises(ValueError):
            bps.DistributedOptimizer(torch.optim.SGD(m.parameters(), lr=0.1), named_parameters=[("w", m.weight),
                                                                                               ("w", m.bias)])
        assert bps.poll(12345) is True and bps.synchronize(12345) is None     # unknown handles
        assert bps.get_pushpull_speed()[1] == -5.0 or bps.get_pushpull_speed()[1] > 0
    finally:
        bps.shutdown()


def test_extension_loader_helpers():
    """get_ext_suffix / get_extension_full_path / check_extension of the reference's common module
    (byteps/common/__init__.py:26-50) against this package's layout."""
    import byteps_b200
    from byteps_b200.common import check_extension, get_ext_suffix, get_extension_full_path

    assert get_ext_suffix().endswith(".so")
    path = get_extension_full_path(byteps_b200.__file__, "_core")
    assert os.path.basename(path).startswith("_core.") and os.path.exists(path)
    check_extension("byteps_b200._core", "BYTEPS_WITHOUT_X", byteps_b200.__file__, "_core")
    with pytest.raises(ImportError, match="has not been built"):
        check_extension("byteps_b200.nope", "BYTEPS_WITHOUT_X", byteps_b200.__file__, "sub", "nope")
