"""Real multi-process / multi-GPU runs of the symmetric-memory path (>= 2 GPUs)."""
import pytest
import torch

from _mp import run_workers

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _pushpull(rank, world):
    import byteps_b200.torch as bps
    from byteps_b200.common import engine

    torch.cuda.set_device(rank)
    bps.init()
    assert engine().backend == "symm"
    for dt in (torch.float32, torch.bfloat16, torch.float16):
        for n in (7, 4096, 1_000_003, 8_000_000):
            x = (torch.arange(n, device="cuda") % 13).to(dt) * (rank + 1)
            x0 = x.clone()
            y = bps.push_pull(x, average=False, name="t%s_%d" % (str(dt)[6:], n))
            tot = sum(r + 1 for r in range(world))
            ref = ((torch.arange(n, device="cuda") % 13).float() * tot).to(dt)
            assert torch.equal(x, x0)
            assert torch.allclose(y.float(), ref.float(), rtol=1e-2 if dt != torch.float32 else 1e-6), (dt, n)
    # many small tensors issued back to back, synchronised at the end (priority order inside the flush)
    ts = [torch.full((1000 + i,), float(rank), device="cuda") for i in range(20)]
    hs = [bps.push_pull_async_inplace(t, average=True, name="many%d" % i, priority=-i) for i, t in enumerate(ts)]
    for h, t in zip(hs, ts):
        out = bps.synchronize(h)
        assert torch.allclose(out, torch.full_like(out, sum(range(world)) / world))
    # integer / CPU tensors fall back to the torch.distributed transport
    z = torch.ones(10, dtype=torch.int64, device="cuda") * (rank + 1)
    assert bps.push_pull(z, average=False, name="ints").tolist() == [sum(r + 1 for r in range(world))] * 10
    c = torch.ones(10) * (rank + 1)
    assert bps.push_pull(c, average=True, name="cpu").tolist() == [sum(r + 1 for r in range(world)) / world] * 10
    bps.shutdown()


def test_pushpull_two_gpus():
    run_workers(_pushpull, world=2, timeout=300)


def _optimizer(rank, world, fused, graph):
    import byteps_b200.torch as bps
    from byteps_b200.torch.graph import GraphedStep

    torch.cuda.set_device(rank)
    bps.init()
    torch.manual_seed(100 + rank)
    mk = lambda: torch.nn.Sequential(torch.nn.Linear(64, 128), torch.nn.ReLU(), torch.nn.Linear(128, 16)).cuda()  # noqa: E731
    model = mk()
    opt = bps.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9),
                                   named_parameters=model.named_parameters(), fused_update=fused)
    bps.broadcast_parameters(model.state_dict(), root_rank=0)
    if not fused:
        bps.broadcast_optimizer_state(opt, root_rank=0)
    ref = mk()
    ref.load_state_dict(model.state_dict())
    ropt = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9)
    torch.manual_seed(7)
    steps = 5
    xs = torch.randn(steps, world * 8, 64, device="cuda")
    ys = torch.randn(steps, world * 8, 16, device="cuda")
    sx = torch.empty(8, 64, device="cuda")
    sy = torch.empty(8, 16, device="cuda")

    def step():
        opt.zero_grad()
        loss = torch.nn.functional.mse_loss(model(sx), sy)
        loss.backward()
        opt.step()
        return loss

    runner = None
    for i in range(steps):
        sx.copy_(xs[i, rank * 8:(rank + 1) * 8])
        sy.copy_(ys[i, rank * 8:(rank + 1) * 8])
        if graph and i == 2:
            runner = GraphedStep(step, warmup=1, pre_replay=opt.refresh_hparams)   # consumes this batch eagerly
        elif runner is not None:
            runner()
        else:
            step()
        ropt.zero_grad()
        torch.nn.functional.mse_loss(ref(xs[i]), ys[i]).backward()
        ropt.step()
    torch.cuda.synchronize()
    for a, b in zip(model.parameters(), ref.parameters()):
        assert torch.allclose(a, b, atol=1e-4), (a - b).abs().max()
    bps.shutdown()


@pytest.mark.parametrize("fused,graph", [(False, False), (True, False), (True, True)])
def test_distributed_optimizer_two_gpus(fused, graph):
    run_workers(_optimizer, world=2, args=(fused, graph), timeout=300)
