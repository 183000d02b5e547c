"""BASELINE config 1: two-process CPU/gloo push-pull plumbing (runs without a GPU)."""
import pytest
import torch

from _mp import run_workers


def _pushpull_100mb(rank, world):
    import byteps_b200.torch as bps

    bps.init()
    assert bps.size() == world and bps.rank() == rank
    n = 25_000_000  # 100 MB of fp32
    g = torch.full((n,), float(rank + 1))
    h = bps.push_pull_async_inplace(g, average=True, name="grad100mb")
    out = bps.synchronize(h)
    expect = sum(range(1, world + 1)) / world
    assert out.data_ptr() == g.data_ptr()
    assert torch.all(out == expect)
    # out-of-place sum, several dtypes
    for dt in (torch.float64, torch.float16, torch.bfloat16, torch.int32, torch.int64, torch.uint8):
        x = torch.arange(1000).to(dt) if dt != torch.uint8 else torch.ones(1000, dtype=dt)
        x0 = x.clone()
        y = bps.push_pull(x, average=False, name="t_%s" % str(dt).split(".")[-1])
        assert torch.equal(y.double(), (x0.float() * world).to(dt).double()), dt
        assert torch.equal(x, x0)   # the input is not modified
    # integer average = floor divide
    z = torch.tensor([3, 5, 7], dtype=torch.int64) * (rank + 1)
    bps.push_pull_inplace(z, average=True, name="intavg")
    tot = sum(r + 1 for r in range(world))
    assert z.tolist() == [3 * tot // world, 5 * tot // world, 7 * tot // world]
    bps.shutdown()


def test_two_process_gloo_pushpull():
    run_workers(_pushpull_100mb, world=2)


def _optimizer(rank, world):
    import byteps_b200.torch as bps

    bps.init()
    torch.manual_seed(1234 + rank)   # different init per rank: broadcast must fix it
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)
    opt = bps.DistributedOptimizer(opt, named_parameters=model.named_parameters())
    bps.broadcast_parameters(model.state_dict(), root_rank=0)
    bps.broadcast_optimizer_state(opt, root_rank=0)
    ref = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    ref.load_state_dict(model.state_dict())
    ref_opt = torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.9)
    torch.manual_seed(99)
    xs = [torch.randn(world * 4, 8) for _ in range(3)]
    ys = [torch.randn(world * 4, 4) for _ in range(3)]
    for x, y in zip(xs, ys):
        # distributed: each rank sees its slice; reference: the full batch
        opt.zero_grad()
        xl, yl = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]
        torch.nn.functional.mse_loss(model(xl), yl).backward()
        opt.step()
        ref_opt.zero_grad()
        torch.nn.functional.mse_loss(ref(x), y).backward()
        ref_opt.step()
    for a, b in zip(model.parameters(), ref.parameters()):
        assert torch.allclose(a, b, atol=1e-5), (a - b).abs().max()
    obj = bps.broadcast_object({"lr": 0.5, "rank": rank}, root_rank=0)
    assert obj == {"lr": 0.5, "rank": 0}
    bps.shutdown()


def test_distributed_optimizer_matches_full_batch():
    run_workers(_optimizer, world=2)


def _accumulate(rank, world):
    import byteps_b200.torch as bps

    bps.init()
    torch.manual_seed(0)
    model = torch.nn.Linear(4, 2)
    opt = bps.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=1.0),
                                   named_parameters=model.named_parameters(), backward_passes_per_step=2)
    bps.broadcast_parameters(model.state_dict(), 0)
    w0 = model.weight.detach().clone()
    opt.zero_grad()
    for i in range(2):
        x = torch.full((1, 4), float(rank + 1 + i))
        model(x).sum().backward()
    opt.step()
    # d/dw sum(Wx+b) = x for each output row; accumulated over 2 passes, averaged over ranks
    expect = sum((r + 1) + (r + 2) for r in range(world)) / world
    assert torch.allclose(w0 - model.weight, torch.full_like(w0, expect))
    # skip_synchronize after an explicit synchronize
    opt.zero_grad()
    for i in range(2):
        model(torch.ones(1, 4)).sum().backward()
    opt.synchronize()
    with opt.skip_synchronize():
        opt.step()
    bps.shutdown()


def test_backward_passes_per_step_and_skip_synchronize():
    run_workers(_accumulate, world=2)


def _cross_barrier_generic(rank, world, optim):
    """CrossBarrier on the generic (per-parameter) path: step() does not wait, forward pre-hooks
    finish exactly their module's parameters; result == the same optimizer on the full batch."""
    import byteps_b200.torch as bps
    from byteps_b200.torch.cross_barrier import CrossBarrier

    bps.init()

    def make():
        torch.manual_seed(7)
        return torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.Tanh(), torch.nn.Linear(8, 3))

    def mk(params):
        if optim == "rmsprop":
            return torch.optim.RMSprop(params, lr=0.01, momentum=0.5)
        if optim == "adam":
            return torch.optim.Adam(params, lr=0.01)
        return torch.optim.SGD(params, lr=0.1, momentum=0.9, nesterov=True)

    model, ref = make(), make()
    steps = 5
    opt = CrossBarrier(model, mk(model.parameters()), model.named_parameters(), num_steps=steps)
    assert opt._generic
    bps.broadcast_parameters(model.state_dict(), root_rank=0)
    ref_opt = mk(ref.parameters())
    gen = torch.Generator().manual_seed(3)
    xs = [torch.randn(world * 4, 6, generator=gen) for _ in range(steps)]
    ys = [torch.randn(world * 4, 3, generator=gen) for _ in range(steps)]
    for it in range(steps):
        if it == 2:                                    # an lr change on the wrapped optimizer reaches the updates
            for g in opt.param_groups:
                g["lr"] *= 0.5
            for g in ref_opt.param_groups:
                g["lr"] *= 0.5
        x, y = xs[it][rank * 4:(rank + 1) * 4], ys[it][rank * 4:(rank + 1) * 4]
        opt.zero_grad()
        torch.nn.functional.mse_loss(model(x), y).backward()
        opt.step()
        if 0 < it < steps - 1:
            assert opt._pending, "step() must not drain in the middle of training"
        ref_opt.zero_grad()
        torch.nn.functional.mse_loss(ref(xs[it]), ys[it]).backward()
        ref_opt.step()
    assert not opt._pending                            # the last step drains
    for a, b in zip(model.parameters(), ref.parameters()):
        assert torch.allclose(a, b, atol=2e-5), (optim, (a - b).abs().max())
    bps.shutdown()


@pytest.mark.parametrize("optim", ["sgd", "adam", "rmsprop"])
def test_cross_barrier_generic_path(optim):
    run_workers(_cross_barrier_generic, world=2, args=(optim,))


def _dynamic_loss_scale(rank, world):
    import byteps_b200.torch as bps
    from byteps_b200.torch.half_optimizer import HalfPrecisionDistributedOptimizer

    bps.init()
    torch.manual_seed(11)
    model = torch.nn.Linear(8, 4).to(torch.bfloat16)
    opt = HalfPrecisionDistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.5, momentum=0.9),
                                            model.named_parameters(), loss_scale=1024.0, dynamic_loss_scale=True,
                                            scale_window=2)
    bps.broadcast_parameters(model.state_dict(), root_rank=0)
    for p, m in opt.master_params().items():
        m.data.copy_(p.data.float())                       # masters follow the broadcast values
    ref_w = {n: p.detach().float().clone() for n, p in model.named_parameters()}
    ref_buf = {n: torch.zeros_like(v) for n, v in ref_w.items()}
    gen = torch.Generator().manual_seed(5)
    xs = [torch.randn(world * 2, 8, generator=gen) for _ in range(4)]
    for it in range(4):
        x = xs[it][rank * 2:(rank + 1) * 2].to(torch.bfloat16)
        if it == 1:
            x = x * float("inf")                           # overflow on every rank: this step must be skipped
        opt.zero_grad()
        scale_before = opt.loss_scale
        # local scaled gradients from a hook-free copy (model.grad is reduced in place, asynchronously)
        import copy

        twin = copy.deepcopy(model)
        gl = torch.autograd.grad(twin(x).float().square().mean() * scale_before, list(twin.parameters()))
        grads = {n: g.detach().float() for (n, _), g in zip(model.named_parameters(), gl)}
        loss = model(x).float().square().mean()
        opt.backward(loss)
        before = [p.detach().clone() for p in model.parameters()]
        opt.step()
        if it == 1:
            assert opt.skipped_steps == 1 and opt.loss_scale == scale_before / 2
            assert all(torch.equal(a, b) for a, b in zip(before, model.parameters()))
            continue
        # reference: SGD+momentum in fp32 on the rank-averaged, unscaled gradients
        for n in ref_w:
            g = grads[n] / scale_before
            gs = [torch.zeros_like(g) for _ in range(world)]
            torch.distributed.all_gather(gs, g)
            g = torch.stack(gs).mean(0)
            ref_buf[n] = g.clone() if ref_buf[n].abs().sum() == 0 and it == 0 else 0.9 * ref_buf[n] + g
            ref_w[n] = ref_w[n] - 0.5 * ref_buf[n]
    masters = {n: opt.master_params()[p] for n, p in model.named_parameters()}
    for n in ref_w:
        assert torch.allclose(masters[n], ref_w[n], atol=2e-2, rtol=2e-2), (n, (masters[n] - ref_w[n]).abs().max())
        assert torch.equal(dict(model.named_parameters())[n].detach(), masters[n].to(torch.bfloat16))
    # 1024 -> 512 on the overflow, then doubled after two clean steps
    assert opt.loss_scale == 1024.0 and opt.skipped_steps == 1
    bps.shutdown()


def test_half_precision_dynamic_loss_scale():
    run_workers(_dynamic_loss_scale, world=2)


def _ddp_generic(rank, world):
    """DistributedDataParallel on the per-parameter path (gloo): averaged gradients with a plain optimizer,
    no_sync accumulation, buffer broadcast, fp16 wire format."""
    import byteps_b200.torch as bps
    from byteps_b200.torch.parallel import DistributedDataParallel as DDP

    bps.init()

    def make(seed):
        torch.manual_seed(seed)
        return torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.BatchNorm1d(8), torch.nn.ReLU(), torch.nn.Linear(8, 2))

    model = DDP(make(100 + rank))                       # different init per rank: the wrapper broadcasts rank 0's
    ref = make(100)
    for a, b in zip(model.module.state_dict().values(), ref.state_dict().values()):
        assert torch.equal(a, b)
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)
    ropt = torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.9)
    gen = torch.Generator().manual_seed(9)
    for it in range(3):
        xs, ys = torch.randn(world * 4, 6, generator=gen), torch.randn(world * 4, 2, generator=gen)
        x, y = xs[rank * 4:(rank + 1) * 4], ys[rank * 4:(rank + 1) * 4]
        opt.zero_grad()
        torch.nn.functional.mse_loss(model(x), y).backward()
        opt.step()
    # BatchNorm uses per-rank batch statistics, so compare against the same computation done rank by rank
    # (gradients averaged by hand) instead of a full-batch forward
    gen = torch.Generator().manual_seed(9)
    replicas = [make(100) for _ in range(world)]
    ropts = [torch.optim.SGD(m.parameters(), lr=0.1, momentum=0.9) for m in replicas]
    for it in range(3):
        xs, ys = torch.randn(world * 4, 6, generator=gen), torch.randn(world * 4, 2, generator=gen)
        for r, m in enumerate(replicas):
            m.load_state_dict({k: (v if "running" not in k and "num_batches" not in k else replicas[0].state_dict()[k])
                               for k, v in m.state_dict().items()})      # buffers follow rank 0, like the wrapper
            ropts[r].zero_grad()
            torch.nn.functional.mse_loss(m(xs[r * 4:(r + 1) * 4]), ys[r * 4:(r + 1) * 4]).backward()
        for ps in zip(*[list(m.parameters()) for m in replicas]):
            g = torch.stack([p.grad for p in ps]).mean(0)
            for p in ps:
                p.grad = g.clone()
        for o in ropts:
            o.step()
    for a, b in zip(model.module.parameters(), replicas[rank].parameters()):
        assert torch.allclose(a, b, atol=1e-5), (a - b).abs().max()
    # no_sync: two local passes, exchanged by the first backward outside the context
    lin = DDP(torch.nn.Linear(3, 1, bias=False), compression=bps.Compression.fp16)
    with torch.no_grad():
        lin.module.weight.fill_(0.0)
    lin.zero_grad()
    with lin.no_sync():
        lin(torch.full((1, 3), float(rank + 1))).sum().backward()
    assert torch.equal(lin.module.weight.grad, torch.full((1, 3), float(rank + 1)))      # local only so far
    lin(torch.full((1, 3), 10.0 * (rank + 1))).sum().backward()
    want = sum(11.0 * (r + 1) for r in range(world)) / world
    assert torch.allclose(lin.module.weight.grad, torch.full((1, 3), want), rtol=1e-3)
    bps.shutdown()
    del ref, ropt


def test_ddp_generic_path():
    run_workers(_ddp_generic, world=2)


def _misc_api(rank, world):
    import byteps_b200.torch as bps

    bps.init()
    # --- broadcast_optimizer_state: Adam moments, step counters and hyper-parameters follow the root
    torch.manual_seed(rank)
    m = torch.nn.Linear(4, 3)
    opt = torch.optim.Adam(m.parameters(), lr=0.01 * (rank + 1), betas=(0.8 + 0.01 * rank, 0.99))
    for _ in range(rank + 1):                              # ranks took a different number of steps
        opt.zero_grad()
        m(torch.randn(2, 4)).sum().backward()
        opt.step()
    bps.broadcast_parameters(m.state_dict(), root_rank=0)
    bps.broadcast_optimizer_state(opt, root_rank=0)
    sd = opt.state_dict()
    flat = [float(sd["param_groups"][0]["lr"]), float(sd["param_groups"][0]["betas"][0])]
    for st in sd["state"].values():
        flat += [float(st["step"])] + st["exp_avg"].flatten().tolist() + st["exp_avg_sq"].flatten().tolist()
    t = torch.tensor(flat, dtype=torch.float64)
    gathered = [torch.zeros_like(t) for _ in range(world)]
    torch.distributed.all_gather(gathered, t)
    assert all(torch.equal(gathered[0], g) for g in gathered)
    assert abs(flat[0] - 0.01) < 1e-12 and abs(flat[1] - 0.8) < 1e-12 and flat[2] == 1.0
    # --- push_pull is differentiable: backward is a push_pull of the incoming gradient
    x = torch.full((3,), float(rank + 1), requires_grad=True)
    y = bps.push_pull(x, average=False, name="ad.x")        # y = sum_r x_r
    (y * torch.tensor([1.0, 2.0, 3.0])).sum().backward()
    assert torch.equal(y.detach(), torch.full((3,), float(sum(range(1, world + 1)))))
    assert torch.equal(x.grad, torch.tensor([1.0, 2.0, 3.0]) * world)
    # --- wire formats on the explicit-cast path
    z = torch.full((5,), 1.0 + rank)
    out = bps.push_pull(z, average=True, name="wire.bf16", compression=bps.Compression.bf16)
    assert out.dtype == torch.float32 and torch.allclose(out, torch.full((5,), sum(range(1, world + 1)) / world))
    # --- argument validation
    with pytest.raises(AssertionError):
        bps.push_pull(z)                                    # manual push_pull needs a name
    with pytest.raises(ValueError):
        bps.DistributedOptimizer(torch.optim.SGD(m.parameters(), lr=0.1),
                                 named_parameters=[("w", m.weight), ("w", m.bias)])
    with pytest.raises(ValueError):
        bps.DistributedOptimizer(torch.optim.SGD(m.parameters(), lr=0.1), named_parameters=[m.weight])
    # --- broadcast_object with an arbitrary picklable payload
    obj = bps.broadcast_object({"epoch": 3 + rank, "lr": [0.1, 0.01]}, root_rank=world - 1, name="ckpt.meta")
    assert obj == {"epoch": 3 + world - 1, "lr": [0.1, 0.01]}
    bps.shutdown()


def test_misc_public_api():
    run_workers(_misc_api, world=2)


def _many_tensors(rank, world):
    import byteps_b200.torch as bps

    bps.init()
    ts = [torch.full((1 + (i * 37) % 3000,), float(rank + i)) for i in range(300)]
    hs = [bps.push_pull_async_inplace(t, average=False, name="many.%d" % i, priority=-(i % 7)) for i, t in enumerate(ts)]
    for i, (h, t) in enumerate(zip(hs, ts)):
        bps.synchronize(h)
        assert torch.all(t == sum(r + i for r in range(world))), (i, t[:2])
    assert bps.push_pull(torch.zeros(0), name="empty").numel() == 0
    one = torch.tensor([3.0 + rank])
    assert bps.push_pull(one, average=True, name="one").item() == sum(3.0 + r for r in range(world)) / world
    big = torch.arange(5_000_001, dtype=torch.float64) * (rank + 1)      # odd size, several partitions
    assert torch.equal(bps.push_pull(big, average=False, name="big"),
                       torch.arange(5_000_001, dtype=torch.float64) * sum(range(1, world + 1)))
    bps.shutdown()


def test_many_tensors_in_flight_and_edge_sizes():
    run_workers(_many_tensors, world=2)


def _unused_parameters(rank, world):
    """Parameters that get no gradient in an iteration (data-independent branch) contribute zeros - also after
    zero_grad(set_to_none=True) dropped their .grad - on the optimizer and the DDP per-parameter paths."""
    import byteps_b200.torch as bps
    from byteps_b200.torch.parallel import DistributedDataParallel as DDP

    bps.init()
    r, n = rank, world
    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Linear(4, 4); self.b = torch.nn.Linear(4, 4); self.c = torch.nn.Linear(4, 2)
        def forward(self, x, use_b):
            h = self.a(x)
            if use_b: h = self.b(h)
            return self.c(h)
    torch.manual_seed(0)
    m = Net()
    opt = bps.DistributedOptimizer(torch.optim.SGD(m.parameters(), lr=0.1), named_parameters=m.named_parameters())
    bps.broadcast_parameters(m.state_dict(), 0)
    for it in range(4):
        opt.zero_grad()
        m(torch.randn(3, 4), use_b=(it % 2 == 0)).sum().backward()     # branch b unused on odd iterations
        opt.step()
    w = torch.cat([p.detach().flatten() for p in m.parameters()])
    ws = [torch.zeros_like(w) for _ in range(n)]
    torch.distributed.all_gather(ws, w)
    assert all(torch.equal(ws[0], x) for x in ws)
    # DDP with an unused branch: rank-dependent usage would deadlock per-parameter schemes; same usage on all ranks works
    d = DDP(Net())
    o2 = torch.optim.SGD(d.parameters(), lr=0.1)
    for it in range(4):
        o2.zero_grad()
        d(torch.randn(3, 4), use_b=(it % 2 == 0)).sum().backward()
        d.synchronize()
        o2.step()
    w = torch.cat([p.detach().flatten() for p in d.parameters()])
    ws = [torch.zeros_like(w) for _ in range(n)]
    torch.distributed.all_gather(ws, w)
    assert all(torch.equal(ws[0], x) for x in ws)
    bps.shutdown()


def test_unused_parameters_on_per_parameter_paths():
    run_workers(_unused_parameters, world=2)


def _same_host_shm(rank, world, mode):
    import os

    os.environ["BYTEPS_HOST_SHM_REDUCE"] = mode
    import byteps_b200.torch as bps
    from byteps_b200.common import engine

    bps.init()
    eng = engine()
    assert eng.backend == "gloo"
    tot = sum(r + 1 for r in range(world))
    for it in range(3):
        hs = []
        for n, dt in ((3, torch.float32), (1_000_003, torch.float32), (65_537, torch.bfloat16), (4099, torch.float64),
                      (1000, torch.int32)):
            g = ((torch.arange(n) + it) % 5).to(dt) * (rank + 1)
            hs.append((bps.push_pull_async_inplace(g, average=False, name="shm_%d_%s" % (n, str(dt)[6:])), g, n, dt))
        for h, g, n, dt in hs:
            bps.synchronize(h)
            assert torch.equal(g, (((torch.arange(n) + it) % 5).double() * tot).to(dt)), (n, dt, it)
        a = torch.full((10_000,), float(rank + 1 + it))
        assert torch.allclose(bps.push_pull(a, average=True, name="shm_avg"), torch.full((10_000,), tot / world + it))
        i = torch.full((33,), rank + 3 * it, dtype=torch.int64)
        assert torch.equal(bps.push_pull(i, average=True, name="shm_iavg"),
                           torch.full((33,), (sum(range(world)) + 3 * it * world) // world, dtype=torch.int64))
    used = eng._hostshm is not None
    assert used == (mode != "0"), (mode, used)
    if used:        # the root heard one READY and one BCAST_READY per follower per exchange
        assert (eng._hostshm.signals_received() > 0) == (rank == world - 1)
    bps.shutdown()


@pytest.mark.parametrize("mode", ["auto", "0"])
def test_same_host_cpu_job_reduces_through_shared_memory(mode):
    """All ranks on one host, CPU tensors: the sum goes through shared-memory slots + the CPU reducer
    (csrc/core/host_reduce.h) instead of gloo's ring over loopback sockets; BYTEPS_HOST_SHM_REDUCE=0 keeps gloo.
    100 MB, 2 processes in the build container: 83 -> 20 ms per push_pull."""
    run_workers(_same_host_shm, world=3, args=(mode,), timeout=240)


def _two_hosts_shm(rank, world):
    """world ranks = world / 2 'hosts' of 2 local ranks each (torchrun-style variables)."""
    import os

    os.environ.update({"LOCAL_RANK": str(rank % 2), "LOCAL_WORLD_SIZE": "2", "GROUP_RANK": str(rank // 2),
                       "BYTEPS_HOST_SHM_REDUCE": "auto"})
    import byteps_b200.torch as bps
    from byteps_b200.common import engine

    bps.init()
    eng = engine()
    assert eng.backend == "gloo" and bps.local_size() == 2 and bps.size() == world
    tot = sum(r + 1 for r in range(world))
    for it in range(3):
        hs = []
        for n, dt in ((5, torch.float32), (300_001, torch.float32), (70_001, torch.bfloat16), (1000, torch.int64)):
            g = ((torch.arange(n) + it) % 5).to(dt) * (rank + 1)
            hs.append((bps.push_pull_async_inplace(g, average=False, name="mh_%d_%s" % (n, str(dt)[6:])), g, n, dt))
        for h, g, n, dt in hs:
            bps.synchronize(h)
            assert torch.equal(g, (((torch.arange(n) + it) % 5).double() * tot).to(dt)), (n, dt, it)
        a = torch.full((4096,), float(rank + 1 + it))
        assert torch.allclose(bps.push_pull(a, average=True, name="mh_avg"), torch.full((4096,), tot / world + it))
    assert eng._hostshm is not None and eng._hostshm_roots is not None
    assert eng._hostshm.is_root() == (rank % 2 == 1)
    bps.shutdown()


def test_cpu_job_on_several_hosts_reduces_inside_each_host_first():
    """Two 'hosts' of two ranks: box sums through shared memory, the two roots all-reduce them over gloo (half the
    traffic of a flat ring), results copied out by the followers."""
    run_workers(_two_hosts_shm, world=4, timeout=240)


def _shm_setup_fails_on_one_rank(rank, world):
    import os

    if rank == 1:      # this rank cannot bind its datagram socket
        os.environ["BYTEPS_SOCKET_PATH"] = "/nonexistent/dir/for/sockets"
    import byteps_b200.torch as bps
    from byteps_b200.common import engine

    bps.init()
    g = torch.full((10_000,), float(rank + 1))
    bps.push_pull_inplace(g, average=False, name="fallback")
    assert torch.all(g == sum(r + 1 for r in range(world)))
    assert engine()._hostshm is None          # agreed on by all ranks: plain gloo
    bps.shutdown()


def test_shared_memory_reduction_falls_back_to_gloo_when_a_rank_cannot_set_it_up():
    run_workers(_shm_setup_fails_on_one_rank, world=2, timeout=120)
