"""CPU-server mode through the public API: scheduler + server processes (the
`import byteps_b200.server` entry point) and two worker processes on 127.0.0.1."""
import os
import subprocess
import sys

import pytest
import torch

from _mp import free_port, run_workers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _spawn_role(role, port, nw, ns, extra=None):
    env = dict(os.environ)
    env.update({"DMLC_ROLE": role, "DMLC_NUM_WORKER": str(nw), "DMLC_NUM_SERVER": str(ns),
                "DMLC_PS_ROOT_URI": "127.0.0.1", "DMLC_PS_ROOT_PORT": str(port), "PYTHONPATH": ROOT})
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    env.update(extra or {})
    return subprocess.Popen([sys.executable, "-c", "import byteps_b200.server"], env=env)


def _worker(rank, world, ps_port, compress):
    # BytePS-style env: every process is its own worker box with one (CPU) device
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE"):
        os.environ.pop(k, None)
    os.environ.update({"DMLC_ROLE": "worker", "DMLC_NUM_WORKER": str(world), "DMLC_NUM_SERVER": "1",
                       "DMLC_WORKER_ID": str(rank), "BYTEPS_LOCAL_RANK": "0", "BYTEPS_LOCAL_SIZE": "1",
                       "DMLC_PS_ROOT_URI": "127.0.0.1", "DMLC_PS_ROOT_PORT": str(ps_port),
                       "BYTEPS_MIN_COMPRESS_BYTES": "0"})
    import byteps_b200.torch as bps
    from byteps_b200.common import engine

    bps.init()
    assert engine().backend == "ps" and bps.size() == world and bps.rank() == rank
    # BASELINE config 1 through the server: 100 MB gradient, 25 partitions of 4 MB
    g = torch.full((25_000_000,), float(rank + 1))
    out = bps.push_pull_inplace(g, average=True, name="grad100mb")
    assert torch.all(out == sum(range(1, world + 1)) / world)
    # several tensors in flight, mixed dtypes
    ts = [torch.arange(1000 + i, dtype=torch.float32) * (rank + 1) for i in range(5)]
    hs = [bps.push_pull_async_inplace(t, average=False, name="t%d" % i, priority=-i) for i, t in enumerate(ts)]
    for i, (h, t) in enumerate(zip(hs, ts)):
        o = bps.synchronize(h)
        assert torch.equal(o, torch.arange(1000 + i, dtype=torch.float32) * sum(r + 1 for r in range(world)))
    # the same name with another size / dtype is a new key generation, not a silent partial sum
    # (server keys are sized by their init push; round 1 summed the first 10 elements and left the rest local)
    for n, dt in ((10, torch.float32), (50, torch.float32), (10, torch.float32), (50, torch.float64)):
        x = torch.arange(n, dtype=dt) * (rank + 1)
        bps.push_pull_inplace(x, average=False, name="resized")
        assert torch.equal(x, torch.arange(n, dtype=dt) * sum(r + 1 for r in range(world))), (n, dt)
    # two dict broadcasts share the default name (type(obj).__name__): sizes differ, both must arrive intact
    small = bps.broadcast_object({"a": 1} if rank == 0 else None, 0)
    big = bps.broadcast_object({"k%d" % i: list(range(i)) for i in range(64)} if rank == 0 else None, 0)
    assert small == {"a": 1} and big["k63"] == list(range(63)) and len(big) == 64
    z = torch.tensor([7, 9], dtype=torch.int64) * (rank + 1)
    bps.push_pull_inplace(z, average=True, name="ints")
    tot = sum(r + 1 for r in range(world))
    assert z.tolist() == [7 * tot // world, 9 * tot // world]
    if compress:
        ps = engine()._ps
        ps.set_compression("Gradient.c", {"compressor_type": "topk", "compressor_k": 100})
        x = torch.zeros(10000)
        x[rank * 50:(rank + 1) * 50] = 5.0 + rank
        bps.push_pull_inplace(x, average=False, name="Gradient.c")
        assert x[:50].eq(5.0).all() and x[50:100].eq(6.0).all() and x[100:].eq(0).all()
        # same thing through declare(**kwargs) and DistributedOptimizer(compression_params=...)
        bps.declare("Gradient.d", byteps_compressor_type="topk", byteps_compressor_k=100)
        y = torch.zeros(10000)
        y[rank * 50:(rank + 1) * 50] = 1.0 + rank
        bps.push_pull_inplace(y, average=False, name="Gradient.d")
        assert y[:50].eq(1.0).all() and y[50:100].eq(2.0).all() and y[100:].eq(0).all()
        torch.manual_seed(0)
        mc = torch.nn.Linear(64, 32, bias=False)
        oc = bps.DistributedOptimizer(torch.optim.SGD(mc.parameters(), lr=0.5, momentum=0.9),
                                      named_parameters=[("cw", mc.weight)],
                                      compression_params={"compressor": "onebit", "scaling": True, "ef": "vanilla"})
        w_before = mc.weight.detach().clone()
        mc(torch.ones(1, 64) * (rank + 1)).sum().backward()
        oc.step()
        # identical inputs up to scale -> every gradient entry is positive; signSGD with scaling keeps the
        # sign and the mean magnitude (worker stage then server stage), so all weights move down equally
        delta = w_before - mc.weight.detach()
        assert (delta > 0).all() and torch.allclose(delta, delta.flatten()[0].expand_as(delta))
    # optimizer end to end
    torch.manual_seed(rank)
    m = torch.nn.Linear(4, 2)
    opt = bps.DistributedOptimizer(torch.optim.SGD(m.parameters(), lr=0.1), named_parameters=m.named_parameters())
    bps.broadcast_parameters(m.state_dict(), 0)
    w0 = m.weight.detach().clone()
    m(torch.full((1, 4), float(rank + 1))).sum().backward()
    opt.step()
    assert torch.allclose(w0 - m.weight, torch.full_like(w0, 0.1 * tot / world))
    assert bps.get_pushpull_speed()[1] is not None
    bps.shutdown()


def _run(compress, env=None):
    port = free_port()
    procs = [_spawn_role("scheduler", port, 2, 1, env), _spawn_role("server", port, 2, 1, env)]
    try:
        run_workers(_worker, world=2, args=(port, compress), env=env, timeout=180)
        for p in procs:
            p.wait(timeout=60)
        assert all(p.returncode == 0 for p in procs)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()


def test_public_api_cpu_server_mode():
    _run(False)


def test_public_api_cpu_server_mode_with_topk():
    _run(True)


def test_public_api_cpu_server_mode_over_shm_van():
    """Same job with DMLC_PS_VAN_TYPE=shm: no sockets between scheduler, server and workers."""
    _run(True, {"DMLC_PS_VAN_TYPE": "shm"})


def _async_worker(rank, world, ps_port):
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE"):
        os.environ.pop(k, None)
    os.environ.update({"DMLC_ROLE": "worker", "DMLC_NUM_WORKER": str(world), "DMLC_NUM_SERVER": "1",
                       "DMLC_WORKER_ID": str(rank), "BYTEPS_LOCAL_RANK": "0", "BYTEPS_LOCAL_SIZE": "1",
                       "DMLC_PS_ROOT_URI": "127.0.0.1", "DMLC_PS_ROOT_PORT": str(ps_port),
                       "BYTEPS_ENABLE_ASYNC": "1"})
    import time

    import byteps_b200.torch as bps

    bps.init()
    torch.manual_seed(0)                                   # identical initial weights on every worker
    model = torch.nn.Linear(5, 3)
    opt = bps.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.1),
                                   named_parameters=model.named_parameters())
    w0 = {n: p.detach().clone() for n, p in model.named_parameters()}
    steps = 4
    total = {n: torch.zeros_like(p) for n, p in model.named_parameters()}
    for it in range(steps):
        opt.zero_grad()
        x = torch.full((2, 5), float(rank + 1))
        model(x).sum().backward()                          # gradients do not depend on the weights
        for n, p in model.named_parameters():
            total[n] += p.grad
        opt.step()                                         # no barrier with the other worker
    time.sleep(1.5)                                        # everybody's deltas have reached the server
    for n, p in model.named_parameters():
        z = torch.zeros_like(p)
        bps.push_pull_inplace(z, average=False, name="AsyncParam." + n)   # zero delta = read the server copy
        # server copy = initial weights + every worker's deltas.  d(sum)/dW = sum of the rows of x = 2 (r + 1) for
        # worker r, d(sum)/db = 2 for everybody: scale my own accumulated gradient accordingly
        if n == "weight":
            g_all = sum(total[n] * (r + 1) / (rank + 1) for r in range(world))
        else:
            g_all = total[n] * world
        assert torch.allclose(z, w0[n] - 0.1 * g_all, atol=1e-5), (n, z, w0[n] - 0.1 * g_all)
    bps.shutdown()


def test_async_mode_through_public_api():
    """BYTEPS_ENABLE_ASYNC=1: the server copy is seeded with the weights and accumulates weight deltas."""
    port = free_port()
    procs = [_spawn_role("scheduler", port, 2, 1, {"BYTEPS_ENABLE_ASYNC": "1"}),
             _spawn_role("server", port, 2, 1, {"BYTEPS_ENABLE_ASYNC": "1"})]
    try:
        run_workers(_async_worker, world=2, args=(port,), timeout=120)
        for p in procs:
            p.wait(timeout=60)
        assert all(p.returncode == 0 for p in procs)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()


def _elastic_worker(rank, world, port_a, port_b):
    """suspend(), then resume() against a NEW scheduler/server set with a different number of servers:
    declared names keep their keys, traffic flows through the new servers."""
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE"):
        os.environ.pop(k, None)
    os.environ.update({"DMLC_ROLE": "worker", "DMLC_NUM_WORKER": "1", "DMLC_NUM_SERVER": "1", "DMLC_WORKER_ID": "0",
                       "BYTEPS_LOCAL_RANK": "0", "BYTEPS_LOCAL_SIZE": "1", "BYTEPS_FORCE_DISTRIBUTED": "1",
                       "DMLC_PS_ROOT_URI": "127.0.0.1", "DMLC_PS_ROOT_PORT": str(port_a)})
    import byteps_b200.torch as bps
    from byteps_b200.common import engine

    bps.init()
    assert engine().backend == "ps"
    for n in ("w2", "w0", "w1"):
        bps.declare(n)
    t = torch.arange(100_000, dtype=torch.float32)
    assert torch.equal(bps.push_pull(t, average=False, name="w1"), t)
    names = engine().registry.declared_names()
    bps.suspend()                                           # the first cluster winds down
    os.environ["DMLC_PS_ROOT_PORT"] = str(port_b)           # a new scheduler with two servers is waiting
    bps.resume(1, 2)
    assert engine().backend == "ps" and engine().registry.declared_names()[:len(names)] == names
    for i in range(3):
        assert torch.equal(bps.push_pull(t * (i + 1), average=True, name="w%d" % i), t * (i + 1))
    bps.shutdown()


def test_elastic_suspend_resume_against_new_cluster():
    pa, pb = free_port(), free_port()
    extra = {"BYTEPS_FORCE_DISTRIBUTED": "1"}
    procs = [_spawn_role("scheduler", pa, 1, 1, extra), _spawn_role("server", pa, 1, 1, extra),
             _spawn_role("scheduler", pb, 1, 2, extra), _spawn_role("server", pb, 1, 2, extra),
             _spawn_role("server", pb, 1, 2, extra)]
    try:
        run_workers(_elastic_worker, world=1, args=(pa, pb), timeout=120)
        for p in procs:
            p.wait(timeout=60)
        assert all(p.returncode == 0 for p in procs)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()


def _many_ps_worker(rank, world, ps_port):
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE"):
        os.environ.pop(k, None)
    os.environ.update({"DMLC_ROLE": "worker", "DMLC_NUM_WORKER": str(world), "DMLC_NUM_SERVER": "2",
                       "DMLC_WORKER_ID": str(rank), "BYTEPS_LOCAL_RANK": "0", "BYTEPS_LOCAL_SIZE": "1",
                       "DMLC_PS_ROOT_URI": "127.0.0.1", "DMLC_PS_ROOT_PORT": str(ps_port)})
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_gloo_pushpull import _many_tensors

    _many_tensors(rank, world)


def test_many_tensors_in_flight_through_servers():
    port = free_port()
    procs = [_spawn_role("scheduler", port, 2, 2), _spawn_role("server", port, 2, 2), _spawn_role("server", port, 2, 2)]
    try:
        run_workers(_many_ps_worker, world=2, args=(port,), timeout=180)
        for p in procs:
            p.wait(timeout=60)
        assert all(p.returncode == 0 for p in procs)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()


def _host_hier_worker(rank, world, ps_port, local_size, ipc):
    """`world` processes = world / local_size boxes of `local_size` local ranks, CPU tensors."""
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE"):
        os.environ.pop(k, None)
    box, lr = divmod(rank, local_size)
    os.environ.update({"DMLC_ROLE": "worker", "DMLC_NUM_WORKER": str(world // local_size), "DMLC_NUM_SERVER": "1",
                       "DMLC_WORKER_ID": str(box), "BYTEPS_LOCAL_RANK": str(lr), "BYTEPS_LOCAL_SIZE": str(local_size),
                       "DMLC_PS_ROOT_URI": "127.0.0.1", "DMLC_PS_ROOT_PORT": str(ps_port),
                       "BYTEPS_FORCE_DISTRIBUTED": "1", "BYTEPS_PARTITION_BYTES": "400000",
                       "BYTEPS_ENABLE_IPC": "1" if ipc else "0"})
    if ipc:      # auto keeps the flat path next to a colocated IPC server: ask for the box-local reduction
        os.environ["BYTEPS_PS_HOST_HIERARCHICAL"] = "1"
    import byteps_b200.torch as bps
    from byteps_b200.common import engine

    bps.init()
    eng = engine()
    assert eng.backend == "ps" and bps.size() == world and bps.local_size() == local_size
    assert eng._ps._hr is not None and eng._ps._hr.is_root() == (lr == local_size - 1)
    tot = sum(r + 1 for r in range(world))
    for it in range(3):
        # several tensors in flight at once (the pool runs them concurrently), 1 and several partitions
        handles = []
        for n, dt in ((1000, torch.float32), (250_000, torch.float32), (70_000, torch.bfloat16), (5000, torch.float64)):
            g = ((torch.arange(n) + it) % 7).to(dt) * (rank + 1)
            handles.append((bps.push_pull_async_inplace(g, average=False, name="hh_%d_%s" % (n, str(dt)[6:])), g, n, dt))
        for h, g, n, dt in handles:
            bps.synchronize(h)
            ref = (((torch.arange(n) + it) % 7).double() * tot).to(dt)
            assert torch.equal(g, ref), (n, dt, it)
        a = torch.full((3000,), float(rank + 1 + it))
        out = bps.push_pull(a, average=True, name="hh_avg")
        assert torch.allclose(out, torch.full((3000,), tot / world + it)), it
        i = torch.full((777,), rank + 10 * it, dtype=torch.int64)
        out = bps.push_pull(i, average=True, name="hh_int")
        assert torch.equal(out, torch.full((777,), (sum(range(world)) + 10 * it * world) // world, dtype=torch.int64))
    # the local ranks really went through the box's root: it heard READY / BCAST signals, followers none
    got = eng._ps._hr.signals_received()
    assert (got > 0) == (lr == local_size - 1), got
    obj = bps.broadcast_object({"lr": 0.1, "step": 7} if rank == 0 else None, root_rank=0, name="hh_obj")
    assert obj == {"lr": 0.1, "step": 7}
    bps.shutdown()


@pytest.mark.parametrize("boxes,local_size,ipc", [(2, 2, True), (1, 3, False)])
def test_cpu_tensors_reduce_inside_the_box_first(boxes, local_size, ipc):
    """Several processes per box + CPU tensors: box-local reduction through shared memory with the reference's
    READY / DO_BROADCAST datagram protocol (csrc/core/host_reduce.h), one push per box to the servers."""
    port = free_port()
    extra = {"BYTEPS_LOCAL_SIZE": str(local_size), "BYTEPS_ENABLE_IPC": "1" if ipc else "0"}
    procs = [_spawn_role("scheduler", port, boxes, 1, extra), _spawn_role("server", port, boxes, 1, extra)]
    try:
        run_workers(_host_hier_worker, world=boxes * local_size, args=(port, local_size, ipc), timeout=240)
        for p in procs:
            p.wait(timeout=60)
        assert all(p.returncode == 0 for p in procs)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
