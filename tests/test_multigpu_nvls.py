"""The default multi-GPU data path, world in {2, 4, 8} x BYTEPS_USE_NVLS in {0, 1}.

Every exchange kernel a training step can launch - in-place, packed (one- and two-shot), fused
optimizer (register and TMA-streamed), the descriptor ring (static order, producer marks, root
scheduling) - and the whole DistributedOptimizer (eager and CUDA-graph) are checked against fp32
``torch.distributed.all_reduce`` / ``torch.optim`` references, with bit-identical replicas
asserted by an all-gather of checksums.  NVLS (multimem.ld_reduce / multimem.st) binds from two
GPUs up; with four or more it is what ``auto`` selects, i.e. what every scaling number runs on.
A negative test checks that a mismatched bucket order ends in a trapped kernel, not a hang.
"""
import os
import struct

import pytest
import torch

from _mp import run_workers

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]

WORLDS = [2, 4, 8]


def _need(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)


def _env(nvls, **extra):
    e = {"BYTEPS_USE_NVLS": "1" if nvls else "0", "BYTEPS_SPIN_TIMEOUT_MS": "20000"}
    e.update(extra)
    return e


def _hp(lr=0.1, wd=0.0, mom=0.0, damp=0.0, b1=0.9, b2=0.999, eps=1e-8, t=1, nesterov=0, adamw=0, first=1, gs=1.0):
    return struct.pack("<9f3if3i", lr, wd, mom, damp, b1, b2, eps, 1 - b1 ** t, 1 - b2 ** t, nesterov, adamw, first,
                       gs, 0, 0, 0)


def _identical_everywhere(t, tag):
    """all ranks hold bit-identical bytes: compare 64-bit checksums gathered with NCCL"""
    import torch.distributed as dist

    raw = t.contiguous().reshape(-1).view(torch.uint8)
    pad = (-raw.numel()) % 8
    if pad:
        raw = torch.cat([raw, torch.zeros(pad, dtype=torch.uint8, device=raw.device)])
    w = raw.view(torch.int64)
    idx = torch.arange(1, w.numel() + 1, device=w.device, dtype=torch.int64)
    chk = torch.stack([(w * idx).sum(), (w ^ idx).sum()])
    out = [torch.zeros_like(chk) for _ in range(dist.get_world_size())]
    dist.all_gather(out, chk)
    for o in out[1:]:
        assert torch.equal(o, out[0]), "replicas differ: %s" % (tag,)


def _kernels(rank, world, nvls):
    import torch.distributed as dist

    import byteps_b200.torch as bps
    from byteps_b200.common import engine
    from byteps_b200.comm.symm import SymmContext, wire_code
    from byteps_b200.ops.ring import RingEntry, RingTable

    torch.cuda.set_device(rank)
    bps.init()
    eng = engine()
    dev = torch.device("cuda", rank)
    ctx = SymmContext(eng.group, dev, 96 << 20, "auto", "1" if nvls else "0")
    assert ctx.nvls == bool(nvls), "multicast binding: wanted %s got %s" % (nvls, ctx.nvls)
    cu, view = ctx.cu, ctx.view
    s = torch.cuda.current_stream().cuda_stream
    tol = {torch.float32: 1e-6, torch.bfloat16: 1e-2, torch.float16: 2e-3}

    def ref_sum(x):
        r = x.float().clone()
        dist.all_reduce(r)
        return r

    # ---- in-place: ragged sizes, several grids
    for dt in (torch.float32, torch.bfloat16, torch.float16):
        es = torch.empty((), dtype=dt).element_size()
        for n, blocks in ((8, 1), (8 * 1031 * 3 + 8, 3), (8 * 1_000_003, 32)):
            torch.manual_seed(1000 * rank + n % 997)
            x = torch.randn(n, device=dev).to(dt)
            ref = ref_sum(x) / world
            ctx.tensor(256, n, dt).copy_(x)
            dist.barrier()
            cu.pushpull_inplace(view, wire_code(dt), 256, n, 1.0 / world, blocks, 512, 0, bool(nvls), s)
            torch.cuda.synchronize()
            got = ctx.tensor(256, n, dt)
            assert torch.allclose(got.float(), ref, atol=tol[dt] * max(1.0, ref.abs().max().item()), rtol=tol[dt]), \
                ("inplace", dt, n, (got.float() - ref).abs().max().item())
            _identical_everywhere(got, ("inplace", dt, n))
            dist.barrier()

    # ---- packed: user tensors -> staging -> user tensors, one-shot and two-shot, with a wire cast
    for ud, wd in ((torch.float32, torch.float32), (torch.float32, torch.bfloat16), (torch.bfloat16, torch.bfloat16)):
        for one_shot in (False, True):
            sizes = [5, 8, 1000, 4097, 33, 70000]
            torch.manual_seed(77 + rank)
            ins = [torch.randn(n, device=dev).to(ud) for n in sizes]
            outs = [torch.full((n,), 7.0, device=dev, dtype=ud) for n in sizes]
            rows, start = [], 0
            for a, b in zip(ins, outs):
                rows.append([a.data_ptr(), b.data_ptr(), start, a.numel()])
                start += (a.numel() + 7) // 8 * 8
            table = torch.tensor(rows, dtype=torch.int64, device=dev)
            refs = [ref_sum(a.to(wd)).to(wd).to(ud) for a in ins]
            dist.barrier()
            cu.pushpull_packed(view, wire_code(ud), wire_code(wd), table.data_ptr(), len(sizes), 1 << 20, start, 1.0, 4,
                               512, 0, bool(nvls) and not one_shot, one_shot, True, s)
            torch.cuda.synchronize()
            for got, ref in zip(outs, refs):
                t = tol[wd] * 2
                assert torch.allclose(got.float(), ref.float(), atol=t * max(1.0, ref.float().abs().max().item()),
                                      rtol=t), ("packed", ud, wd, one_shot)
                _identical_everywhere(got, ("packed", ud, wd, one_shot))
            dist.barrier()

    # ---- fused optimizer kernels (register and TMA engines) and ring descriptors vs torch.optim
    def fused_case(kind, dt, engine_name):
        n = 8 * 50_021
        es = torch.empty((), dtype=dt).element_size()
        goff, poff = 0, (n * es + 255) // 256 * 256
        torch.manual_seed(5)
        w0 = torch.randn(n, device=dev).to(dt)
        ref_w = torch.nn.Parameter(w0.float().clone())
        if kind == "sgd":
            ropt = torch.optim.SGD([ref_w], lr=0.1, momentum=0.9, weight_decay=0.01)
            code, rkind = cu.OPT_SGD, cu.RING_SGD
        else:
            ropt = torch.optim.AdamW([ref_w], lr=0.01, weight_decay=0.01)
            code, rkind = cu.OPT_ADAM, cu.RING_ADAM
        b, e = cu.shard_units(n // 8, world, rank)
        master = torch.zeros(max((e - b) * 8, 8), device=dev)
        master[:(e - b) * 8] = w0.float()[b * 8:e * 8]
        s0, s1 = torch.zeros_like(master), torch.zeros_like(master)
        ctx.tensor(poff, n, dt).copy_(w0)
        hp = torch.zeros(64, dtype=torch.uint8, device=dev)
        table = RingTable([RingEntry(grad_off=goff, param_off=poff, numel=n, wire=wire_code(dt), slot=7, kind=rkind,
                                     scale=1.0 / world, master=master.data_ptr(), state0=s0.data_ptr(),
                                     state1=s1.data_ptr(), hp=hp.data_ptr())], dev)
        for step in range(1, 4):
            torch.manual_seed(step * 31 + rank)
            g = torch.randn(n, device=dev).to(dt)
            ctx.tensor(goff, n, dt).copy_(g)
            blob = (_hp(lr=0.1, wd=0.01, mom=0.9, first=int(step == 1)) if kind == "sgd"
                    else _hp(lr=0.01, wd=0.01, t=step, adamw=1, first=int(step == 1)))
            cu.write_blob(hp.data_ptr(), blob, s)
            ref_w.grad = ref_sum(g) / world
            ropt.step()
            dist.barrier()
            if engine_name == "lsu":
                cu.pushpull_fused_opt(view, wire_code(dt), wire_code(dt), wire_code(dt), code, 0, 0, goff, poff, n,
                                      1.0 / world, master.data_ptr(), s0.data_ptr(), s1.data_ptr(), hp.data_ptr(), 8,
                                      512, 0, bool(nvls), s)
            elif engine_name == "tma":
                cu.pushpull_fused_opt_tma(view, wire_code(dt), code, goff, poff, n, 1.0 / world, master.data_ptr(),
                                          s0.data_ptr(), s1.data_ptr(), hp.data_ptr(), 8, 3, bool(nvls), 0, s)
            else:
                table.launch(view, 8, s, nvls=bool(nvls))
            torch.cuda.synchronize()
            got = ctx.tensor(poff, n, dt)
            t = 1e-5 if dt == torch.float32 else 1.2e-2
            assert torch.allclose(got.float(), ref_w.detach(), atol=t, rtol=t), \
                ("fused", kind, dt, engine_name, step, (got.float() - ref_w.detach()).abs().max().item())
            _identical_everywhere(got, ("fused", kind, dt, engine_name, step))
            dist.barrier()

    for kind in ("sgd", "adamw"):
        for dt in (torch.float32, torch.bfloat16):
            for engine_name in ("lsu", "tma", "ring"):
                fused_case(kind, dt, engine_name)

    # ---- ring: many buckets per launch, producer marks from a side stream, root scheduling
    dt = torch.bfloat16
    sizes = [8 * 1031, 8, 8 * 400_007, 8 * 513, 8 * 3, 8 * 100_003]
    offs, off = [], 256
    for n in sizes:
        offs.append(off)
        off = (off + n * 2 + 255) // 256 * 256
    entries = [RingEntry(grad_off=o, numel=n, wire=wire_code(dt), slot=10 + i, scale=1.0 / world, priority=i % 3)
               for i, (o, n) in enumerate(zip(offs, sizes))]
    table = RingTable(entries, dev)
    side = torch.cuda.Stream()
    for variant in ("self", "marks", "sched", "sched_credit"):
        torch.manual_seed(rank * 13 + len(variant))
        ins = [torch.randn(n, device=dev).to(dt) for n in sizes]
        refs = [ref_sum(x) / world for x in ins]
        for o, n, x in zip(offs, sizes, ins):
            ctx.tensor(o, n, dt).copy_(x)
        torch.cuda.synchronize()
        dist.barrier()
        if variant == "self":
            table.launch(view, 16, s, nvls=bool(nvls))
        else:
            order = list(range(len(sizes)))
            order = order[rank % len(order):] + order[:rank % len(order)]     # a different mark order on every rank
            side.wait_stream(torch.cuda.current_stream())
            table.launch(view, 16, s, nvls=bool(nvls), self_mark=False, sched=variant != "marks",
                         credit_bytes=(1 << 20) if variant == "sched_credit" else 0)
            for i in order:
                cu.ring_mark(view, [entries[i].slot], side.cuda_stream)
        torch.cuda.synchronize()
        for o, n, ref in zip(offs, sizes, refs):
            got = ctx.tensor(o, n, dt)
            assert torch.allclose(got.float(), ref, atol=1e-2 * max(1.0, ref.abs().max().item()), rtol=1e-2), \
                ("ring", variant, n)
            _identical_everywhere(got, ("ring", variant, n))
        dist.barrier()
    torch.cuda.synchronize()
    dist.barrier()
    ctx.close()
    bps.shutdown()


@pytest.mark.parametrize("nvls", [0, 1])
@pytest.mark.parametrize("world", WORLDS)
def test_exchange_kernels(world, nvls):
    _need(world)
    run_workers(_kernels, world=world, args=(nvls,), env=_env(nvls), timeout=600)


def _optimizer_case(rank, world, nvls, fused, graph, ring, opt_name):
    import byteps_b200.torch as bps
    from byteps_b200.torch.graph import GraphedStep

    os.environ["BYTEPS_RING"] = ring
    torch.manual_seed(100 + rank)
    mk = lambda: torch.nn.Sequential(torch.nn.Linear(64, 256), torch.nn.ReLU(), torch.nn.Linear(256, 256),  # noqa: E731
                                     torch.nn.ReLU(), torch.nn.Linear(256, 16)).cuda()
    model = mk()
    if opt_name == "sgd":
        base = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9)
    else:
        base = torch.optim.AdamW(model.parameters(), lr=0.01, weight_decay=0.01)
    opt = bps.DistributedOptimizer(base, named_parameters=model.named_parameters(), fused_update=fused,
                                   bucket_bytes=64 << 10)      # several buckets
    gs = opt.grad_sync
    tag = (world, nvls, fused, graph, ring, opt_name)
    assert gs is not None and gs.ctx.nvls == bool(nvls), tag
    assert gs._ring_mode == ring, (gs._ring_mode, tag)
    assert len(gs.buckets) >= 3
    bps.broadcast_parameters(model.state_dict(), root_rank=0)
    if not fused:
        bps.broadcast_optimizer_state(opt, root_rank=0)
    ref = mk()
    ref.load_state_dict(model.state_dict())
    if opt_name == "sgd":
        ropt = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9)
    else:
        ropt = torch.optim.AdamW(ref.parameters(), lr=0.01, weight_decay=0.01)
    torch.manual_seed(7)
    steps = 6
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
        assert torch.allclose(a, b, atol=2e-4), (tag, (a - b).abs().max())
        _identical_everywhere(a.detach(), ("parameters after training",) + tag)
    ms = gs.exposed_comm_ms()
    assert ms is not None and 0.0 <= ms < 1000.0, (tag, ms)
    if ring != "off":
        spans = gs.ring_spans()
        assert spans and all(t1 >= t0 > 0 for _, _, t0, t1 in spans), (tag, spans)
    del runner
    gs.close()


OPT_CASES = [(False, False, "sgd"), (True, False, "sgd"), (True, True, "sgd"), (True, True, "adamw")]


def _optimizer_matrix(rank, world, nvls):
    import torch.distributed as dist

    import byteps_b200.torch as bps

    torch.cuda.set_device(rank)
    bps.init()
    for ring in ("off", "batch", "persistent"):
        for fused, graph, opt_name in OPT_CASES:
            _optimizer_case(rank, world, nvls, fused, graph, ring, opt_name)
            dist.barrier()
    bps.shutdown()


@pytest.mark.parametrize("nvls", [0, 1])
@pytest.mark.parametrize("world", WORLDS)
def test_distributed_optimizer(world, nvls):
    """DistributedOptimizer x {unfused, fused, fused+graph SGD, fused+graph AdamW} x BYTEPS_RING in
    {off, batch, persistent}: trained parameters equal full-batch torch training, bit-identical
    on every rank."""
    _need(world)
    run_workers(_optimizer_matrix, world=world, args=(nvls,), env=_env(nvls), timeout=900)


def _priority_training(rank, world):
    """BYTEPS_SCHEDULING_CREDIT on the training path: the ring's root scheduler runs, results
    stay exact, and the trace shows every bucket was processed once."""
    import byteps_b200.torch as bps

    torch.cuda.set_device(rank)
    bps.init()
    torch.manual_seed(3)
    model = torch.nn.Sequential(*[torch.nn.Linear(128, 128) for _ in range(6)]).cuda()
    opt = bps.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.01), bucket_bytes=32 << 10,
                                   named_parameters=model.named_parameters(), fused_update=True)
    gs = opt.grad_sync
    assert gs._ring_sched and gs._ring_mode == "persistent"
    prios = [b.priority for b in gs.buckets]
    assert prios == sorted(prios), "buckets are laid out in backward order: later buckets hold earlier layers"
    bps.broadcast_parameters(model.state_dict(), root_rank=0)
    ref = torch.nn.Sequential(*[torch.nn.Linear(128, 128) for _ in range(6)]).cuda()
    ref.load_state_dict(model.state_dict())
    ropt = torch.optim.SGD(ref.parameters(), lr=0.01)
    torch.manual_seed(11)
    xs = torch.randn(3, world * 4, 128, device="cuda")
    for i in range(3):
        opt.zero_grad()
        model(xs[i, rank * 4:(rank + 1) * 4]).square().mean().backward()
        opt.step()
        ropt.zero_grad()
        ref(xs[i]).square().mean().backward()
        ropt.step()
    torch.cuda.synchronize()
    for a, b in zip(model.parameters(), ref.parameters()):
        assert torch.allclose(a, b, atol=1e-5), (a - b).abs().max()
    pos = sorted(p for _, p, _, _ in gs.ring_spans())
    assert pos == list(range(len(gs.buckets)))
    bps.shutdown()


@pytest.mark.parametrize("world", [2, 8])
def test_priority_scheduling_on_the_training_path(world):
    _need(world)
    run_workers(_priority_training, world=world, env=_env(world >= 4, BYTEPS_SCHEDULING_CREDIT="1",
                                                           BYTEPS_RING="persistent"), timeout=600)


def _mismatched_order(rank, world):
    """Rank 1 launches its two buckets in the opposite order: the per-bucket kernels wait on each
    other's barrier slots forever.  The watchdog must turn that into a CUDA error quickly."""
    import byteps_b200.torch as bps
    from byteps_b200.common import engine
    from byteps_b200.comm.symm import SymmContext, wire_code
    from byteps_b200.ops.ring import RingEntry, RingTable

    torch.cuda.set_device(rank)
    bps.init()
    dev = torch.device("cuda", rank)
    ctx = SymmContext(engine().group, dev, 1 << 20, "auto", "0")
    e = [RingEntry(grad_off=0, numel=8 * 1024, wire=wire_code(torch.float32), slot=0, scale=1.0),
         RingEntry(grad_off=1 << 16, numel=8 * 1024, wire=wire_code(torch.float32), slot=1, scale=1.0)]
    ta, tb = RingTable(e[:1], dev), RingTable(e[1:], dev)
    first, second = (ta, tb) if rank == 0 else (tb, ta)
    s = torch.cuda.current_stream().cuda_stream
    first.launch(ctx.view, 1, s)
    second.launch(ctx.view, 1, s)
    try:
        torch.cuda.synchronize()
    except Exception as exc:  # noqa: BLE001
        print("TRAPPED: %s" % str(exc).splitlines()[0], flush=True)
        os._exit(42)
    os._exit(0)


def test_mismatched_launch_order_traps_instead_of_hanging():
    """all ranks must launch the same exchanges in the same order.  When they
    do not, every rank gets a trapped kernel within BYTEPS_SPIN_TIMEOUT_MS - not a hung GPU."""
    _need(2)
    import time

    t0 = time.time()
    with pytest.raises(AssertionError) as info:
        run_workers(_mismatched_order, world=2, env={"BYTEPS_SPIN_TIMEOUT_MS": "1500", "BYTEPS_USE_NVLS": "0"},
                    timeout=120)
    assert "exit code 42" in str(info.value), str(info.value)
    assert time.time() - t0 < 100
