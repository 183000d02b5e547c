"""The native torch adapter (csrc/torch/native_ops.cc) on one GPU: at::Tensor in, handles out,
flush windows, priority order, poll/synchronize - the role of the reference's
byteps/torch/ops.cc + handle_manager.cc + ready_event.cc.  Multi-rank numerics of the same path
run through the public API in tests/test_multigpu.py (every CUDA push_pull goes through it)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(flush_bytes=1 << 20, partition_bytes=4096000, arena=64 << 20):
    from byteps_b200 import _native

    cu, mod = _native.cuda(), _native.torch_ops()
    assert mod is not None, "the native adapter must be built on a GPU box"
    mem = cu.SymmMem(0, 1, torch.cuda.current_device(), arena, "local", "t")
    view = mem.view()
    stream = torch.cuda.Stream()
    ops = mod.NativeSymmOps(data=[view.data_ptr(0)], sig=[view.sig_ptr(0)], mc=0, epoch=view.epoch_ptr, rank=0,
                            world=1, arena_bytes=mem.data_bytes, comm_stream=stream.cuda_stream,
                            device=torch.cuda.current_device(), partition_bytes=partition_bytes,
                            group_bytes=32 << 20, one_shot_bytes=256 << 10, flush_bytes=flush_bytes, credit_bytes=0,
                            blocks=0, threads=512, nvls=False, wire_override=-1)
    return ops, mem, stream


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
def test_native_push_pull_roundtrip(dt):
    ops, mem, stream = _make()
    torch.manual_seed(0)
    sizes = [1, 7, 1000, 4097, 70001, 3_000_001]        # the last one spans three partitions
    ins = [torch.randn(n, device="cuda").to(dt) for n in sizes]
    outs = [torch.full((n,), 7.0, device="cuda", dtype=dt) for n in sizes]
    hs = [ops.push_pull_async(a, b, True, "byteps.t%d" % i, -i, 0) for i, (a, b) in enumerate(zip(ins, outs))]
    assert ops.outstanding() == len(sizes)
    for h, a, b in zip(hs, ins, outs):
        got = ops.synchronize(h, False)
        assert got.data_ptr() == b.data_ptr()
    torch.cuda.synchronize()
    for a, b in zip(ins, outs):
        assert torch.equal(a, b)                          # one rank: average of one = itself, inputs untouched
    assert ops.outstanding() == 0 and ops.launches >= 1
    with pytest.raises(Exception):
        ops.synchronize(hs[0], False)                     # released
    del ops


def test_native_flush_windows_and_poll():
    ops, mem, stream = _make(flush_bytes=1 << 20)
    small = [torch.ones(1000, device="cuda") * i for i in range(8)]
    hs = [ops.push_pull_async(t, t, False, "byteps.s%d" % i, 0, 0) for i, t in enumerate(small)]
    assert ops.launches == 0                              # 32 KB pending: below the flush window
    assert not ops.poll(hs[0])                            # polling never flushes (flush points must match on all ranks)
    big = torch.ones(1 << 20, device="cuda")              # 4 MB: crosses the window -> everything pending goes out
    hb = ops.push_pull_async(big, big, False, "byteps.big", 0, 0)
    n1 = ops.launches
    assert n1 >= 1
    torch.cuda.synchronize()
    assert all(ops.poll(h) for h in hs + [hb])
    # in-place sum over one rank leaves the data alone
    assert all(torch.all(t == i) for i, t in enumerate(small))
    # the fast entry point: prefixes the name, declines what it cannot move
    assert ops.try_push_pull_async(torch.ones(4), torch.ones(4), True, "cpu", 0, 0, True) == -1
    ints = torch.ones(4, dtype=torch.int64, device="cuda")
    assert ops.try_push_pull_async(ints, ints, True, "ints", 0, 0, True) == -1
    nc = torch.ones(8, 8, device="cuda").t()
    assert ops.try_push_pull_async(nc, nc, True, "nc", 0, 0, True) == -1
    x = torch.arange(2_000_000, device="cuda", dtype=torch.float32)
    h = ops.try_push_pull_async(x, x, True, "eager", 0, 0, True)   # >= one partition: launched immediately
    assert h >= 0 and ops.launches > n1
    assert torch.equal(ops.synchronize(h, True), torch.arange(2_000_000, device="cuda", dtype=torch.float32))
    assert ops.take_bytes() == 8_000_000 and ops.take_bytes() == 0
    del ops


def test_native_ops_cost_per_call():
    """The point of the adapter: a push_pull call must cost microseconds, not tens of them."""
    import time

    ops, mem, stream = _make(flush_bytes=16 << 20)
    ts = [torch.ones(4096, device="cuda") for _ in range(161)]
    for rep in range(3):
        t0 = time.perf_counter()
        hs = [ops.try_push_pull_async(t, t, True, "g%d" % i, -i, 0, True) for i, t in enumerate(ts)]
        ops.flush()
        for h in hs:
            ops.synchronize(h, False)
        dt = time.perf_counter() - t0
    torch.cuda.synchronize()
    per_call_us = dt / len(ts) * 1e6
    print("native push_pull + synchronize: %.2f us per tensor" % per_call_us)
    assert per_call_us < 10.0, per_call_us
    del ops
