"""GPU compressors vs the CPU reference implementations (two-stage contract:
worker compress -> 'server' decompress+sum+recompress -> worker decompress)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class _RankCtx:
    """SymmContext look-alike for one virtual rank."""

    def __init__(self, vc, r):
        self.cu, self.device, self.world, self.rank = vc.cu, vc.device, vc.world, r
        self.view, self.arena, self.data_bytes = vc.views[r], vc.arenas[r], vc.data_bytes


def _run_lockstep(vc, comps, grads, outs, average):
    """Issue every pipeline phase for ALL virtual ranks before the next phase (see GpuCompressor.phases)."""
    cur = torch.cuda.current_stream()
    for st in vc.streams:
        st.wait_stream(cur)
    plans = []
    for r, c in enumerate(comps):
        with torch.cuda.stream(vc.streams[r]):
            plans.append(c.phases(grads[r], outs[r], average, vc.streams[r].cuda_stream))
    for i in range(len(plans[0])):
        for r in range(len(comps)):
            with torch.cuda.stream(vc.streams[r]):
                plans[r][i]()
    for st in vc.streams:
        cur.wait_stream(st)


def _cpu_two_stage(kw, grads_per_rank_per_step, n):
    from byteps_b200 import _native

    c = _native.core()
    world = len(grads_per_rank_per_step)
    wcomp = [c.Compressor(kw, n * 4, c.F32, False) for _ in range(world)]
    scomp = c.Compressor(kw, n * 4, c.F32, True)
    buf = np.zeros(wcomp[0].max_compressed_bytes() + 64, dtype=np.uint8)
    outs = []
    for it in range(len(grads_per_rank_per_step[0])):
        total = np.zeros(n, dtype=np.float32)
        for r in range(world):
            g = grads_per_rank_per_step[r][it].copy()
            m = wcomp[r].compress(g.ctypes.data, buf.ctypes.data)
            d = np.zeros(n, dtype=np.float32)
            scomp.decompress(buf.ctypes.data, m, d.ctypes.data)
            total += d
        m = scomp.compress(total.ctypes.data, buf.ctypes.data)
        final = np.zeros(n, dtype=np.float32)
        wcomp[0].decompress(buf.ctypes.data, m, final.ctypes.data)
        outs.append(final)
    return outs


@pytest.mark.parametrize("world", [1, 4])
@pytest.mark.parametrize("kw", [
    {"compressor_type": "onebit"},
    {"compressor_type": "onebit", "compressor_onebit_scaling": "true", "ef_type": "vanilla"},
    {"compressor_type": "topk", "compressor_k": "37"},
    {"compressor_type": "topk", "compressor_k": "0.01", "ef_type": "vanilla", "momentum_type": "nesterov",
     "momentum_mu": "0.9"},
    {"compressor_type": "randomk", "compressor_k": "50", "seed": "17"},
    {"compressor_type": "randomk", "compressor_k": "50", "seed": "17", "ef_type": "vanilla"},
])
def test_gpu_compressor_matches_cpu_reference(world, kw):
    from byteps_b200.comm.symm import VirtualCluster
    from byteps_b200.ops.compress import GpuCompressor

    n, steps = 5000, 3
    vc = VirtualCluster(world, "cuda:0", 1 << 20)
    comps = [GpuCompressor(_RankCtx(vc, r), kw, n, torch.float32, payload_off=0) for r in range(world)]
    rng = np.random.RandomState(5)
    grads = [[rng.randn(n).astype(np.float32) for _ in range(steps)] for _ in range(world)]
    ref = _cpu_two_stage(kw, grads, n)
    for it in range(steps):
        gs = [torch.from_numpy(grads[r][it]).cuda() for r in range(world)]
        outs = [torch.empty(n, device="cuda") for _ in range(world)]
        _run_lockstep(vc, comps, gs, outs, False)
        torch.cuda.synchronize()
        for r in range(world):
            assert torch.equal(outs[r], outs[0]), "ranks must agree bit for bit"
        np.testing.assert_allclose(outs[0].cpu().numpy(), ref[it], rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("partition,normalize", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gpu_dithering_is_unbiased_and_bounded(partition, normalize):
    from byteps_b200.comm.symm import VirtualCluster
    from byteps_b200.ops.compress import GpuCompressor

    n, world = 4096, 2
    kw = {"compressor_type": "dithering", "compressor_k": "8", "seed": "3", "dithering_partition": str(partition),
          "dithering_normalize": str(normalize)}
    vc = VirtualCluster(world, "cuda:0", 1 << 20)
    comps = [GpuCompressor(_RankCtx(vc, r), kw, n, torch.float32, two_stage=False) for r in range(world)]
    torch.manual_seed(0)
    gs = [torch.randn(n, device="cuda") for _ in range(world)]
    exact = (gs[0] + gs[1])
    acc = torch.zeros(n, device="cuda")
    trials = 60
    for _ in range(trials):
        outs = [torch.empty(n, device="cuda") for _ in range(world)]
        _run_lockstep(vc, comps, [g.clone() for g in gs], outs, False)
        torch.cuda.synchronize()
        assert torch.equal(outs[0], outs[1])
        acc += outs[0]
    mean = acc / trials
    scale = max(g.abs().max().item() if normalize == 0 else g.norm().item() for g in gs)
    # stochastic rounding is unbiased: the mean converges to the exact sum
    assert (mean - exact).abs().mean().item() < 0.15 * scale / (8 if partition == 0 else 2) + 0.02


def test_fp16_and_bf16_inputs():
    from byteps_b200.comm.symm import VirtualCluster
    from byteps_b200.ops.compress import GpuCompressor

    for dt in (torch.bfloat16, torch.float16):
        vc = VirtualCluster(2, "cuda:0", 1 << 20)
        kw = {"compressor_type": "topk", "compressor_k": "16"}
        comps = [GpuCompressor(_RankCtx(vc, r), kw, 2048, dt, two_stage=False) for r in range(2)]
        gs = [torch.zeros(2048, device="cuda", dtype=dt) for _ in range(2)]
        gs[0][:16] = 3.0
        gs[1][16:32] = -2.0
        _run_lockstep(vc, comps, gs, [None, None], True)
        torch.cuda.synchronize()
        exp = torch.zeros(2048, device="cuda", dtype=dt)
        exp[:16] = 1.5
        exp[16:32] = -1.0
        assert torch.equal(gs[0], exp) and torch.equal(gs[1], exp)


@pytest.mark.parametrize("case", ["random", "adversarial_sample", "ties", "large_fraction"])
def test_fused_topk_select_is_exact(case):
    """Radix select with the sampled first-level filter == torch.topk on |x| (values and index set),
    including the fallback when the strided sample over-estimates the threshold."""
    from byteps_b200 import _native

    cu = _native.cuda()
    torch.manual_seed(3)
    n = 131072 * 2 + 5
    if case == "random":
        x, k = torch.randn(n, device="cuda"), 2621
    elif case == "adversarial_sample":
        # every large value sits exactly on a sampled position (stride = n // 32768 = 8): the sample holds
        # 16x the true fraction of large values, the guessed threshold is too high, the kernel must notice
        x = torch.rand(n, device="cuda") * 1e-3
        pos = torch.arange(0, 2000, device="cuda") * (n // 32768)
        x[pos] = 10.0 + torch.rand(2000, device="cuda")
        k = 2621
    elif case == "ties":
        x = torch.ones(n, device="cuda")
        x[::7] = 2.0
        k = 40000           # all 2.0 entries (37450) plus some of the tied 1.0 entries
    else:
        x, k = torch.randn(n, device="cuda"), n // 3
    x0 = x.clone()
    pairs = torch.zeros(2 * k, dtype=torch.int32, device="cuda")
    scratch = torch.zeros(cu.TOPK_SCRATCH_BYTES // 4, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    cu.topk_finish(x.data_ptr(), n, k, 0, pairs.data_ptr(), True, scratch.data_ptr(), s)
    torch.cuda.synchronize()
    idx = pairs[0::2].long()
    vals = pairs[1::2].view(torch.float32)
    assert idx.unique().numel() == k, "indices must be unique"
    assert torch.equal(vals, x0[idx])
    ref_vals, _ = torch.topk(x0.abs(), k)
    assert torch.equal(vals.abs().sort(descending=True).values, ref_vals)
    # kept entries were zeroed in place (the error-feedback update), everything else untouched
    exp = x0.clone()
    exp[idx] = 0
    assert torch.equal(x, exp)


def test_device_xorshift_stream_matches_cpu_generator():
    """random-k indices are drawn on the device by jumping ahead in the CPU compressor's xorshift128+
    stream (linear over GF(2)); consecutive calls continue the stream."""
    from byteps_b200 import _native

    cu, core = _native.cuda(), _native.core()
    seed, n = 0x1234567, 25_000_000
    rng = core.XorShift128Plus()
    rng.set_seed(seed)
    state = torch.tensor([seed, seed], dtype=torch.int64, device="cuda")
    jump = torch.frombuffer(bytearray(cu.xorshift_jump_table()), dtype=torch.int64).cuda()
    s = torch.cuda.current_stream().cuda_stream
    for k in (1, 63, 64, 1000, 250_000):
        idx = torch.zeros(k, dtype=torch.int32, device="cuda")
        cu.randomk_draw(state.data_ptr(), jump.data_ptr(), k, n, idx.data_ptr(), s)
        torch.cuda.synchronize()
        want = [rng.randint(0, n) for _ in range(k)]
        assert idx.tolist() == want, k
