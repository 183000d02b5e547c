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
