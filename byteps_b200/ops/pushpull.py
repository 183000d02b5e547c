"""Thin functional wrappers over the fused exchange kernels for code that owns
its symmetric arena (benchmarks, the CPU-server hierarchical mode, tests)."""
import torch

from ..comm.symm import SymmContext, pick_blocks, wire_code


def _cfg(ctx: SymmContext, nbytes: int, blocks, threads):
    shard = (nbytes + ctx.world - 1) // ctx.world
    return blocks or pick_blocks(shard, threads, 32, cap=64)


def pushpull_inplace(ctx: SymmContext, off: int, numel: int, dtype: torch.dtype, average: bool = True, stream=None,
                     blocks: int = 0, threads: int = 512, nvls=None):
    """arena[off : off+numel] on every rank becomes the (averaged) sum.  One kernel."""
    s = stream or torch.cuda.current_stream(ctx.device)
    es = torch.empty((), dtype=dtype).element_size()
    ctx.cu.pushpull_inplace(ctx.view, wire_code(dtype), off, numel, (1.0 / ctx.world) if average else 1.0,
                            _cfg(ctx, numel * es, blocks, threads), threads, 0,
                            ctx.nvls if nvls is None else nvls, s.cuda_stream)


def reduce_scatter(ctx: SymmContext, off: int, numel: int, dtype: torch.dtype, stream=None, blocks: int = 0,
                   threads: int = 512):
    """REDUCE stage of the CPU-server mode: my shard of the window becomes the box-local sum."""
    s = stream or torch.cuda.current_stream(ctx.device)
    es = torch.empty((), dtype=dtype).element_size()
    ctx.cu.reduce_scatter(ctx.view, wire_code(dtype), off, numel, _cfg(ctx, numel * es, blocks, threads), threads, 0,
                          False, s.cuda_stream)


def all_gather(ctx: SymmContext, off: int, numel: int, dtype: torch.dtype, scale: float = 1.0, stream=None,
               blocks: int = 0, threads: int = 512):
    """BROADCAST stage of the CPU-server mode: my shard is pushed (scaled) to every peer."""
    s = stream or torch.cuda.current_stream(ctx.device)
    es = torch.empty((), dtype=dtype).element_size()
    ctx.cu.all_gather(ctx.view, wire_code(dtype), off, numel, scale, _cfg(ctx, numel * es, blocks, threads), threads,
                      0, False, s.cuda_stream)


def shard_elems(ctx: SymmContext, numel: int):
    """[begin, end) element range of the shard this rank owns in a window of `numel` elements."""
    b, e = ctx.cu.shard_units((numel + 7) // 8, ctx.world, ctx.rank)
    return b * 8, min(e * 8, numel)
