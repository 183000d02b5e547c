"""End-to-end compressor training equivalence, the reference's own test idea
(/root/reference/tests/test_onebit.py:32-116, test_topk.py:33-115, test_randomk.py:33-128,
test_dithering.py:45-178 and their harness meta_test.py:26-85): train a model through the public API with ``compression_params`` on a
1-worker + 1-server job (BYTEPS_FORCE_DISTRIBUTED=1, BYTEPS_MIN_COMPRESS_BYTES=0), re-implement the
compressor in numpy with the SAME xorshift128+ stream, apply it TWICE to the recorded gradients
(worker stage then server stage, each with its own error-feedback state) and require the final
weights to agree to fp32 round-off.  Runs on the CPU: scheduler and server are separate processes.
"""
import os

import numpy as np
import pytest
import torch

from _mp import free_port, run_workers
from test_ps_api import _spawn_role

LR, STEPS = 0.05, 6


class _XorShift:
    """xorshift128+ exactly as csrc/compress/compressor.h (reference utils.h:74-113): a = b = seed"""

    M = (1 << 64) - 1

    def __init__(self, seed):
        self.a = self.b = seed & self.M

    def next(self):
        t, s = self.a, self.b
        self.a = s
        t ^= (t << 23) & self.M
        t ^= t >> 17
        t ^= s ^ (s >> 26)
        self.b = t
        return (t + s) & self.M

    def randint(self, lo, hi):
        return self.next() % (hi - lo) + lo

    def bernoulli(self, p):
        """double(next()) < p * double(2^64 - 1), compressor.h"""
        return float(self.next()) < p * float(self.M)


def _round_next_pow2(v):
    v = (v - 1) & 0xFFFFFFFF
    for sh in (1, 2, 4, 8, 16):
        v |= v >> sh
    return (v + 1) & 0xFFFFFFFF


class _NumpyStage:
    """One compression stage (worker or server) of one tensor: optional vanilla error feedback around
    onebit / topk / randomk, returning the DEcompressed tensor the other side sees."""

    def __init__(self, kind, n, k=None, scaling=False, ef=False, seed=0, partition="linear", normalize="max"):
        self.kind, self.n, self.k, self.scaling, self.ef = kind, n, k, scaling, ef
        self.partition, self.normalize = partition, normalize
        self.err = np.zeros(n, dtype=np.float32)
        self.rng = _XorShift(seed) if kind in ("randomk", "dithering") else None

    def _dither(self, p):
        """stochastic quantisation to k levels (linear) or 2^(k-1) geometric levels (natural) of |x| / scale,
        one xorshift draw per element in index order; the other side sees sign * q * scale / levels"""
        s = self.k
        if self.normalize == "max":
            scale = float(np.abs(p).max()) if self.n else 0.0
        else:
            scale = 0.0
            for v in p.tolist():                 # sequential double sum, like the compressor
                scale += v * v
            scale = scale ** 0.5
        d = np.zeros(self.n, dtype=np.float32)
        if scale <= 0:
            return d
        fscale = np.float32(scale)
        levels = s if self.partition == "linear" else 1 << (s - 1)
        for i, x in enumerate(p.tolist()):
            ax = abs(x)
            if self.partition == "linear":
                nrm = np.float32((ax / scale) * s)
                fl = np.floor(nrm)
                q = int(fl) + (1 if self.rng.bernoulli(float(np.float32(nrm - fl))) else 0)
            else:
                nrm = (ax / scale) * levels
                fl = _round_next_pow2(int(np.ceil(nrm))) >> 1
                length = fl if fl else 1
                q = fl + length * (1 if self.rng.bernoulli((nrm - fl) / length) else 0)
            if q:
                num = np.float32(np.float32(q) * fscale) / np.float32(levels)
                d[i] = -num if np.signbit(np.float32(x)) else num
        return d

    def __call__(self, g):
        p = (g + self.err).astype(np.float32) if self.ef else g.astype(np.float32)
        if self.kind == "onebit":
            scale = np.float32(np.abs(p).astype(np.float64).sum() / self.n) if self.scaling else np.float32(1.0)
            d = np.where(p < 0, -scale, scale).astype(np.float32)
        elif self.kind == "dithering":
            d = self._dither(p)
        elif self.kind == "topk":
            idx = np.argsort(-np.abs(p), kind="stable")[:self.k]
            d = np.zeros(self.n, dtype=np.float32)
            d[idx] = p[idx]
        else:
            d = np.zeros(self.n, dtype=np.float32)
            for _ in range(self.k):
                i = self.rng.randint(0, self.n)
                d[i] = p[i]
        if self.ef:
            self.err = (p - d).astype(np.float32)
        return d


CASES = {
    "onebit": dict(params={"compressor": "onebit", "scaling": True}, stage=dict(kind="onebit", scaling=True)),
    "onebit_ef": dict(params={"compressor": "onebit", "scaling": True, "ef": "vanilla"},
                      stage=dict(kind="onebit", scaling=True, ef=True)),
    "topk_ef": dict(params={"compressor": "topk", "k": 0.05, "ef": "vanilla"}, stage=dict(kind="topk", ef=True)),
    "randomk_ef": dict(params={"compressor": "randomk", "k": 0.05, "ef": "vanilla", "seed": 2020},
                       stage=dict(kind="randomk", ef=True, seed=2020)),
    # the reference's dithering matrix (test_dithering.py:166-178): levels x partition x normalisation x seed
    "dithering_linear_max": dict(params={"compressor": "dithering", "k": 4, "partition": "linear", "normalize": "max",
                                         "seed": 2020},
                                 stage=dict(kind="dithering", k=4, partition="linear", normalize="max", seed=2020)),
    "dithering_linear_l2": dict(params={"compressor": "dithering", "k": 8, "partition": "linear", "normalize": "l2",
                                        "seed": 2020},
                                stage=dict(kind="dithering", k=8, partition="linear", normalize="l2", seed=2020)),
    "dithering_natural_max": dict(params={"compressor": "dithering", "k": 3, "partition": "natural",
                                          "normalize": "max", "seed": 2021},
                                  stage=dict(kind="dithering", k=3, partition="natural", normalize="max", seed=2021)),
    "dithering_natural_l2_ef": dict(params={"compressor": "dithering", "k": 4, "partition": "natural",
                                            "normalize": "l2", "seed": 7, "ef": "vanilla"},
                                    stage=dict(kind="dithering", k=4, partition="natural", normalize="l2", seed=7,
                                               ef=True)),
}


def _worker(rank, world, ps_port, case):
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE"):
        os.environ.pop(k, None)
    os.environ.update({"DMLC_ROLE": "worker", "DMLC_NUM_WORKER": "1", "DMLC_NUM_SERVER": "1", "DMLC_WORKER_ID": "0",
                       "BYTEPS_LOCAL_RANK": "0", "BYTEPS_LOCAL_SIZE": "1", "DMLC_PS_ROOT_URI": "127.0.0.1",
                       "DMLC_PS_ROOT_PORT": str(ps_port), "BYTEPS_FORCE_DISTRIBUTED": "1",
                       "BYTEPS_MIN_COMPRESS_BYTES": "0", "BYTEPS_PARTITION_BYTES": "2147483647"})
    import byteps_b200.torch as bps
    from byteps_b200.common import engine

    bps.init()
    assert engine().backend == "ps" and bps.size() == 1
    spec = CASES[case]
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(20, 64), torch.nn.Tanh(), torch.nn.Linear(64, 10))
    w0 = {n: p.detach().numpy().copy() for n, p in model.named_parameters()}
    opt = bps.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=LR), named_parameters=model.named_parameters(),
                                   compression_params=spec["params"])
    bps.broadcast_parameters(model.state_dict(), root_rank=0)
    # the optimizer's hooks start the (asynchronous, in-place) push_pull of a gradient the moment autograd has
    # accumulated it, so the raw gradient has to be captured by a tensor hook, which runs before that
    recorded, current = [], {}
    for name, p in model.named_parameters():
        p.register_hook(lambda g, name=name: current.__setitem__(name, g.detach().numpy().copy()))
    torch.manual_seed(1)
    xs, ys = torch.randn(STEPS, 16, 20), torch.randint(0, 10, (STEPS, 16))
    for i in range(STEPS):
        opt.zero_grad()
        torch.nn.functional.cross_entropy(model(xs[i]), ys[i]).backward()
        recorded.append(dict(current))
        opt.step()
    # ---- the numpy model: worker stage, then server stage (sum over ONE worker), then plain SGD
    for name, p in model.named_parameters():
        n = p.numel()
        kw = dict(spec["stage"])
        if kw["kind"] in ("topk", "randomk"):
            kw["k"] = max(1, int(spec["params"]["k"] * n))
        worker, server = _NumpyStage(n=n, **kw), _NumpyStage(n=n, **kw)
        w = w0[name].reshape(-1).copy()
        for i in range(STEPS):
            g = recorded[i][name].reshape(-1)
            w = (w - np.float32(LR) * server(worker(g))).astype(np.float32)
        got = p.detach().numpy().reshape(-1)
        np.testing.assert_allclose(got, w, rtol=2e-5, atol=2e-6, err_msg="%s / %s" % (case, name))
    bps.shutdown()


@pytest.mark.parametrize("case", sorted(CASES))
def test_training_matches_two_stage_numpy_model(case):
    port = free_port()
    procs = [_spawn_role("scheduler", port, 1, 1), _spawn_role("server", port, 1, 1)]
    try:
        run_workers(_worker, world=1, args=(port, case), timeout=180)
        for p in procs:
            p.wait(timeout=60)
        assert all(p.returncode == 0 for p in procs)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
