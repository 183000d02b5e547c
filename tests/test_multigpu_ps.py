"""CPU-server mode with GPUs: flat (one GPU per worker box) and hierarchical
(all GPUs of the box form one worker: NVLink reduce-scatter -> 8 parallel host
paths through the server -> NVLink all-gather)."""
import os
import subprocess
import sys

import pytest
import torch

from _mp import free_port, run_workers

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _spawn_role(role, port, nw, ns, local_size, ipc=False):
    env = dict(os.environ)
    env["BYTEPS_ENABLE_IPC"] = "1" if ipc else "0"
    env.update({"DMLC_ROLE": role, "DMLC_NUM_WORKER": str(nw), "DMLC_NUM_SERVER": str(ns),
                "BYTEPS_LOCAL_SIZE": str(local_size), "DMLC_PS_ROOT_URI": "127.0.0.1",
                "DMLC_PS_ROOT_PORT": str(port), "PYTHONPATH": ROOT})
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    return subprocess.Popen([sys.executable, "-c", "import byteps_b200.server"], env=env)


def _worker(rank, world, ps_port, hier, ipc=False):
    torch.cuda.set_device(rank)
    os.environ["BYTEPS_ENABLE_IPC"] = "1" if ipc else "0"
    if hier:   # one box, `world` GPUs
        os.environ.update({"BYTEPS_LOCAL_RANK": str(rank), "BYTEPS_LOCAL_SIZE": str(world), "DMLC_NUM_WORKER": "1",
                           "DMLC_WORKER_ID": "0", "BYTEPS_FORCE_DISTRIBUTED": "1"})
    else:      # `world` boxes with one GPU each
        os.environ.update({"BYTEPS_LOCAL_RANK": "0", "BYTEPS_LOCAL_SIZE": "1", "DMLC_NUM_WORKER": str(world),
                           "DMLC_WORKER_ID": str(rank)})
    os.environ.update({"DMLC_ROLE": "worker", "DMLC_NUM_SERVER": "1", "DMLC_PS_ROOT_URI": "127.0.0.1",
                       "DMLC_PS_ROOT_PORT": str(ps_port)})
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE"):
        os.environ.pop(k, None)
    os.environ["MASTER_PORT"] = str(ps_port + 1)
    import byteps_b200.torch as bps
    from byteps_b200.common import engine

    bps.init()
    assert engine().backend == "ps" and bps.size() == world
    tot = sum(r + 1 for r in range(world))
    # several rounds per name: with IPC the pull is answered by reference into the server's double-buffered store
    # (kv_app.h / server.cc), so round k+1 must not disturb what round k is still copying to the GPU
    for it in range(3 if ipc else 1):
        for dt in (torch.float32, torch.bfloat16):
            for n in (1000, 3_000_000):
                g = ((torch.arange(n, device="cuda") + it) % 11).to(dt) * (rank + 1)
                out = bps.push_pull_inplace(g, average=False, name="g_%s_%d" % (str(dt)[6:], n))
                ref = (((torch.arange(n, device="cuda") + it) % 11).float() * tot).to(dt)
                assert torch.allclose(out.float(), ref.float(), rtol=1e-2), (dt, n, it)
                g2 = torch.ones(n, device="cuda", dtype=dt) * (rank + 1 + it)
                out = bps.push_pull(g2, average=True, name="a_%s_%d" % (str(dt)[6:], n))
                assert torch.allclose(out.float(), torch.full((n,), tot / world + it, device="cuda"), rtol=1e-2), (dt, n, it)
    bps.shutdown()


@pytest.mark.parametrize("hier,ipc", [(False, False), (True, False), (False, True), (True, True)])
def test_cpu_server_mode_two_gpus(hier, ipc):
    port = free_port()
    nw_nodes = 2
    procs = [_spawn_role("scheduler", port, 1 if hier else 2, 1, 2 if hier else 1, ipc),
             _spawn_role("server", port, 1 if hier else 2, 1, 2 if hier else 1, ipc)]
    try:
        run_workers(_worker, world=nw_nodes, args=(port, hier, ipc), timeout=240)
        for p in procs:
            p.wait(timeout=60)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
