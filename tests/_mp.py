"""Helpers to run a function in N fresh processes with a torchrun-style env."""
import os
import socket
import sys
import traceback

import torch.multiprocessing as mp


def free_port():
    """A free port whose successor is free too (DMLC_PS_ROOT_PORT + 1 hosts the torch.distributed rendezvous)."""
    for _ in range(64):
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        p = s.getsockname()[1]
        s2 = socket.socket()
        try:
            s2.bind(("127.0.0.1", p + 1))
            return p
        except OSError:
            continue
        finally:
            s2.close()
            s.close()
    return p


def _entry(rank, world, port, fn, args, extra_env, errq):
    try:
        os.environ.update({"RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank),
                           "LOCAL_WORLD_SIZE": str(world), "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
        for k in ("BYTEPS_LOCAL_RANK", "BYTEPS_LOCAL_SIZE", "DMLC_NUM_WORKER", "DMLC_WORKER_ID"):
            os.environ.pop(k, None)
        os.environ.update(extra_env or {})
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        if root not in sys.path:
            sys.path.insert(0, root)
        fn(rank, world, *args)
    except Exception:  # noqa: BLE001
        errq.put((rank, traceback.format_exc()))
        raise


def run_workers(fn, world=2, args=(), env=None, timeout=120):
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = free_port()
    procs = [ctx.Process(target=_entry, args=(r, world, port, fn, args, env, errq)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout)
    errs = []
    while not errq.empty():
        errs.append(errq.get())
    for p in procs:
        if p.is_alive():
            p.kill()
            errs.append((-1, "timeout"))
        elif p.exitcode != 0:
            errs.append((-1, "exit code %s" % p.exitcode))
    assert not errs, "\n".join("rank %s: %s" % e for e in errs)
