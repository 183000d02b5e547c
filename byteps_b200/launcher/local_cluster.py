"""Run a whole BytePS job on ONE host: scheduler + S servers + N workers as local processes.

The counterpart of the reference's dmlc tracker ``local.py`` (/root/reference/3rdparty/ps-lite/
tracker) and of the harness its tests use (tests/meta_test.py:26-85): every worker
is its own one-device "box", so the full worker -> server -> worker path runs without a
cluster.

    python -m byteps_b200.launcher.local_cluster -n 2 -s 1 python train.py
    python -m byteps_b200.launcher.local_cluster -n 2 -s 2 --gpus-per-worker 4 python train.py
"""
from __future__ import annotations

import argparse
import os
import socket
import subprocess
import sys
import time
from typing import List


def _free_port() -> int:
    """A port p with p + 1 free as well: p is the scheduler's, p + 1 the workers' torch.distributed rendezvous."""
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


def die_with_parent():
    """preexec_fn: the child gets SIGTERM when the launcher dies (even by SIGKILL), so an aborted job leaves no
    scheduler / server / worker behind (Linux prctl(PR_SET_PDEATHSIG))."""
    try:
        import ctypes
        import signal

        ctypes.CDLL(None, use_errno=True).prctl(1, signal.SIGTERM, 0, 0, 0)      # PR_SET_PDEATHSIG = 1
    except Exception:  # noqa: BLE001 - not Linux / no libc: keep going without it
        pass


def build_envs(num_workers: int, num_servers: int, port: int, gpus_per_worker: int = 1, base=None):
    """[(role, env)] for the scheduler, the servers and every worker process."""
    base = dict(os.environ if base is None else base)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        base.pop(k, None)
    # every GPU process is a transport-level worker node, so scheduler and servers need the per-box process count too
    common = {"DMLC_NUM_WORKER": str(num_workers), "DMLC_NUM_SERVER": str(num_servers),
              "DMLC_PS_ROOT_URI": "127.0.0.1", "DMLC_PS_ROOT_PORT": str(port),
              "BYTEPS_LOCAL_SIZE": str(gpus_per_worker)}
    out = [("scheduler", dict(base, DMLC_ROLE="scheduler", **common))]
    for _ in range(num_servers):
        out.append(("server", dict(base, DMLC_ROLE="server", **common)))
    for w in range(num_workers):
        for lr in range(gpus_per_worker):
            env = dict(base, DMLC_ROLE="worker", DMLC_WORKER_ID=str(w), BYTEPS_LOCAL_RANK=str(lr), **common)
            if gpus_per_worker * num_workers > 1 or num_servers > 0:
                env.setdefault("BYTEPS_FORCE_DISTRIBUTED", "1")
            if num_servers > 0:
                env.setdefault("OMP_WAIT_POLICY", "passive")     # see launcher/launch.py::worker_command
            out.append(("worker", env))
    return out


def main(argv: List[str] = None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-n", "--num-workers", type=int, default=2)
    ap.add_argument("-s", "--num-servers", type=int, default=1)
    ap.add_argument("--gpus-per-worker", type=int, default=1)
    ap.add_argument("--port", type=int, default=0)
    ap.add_argument("command", nargs=argparse.REMAINDER)
    args = ap.parse_args(argv)
    if not args.command:
        ap.error("no worker command given")
    port = args.port or _free_port()
    procs = []
    for role, env in build_envs(args.num_workers, args.num_servers, port, args.gpus_per_worker):
        cmd = args.command if role == "worker" else [sys.executable, "-c", "import byteps_b200.server"]
        procs.append((role, subprocess.Popen(cmd, env=env, preexec_fn=die_with_parent)))
    rc = 0
    try:
        workers = [p for role, p in procs if role == "worker"]
        while any(p.poll() is None for p in workers):
            failed = [p.returncode for p in workers if p.poll() not in (None, 0)]
            if failed:       # a dead worker never reaches the collectives the others wait in: stop the job
                rc = failed[0]
                print("local_cluster: a worker exited with code %d; stopping the job" % rc, file=sys.stderr)
                for _, p in procs:
                    if p.poll() is None:
                        p.terminate()
                break
            time.sleep(0.05)
        for p in workers:
            try:
                rc = p.wait(timeout=10) or rc
            except subprocess.TimeoutExpired:
                p.kill()
                rc = rc or 1
        for role, p in procs:
            if role != "worker":
                try:
                    p.wait(timeout=30)        # servers/scheduler leave once every worker said goodbye
                except subprocess.TimeoutExpired:
                    p.terminate()
    finally:
        for _, p in procs:
            if p.poll() is None:
                p.kill()
    return rc


if __name__ == "__main__":
    sys.exit(main())
