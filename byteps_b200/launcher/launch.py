#!/usr/bin/env python
"""bpslaunch - start one training process per GPU (worker role) or run the
server / scheduler role.

Parity: /root/reference/launcher/launch.py:49-277 - role dispatch from
``DMLC_ROLE``, one subprocess per entry of ``NVIDIA_VISIBLE_DEVICES`` with
``BYTEPS_LOCAL_RANK`` / ``BYTEPS_LOCAL_SIZE``, NUMA-aware CPU pinning through
``numactl --physcpubind`` (``BYTEPS_NUMA_ON``, ``BYTEPS_VISIBLE_CPU_CORES``,
``BYTEPS_CPU_BLACKLIST``, ``BYTEPS_NUMA_DEFAULT_QUOTA``), optional gdb wrapper
(``BYTEPS_ENABLE_GDB``), trace directory creation, server/scheduler =
``import byteps_b200.server``.

Our per-GPU processes are symmetric (there is no "root" GPU that alone talks
to the servers), so CPU cores are split evenly instead of giving the root an
extra quota.

    bpslaunch python train.py --arg ...
"""
import os
import re
import shutil
import subprocess
import sys
import threading

from .local_cluster import die_with_parent

NUMA_PATH = "/sys/devices/system/node"


def numa_nodes():
    """[[cpu ids of node 0], [cpu ids of node 1], ...] (physical view of sysfs)."""
    out = []
    if not os.path.isdir(NUMA_PATH):
        return out
    for node in sorted(d for d in os.listdir(NUMA_PATH) if re.fullmatch(r"node\d+", d)):
        cpus = sorted(int(m.group(1)) for m in (re.fullmatch(r"cpu(\d+)", f) for f in
                                                os.listdir(os.path.join(NUMA_PATH, node))) if m)
        if cpus:
            out.append(cpus)
    return out


def allocate_cpu(local_size, nodes=None, multithreaded=None, blacklist=None, quota=None):
    """Core lists, one per local rank.  Ranks are spread over NUMA nodes
    round-robin; with SMT only the first half of each node's ids (the physical
    cores) is used, like the reference does."""
    nodes = [list(n) for n in (nodes if nodes is not None else numa_nodes())]
    if not nodes or local_size <= 0:
        return None
    if multithreaded is None:
        multithreaded = os.getenv("BYTEPS_MULTITHREADED_CPU", "1").lower() in ("1", "true")
    if blacklist is None:
        blacklist = {int(x) for x in os.getenv("BYTEPS_CPU_BLACKLIST", "-1").split(",") if x.strip()}
    if multithreaded:
        nodes = [n[:max(1, len(n) // 2)] for n in nodes]
    nodes = [[c for c in n if c not in blacklist] for n in nodes]
    per_node = [0] * len(nodes)
    for r in range(local_size):
        per_node[r % len(nodes)] += 1
    out, cursor = [], [0] * len(nodes)
    env_quota = int(os.getenv("BYTEPS_NUMA_DEFAULT_QUOTA", "0"))
    # the reference gives its "root" rank (the last one: it drives NCCL and the PS traffic) a larger quota;
    # ranks are symmetric here, the knob is honoured for the last rank when set
    root_quota = int(os.getenv("BYTEPS_NUMA_ROOT_QUOTA", "0"))
    for r in range(local_size):
        ni = r % len(nodes)
        q = quota or env_quota or max(1, len(nodes[ni]) // max(1, per_node[ni]))
        if root_quota and r == local_size - 1:
            q = root_quota
        cores = nodes[ni][cursor[ni]:cursor[ni] + q]
        cursor[ni] += q
        out.append(cores)
    return out


def check_env():
    role = os.environ.get("DMLC_ROLE", "worker").lower()
    assert role in ("worker", "server", "scheduler"), "DMLC_ROLE must be worker|server|scheduler"
    required = []
    if role != "worker" or int(os.environ.get("DMLC_NUM_WORKER", "1")) > 1 or \
            os.environ.get("BYTEPS_FORCE_DISTRIBUTED", "0") not in ("0", ""):
        required = ["DMLC_NUM_WORKER", "DMLC_NUM_SERVER", "DMLC_PS_ROOT_URI", "DMLC_PS_ROOT_PORT"]
    if role == "worker" and int(os.environ.get("DMLC_NUM_WORKER", "1")) > 1:
        required.append("DMLC_WORKER_ID")
    missing = [e for e in required if e not in os.environ]
    if missing:
        print("bpslaunch: missing environment variables: " + ", ".join(missing))
        sys.exit(1)
    return role


def visible_gpus():
    v = os.environ.get("NVIDIA_VISIBLE_DEVICES") or os.environ.get("CUDA_VISIBLE_DEVICES")
    if v and v not in ("all", "void", "none"):
        return [x for x in v.split(",") if x != ""]
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
        n = len([ln for ln in out.splitlines() if ln.startswith("GPU ")])
        return [str(i) for i in range(n)] or ["0"]
    except Exception:  # noqa: BLE001
        return ["0"]


def worker_command(local_rank, local_size, argv, cores=None):
    env = dict(os.environ)
    env["BYTEPS_LOCAL_RANK"] = str(local_rank)
    env["BYTEPS_LOCAL_SIZE"] = str(local_size)
    env.setdefault("DMLC_ROLE", "worker")
    if int(env.get("DMLC_NUM_SERVER", "0") or 0) > 0:
        # CPU-server mode: the framework's OpenMP team must not spin between parallel regions - it would take the
        # cores the transport / server threads need (1 worker + 1 colocated server, 10 MB CPU tensor: 55 -> 5 ms per
        # push_pull on an 8-core host).  libgomp reads this when it is loaded, so it has to be in the environment.
        env.setdefault("OMP_WAIT_POLICY", "passive")
    if env.get("BYTEPS_TRACE_ON", "") == "1":
        d = os.path.join(env.get("BYTEPS_TRACE_DIR", "./trace"), str(local_rank))
        os.makedirs(d, exist_ok=True)
    cmd = list(argv)
    if env.get("BYTEPS_ENABLE_GDB", "0") == "1":
        cmd = ["gdb", "-ex", "run", "-ex", "bt", "-batch", "--args"] + cmd
    if cores and env.get("BYTEPS_NUMA_ON", "1") == "1":
        vis = env.get("BYTEPS_VISIBLE_CPU_CORES")
        bind = vis.split(":")[local_rank] if vis else ",".join(str(c) for c in cores)
        if shutil.which("numactl"):
            cmd = ["numactl", "--physcpubind", bind] + cmd
        else:
            # no numactl in the image: the launcher pins the child itself (sched_setaffinity before exec)
            env["_BYTEPS_CPU_AFFINITY"] = bind
    return cmd, env


def _parse_cpu_list(spec):
    cores = set()
    for part in spec.split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            lo, hi = part.split("-", 1)
            cores.update(range(int(lo), int(hi) + 1))
        else:
            cores.add(int(part))
    return cores


def _pin_self(spec):
    try:
        allowed = os.sched_getaffinity(0)
        want = _parse_cpu_list(spec) & allowed
        if want:
            os.sched_setaffinity(0, want)
    except (AttributeError, OSError, ValueError):
        pass


def launch_workers(argv):
    gpus = visible_gpus()
    local_size = int(os.environ.get("BYTEPS_LOCAL_SIZE", len(gpus)))
    alloc = allocate_cpu(local_size) if os.environ.get("BYTEPS_NUMA_ON", "1") == "1" else None
    procs, codes = [], [0] * local_size

    lock = threading.Lock()

    def run(i):
        cmd, env = worker_command(i, local_size, argv, alloc[i] if alloc else None)
        aff = env.pop("_BYTEPS_CPU_AFFINITY", None)
        def pre(a=aff):
            die_with_parent()
            if a:
                _pin_self(a)

        p = subprocess.Popen(cmd, env=env, preexec_fn=pre)
        with lock:
            procs.append(p)
        codes[i] = p.wait()
        if codes[i] != 0 and os.environ.get("BYTEPS_LAUNCH_KEEP_GOING", "0") != "1":
            # a dead rank never reaches the collectives its siblings wait in: stop them instead of hanging the job
            # (the reference's launcher just joins its threads)
            with lock:
                others = [q for q in procs if q is not p and q.poll() is None]
            if others:
                print("bpslaunch: local rank %d exited with code %d; stopping the other %d local worker(s)" % (
                    i, codes[i], len(others)), file=sys.stderr)
            for q in others:
                q.terminate()

    ts = [threading.Thread(target=run, args=(i,)) for i in range(local_size)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    first_bad = [c for c in codes if c not in (0, -15)]      # -15: a sibling we terminated ourselves
    return first_bad[0] if first_bad else max(codes, key=abs)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    role = check_env()
    if role == "worker":
        if not argv:
            print("usage: bpslaunch COMMAND [ARGS...]")
            return 2
        return launch_workers(argv)
    # server / scheduler
    env = dict(os.environ)
    cmd = [sys.executable, "-c", "import byteps_b200.server"]
    if env.get("BYTEPS_ENABLE_GDB", "0") == "1":
        cmd = ["gdb", "-ex", "run", "-ex", "bt", "-batch", "--args"] + cmd
    return subprocess.call(cmd, env=env, preexec_fn=die_with_parent)


if __name__ == "__main__":
    sys.exit(main())
