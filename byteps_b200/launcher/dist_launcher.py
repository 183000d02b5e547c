#!/usr/bin/env python
"""Launch a whole job over ssh from host files.

Parity: /root/reference/launcher/dist_launcher.py:78-151 - one scheduler (on
the first server host), ``len(server_hosts)`` servers and ``len(worker_hosts)``
workers, each started with the right ``DMLC_*`` environment through ssh; extra
environment can be forwarded with ``--env KEY:VALUE``.

    python -m byteps_b200.launcher.dist_launcher -WH workers.txt -SH servers.txt \\
        --scheduler-ip 10.0.0.1 --scheduler-port 1234 bpslaunch python train.py
"""
import argparse
import shlex
import subprocess
import sys
import threading


def read_hosts(path):
    hosts = []
    with open(path) as f:
        for line in f:
            h = line.strip()
            if h and not h.startswith("#"):
                hosts.append(h)
    return hosts


def build_env(role, num_workers, num_servers, sched_ip, sched_port, worker_id=None, extra=None):
    env = {"DMLC_ROLE": role, "DMLC_NUM_WORKER": str(num_workers), "DMLC_NUM_SERVER": str(num_servers),
           "DMLC_PS_ROOT_URI": sched_ip, "DMLC_PS_ROOT_PORT": str(sched_port)}
    if worker_id is not None:
        env["DMLC_WORKER_ID"] = str(worker_id)
    env.update(extra or {})
    return env


def ssh_command(host, env, command, username=None, port=22):
    exports = " ".join("export %s=%s;" % (k, shlex.quote(str(v))) for k, v in env.items())
    target = ("%s@%s" % (username, host)) if username else host
    return ["ssh", "-o", "StrictHostKeyChecking=no", "-p", str(port), target, exports + " " + command]


def plan(args):
    workers, servers = read_hosts(args.worker_hostfile), read_hosts(args.server_hostfile)
    extra = dict(kv.split(":", 1) for kv in args.env)
    if args.gpus_per_worker:
        # every GPU process of a worker box is a transport-level node: scheduler and servers must expect
        # DMLC_NUM_WORKER x BYTEPS_LOCAL_SIZE of them (in the reference only the box's root process talks to servers)
        extra.setdefault("BYTEPS_LOCAL_SIZE", str(args.gpus_per_worker))
    nw, ns = len(workers), len(servers)
    jobs = [(servers[0] if servers else args.scheduler_ip,
             build_env("scheduler", nw, ns, args.scheduler_ip, args.scheduler_port, extra=extra),
             "python -c 'import byteps_b200.server'")]
    for h in servers:
        jobs.append((h, build_env("server", nw, ns, args.scheduler_ip, args.scheduler_port, extra=extra),
                     "python -c 'import byteps_b200.server'"))
    cmd = " ".join(shlex.quote(c) for c in args.command)
    for i, h in enumerate(workers):
        jobs.append((h, build_env("worker", nw, ns, args.scheduler_ip, args.scheduler_port, worker_id=i, extra=extra),
                     cmd))
    return jobs


def main(argv=None):
    ap = argparse.ArgumentParser(description="Launch a distributed byteps_b200 job over ssh")
    ap.add_argument("-WH", "--worker-hostfile", required=True)
    ap.add_argument("-SH", "--server-hostfile", required=True)
    ap.add_argument("--scheduler-ip", required=True)
    ap.add_argument("--scheduler-port", type=int, required=True)
    ap.add_argument("--username", default=None)
    ap.add_argument("--ssh-port", type=int, default=22)
    ap.add_argument("--env", action="append", default=[], help="KEY:VALUE forwarded to every process")
    ap.add_argument("--gpus-per-worker", type=int, default=0,
                    help="GPU processes per worker box (exported as BYTEPS_LOCAL_SIZE to every role; needed as soon "
                         "as a box runs more than one)")
    ap.add_argument("--dry-run", action="store_true")
    ap.add_argument("command", nargs=argparse.REMAINDER)
    args = ap.parse_args(argv)
    jobs = plan(args)
    cmds = [ssh_command(h, env, c, args.username, args.ssh_port) for h, env, c in jobs]
    if args.dry_run:
        for c in cmds:
            print(" ".join(shlex.quote(x) for x in c))
        return 0
    codes = [0] * len(cmds)

    def run(i):
        codes[i] = subprocess.call(cmds[i])

    ts = [threading.Thread(target=run, args=(i,)) for i in range(len(cmds))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return max(codes, key=abs)


if __name__ == "__main__":
    sys.exit(main())
