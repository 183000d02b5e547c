import os
import subprocess
import sys

from byteps_b200.launcher import dist_launcher, launch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpu_allocation_splits_numa_nodes():
    nodes = [list(range(0, 8)), list(range(8, 16))]
    alloc = launch.allocate_cpu(4, nodes=nodes, multithreaded=False, blacklist=set())
    assert alloc == [[0, 1, 2, 3], [8, 9, 10, 11], [4, 5, 6, 7], [12, 13, 14, 15]]
    alloc = launch.allocate_cpu(2, nodes=nodes, multithreaded=True, blacklist={1})
    assert alloc == [[0, 2, 3], [8, 9, 10, 11]]          # SMT: first half of each node, minus the blacklist
    assert launch.allocate_cpu(2, nodes=[], multithreaded=False, blacklist=set()) is None


def test_worker_command_sets_local_rank_env(monkeypatch):
    monkeypatch.setenv("BYTEPS_NUMA_ON", "0")
    cmd, env = launch.worker_command(3, 8, ["python", "train.py"], cores=[1, 2])
    assert cmd == ["python", "train.py"]
    assert env["BYTEPS_LOCAL_RANK"] == "3" and env["BYTEPS_LOCAL_SIZE"] == "8" and env["DMLC_ROLE"] == "worker"


def test_bpslaunch_spawns_one_process_per_gpu(tmp_path):
    env = dict(os.environ, NVIDIA_VISIBLE_DEVICES="0,1,2", BYTEPS_NUMA_ON="0", DMLC_ROLE="worker", PYTHONPATH=ROOT)
    script = "import os; open(os.path.join(r'%s', os.environ['BYTEPS_LOCAL_RANK']), 'w').write(os.environ['BYTEPS_LOCAL_SIZE'])" % tmp_path
    rc = subprocess.call([sys.executable, os.path.join(ROOT, "bin", "bpslaunch"), sys.executable, "-c", script], env=env)
    assert rc == 0
    assert sorted(os.listdir(tmp_path)) == ["0", "1", "2"]
    assert open(os.path.join(tmp_path, "1")).read() == "3"


def test_dist_launcher_plan(tmp_path, capsys):
    wh, sh = tmp_path / "w.txt", tmp_path / "s.txt"
    wh.write_text("10.0.0.2\n10.0.0.3\n")
    sh.write_text("10.0.0.1\n")
    rc = dist_launcher.main(["-WH", str(wh), "-SH", str(sh), "--scheduler-ip", "10.0.0.1", "--scheduler-port", "1234",
                             "--env", "BYTEPS_LOG_LEVEL:INFO", "--dry-run", "bpslaunch", "python", "train.py"])
    assert rc == 0
    out = capsys.readouterr().out.strip().splitlines()
    assert len(out) == 4                                    # scheduler + 1 server + 2 workers
    assert "DMLC_ROLE=scheduler" in out[0] and "DMLC_ROLE=server" in out[1]
    assert "DMLC_WORKER_ID=1" in out[3] and "DMLC_NUM_WORKER=2" in out[3] and "BYTEPS_LOG_LEVEL=INFO" in out[3]
    assert out[3].endswith("bpslaunch python train.py'") or "bpslaunch python train.py" in out[3]


def test_local_cluster_runs_full_ps_job(tmp_path):
    """scheduler + server + 2 one-device workers on this host through the launcher."""
    from byteps_b200.launcher import local_cluster

    envs = local_cluster.build_envs(2, 1, 4321, gpus_per_worker=2, base={})
    assert [r for r, _ in envs] == ["scheduler", "server", "worker", "worker", "worker", "worker"]
    assert envs[-1][1]["DMLC_WORKER_ID"] == "1" and envs[-1][1]["BYTEPS_LOCAL_RANK"] == "1"
    script = tmp_path / "train.py"
    script.write_text(
        "import os, torch\n"
        "import byteps_b200.torch as bps\n"
        "bps.init()\n"
        "t = torch.full((1000,), float(bps.rank() + 1))\n"
        "bps.push_pull_inplace(t, average=False, name='x')\n"
        "assert torch.all(t == 3.0), t[:3]\n"
        "open(os.path.join(r'%s', 'ok%%d' %% bps.rank()), 'w').write(str(bps.size()))\n"
        "bps.shutdown()\n" % tmp_path)
    env = dict(os.environ, PYTHONPATH=ROOT)
    rc = subprocess.call([sys.executable, "-m", "byteps_b200.launcher.local_cluster", "-n", "2", "-s", "1",
                          sys.executable, str(script)], env=env, timeout=120)
    assert rc == 0
    assert sorted(f for f in os.listdir(tmp_path) if f.startswith("ok")) == ["ok0", "ok1"]


def test_bpslaunch_pins_cpus_without_numactl(tmp_path, monkeypatch):
    """With BYTEPS_NUMA_ON=1 and no numactl binary the launcher pins each worker itself."""
    import shutil

    if shutil.which("numactl"):
        import pytest

        pytest.skip("numactl present: the launcher delegates to it")
    allowed = sorted(os.sched_getaffinity(0))
    if len(allowed) < 2:
        import pytest

        pytest.skip("needs two CPUs")
    a, b = allowed[0], allowed[1]
    env = dict(os.environ, NVIDIA_VISIBLE_DEVICES="0,1", BYTEPS_NUMA_ON="1", DMLC_ROLE="worker", PYTHONPATH=ROOT,
               BYTEPS_VISIBLE_CPU_CORES="%d:%d" % (a, b))
    script = ("import os; open(os.path.join(r'%s', os.environ['BYTEPS_LOCAL_RANK']), 'w')"
              ".write(','.join(map(str, sorted(os.sched_getaffinity(0)))))" % tmp_path)
    rc = subprocess.call([sys.executable, os.path.join(ROOT, "bin", "bpslaunch"), sys.executable, "-c", script], env=env)
    assert rc == 0
    assert open(os.path.join(tmp_path, "0")).read() == str(a)
    assert open(os.path.join(tmp_path, "1")).read() == str(b)


def test_doctor_report(monkeypatch):
    from byteps_b200 import doctor

    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "BYTEPS_LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("DMLC_NUM_WORKER", "2")
    monkeypatch.setenv("DMLC_NUM_SERVER", "1")
    info = doctor.collect()
    assert info["selected_backend"] == "ps" and info["world"]["size"] == 2 and info["world"]["distributed"]
    assert "topk_compressor_type" in info["compressors"] and info["core_module"].endswith(".so")
    assert info["env"]["DMLC_NUM_SERVER"] == "1"
    assert doctor.main([]) == 0


def test_local_size_reaches_every_role(tmp_path, capsys):
    from byteps_b200.launcher import local_cluster

    envs = local_cluster.build_envs(1, 1, 4000, gpus_per_worker=4, base={})
    assert all(e["BYTEPS_LOCAL_SIZE"] == "4" for _, e in envs)
    wh, sh = tmp_path / "w.txt", tmp_path / "s.txt"
    wh.write_text("10.0.0.2\n")
    sh.write_text("10.0.0.1\n")
    dist_launcher.main(["-WH", str(wh), "-SH", str(sh), "--scheduler-ip", "10.0.0.1", "--scheduler-port", "1",
                        "--gpus-per-worker", "8", "--dry-run", "bpslaunch", "python", "t.py"])
    out = capsys.readouterr().out.strip().splitlines()
    assert len(out) == 3 and all("BYTEPS_LOCAL_SIZE=8" in line for line in out)


def test_local_cluster_stops_the_job_when_a_worker_dies(tmp_path):
    """One worker crashes before init: the launcher must not leave the others waiting in the rendezvous forever."""
    import time

    script = tmp_path / "train.py"
    script.write_text(
        "import os, sys\n"
        "if os.environ['DMLC_WORKER_ID'] == '1':\n"
        "    sys.exit(7)\n"
        "import byteps_b200.torch as bps\n"
        "bps.init()\n"
        "bps.shutdown()\n")
    env = dict(os.environ, PYTHONPATH=ROOT)
    t0 = time.time()
    rc = subprocess.call([sys.executable, "-m", "byteps_b200.launcher.local_cluster", "-n", "2", "-s", "1",
                          sys.executable, str(script)], env=env, timeout=100)
    assert rc == 7
    assert time.time() - t0 < 60


def test_bpslaunch_stops_siblings_when_a_rank_dies(tmp_path):
    import time

    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys, time\n"
        "if os.environ['BYTEPS_LOCAL_RANK'] == '1':\n"
        "    sys.exit(5)\n"
        "time.sleep(120)\n")
    env = dict(os.environ, PYTHONPATH=ROOT, DMLC_ROLE="worker", DMLC_NUM_WORKER="1", DMLC_NUM_SERVER="0",
               DMLC_WORKER_ID="0", DMLC_PS_ROOT_URI="127.0.0.1", DMLC_PS_ROOT_PORT="1", NVIDIA_VISIBLE_DEVICES="0,1",
               BYTEPS_NUMA_ON="0")
    t0 = time.time()
    rc = subprocess.call([sys.executable, "-m", "byteps_b200.launcher.launch", sys.executable, str(script)], env=env,
                         timeout=100)
    assert rc == 5 and time.time() - t0 < 30


def test_children_die_with_a_killed_launcher(tmp_path):
    """kill -9 of the launcher must not leave scheduler / server / worker processes behind."""
    import signal
    import time

    script = tmp_path / "w.py"
    script.write_text("import os, time\n"
                      "open(os.path.join(r'%s', 'pid%%d' %% os.getpid()), 'w').close()\n"
                      "time.sleep(300)\n" % tmp_path)
    env = dict(os.environ, PYTHONPATH=ROOT)
    p = subprocess.Popen([sys.executable, "-m", "byteps_b200.launcher.local_cluster", "-n", "2", "-s", "1",
                          sys.executable, str(script)], env=env)
    deadline = time.time() + 60
    while time.time() < deadline and len([f for f in os.listdir(tmp_path) if f.startswith("pid")]) < 2:
        time.sleep(0.1)
    pids = [int(f[3:]) for f in os.listdir(tmp_path) if f.startswith("pid")]
    assert len(pids) == 2
    children = subprocess.run(["ps", "-o", "pid=", "--ppid", str(p.pid)], capture_output=True, text=True).stdout.split()
    assert len(children) == 4          # scheduler, server, two workers
    p.send_signal(signal.SIGKILL)
    p.wait()
    deadline = time.time() + 20
    alive = children
    while time.time() < deadline and alive:
        alive = [c for c in children if os.path.exists("/proc/%s" % c)
                 and "Z" not in open("/proc/%s/stat" % c).read().split(")")[-1].split()[0]]
        time.sleep(0.2)
    assert not alive, alive
