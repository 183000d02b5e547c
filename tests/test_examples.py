"""The shipped example scripts run end to end on CPU (2 processes, gloo)."""
import os
import subprocess
import sys

import pytest

from _mp import free_port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script, *args, nproc=2, timeout=240):
    env = dict(os.environ, PYTHONPATH=ROOT, OMP_NUM_THREADS="1")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "DMLC_ROLE", "DMLC_NUM_WORKER", "DMLC_NUM_SERVER"):
        env.pop(k, None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "examples", "pytorch", script)] + list(args)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    assert r.returncode == 0, r.stdout[-1500:] + "\n" + r.stderr[-3000:]
    return r.stdout


def test_mnist_example():
    out = _torchrun("train_mnist_byteps.py", "--epochs", "1", "--no-cuda")
    assert "loss" in out.lower()


def test_gradient_compression_example():
    out = _torchrun("train_gc_byteps.py", "--no-cuda", "--steps", "11", "--compressor", "topk", "--k", "0.05")
    assert "step  10" in out


def test_ddp_imagenet_example():
    out = _torchrun("train_imagenet_resnet_byteps_ddp.py", "--no-cuda", "--model", "resnet18", "--image-size", "32",
                    "--synthetic-samples", "16", "--batch-size", "4", "--epochs", "1")
    assert "mean loss" in out


@pytest.mark.parametrize("script", ["benchmark_byteps.py", "benchmark_byteps_ddp.py",
                                    "benchmark_cross_barrier_byteps.py"])
def test_synthetic_benchmarks(script):
    out = _torchrun(script, "--no-cuda", "--model", "resnet18", "--batch-size", "2", "--num-warmup-batches", "1",
                    "--num-batches-per-iter", "1", "--num-iters", "2", timeout=400)
    assert "img/sec" in out.lower()


def test_imagenet_example_checkpoint_resume(tmp_path):
    """Rank 0 saves; a second run finds the checkpoint, rank 0 loads it and everyone continues from it."""
    fmt = str(tmp_path / "ckpt-{epoch}.pt")
    common = ["--no-cuda", "--model", "resnet18", "--image-size", "32", "--batch-size", "2", "--steps-per-epoch", "2",
              "--batches-per-pushpull", "2", "--checkpoint-format", fmt]
    out1 = _torchrun("train_imagenet_resnet50_byteps.py", "--epochs", "1", *common)
    assert "epoch 1" in out1 and os.path.exists(fmt.format(epoch=1))
    out2 = _torchrun("train_imagenet_resnet50_byteps.py", "--epochs", "2", *common)
    assert "epoch 2" in out2 and "resumed from epoch 1" in out2 and "epoch 1:" not in out2


def test_bert_and_elastic_examples():
    out = _torchrun("train_bert_byteps.py", "--no-cuda", "--model", "bert_tiny", "--batch-size", "2", "--seq-len", "16",
                    "--steps", "2", "--warmup-steps", "1")
    assert "tokens/sec" in out
    out = _torchrun("elastic_benchmark_byteps.py")
    assert "before: 1.5" in out and "after : 1.5" in out


def test_bpslaunch_single_box_flow():
    """The reference's way of starting a job: DMLC_* variables + bpslaunch, one process per visible device."""
    env = dict(os.environ, PYTHONPATH=ROOT, NVIDIA_VISIBLE_DEVICES="0,1", BYTEPS_NUMA_ON="0", DMLC_ROLE="worker",
               DMLC_NUM_WORKER="1", DMLC_NUM_SERVER="0", DMLC_WORKER_ID="0", DMLC_PS_ROOT_URI="127.0.0.1",
               DMLC_PS_ROOT_PORT=str(free_port()), OMP_NUM_THREADS="1")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "BYTEPS_LOCAL_RANK", "BYTEPS_LOCAL_SIZE"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bin", "bpslaunch"), sys.executable,
                        os.path.join(ROOT, "examples", "pytorch", "train_mnist_byteps.py"), "--no-cuda", "--epochs", "1"],
                       capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-2000:]
    assert "averaged over 2 workers" in r.stdout


def test_bench_reference_arm_reports_unavailable():
    """The driver runs `bench.py --impl reference` first; without an installable reference it must print one JSON
    line with "unavailable" and exit 0 (also under torchrun-style environment variables)."""
    import json

    env = dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", MASTER_ADDR="127.0.0.1", MASTER_PORT="29999")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                          "--steps", "2", "--warmup", "1"], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and "unavailable" in d and "\n" not in d["unavailable"]
