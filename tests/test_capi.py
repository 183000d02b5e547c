"""The C API (csrc/capi/byteps_c_api.h): a whole job - scheduler, server, two workers - as plain C processes linked
against libbyteps_b200.so, no Python in any role; and the same entry points through ctypes in a one-process job."""
import ctypes
import os
import subprocess
import sys

import pytest

from _mp import free_port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "byteps_b200", "libbyteps_b200.so")


def _build_lib():
    sys.path.insert(0, ROOT)
    from byteps_b200 import _build

    _build.build_core()         # incremental: a no-op when the objects are up to date
    assert os.path.exists(LIB)


@pytest.mark.parametrize("van", ["tcp", "shm"])
@pytest.mark.skipif(any(k in os.environ.get("LD_PRELOAD", "") for k in ("asan", "tsan", "ubsan")),
                    reason="tools/sanitize.sh: a plain C program cannot link the instrumented library")
def test_c_job_without_python(tmp_path, van):
    _build_lib()
    exe = str(tmp_path / "capi_job")
    subprocess.check_call(["gcc", "-O1", "-o", exe, os.path.join(ROOT, "tests", "native", "capi_job.c"),
                           "-I" + os.path.join(ROOT, "byteps_b200", "csrc"), "-L" + os.path.dirname(LIB),
                           "-lbyteps_b200", "-lm", "-Wl,-rpath," + os.path.dirname(LIB)])
    out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libpython" not in out
    port = free_port()
    base = dict(os.environ, DMLC_NUM_WORKER="2", DMLC_NUM_SERVER="1", DMLC_PS_ROOT_URI="127.0.0.1",
                DMLC_PS_ROOT_PORT=str(port), DMLC_PS_VAN_TYPE=van, BYTEPS_MIN_COMPRESS_BYTES="0")
    procs = [subprocess.Popen([exe, "server"], env=dict(base, DMLC_ROLE=r)) for r in ("scheduler", "server")]
    workers = [subprocess.Popen([exe, "worker"], env=dict(base, DMLC_ROLE="worker", DMLC_WORKER_ID=str(w)),
                                stdout=subprocess.PIPE, text=True) for w in range(2)]
    try:
        outs = [w.communicate(timeout=120)[0] for w in workers]
        assert [w.returncode for w in workers] == [0, 0], outs
        assert all("ok" in o for o in outs)
        for p in procs:
            assert p.wait(timeout=60) == 0
    finally:
        for p in procs + workers:
            if p.poll() is None:
                p.kill()


def test_c_api_single_process_through_ctypes(monkeypatch):
    _build_lib()
    for k in ("DMLC_NUM_WORKER", "DMLC_NUM_SERVER", "DMLC_WORKER_ID", "DMLC_ROLE", "BYTEPS_LOCAL_RANK",
              "BYTEPS_LOCAL_SIZE"):
        monkeypatch.delenv(k, raising=False)
    lib = ctypes.CDLL(LIB)
    lib.byteps_last_error.restype = ctypes.c_char_p
    lib.byteps_push_pull.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int, ctypes.c_int]
    assert lib.byteps_push_pull(b"early", None, 0, 0, 1, 0, 0) < 0
    assert b"byteps_init" in lib.byteps_last_error()
    assert lib.byteps_init() == 0 and lib.byteps_size() == 1 and lib.byteps_rank() == 0
    k0, k1 = lib.byteps_declare_tensor(b"a"), lib.byteps_declare_tensor(b"b")
    assert k1 == k0 + 1 and lib.byteps_declare_tensor(b"a") == k0
    buf = (ctypes.c_float * 8)(*range(8))
    h = lib.byteps_push_pull(b"a", buf, 32, 0, 1, 0, 0)
    assert h <= -2 and lib.byteps_poll(h) == 1 and lib.byteps_wait(h) == 0
    assert list(buf) == list(range(8))                   # one worker: the average is the input
    assert lib.byteps_push_pull(b"a", buf, 30, 0, 1, 0, 0) < 0      # 30 bytes is not a whole number of floats
    assert lib.byteps_shutdown() == 0


@pytest.mark.gpu
def test_c_job_with_gpu_tensors(tmp_path):
    """byteps_push_pull_device: GPU buffers through the C API, no Python in any role (two workers on the
    same GPU when the box has one)."""
    _build_lib()
    sys.path.insert(0, ROOT)
    from byteps_b200 import _build

    _build.build_cuda()
    libdir = os.path.dirname(LIB)
    assert os.path.exists(os.path.join(libdir, "libbyteps_b200_cuda.so"))
    cudart = [d for d in _build._cudart_dirs() if os.path.exists(os.path.join(d, "libcudart.so.12"))][0]
    exe = str(tmp_path / "capi_gpu_job")
    subprocess.check_call(["gcc", "-O1", "-o", exe, os.path.join(ROOT, "tests", "native", "capi_gpu_job.c"),
                           "-I" + os.path.join(ROOT, "byteps_b200", "csrc"), "-I/usr/local/cuda/include",
                           "-L" + libdir, "-lbyteps_b200", "-L" + cudart, "-l:libcudart.so.12", "-lm",
                           "-Wl,-rpath," + libdir, "-Wl,-rpath," + cudart])
    assert "libpython" not in subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    port = free_port()
    base = dict(os.environ, DMLC_NUM_WORKER="2", DMLC_NUM_SERVER="1", DMLC_PS_ROOT_URI="127.0.0.1",
                DMLC_PS_ROOT_PORT=str(port), BYTEPS_ENABLE_IPC="1")
    procs = [subprocess.Popen([exe, "server"], env=dict(base, DMLC_ROLE=r)) for r in ("scheduler", "server")]
    workers = [subprocess.Popen([exe, "worker"], env=dict(base, DMLC_ROLE="worker", DMLC_WORKER_ID=str(w)),
                                stdout=subprocess.PIPE, text=True) for w in range(2)]
    try:
        outs = [w.communicate(timeout=180)[0] for w in workers]
        assert [w.returncode for w in workers] == [0, 0], outs
        assert all("ok" in o for o in outs)
        for p in procs:
            assert p.wait(timeout=60) == 0
    finally:
        for p in procs + workers:
            if p.poll() is None:
                p.kill()


def test_cuda_helper_library_has_no_undefined_symbols():
    """libbyteps_b200_cuda.so is dlopen'ed by the C API on a GPU box only; resolve every symbol NOW on the CPU box
    so a kernel it calls but does not link (gpu_stage.cc -> kernels/misc.cu) is caught without a GPU."""
    import torch  # noqa: F401  (puts the bundled libcudart on the loader's path)

    path = os.path.join(ROOT, "byteps_b200", "libbyteps_b200_cuda.so")
    if not os.path.exists(path):
        from byteps_b200 import _build

        _build.build_all(verbose=False)
    lib = ctypes.CDLL(path, mode=os.RTLD_NOW)
    for sym in ("byteps_cuda_stage_fns", "byteps_cuda_stage_create", "byteps_cuda_stage_destroy", "byteps_cuda_host_alloc"):
        assert hasattr(lib, sym), sym
