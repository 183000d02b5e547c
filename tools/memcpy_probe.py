"""How fast can ONE CPU thread copy into / out of the staging memory the CPU-server mode uses?
heap -> heap, heap -> POSIX shm, and the same with the shm window cudaHostRegister'ed (what the worker does)."""
import ctypes
import mmap
import os
import time

import numpy as np
import torch

N = 100 * 1000 * 1000


def shm(name):
    fd = os.open("/dev/shm/" + name, os.O_CREAT | os.O_RDWR, 0o600)
    os.ftruncate(fd, N)
    m = mmap.mmap(fd, N)
    return fd, m, np.frombuffer(m, dtype=np.uint8)


def rate(dst, src, reps=5):
    dst[:] = src
    t0 = time.perf_counter()
    for _ in range(reps):
        dst[:] = src
    return N * reps / (time.perf_counter() - t0) / 1e9


a = np.ones(N, dtype=np.uint8)
b = np.zeros(N, dtype=np.uint8)
print("heap -> heap           %.1f GB/s" % rate(b, a))
fd, m, w = shm("bps_probe_a")
print("heap -> shm            %.1f GB/s" % rate(w, a))
print("shm  -> heap           %.1f GB/s" % rate(b, w))
if torch.cuda.is_available():
    torch.cuda.init()
    addr = ctypes.addressof(ctypes.c_char.from_buffer(m))
    rc = torch.cuda.cudart().cudaHostRegister(addr, N, 0)
    print("cudaHostRegister rc", int(rc))
    print("heap -> pinned shm     %.1f GB/s" % rate(w, a))
    print("pinned shm -> heap     %.1f GB/s" % rate(b, w))
    # a second mapping of the same file (what the server process sees)
    m2 = mmap.mmap(fd, N)
    w2 = np.frombuffer(m2, dtype=np.uint8)
    print("heap -> 2nd mapping    %.1f GB/s" % rate(w2, a))
    print("2nd mapping -> heap    %.1f GB/s" % rate(b, w2))
    g = torch.empty(N, dtype=torch.uint8, device="cuda")
    t = torch.from_numpy(w)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        t.copy_(g, non_blocking=True)
    torch.cuda.synchronize()
    print("D2H into pinned shm    %.1f GB/s" % (N * 5 / (time.perf_counter() - t0) / 1e9))
    t0 = time.perf_counter()
    for _ in range(5):
        g.copy_(t, non_blocking=True)
    torch.cuda.synchronize()
    print("H2D from pinned shm    %.1f GB/s" % (N * 5 / (time.perf_counter() - t0) / 1e9))
os.unlink("/dev/shm/bps_probe_a")
os._exit(0)
