"""How fast can ONE CPU thread copy into / out of the staging memory the CPU-server mode uses?
heap -> heap, heap -> POSIX shm, and the same with the shm window cudaHostRegister'ed (what the worker does);
then the PCIe rates of 4 MB staged copies from / to that memory, one direction and both at once."""
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
    # PCIe staging rates with the raw runtime (torch would bounce a copy from memory it does not know is pinned)
    from cuda.bindings import runtime as rt

    def ck(r):
        assert int(r[0]) == 0, r
        return r[1] if len(r) == 2 else r[1:]

    dev = ck(rt.cudaMalloc(N))
    dev2 = ck(rt.cudaMalloc(N))
    host = ck(rt.cudaHostAlloc(N, 0))
    s1, s2 = ck(rt.cudaStreamCreate()), ck(rt.cudaStreamCreate())
    H2D, D2H = rt.cudaMemcpyKind.cudaMemcpyHostToDevice, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost

    def timed(jobs, reps=5):
        """jobs: [(dst, src, kind, stream, chunk)] issued together; wall time until all streams are idle"""
        best = 1e9
        for _ in range(reps):
            for st in (s1, s2):
                rt.cudaStreamSynchronize(st)
            t0 = time.perf_counter()
            for dst, src, kind, st, chunk in jobs:
                for off in range(0, N, chunk):
                    rt.cudaMemcpyAsync(dst + off, src + off, min(chunk, N - off), kind, st)
            for st in (s1, s2):
                rt.cudaStreamSynchronize(st)
            best = min(best, time.perf_counter() - t0)
        return N / best / 1e9

    CH = 4096000
    print("H2D cudaHostAlloc, one copy      %.1f GB/s" % timed([(dev, host, H2D, s1, N)]))
    print("H2D cudaHostAlloc, 4 MB copies   %.1f GB/s" % timed([(dev, host, H2D, s1, CH)]))
    print("D2H cudaHostAlloc, 4 MB copies   %.1f GB/s" % timed([(host, dev, D2H, s1, CH)]))
    print("H2D registered shm, 4 MB copies  %.1f GB/s" % timed([(dev, addr, H2D, s1, CH)]))
    print("D2H registered shm, 4 MB copies  %.1f GB/s" % timed([(addr, dev, D2H, s1, CH)]))
    print("duplex (D2H + H2D together), per direction, cudaHostAlloc   %.1f GB/s" % timed(
        [(host, dev, D2H, s1, CH), (dev2, host, H2D, s2, CH)]))
    print("duplex (D2H + H2D together), per direction, registered shm  %.1f GB/s" % timed(
        [(addr, dev, D2H, s1, CH), (dev2, addr, H2D, s2, CH)]))
os.unlink("/dev/shm/bps_probe_a")
os._exit(0)
