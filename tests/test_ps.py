"""Parameter-server path: KV transport, summation server state machine, worker pipeline."""
import ctypes
import os

import numpy as np
import pytest

from _cluster import Cluster


def _core():
    from byteps_b200 import _native

    return _native.core()


def test_meta_codec_roundtrip():
    c = _core()
    ok, head, body, key, cmd, push, req, ctrl, port, n = c.meta_roundtrip(7, "hello world", 2 ** 40 + 3, 11)
    assert ok and head == 7 and body == "hello world" and key == 2 ** 40 + 3 and cmd == 11
    assert push and req and port == 7 and n > 0


def test_id_arithmetic_and_key_ranges():
    c = _core()
    assert c.Postoffice.worker_rank_to_id(0) == 9 and c.Postoffice.server_rank_to_id(0) == 8
    assert c.Postoffice.id_to_rank(9 + 2 * 5) == 5 and c.Postoffice.id_to_rank(8 + 2 * 3) == 3


@pytest.mark.parametrize("nw,ns", [(1, 1), (2, 2), (4, 2)])
def test_sync_push_pull_sums(nw, ns):
    c = _core()
    cl = Cluster(nw, ns).start()
    n = 300_000
    parts = [(c.make_key(0, 0), 0, 600_000), (c.make_key(0, 1), 600_000, n * 4 - 600_000)]
    results = {}

    def work(rank, w, po):
        for key, off, ln in parts:
            z = np.zeros(ln // 4, dtype=np.float32)
            w.init_key(key, z.ctypes.data, ln, c.F32)
        for it in range(3):
            x = (np.arange(n, dtype=np.float32) % 97) * (rank + 1) + it
            h = w.push_pull("g", x.ctypes.data, c.F32, parts, 0, 0, 1.0 / nw)
            assert w.wait(h)
            results[(rank, it)] = x
    cl.run_workers(work)
    for it in range(3):
        expect = sum((np.arange(n, dtype=np.float32) % 97) * (r + 1) + it for r in range(nw)) / nw
        for r in range(nw):
            np.testing.assert_allclose(results[(r, it)], expect, rtol=1e-6)
    assert sum(s.num_keys() for s in cl.servers) == 2
    # every pull response (2 partitions x 3 rounds per worker, + the init round) landed in place
    assert all(po.direct_recvs() >= 6 for po in cl.worker_pos)
    cl.stop()


@pytest.mark.parametrize("dtype", ["F16", "BF16", "F64", "I32", "I64", "U8"])
def test_dtypes(dtype):
    import torch

    c = _core()
    tdt = {"F16": torch.float16, "BF16": torch.bfloat16, "F64": torch.float64, "I32": torch.int32,
           "I64": torch.int64, "U8": torch.uint8}[dtype]
    cl = Cluster(2, 1).start()
    out = {}

    def work(rank, w, po):
        x = (torch.arange(1000) % 7 + rank).to(tdt)
        key = c.make_key(3, 0)
        w.init_key(key, x.data_ptr(), x.numel() * x.element_size(), getattr(c, dtype))
        h = w.push_pull("t", x.data_ptr(), getattr(c, dtype), [(key, 0, x.numel() * x.element_size())], 0, 0, 1.0)
        assert w.wait(h)
        out[rank] = x
    cl.run_workers(work)
    expect = ((torch.arange(1000) % 7) * 2 + 1).to(tdt)
    assert torch.equal(out[0], expect) and torch.equal(out[1], expect)
    cl.stop()


def test_priority_and_credits_order():
    """With a one-partition credit window, partitions leave in (priority desc, key asc) order."""
    c = _core()
    cl = Cluster(1, 1, extra={}, worker_kwargs={"credit_bytes": 4096}).start()
    order = []

    def work(rank, w, po):
        bufs = []
        for i in range(6):
            x = np.full(1024, float(i), dtype=np.float32)
            w.init_key(c.make_key(i, 0), x.ctypes.data, 4096, c.F32)
            bufs.append(x)
        hs = [w.push_pull("t%d" % i, bufs[i].ctypes.data, c.F32, [(c.make_key(i, 0), 0, 4096)], i % 3, 0, 1.0)
              for i in range(6)]
        for h in hs:
            assert w.wait(h)
        for i in range(6):
            assert np.all(bufs[i] == float(i))
        order.append(w.bytes_pushed())
    cl.run_workers(work)
    assert order[0] == 6 * 4096
    cl.stop()


def test_async_mode_accumulates():
    c = _core()
    cl = Cluster(2, 1, server_kwargs={"sync_mode": False}).start()
    out = {}

    def work(rank, w, po):
        key = c.make_key(0, 0)
        z = np.zeros(256, dtype=np.float32)
        w.init_key(key, z.ctypes.data, 1024, c.F32)
        po.barrier(0, c.GROUP_WORKER)
        for it in range(4):
            d = np.full(256, 1.0, dtype=np.float32)       # "weight delta"
            assert w.wait(w.push_pull("w", d.ctypes.data, c.F32, [(key, 0, 1024)], 0, 0, 1.0))
        po.barrier(0, c.GROUP_WORKER)
        d = np.zeros(256, dtype=np.float32)
        assert w.wait(w.push_pull("w", d.ctypes.data, c.F32, [(key, 0, 1024)], 0, 0, 1.0))
        out[rank] = d
    cl.run_workers(work)
    # the store accumulated every delta from both workers (8 pushes of 1.0) + zeros
    for r in (0, 1):
        assert np.all(out[r] == 8.0), out[r][:4]
    cl.stop()


@pytest.mark.parametrize("opts", [{"enable_schedule": True}, {"engine_blocking": True}, {"engine_threads": 1}])
def test_server_engine_variants(opts):
    c = _core()
    cl = Cluster(3, 1, server_kwargs=opts).start()
    out = {}

    def work(rank, w, po):
        keys = [c.make_key(k, 0) for k in range(8)]
        bufs = [np.full(5000, float(rank + k), dtype=np.float32) for k in range(8)]
        for k, b in zip(keys, bufs):
            w.init_key(k, b.ctypes.data, b.nbytes, c.F32)
        for it in range(3):
            for k in range(8):
                bufs[k][:] = rank + k + it
            hs = [w.push_pull("k%d" % k, bufs[k].ctypes.data, c.F32, [(keys[k], 0, bufs[k].nbytes)], -k, 0, 1.0)
                  for k in range(8)]
            for h in hs:
                assert w.wait(h)
            for k in range(8):
                assert np.all(bufs[k] == sum(r + k + it for r in range(3))), (k, it, bufs[k][:3])
        out[rank] = True
    cl.run_workers(work)
    assert len(out) == 3
    cl.stop()


@pytest.mark.parametrize("kw", [
    {"compressor_type": "onebit", "compressor_onebit_scaling": "true"},
    {"compressor_type": "topk", "compressor_k": "64"},
    {"compressor_type": "randomk", "compressor_k": "64", "seed": "13"},
    {"compressor_type": "dithering", "compressor_k": "4", "seed": "13"},
    {"compressor_type": "topk", "compressor_k": "0.01", "ef_type": "vanilla", "momentum_type": "nesterov",
     "momentum_mu": "0.9"},
])
def test_compressed_push_pull_matches_double_application(kw):
    """Worker compresses, server decompresses+sums+recompresses, worker decompresses:
    compare against applying independent compressor instances the same way
    (the reference's tests check the same two-stage contract in numpy)."""
    c = _core()
    nw, n = 2, 8192
    cl = Cluster(nw, 1, worker_kwargs={"min_compress_bytes": 0}).start()
    rng = np.random.RandomState(0)
    grads = [[rng.randn(n).astype(np.float32) for _ in range(3)] for _ in range(nw)]
    out = {}

    def work(rank, w, po):
        key = c.make_key(0, 0)
        z = np.zeros(n, dtype=np.float32)
        w.init_key(key, z.ctypes.data, n * 4, c.F32)
        w.register_compressor(key, kw, n * 4, c.F32)
        assert w.has_compressor(key)
        res = []
        for it in range(3):
            g = grads[rank][it].copy()
            assert w.wait(w.push_pull("g", g.ctypes.data, c.F32, [(key, 0, n * 4)], 0, 0, 1.0))
            res.append(g)
        out[rank] = res
    cl.run_workers(work)
    cl.stop()
    # model: per-worker compressor (with momentum/EF), server-side compressor (EF only, no momentum)
    wcomp = [c.Compressor(kw, n * 4, c.F32, False) for _ in range(nw)]
    scomp = c.Compressor(kw, n * 4, c.F32, True)
    buf = np.zeros(max(wcomp[0].max_compressed_bytes(), 64) + 64, dtype=np.uint8)
    for it in range(3):
        total = np.zeros(n, dtype=np.float32)
        for r in range(nw):
            g = grads[r][it].copy()
            m = wcomp[r].compress(g.ctypes.data, buf.ctypes.data)
            d = np.zeros(n, dtype=np.float32)
            scomp.decompress(buf.ctypes.data, m, d.ctypes.data)
            total += d
        m = scomp.compress(total.ctypes.data, buf.ctypes.data)
        final = np.zeros(n, dtype=np.float32)
        wcomp[0].decompress(buf.ctypes.data, m, final.ctypes.data)
        for r in range(nw):
            if kw["compressor_type"] in ("randomk", "dithering"):
                # server-side RNG stream is shared across pushes in arrival order: only check structure
                assert np.isfinite(out[r][it]).all()
                assert np.array_equal(out[0][it], out[r][it])
            else:
                np.testing.assert_allclose(out[r][it], final, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5])
def test_randomised_cluster_schedules(seed):
    """Random topology (1-3 workers, 1-3 servers), random tensor shapes/dtypes/partitions, priorities, lanes,
    scheduling credits and server options, tensors issued in shuffled priority order with several in flight:
    every round must produce the exact sum on every worker."""
    rng = np.random.RandomState(seed)
    c = _core()
    nw, ns = int(rng.randint(1, 4)), int(rng.randint(1, 4))
    extra = {"num_lanes": int(rng.randint(1, 5))}
    if rng.rand() < 0.5:
        extra.update(resend=True, resend_timeout_ms=300)
    server_kwargs = [{}, {"enable_schedule": True}, {"engine_threads": 1}, {"engine_blocking": True}][rng.randint(0, 4)]
    if seed % 2 == 0:
        extra["van_type"] = "shm"       # the socket-free transport under the same schedules
    elif seed == 5:
        extra["local"] = True           # Unix-domain sockets
    cl = Cluster(nw, ns, extra=extra, server_kwargs=server_kwargs).start()
    dtypes = [("F32", np.float32), ("F64", np.float64), ("I32", np.int32), ("I64", np.int64)]
    tensors = []
    for t in range(int(rng.randint(2, 7))):
        code, npdt = dtypes[rng.randint(0, len(dtypes))]
        n = int(rng.randint(1, 60000))
        es = np.dtype(npdt).itemsize
        bound = int(rng.choice([1 << 12, 1 << 15, 1 << 18])) // es * es
        parts, off, i = [], 0, 0
        while off < n * es:
            ln = min(bound, n * es - off)
            parts.append((c.make_key(t, i), off, ln))
            off += ln
            i += 1
        tensors.append((t, code, npdt, n, parts, int(rng.randint(-5, 5))))
    rounds = 3
    errs = []

    def work(rank, w, po):
        bufs = {}
        for t, code, npdt, n, parts, prio in tensors:
            for key, off, ln in parts:
                z = np.zeros(ln, dtype=np.uint8)
                w.init_key(key, z.ctypes.data, ln, getattr(c, code))
            bufs[t] = np.zeros(n, dtype=npdt)
        order = list(range(len(tensors)))
        for it in range(rounds):
            np.random.RandomState(seed * 100 + it).shuffle(order)       # same order on every worker
            hs = []
            for ti in order:
                t, code, npdt, n, parts, prio = tensors[ti]
                bufs[t][:] = (np.arange(n) % 17 + rank + it + t).astype(npdt)
                hs.append((t, w.push_pull("t%d" % t, bufs[t].ctypes.data, getattr(c, code), parts, prio, it, 1.0)))
            for t, h in hs:
                assert w.wait(h, 60_000)
            for t, code, npdt, n, parts, prio in tensors:
                want = sum((np.arange(n) % 17 + r + it + t) for r in range(nw)).astype(npdt)
                if not np.array_equal(bufs[t], want):
                    errs.append((rank, it, t, code, bufs[t][:4].tolist(), want[:4].tolist()))
    cl.run_workers(work)
    cl.stop()
    assert not errs, errs[:3]


@pytest.mark.parametrize("refuse_register", [False, True])
def test_device_pipeline_pulls_by_reference(refuse_register):
    """Colocated CPU-server mode, device tensors: the pull is answered with a REFERENCE into the server's
    shared-memory store, the worker 'DMAs' from there (no server->worker copy) and scales on the device side.
    Driven on the CPU through the host-memory stage table (core_bind_ext.cc: host_stage_fns)."""
    c = _core()
    nw = 2
    cl = Cluster(nw, 1, extra={"enable_ipc": True}).start()
    n = 500_000
    cut = 1_200_000
    parts = [(c.make_key(3, 0), 0, cut), (c.make_key(3, 1), cut, n * 4 - cut)]
    c.host_stage_reset(refuse_register)
    before = c.ipc_stats()["ref_responses"]
    results, stagings = {}, {}

    def work(rank, w, po):
        w.set_gpu_stage(c.host_stage_fns())
        for key, off, ln in parts:
            z = np.zeros(ln // 4, dtype=np.float32)
            w.init_key(key, z.ctypes.data, ln, c.F32)
        staging = np.zeros(n, dtype=np.float32)
        stagings[rank] = (staging.ctypes.data, staging.nbytes)
        for it in range(4):
            x = (np.arange(n, dtype=np.float32) % 89) * (rank + 1) + it
            out = np.zeros_like(x)
            h = w.push_pull_device("g", x.ctypes.data, out.ctypes.data, staging.ctypes.data, c.F32, parts, 0, 0,
                                   1.0 / nw)
            assert w.wait(h)
            results[(rank, it)] = out
    cl.run_workers(work)
    for it in range(4):
        expect = sum((np.arange(n, dtype=np.float32) % 89) * (r + 1) + it for r in range(nw)) / nw
        for r in range(nw):
            np.testing.assert_allclose(results[(r, it)], expect, rtol=1e-6)
    st = c.host_stage_stats()
    refs = c.ipc_stats()["ref_responses"] - before
    assert refs == nw * len(parts) * 4, refs                      # every pull was answered by reference
    assert st["h2d"] == nw * len(parts) * 4 and st["scaled"] == st["h2d"]
    inside = sum(1 for src, ln in st["h2d_sources"]
                 if any(base <= src < base + size for base, size in stagings.values()))
    if refuse_register:
        assert inside == st["h2d"]            # page-locking refused: staged through the worker's own window
    else:
        assert inside == 0                    # copied straight out of the server's store
        assert 1 <= st["registered"] <= 2 * len(parts) * nw       # once per mapping, not per round
    cl.stop()


def test_pull_by_reference_can_be_disabled(monkeypatch):
    c = _core()
    monkeypatch.setenv("BYTEPS_PS_PULL_BY_REF", "0")
    cl = Cluster(1, 1, extra={"enable_ipc": True}).start()
    n = 100_000
    parts = [(c.make_key(4, 0), 0, n * 4)]
    c.host_stage_reset()
    before = c.ipc_stats()["ref_responses"]

    def work(rank, w, po):
        w.set_gpu_stage(c.host_stage_fns())
        z = np.zeros(n, dtype=np.float32)
        w.init_key(parts[0][0], z.ctypes.data, n * 4, c.F32)
        x = np.arange(n, dtype=np.float32)
        out = np.zeros_like(x)
        staging = np.zeros_like(x)
        h = w.push_pull_device("g", x.ctypes.data, out.ctypes.data, staging.ctypes.data, c.F32, parts)
        assert w.wait(h)
        np.testing.assert_array_equal(out, x)
    cl.run_workers(work)
    assert c.ipc_stats()["ref_responses"] == before
    cl.stop()


def test_numa_hint_travels_with_the_init_push(monkeypatch):
    """BYTEPS_NUMA_AWARE=1 on a (faked) 2-node host: workers announce their GPU's node in the init push, the server
    places the store (mbind may be refused here: counted, not fatal) and the pushers count still decodes."""
    c = _core()
    monkeypatch.setenv("BYTEPS_NUMA_AWARE", "1")
    monkeypatch.setenv("BYTEPS_NUMA_FAKE_NODES", "2")
    nw = 2
    cl = Cluster(nw, 1, extra={"enable_ipc": True}).start()
    n = 200_000
    parts = [(c.make_key(5, 0), 0, n * 4)]
    results = {}

    def work(rank, w, po):
        w.set_numa_node(rank % 2)
        assert w.numa_node() == rank % 2
        z = np.zeros(n, dtype=np.float32)
        w.init_key(parts[0][0], z.ctypes.data, n * 4, c.F32)
        for it in range(2):
            x = np.full(n, rank + 1 + it, dtype=np.float32)
            h = w.push_pull("g", x.ctypes.data, c.F32, parts, 0, 0, 1.0)
            assert w.wait(h)
            results[(rank, it)] = x
    cl.run_workers(work)
    for it in range(2):
        for r in range(nw):
            np.testing.assert_array_equal(results[(r, it)], np.full(n, 3 + 2 * it, dtype=np.float32))
    cl.stop()


def _non_loopback_ipv4():
    import socket

    s = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    try:
        s.connect(("10.255.255.255", 1))
        ip = s.getsockname()[0]
    except OSError:
        ip = ""
    finally:
        s.close()
    return "" if ip.startswith("127.") else ip


def test_shared_memory_names_stay_on_the_host():
    """BYTEPS_ENABLE_IPC=1 with a server that advertises another address (= lives on another host as far as the
    worker can tell): no window is announced and no reference is asked for, the payload travels by value."""
    ip = _non_loopback_ipv4()
    if not ip:
        pytest.skip("no non-loopback IPv4 address to stand in for a remote host")
    c = _core()
    cl = Cluster(1, 1, extra={"enable_ipc": True}, server_host=ip).start()
    n = 300_000
    parts = [(c.make_key(6, 0), 0, n * 4)]
    c.host_stage_reset()
    before = c.ipc_stats()

    def work(rank, w, po):
        w.set_gpu_stage(c.host_stage_fns())
        z = np.zeros(n, dtype=np.float32)
        w.init_key(parts[0][0], z.ctypes.data, n * 4, c.F32)
        staging_name = "BytePS_ShM_%d_remote_test" % __import__("os").getpid()
        ptr = c.shm_create(staging_name, n * 4)       # a registered window, like comm/ps.py::_Staging
        try:
            for it in range(2):
                x = np.arange(n, dtype=np.float32) + it
                out = np.zeros_like(x)
                h = w.push_pull_device("g", x.ctypes.data, out.ctypes.data, ptr, c.F32, parts)
                assert w.wait(h)
                np.testing.assert_array_equal(out, x)
        finally:
            c.shm_release(staging_name)
    cl.run_workers(work)
    after = c.ipc_stats()
    assert after["ref_responses"] == before["ref_responses"]
    assert after["shm_responses"] == before["shm_responses"]
    assert after["payload_responses"] > before["payload_responses"]
    cl.stop()


@pytest.mark.parametrize("ipc", [True, False])
def test_host_push_pull_delivers_into_a_separate_output(ipc):
    """push_pull(ptr, ..., out=...): the staged input window is only read, the averaged result lands in `out`
    partition by partition - copied straight out of the colocated server's shared-memory store when IPC is on
    (pull by reference), out of the response window otherwise."""
    c = _core()
    nw = 2
    cl = Cluster(nw, 1, extra={"enable_ipc": ipc}).start()
    n = 400_000
    cut = 1_000_000
    parts = [(c.make_key(8, 0), 0, cut), (c.make_key(8, 1), cut, n * 4 - cut)]
    before = c.ipc_stats()["ref_responses"]
    results = {}

    def work(rank, w, po):
        name = "BytePS_ShM_%d_outtest%d" % (os.getpid(), rank)
        ptr = c.shm_create(name, n * 4)           # registered window, like comm/ps.py::_Staging
        try:
            win = np.frombuffer((ctypes.c_uint8 * (n * 4)).from_address(ptr), dtype=np.float32)
            for key, off, ln in parts:
                z = np.zeros(ln // 4, dtype=np.float32)
                w.init_key(key, z.ctypes.data, ln, c.F32)
            for it in range(3):
                win[:] = (np.arange(n, dtype=np.float32) % 31) * (rank + 1) + it
                out = np.full(n, -1.0, dtype=np.float32)
                h = w.push_pull("g", ptr, c.F32, parts, 0, 0, 1.0 / nw, 0, out.ctypes.data)
                assert w.wait(h)
                results[(rank, it)] = out
        finally:
            c.shm_release(name)
    cl.run_workers(work)
    for it in range(3):
        expect = sum((np.arange(n, dtype=np.float32) % 31) * (r + 1) + it for r in range(nw)) / nw
        for r in range(nw):
            np.testing.assert_allclose(results[(r, it)], expect, rtol=1e-6)
    refs = c.ipc_stats()["ref_responses"] - before
    assert refs == (nw * len(parts) * 3 if ipc else 0), refs
    cl.stop()
