"""In-process PS cluster for tests: scheduler + servers + workers, each with its
own Postoffice, all in threads of one process over 127.0.0.1 (the reference
needs separate bpslaunch processes for this, tests/meta_test.py:26-85)."""
import threading

from _mp import free_port


class Cluster:
    def __init__(self, num_workers=2, num_servers=1, extra=None, server_kwargs=None, worker_kwargs=None,
                 server_host="127.0.0.1"):
        from byteps_b200 import _native

        self.core = _native.core()
        self.port = free_port()
        self.nw, self.ns = num_workers, num_servers
        self.extra = extra or {}
        self.server_host = server_host       # what the servers advertise (a non-loopback address = "another host")
        self.server_kwargs = server_kwargs or {}
        self.worker_kwargs = worker_kwargs or {}
        self.sched = None
        self.servers, self.server_pos = [], []
        self.workers, self.worker_pos = [None] * num_workers, [None] * num_workers

    def _po(self, role, rank=-1):
        host = self.server_host if role == "server" else "127.0.0.1"
        return self.core.Postoffice(role, self.nw, self.ns, "127.0.0.1", self.port, host, rank, self.extra)

    def start(self, make_worker=True):
        errs = []

        def run_sched():
            try:
                self.sched = self._po("scheduler")
                self.sched.start(0, True)
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        def run_server(i):
            try:
                po = self._po("server", i)
                srv = self.core.SumServer(po, **self.server_kwargs)
                self.server_pos.append(po)
                self.servers.append(srv)
                po.start(0, True)
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        def run_worker(i):
            try:
                po = self._po("worker", i)
                self.worker_pos[i] = po
                if make_worker:
                    self.workers[i] = self.core.PSWorker(po, **self.worker_kwargs)
                po.start(0, True)
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        ts = [threading.Thread(target=run_sched)]
        ts += [threading.Thread(target=run_server, args=(i,)) for i in range(self.ns)]
        ts += [threading.Thread(target=run_worker, args=(i,)) for i in range(self.nw)]
        for t in ts:
            t.start()
        for t in ts:
            t.join(60)
        assert not errs, errs
        assert all(not t.is_alive() for t in ts), "cluster start timed out"
        return self

    def stop(self):
        def fin(po):
            po.finalize(0, True)

        for w in self.workers:
            if w is not None:
                w.stop()
        ts = [threading.Thread(target=fin, args=(po,)) for po in [self.sched] + self.server_pos + self.worker_pos]
        for t in ts:
            t.start()
        for t in ts:
            t.join(60)
        for s in self.servers:
            s.stop()

    def run_workers(self, fn):
        """fn(rank, psworker, postoffice) in one thread per worker."""
        errs = []

        def wrap(i):
            try:
                fn(i, self.workers[i], self.worker_pos[i])
            except Exception as e:  # noqa: BLE001
                import traceback

                errs.append(traceback.format_exc())
                del e

        ts = [threading.Thread(target=wrap, args=(i,)) for i in range(self.nw)]
        for t in ts:
            t.start()
        for t in ts:
            t.join(120)
        assert not errs, "\n".join(errs)
        assert all(not t.is_alive() for t in ts), "workers timed out"
