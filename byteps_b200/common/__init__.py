"""Process-wide lifecycle: init / shutdown / suspend / resume / ranks.

Parity: ``BytePSBasics`` (/root/reference/byteps/common/__init__.py:52-139) and
the C API behind it (/root/reference/byteps/common/operations.cc:41-119).  The
reference keeps its state in C++ statics reached through ctypes; here the state
is one explicit python object owning a native engine, so suspend/resume is
"drop the engine, keep the registry's declaration order, build a new one".
"""
from __future__ import annotations

import atexit
import os
import threading
from typing import Optional

from ..config import Config

_state_lock = threading.RLock()


class _Global:
    def __init__(self):
        self.cfg: Optional[Config] = None
        self.engine = None
        self.group = None
        self.pg = None
        self.initialized = False
        self.owns_pg = False
        self.declared_order = []   # survives suspend/resume
        self.ps = None


_G = _Global()
_pg_generation = [0]     # how many times this process has created the default process group


def _setup_process_group(cfg: Config):
    """torch.distributed is only the bootstrap/control plane (and the gloo/NCCL
    fallback transports); DMLC_* variables are mapped onto its rendezvous."""
    import torch
    import torch.distributed as dist

    from ..comm.group import SoloGroup, TorchGroup

    if cfg.size == 1:
        return SoloGroup(), None, False
    owns = False
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", cfg.root_uri)
        os.environ.setdefault("MASTER_PORT", str(cfg.root_port + 1))
        backend = "cpu:gloo,cuda:nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if torch.cuda.is_available():
            # respect a device the user already selected (e.g. one-GPU "boxes" sharing a host);
            # otherwise pin to the local rank like the reference's examples do
            if torch.cuda.current_device() == 0:
                torch.cuda.set_device(cfg.local_rank % max(1, torch.cuda.device_count()))
            kwargs["device_id"] = torch.device("cuda", torch.cuda.current_device())
        gen = _pg_generation[0]
        _pg_generation[0] += 1
        if gen > 0 and os.environ.get("TORCHELASTIC_USE_AGENT_STORE", "") == "True":
            # resume() under torchrun: the rendezvous store lives in the elastic agent and still holds the keys
            # of the previous default group, so a second env:// init would wait forever.  Re-initialise over the
            # same store under a fresh prefix instead.
            base = dist.TCPStore(os.environ["MASTER_ADDR"], int(os.environ["MASTER_PORT"]), cfg.size, False)
            kwargs["store"] = dist.PrefixStore("byteps_b200_gen%d" % gen, base)
        try:
            dist.init_process_group(backend=backend, rank=cfg.rank, world_size=cfg.size, **kwargs)
        except TypeError:
            kwargs.pop("device_id", None)
            dist.init_process_group(backend=backend, rank=cfg.rank, world_size=cfg.size, **kwargs)
        owns = True
    return TorchGroup(None), None, owns


def get_ext_suffix() -> str:
    """Suffix of a native extension module for this interpreter (".cpython-312-x86_64-linux-gnu.so")."""
    import sysconfig

    return sysconfig.get_config_var("EXT_SUFFIX") or sysconfig.get_config_var("SO") or ".so"


def get_extension_full_path(pkg_path: str, *args) -> str:
    """Where the extension `args[-1]` of the package that contains `pkg_path` lives, e.g.
    ``get_extension_full_path(byteps_b200.__file__, "_core")`` (the reference's loader helper,
    /root/reference/byteps/common/__init__.py:39-43; its plugins call it with "c_lib")."""
    assert len(args) >= 1
    return os.path.join(os.path.dirname(pkg_path), *args[:-1], args[-1] + get_ext_suffix())


def check_extension(ext_name: str, ext_env_var: str, pkg_path: str, *args):
    """ImportError with the rebuild hint when a native extension has not been built."""
    full_path = get_extension_full_path(pkg_path, *args)
    if not os.path.exists(full_path):
        raise ImportError("Extension %s has not been built (%s is missing).  Run `python __graft_entry__.py` "
                          "(`python -m byteps_b200._build` prints the compiler output); %s is accepted for "
                          "compatibility and ignored." % (ext_name, full_path, ext_env_var))


class BytePSBasics:
    """Same surface as the reference's BytePSBasics."""

    def init(self, lazy: bool = True):
        with _state_lock:
            if _G.initialized:
                return
            from .. import _native
            from ..comm.engine import PushPullEngine

            cfg = Config.from_env()
            core = _native.core()
            lvl = {"TRACE": 0, "DEBUG": 1, "INFO": 2, "WARNING": 3, "ERROR": 4, "FATAL": 5}.get(
                cfg.log_level.upper(), 3)
            core.set_log_level(lvl)
            group, pg, owns = _setup_process_group(cfg)
            _G.cfg, _G.group, _G.pg, _G.owns_pg = cfg, group, pg, owns
            _G.engine = PushPullEngine(cfg, group, pg)
            # re-declare tensors in the original order so keys stay stable
            # (operations.cc:96-112, global.cc:431-436)
            for name in _G.declared_order:
                _G.engine.registry.declare(name)
            if cfg.is_distributed and cfg.role == "worker" and cfg.num_server > 0:
                from ..comm.ps import PSClient

                _G.ps = PSClient(cfg, _G.engine)
                _G.engine.attach_ps(_G.ps)
                if cfg.backend in ("auto", "ps"):
                    _G.engine.backend = "ps"
            _G.initialized = True

    def shutdown(self):
        with _state_lock:
            if not _G.initialized:
                return
            eng = _G.engine
            _G.declared_order = list(eng.registry.declared_names())
            eng.shutdown()
            if _G.ps is not None:
                _G.ps.close()
                _G.ps = None
            _G.engine = None
            _G.initialized = False
            if _G.owns_pg:
                import torch.distributed as dist

                if dist.is_initialized():
                    try:
                        dist.destroy_process_group()
                    except Exception:  # noqa: BLE001
                        pass
                _G.owns_pg = False

    def suspend(self):
        """Tear the engine down but remember declared tensors (operations.cc:114-119)."""
        self.shutdown()

    def resume(self, num_workers: int, num_servers: int, global_rank: int = -1):
        """Rebuild with a new topology (common/__init__.py:75-81 in the reference)."""
        os.environ["DMLC_NUM_WORKER"] = str(num_workers)
        os.environ["DMLC_NUM_SERVER"] = str(num_servers)
        if global_rank >= 0:
            os.environ["BYTEPS_GLOBAL_RANK"] = str(global_rank)
        self.init()

    def _cfg(self) -> Config:
        if not _G.initialized:
            raise ValueError("BytePS has not been initialized; use bps.init().")
        return _G.cfg

    def rank(self):
        return self._cfg().rank

    def size(self):
        return self._cfg().size

    def local_rank(self):
        return self._cfg().local_rank

    def local_size(self):
        return self._cfg().local_size

    def get_pushpull_speed(self):
        """(timestamp_ms, MB/s) of the oldest unread telemetry sample; (0, -5.0) if none."""
        if not _G.initialized:
            return (0, -5.0)
        return _G.engine.telemetry.get()


def engine():
    if not _G.initialized:
        raise ValueError("BytePS has not been initialized; use bps.init().")
    return _G.engine


def is_initialized() -> bool:
    return _G.initialized


def config() -> Config:
    return _G.cfg


_declared_set = set()


def remember_declared(name: str):
    # called on every push_pull: membership must not scan the list (it was 5 us per call with a few hundred names)
    if name not in _declared_set:
        if name not in _G.declared_order:
            _G.declared_order.append(name)
        _declared_set.add(name)


@atexit.register
def _cleanup():
    try:
        if _G.initialized:
            BytePSBasics().shutdown()
    except Exception:  # noqa: BLE001
        pass
