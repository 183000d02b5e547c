"""Start-up phase stamps on stderr (``BYTEPS_TIMING=1``): where the seconds between
``init()`` and the first step go (symmetric-memory mapping, NVLS binding, cuDNN
autotune in the warm-up steps, graph capture)."""
from __future__ import annotations

import os
import sys
import time

_ON = os.environ.get("BYTEPS_TIMING", "0") not in ("0", "")
_T0 = time.time()
_LAST = [_T0]


def stamp(label: str):
    if not _ON:
        return
    now = time.time()
    sys.stderr.write("[bps-timing r%s +%.2fs (%.2fs)] %s\n" % (
        os.environ.get("RANK", os.environ.get("BYTEPS_GLOBAL_RANK", "0")), now - _T0, now - _LAST[0], label))
    sys.stderr.flush()
    _LAST[0] = now
