"""User-facing compression parameters -> per-tensor compressor kwargs.

The reference accepts a ``compression_params`` dict only in its MXNet trainer and
turns it into ``byteps_*`` attributes on every parameter
(/root/reference/byteps/mxnet/__init__.py:236-317, docs/gradient-compression.md:39-60);
this is the same mapping as one function every front end shares.

    {"compressor": "topk", "k": 0.01, "ef": "vanilla", "momentum": "nesterov"}
      -> {"compressor_type": "topk", "compressor_k": "0.01", "ef_type": "vanilla",
          "momentum_type": "nesterov", "momentum_mu": "<optimizer momentum>"}
"""
from __future__ import annotations

from typing import Dict, Optional

_NEEDS_K = ("topk", "randomk", "dithering")


def translate(compression_params: Optional[dict], optimizer_params: Optional[dict] = None) -> Dict[str, str]:
    cp = dict(compression_params or {})
    out: Dict[str, str] = {}
    if "compressor" not in cp:
        return out
    for item in ("compressor", "ef", "momentum"):
        v = cp.get(item)
        if v:
            if not isinstance(v, str):
                raise TypeError("%s should be str" % item)
            out["%s_type" % item] = v
    comp = cp["compressor"]
    if comp == "onebit":
        out["compressor_onebit_scaling"] = str(bool(cp.get("scaling", False)))
    elif comp in _NEEDS_K:
        out["compressor_k"] = str(cp["k"])          # KeyError if missing, like the reference
    else:
        raise ValueError("unknown compressor %r" % comp)
    if cp.get("momentum"):
        mu = (optimizer_params or {}).get("momentum")
        if mu is None:
            raise KeyError("compression momentum needs the optimizer's 'momentum'")
        out["momentum_mu"] = str(mu)
    if cp.get("seed") is not None:
        out["seed"] = str(cp["seed"])
    part = cp.get("partition")
    if part:
        if part not in ("linear", "natural"):
            raise ValueError("Unsupported partition")
        out["dithering_partition"] = "0" if part == "linear" else "1"
    norm = cp.get("normalize")
    if norm:
        if norm not in ("max", "l2"):
            raise ValueError("Unsupported normalization")
        out["dithering_normalize"] = "0" if norm == "max" else "1"
    return out
