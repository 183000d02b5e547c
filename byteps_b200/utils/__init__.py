"""Small host-side helpers shared by the python layers (timing stamps, env parsing)."""
from .timing import stamp  # noqa: F401


def env_int(name: str, default: int) -> int:
    import os

    v = os.environ.get(name)
    return int(v) if v not in (None, "") else default


def env_flag(name: str, default: bool = False) -> bool:
    import os

    v = os.environ.get(name)
    if v in (None, ""):
        return default
    return v.lower() not in ("0", "false", "no", "off")
