"""Process-wide default Engine (one bp_context per device) for the drop-in classes."""
import os
import threading

from ._native import Engine

_engines = {}
_lock = threading.Lock()
_default_device = None


def parse_device(device):
    """'cuda', 'cuda:3', 3, None -> device index.  Anything else (e.g. 'cpu') is an error:
    the engine has no CPU path."""
    if device is None:
        return int(os.environ.get("VISAPI_DEVICE", "0"))
    if isinstance(device, int):
        return device
    s = str(device).strip().lower()
    if s == "cuda":
        return 0
    if s.startswith("cuda:"):
        return int(s.split(":", 1)[1])
    if s.isdigit():
        return int(s)
    raise ValueError("--device %r: this engine runs on CUDA devices only (cuda, cuda:N)" % (device,))


def set_default_device(device):
    global _default_device
    _default_device = parse_device(device)


def get_engine(device=None):
    idx = parse_device(device if device is not None else _default_device)
    with _lock:
        eng = _engines.get(idx)
        if eng is None:
            eng = Engine(idx)
            _engines[idx] = eng
    return eng


def close_all():
    with _lock:
        for eng in _engines.values():
            eng.close()
        _engines.clear()
