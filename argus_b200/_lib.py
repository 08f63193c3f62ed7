"""ctypes binding of libargus_b200.so (the C ABI declared in include/argus_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
import re
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_void_p
from pathlib import Path

_PKG_DIR = Path(__file__).resolve().parent
# ARGUS_B200_LIB selects another build of the same library (A/B experiments); the default is the in-tree build
LIB_PATH = Path(os.environ["ARGUS_B200_LIB"]).resolve() if os.environ.get("ARGUS_B200_LIB") else _PKG_DIR / "libargus_b200.so"
HEADER_PATH = _PKG_DIR.parent / "include" / "argus_b200.h"


class ArgusError(RuntimeError):
    """Raised when a C-ABI call returns a non-zero status."""


_lib = None


def declared_symbols() -> list[str]:
    """All function names declared in include/argus_b200.h (used by the CPU symbol-export test)."""
    text = HEADER_PATH.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(argus_[a-z0-9_]+)\s*\(", text)))


def load() -> ctypes.CDLL:
    """Load the shared library (built in-tree by `make` / `__graft_entry__.build()`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ArgusError(
            f"{LIB_PATH} not found: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "argus_b200 has no CPU or PyTorch fallback."
        )
    lib = ctypes.CDLL(str(LIB_PATH), mode=os.RTLD_GLOBAL if hasattr(os, "RTLD_GLOBAL") else ctypes.DEFAULT_MODE)
    lib.argus_last_error_string.restype = c_char_p
    lib.argus_last_error_string.argtypes = []
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().argus_last_error_string()
        raise ArgusError(msg.decode("utf-8", "replace") if msg else f"argus call failed with status {status}")


def ptr(t) -> c_void_p:
    """Device (or host) pointer of a torch tensor, or NULL for None."""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def stream_ptr(stream=None) -> c_void_p:
    import torch

    s = stream if stream is not None else torch.cuda.current_stream()
    return c_void_p(s.cuda_stream)


def call(name: str, *args) -> None:
    """Call `int argus_xxx(...)` and raise on failure. Arguments are converted by type:
    tensors/None -> void*, int -> int, float -> float (use c_double(...) explicitly for doubles)."""
    import torch

    fn = getattr(load(), name)
    conv = []
    for a in args:
        if a is None or isinstance(a, torch.Tensor):
            conv.append(ptr(a))
        elif isinstance(a, bool):
            conv.append(c_int(int(a)))
        elif isinstance(a, int):
            conv.append(c_int(a))
        elif isinstance(a, float):
            conv.append(c_float(a))
        else:
            conv.append(a)
    fn.restype = c_int
    check(fn(*conv))


__all__ = ["ArgusError", "load", "check", "call", "ptr", "stream_ptr", "declared_symbols", "c_int64", "c_double",
           "c_float", "c_int", "c_void_p"]
