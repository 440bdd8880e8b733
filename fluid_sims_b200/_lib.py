"""ctypes binding of libtau_b200.so (the C-ABI declared in include/tau_b200.h).

The library is the product: there is no Python/CPU fallback.  Importing this module on a box where
the shared library has not been built raises immediately; creating a solver handle on a box without
a CUDA device fails with -ENODEV from the library itself.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# TAU_B200_LIB: load another build of the same library (kernel experiments); never a fallback
LIB_PATH = os.environ.get("TAU_B200_LIB") or os.path.join(_HERE, "libtau_b200.so")


class TauError(RuntimeError):
    """Non-zero return code from the C-ABI; carries the library's message."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"[tau_b200 rc={code}] {msg}")
        self.code = code


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or `make -C fluid_sims_b200/csrc`).  There is no CPU fallback."
        )
    return C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)


lib = _load()
lib.tau_last_error.restype = C.c_char_p
lib.tau_abi_version.restype = C.c_int
lib.tau_device_count.restype = C.c_int


def check(rc: int) -> None:
    if rc != 0:
        raise TauError(rc, lib.tau_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    return int(lib.tau_device_count())


def declare(name: str, argtypes, restype=C.c_int):
    fn = getattr(lib, name)
    fn.argtypes = argtypes
    fn.restype = restype
    return fn
