"""`Engine`: the device constraint store behind libpcp's `Space` surface.

Loads pcp_b200/libpcp_b200.so (built in-tree by `__graft_entry__.build()` /
`make -C pcp_b200/csrc`) and fails loudly when it is missing or when no B200 is
usable -- there is no CPU path in the product.
"""
from __future__ import annotations

import ctypes as C
import os

from ._capi import Config, EngineBase, PcpError, bind

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpcp_b200.so")

FLAG_INCREMENTAL = 1

# every symbol include/pcp_b200.h declares
ABI_SYMBOLS = [
    "pcp_engine_create", "pcp_engine_destroy", "pcp_last_error", "pcp_set_timing", "pcp_vars_alloc",
    "pcp_sum_alloc", "pcp_prop_alloc", "pcp_props_alloc", "pcp_consistency", "pcp_domains_read",
    "pcp_var_update", "pcp_active_read", "pcp_label", "pcp_restore", "pcp_num_vars", "pcp_num_props",
    "pcp_search_run",
]

_lib = None


def load_library() -> C.CDLL:
    """dlopen the C-ABI library; raises if the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the CUDA extension first "
                "(python -c 'import __graft_entry__ as g; g.build()' or make -C pcp_b200/csrc). "
                "pcp_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        bind(lib, "pcp_")
        lib.pcp_engine_create.restype = C.c_int
        lib.pcp_engine_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
        lib.pcp_set_timing.restype = C.c_int
        lib.pcp_set_timing.argtypes = [C.c_void_p, C.c_int32]
        _lib = lib
    return _lib


class Engine(EngineBase):
    """One engine = one (vstore, cstore) pair on one B200, driven from one host thread."""

    _prefix = "pcp_"

    def __init__(self, device: int = 0, incremental: bool = False, max_labels: int = 0, tail_limit: int = 0,
                 timing: bool = False):
        self._lib = load_library()
        cfg = Config(device, FLAG_INCREMENTAL if incremental else 0, max_labels, tail_limit)
        h = C.c_void_p()
        rc = self._lib.pcp_engine_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise PcpError(rc, "pcp_engine_create failed (no usable sm_100 CUDA device?) -- see stderr")
        self._h = h
        if timing:
            self.set_timing(True)

    def set_timing(self, enabled: bool) -> None:
        self._check(self._lib.pcp_set_timing(self._h, int(enabled)))
