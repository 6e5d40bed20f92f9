"""`Engine`: the device constraint store behind libpcp's `Space` surface.

Loads pcp_b200/libpcp_b200.so (built in-tree by `__graft_entry__.build()` /
`make -C pcp_b200/csrc`) and fails loudly when it is missing or when no B200 is
usable -- there is no CPU path in the product.
"""
from __future__ import annotations

import ctypes as C
import os

from ._capi import Config, EngineBase, PcpError, SearchConfig, SearchResult, Stats, bind

_HERE = os.path.dirname(os.path.abspath(__file__))
# PCP_B200_LIB: development override (kernel variants built side by side for A/B timing)
LIB_PATH = os.environ.get("PCP_B200_LIB") or os.path.join(_HERE, "libpcp_b200.so")

FLAG_INCREMENTAL = 1
FLAG_HOST_SEARCH = 2
FLAG_INTERVAL_SET = 4

# every symbol include/pcp_b200.h declares
ABI_SYMBOLS = [
    "pcp_engine_create", "pcp_engine_fork", "pcp_engine_destroy", "pcp_last_error", "pcp_set_timing", "pcp_set_grid_limit", "pcp_stream", "pcp_vars_alloc",
    "pcp_sum_alloc", "pcp_prop_alloc", "pcp_props_alloc", "pcp_formula_alloc", "pcp_consistency", "pcp_consistency_batch", "pcp_domains_read",
    "pcp_domains_size_read", "pcp_domains_read_bits",
    "pcp_var_update", "pcp_active_read", "pcp_label", "pcp_restore", "pcp_num_vars", "pcp_num_props",
    "pcp_search_run", "pcp_search_open", "pcp_search_step", "pcp_search_step_many", "pcp_search_set_incumbent", "pcp_search_close",
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
        lib.pcp_engine_fork.restype = C.c_int
        lib.pcp_engine_fork.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        lib.pcp_set_timing.restype = C.c_int
        lib.pcp_set_timing.argtypes = [C.c_void_p, C.c_int32]
        lib.pcp_set_grid_limit.restype = C.c_int
        lib.pcp_set_grid_limit.argtypes = [C.c_void_p, C.c_int32]
        lib.pcp_stream.restype = C.c_int
        lib.pcp_stream.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        i32p, u64p = C.POINTER(C.c_int32), C.POINTER(C.c_uint64)
        lib.pcp_search_open.restype = C.c_int
        lib.pcp_search_open.argtypes = [C.c_void_p, C.POINTER(SearchConfig), i32p, u64p, i32p, i32p, C.c_uint64,
                                        C.POINTER(C.c_void_p)]
        lib.pcp_search_step.restype = C.c_int
        lib.pcp_search_step.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(SearchResult)]
        lib.pcp_consistency_batch.restype = C.c_int
        lib.pcp_consistency_batch.argtypes = [C.POINTER(C.c_void_p), C.c_int32, i32p, C.POINTER(Stats)]
        lib.pcp_search_step_many.restype = C.c_int
        lib.pcp_search_step_many.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_uint64, C.POINTER(SearchResult)]
        lib.pcp_search_set_incumbent.restype = C.c_int
        lib.pcp_search_set_incumbent.argtypes = [C.c_void_p, C.c_int32]
        lib.pcp_search_close.restype = None
        lib.pcp_search_close.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


class Engine(EngineBase):
    """One engine = one (vstore, cstore) pair on one B200, driven from one host thread."""

    _prefix = "pcp_"

    def __init__(self, device: int = 0, incremental: bool = False, max_labels: int = 0, tail_limit: int = 0,
                 timing: bool = False, host_search: bool = False, interval_set: bool = False):
        self._lib = load_library()
        flags = ((FLAG_INCREMENTAL if incremental else 0) | (FLAG_HOST_SEARCH if host_search else 0) |
                 (FLAG_INTERVAL_SET if interval_set else 0))
        cfg = Config(device, flags, max_labels, tail_limit)
        h = C.c_void_p()
        rc = self._lib.pcp_engine_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise PcpError(rc, "pcp_engine_create failed (no usable sm_100 CUDA device?) -- see stderr")
        self._h = h
        if timing:
            self.set_timing(True)

    def fork(self) -> "Engine":
        """pcp_engine_fork: another (vstore, cstore) pair in this engine's current state that shares
        the static model (descriptors, reactor) on the device."""
        child = Engine.__new__(Engine)
        child._lib = self._lib
        h = C.c_void_p()
        self._check(self._lib.pcp_engine_fork(self._h, C.byref(h)))
        child._h = h
        return child

    def set_timing(self, enabled: bool) -> None:
        self._check(self._lib.pcp_set_timing(self._h, int(enabled)))

    def set_grid_limit(self, max_ctas: int) -> None:
        """Cap the CTAs of this engine's launches (engines running side by side on one GPU)."""
        self._check(self._lib.pcp_set_grid_limit(self._h, int(max_ctas)))

    def cuda_stream(self) -> int:
        """The engine's cudaStream_t as an integer (torch.cuda.ExternalStream(ptr) wraps it)."""
        p = C.c_void_p()
        self._check(self._lib.pcp_stream(self._h, C.byref(p)))
        return int(p.value or 0)

    def search_open(self, node_limit: int = 0, all_solutions: bool = False, var_sel: int = 0, val_sel: int = 0,
                    distributor: int = 0, bb_mode: int = 0, bb_var: int = 0, warmup_nodes: int = 0) -> "SearchHandle":
        """Resumable search (pcp_search_open/step/close): the host can exchange a stop /
        incumbent word with other ranks between budget slices."""
        cfg = SearchConfig(node_limit, int(all_solutions), var_sel, val_sel, distributor, bb_mode, bb_var, 0,
                           warmup_nodes)
        return SearchHandle(self, cfg)


def consistency_batch(engines):
    """pcp_consistency_batch: the fixpoints of several engines launched together (they run side by
    side on the GPU).  Returns (statuses, stats) with one entry per engine."""
    n = len(engines)
    hs = (C.c_void_p * n)(*[e._h for e in engines])
    status = (C.c_int32 * n)()
    stats = (Stats * n)()
    rc = engines[0]._lib.pcp_consistency_batch(hs, n, status, stats)
    if rc != 0:
        _raise_from(engines, rc)
    return [int(x) for x in status], list(stats)


def _raise_from(engines, rc):
    """the error text lives in the engine that failed: report that one"""
    for e in engines:
        msg = e._fn("last_error")(e._h)
        if msg:
            e._check(rc)
    engines[0]._check(rc)


def search_step_many(handles, max_nodes: int = 0):
    """pcp_search_step_many: advance several open searches (one engine each) together."""
    n = len(handles)
    hs = (C.c_void_p * n)(*[h._h for h in handles])
    res = (SearchResult * n)()
    rc = handles[0]._lib.pcp_search_step_many(hs, n, max_nodes, res)
    if rc != 0:
        _raise_from([h._engine for h in handles], rc)
    return list(res)


class SearchHandle:
    def __init__(self, engine: Engine, cfg: SearchConfig):
        self._engine = engine
        self._lib = engine._lib
        self._cfg = cfg
        self._h = C.c_void_p()
        rc = self._lib.pcp_search_open(engine._h, C.byref(cfg), None, None, None, None, 0, C.byref(self._h))
        engine._check(rc)

    def step(self, max_nodes: int = 0) -> SearchResult:
        """Run at most `max_nodes` more nodes; `.status` 0 = still open (see pcp_b200.h)."""
        res = SearchResult()
        self._engine._check(self._lib.pcp_search_step(self._h, max_nodes, C.byref(res)))
        return res

    def set_incumbent(self, value: int) -> None:
        """Adopt another rank's incumbent (pcp_search_set_incumbent)."""
        self._engine._check(self._lib.pcp_search_set_incumbent(self._h, int(value)))

    def close(self) -> None:
        if self._h:
            self._lib.pcp_search_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
