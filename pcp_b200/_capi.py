"""ctypes binding of the C ABI declared in include/pcp_b200.h.

`bind(lib, prefix)` attaches argtypes/restypes; `EngineBase` is the thin Python mirror of
the reference's store surface (VStore::alloc, Store::alloc, Consistency::consistency,
Snapshot::label/restore) on top of those symbols.  The symbol prefix is a parameter only
so that the test-side oracle (oracle/oracle_api.py), which exports the same shapes under
`pcpo_`, can reuse the glue; the product always binds `pcp_` from libpcp_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

FALSE, UNKNOWN, TRUE = -1, 0, 1

ERRORS = {0: "OK", -1: "PCP_ERR_INVALID", -2: "PCP_ERR_CUDA", -3: "PCP_ERR_NOMEM", -4: "PCP_ERR_UNSUPPORTED"}


class PcpError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


class ContractViolation(PcpError):
    """A call that panics in the reference (assert!) -- PCP_ERR_INVALID."""


class Operand(C.Structure):
    _fields_ = [("var", C.c_int32), ("off", C.c_int32)]


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("flags", C.c_uint32), ("max_labels", C.c_uint32), ("tail_limit", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("propagations", C.c_uint64), ("iterations", C.c_uint32), ("active_props", C.c_uint32),
                ("kernel_ms", C.c_float), ("launches", C.c_uint32)]


class SearchConfig(C.Structure):
    _fields_ = [("node_limit", C.c_uint64), ("all_solutions", C.c_int32), ("var_sel", C.c_int32),
                ("val_sel", C.c_int32), ("distributor", C.c_int32), ("bb_mode", C.c_int32),
                ("bb_var", C.c_int32), ("trace_domains", C.c_int32), ("warmup_nodes", C.c_int32)]


class SearchResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("has_bb_value", C.c_int32), ("bb_value", C.c_int32),
                ("reserved", C.c_int32), ("num_nodes", C.c_uint64), ("num_solution", C.c_uint64),
                ("num_failed_node", C.c_uint64), ("num_prune", C.c_uint64), ("propagations", C.c_uint64),
                ("iterations", C.c_uint64), ("seconds", C.c_double), ("kernel_seconds", C.c_double)]


_i32p = C.POINTER(C.c_int32)
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)
_u64p = C.POINTER(C.c_uint64)
_opp = C.POINTER(Operand)

# name -> (restype, argtypes) without the engine-handle-specific create/destroy
_COMMON = {
    "engine_destroy": (None, [C.c_void_p]),
    "last_error": (C.c_char_p, [C.c_void_p]),
    "vars_alloc": (C.c_int, [C.c_void_p, _i32p, _i32p, C.c_int32, _i32p]),
    "sum_alloc": (C.c_int, [C.c_void_p, _opp, C.c_int32, _i32p]),
    "prop_alloc": (C.c_int, [C.c_void_p, C.c_int32, _opp, C.c_int32, _i32p]),
    "props_alloc": (C.c_int, [C.c_void_p, C.c_int32, _opp, C.c_int32, C.c_int64, _i32p]),
    "formula_alloc": (C.c_int, [C.c_void_p, _i32p, C.c_int32, _i32p]),
    "consistency": (C.c_int, [C.c_void_p, _i32p, C.POINTER(Stats)]),
    "domains_size_read": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _u32p]),
    "domains_read_bits": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _u32p]),
    "domains_read": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _i32p, _i32p]),
    "var_update": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, _i32p]),
    "active_read": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _u8p]),
    "label": (C.c_int, [C.c_void_p, _u64p]),
    "restore": (C.c_int, [C.c_void_p, C.c_uint64]),
    "num_vars": (C.c_int, [C.c_void_p, _i32p]),
    "num_props": (C.c_int, [C.c_void_p, _i32p]),
    "search_run": (C.c_int, [C.c_void_p, C.POINTER(SearchConfig), C.POINTER(SearchResult), _i32p, _u64p,
                             _i32p, _i32p, C.c_uint64]),
}


def bind(lib: C.CDLL, prefix: str) -> None:
    for name, (res, args) in _COMMON.items():
        fn = getattr(lib, prefix + name)  # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args


def _ops_array(ops) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(ops, dtype=np.int32))
    assert a.shape[-1] == 2, "operands are (var, off) pairs"
    return a


class EngineBase:
    """Python mirror of the store surface; subclasses provide `_lib`, `_prefix`, `_h`."""

    _lib: C.CDLL
    _prefix: str
    _h: Optional[C.c_void_p]

    def _fn(self, name):
        return getattr(self._lib, self._prefix + name)

    def _check(self, rc: int) -> None:
        if rc != 0:
            msg = self._fn("last_error")(self._h)
            msg = msg.decode() if msg else ""
            raise (ContractViolation if rc == -1 else PcpError)(rc, msg)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._fn("engine_destroy")(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # --- VStore::alloc (variable/store.rs:135-140)
    def vars_alloc(self, lo, hi) -> int:
        lo = np.ascontiguousarray(np.asarray(lo, dtype=np.int32).reshape(-1))
        hi = np.ascontiguousarray(np.asarray(hi, dtype=np.int32).reshape(-1))
        assert lo.shape == hi.shape
        first = C.c_int32(-1)
        self._check(self._fn("vars_alloc")(self._h, lo.ctypes.data_as(_i32p), hi.ctypes.data_as(_i32p),
                                           lo.shape[0], C.byref(first)))
        return first.value

    # --- Sum::new (term/sum.rs:28-32)
    def sum_alloc(self, terms) -> int:
        a = _ops_array(terms).reshape(-1, 2)
        sid = C.c_int32(-1)
        self._check(self._fn("sum_alloc")(self._h, a.ctypes.data_as(_opp), a.shape[0], C.byref(sid)))
        return sid.value

    # --- Store::alloc (propagation/store.rs:223-230)
    def prop_alloc(self, kind: int, ops) -> int:
        a = _ops_array(ops).reshape(-1, 2)
        idx = C.c_int32(-1)
        self._check(self._fn("prop_alloc")(self._h, kind, a.ctypes.data_as(_opp), a.shape[0], C.byref(idx)))
        return idx.value

    def props_alloc(self, kind: int, ops) -> int:
        a = _ops_array(ops)
        assert a.ndim == 3
        first = C.c_int32(-1)
        self._check(self._fn("props_alloc")(self._h, kind, a.ctypes.data_as(_opp), a.shape[1], a.shape[0],
                                            C.byref(first)))
        return first.value

    # --- a formula tree as one propagator (logic/*.rs; prefix words, see pcp_formula_alloc)
    def formula_alloc(self, words) -> int:
        a = np.ascontiguousarray(np.asarray(words, dtype=np.int32).reshape(-1))
        idx = C.c_int32(-1)
        self._check(self._fn("formula_alloc")(self._h, a.ctypes.data_as(_i32p), a.shape[0], C.byref(idx)))
        return idx.value

    # --- Consistency::consistency (propagation/store.rs:247-257)
    def consistency(self) -> Tuple[int, Stats]:
        st = C.c_int32(0)
        stats = Stats()
        self._check(self._fn("consistency")(self._h, C.byref(st), C.byref(stats)))
        return st.value, stats

    def domains(self, first: int = 0, n: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray]:
        if n is None:
            n = self.num_vars - first
        lo = np.empty(n, np.int32)
        hi = np.empty(n, np.int32)
        self._check(self._fn("domains_read")(self._h, first, n, lo.ctypes.data_as(_i32p), hi.ctypes.data_as(_i32p)))
        return lo, hi

    def domain_sizes(self, first: int = 0, n: Optional[int] = None) -> np.ndarray:
        """Cardinality::size() per variable (first_smallest_var.rs:30-39)."""
        if n is None:
            n = self.num_vars - first
        out = np.empty(n, np.uint32)
        self._check(self._fn("domains_size_read")(self._h, first, n, out.ctypes.data_as(_u32p)))
        return out

    def domain_bits(self, base: int, words: int, first: int = 0, n: Optional[int] = None) -> np.ndarray:
        """The values of each domain as a bit window [n, words] starting at value `base`."""
        if n is None:
            n = self.num_vars - first
        out = np.zeros((n, words), np.uint32)
        self._check(self._fn("domains_read_bits")(self._h, first, n, base, words, out.ctypes.data_as(_u32p)))
        return out

    # --- MonotonicUpdate::update (variable/store.rs:151-166)
    def var_update(self, idx: int, lo: int, hi: int) -> bool:
        ok = C.c_int32(0)
        self._check(self._fn("var_update")(self._h, idx, lo, hi, C.byref(ok)))
        return bool(ok.value)

    def active(self, first: int = 0, n: Optional[int] = None) -> np.ndarray:
        if n is None:
            n = self.num_props - first
        out = np.empty(n, np.uint8)
        self._check(self._fn("active_read")(self._h, first, n, out.ctypes.data_as(_u8p)))
        return out

    # --- Snapshot (kernel/restoration.rs:20-30)
    def label(self) -> int:
        l = C.c_uint64(0)
        self._check(self._fn("label")(self._h, C.byref(l)))
        return l.value

    def restore(self, label: int) -> None:
        self._check(self._fn("restore")(self._h, label))

    @property
    def num_vars(self) -> int:
        n = C.c_int32(0)
        self._check(self._fn("num_vars")(self._h, C.byref(n)))
        return n.value

    @property
    def num_props(self) -> int:
        n = C.c_int32(0)
        self._check(self._fn("num_props")(self._h, C.byref(n)))
        return n.value

    # --- search driver (search/mod.rs:45-52 and friends)
    def search(self, node_limit: int = 0, all_solutions: bool = False, var_sel: int = 0, val_sel: int = 0,
               distributor: int = 0, bb_mode: int = 0, bb_var: int = 0, trace: int = 0,
               trace_domains: bool = False, warmup_nodes: int = 0):
        cfg = SearchConfig(node_limit, int(all_solutions), var_sel, val_sel, distributor, bb_mode, bb_var,
                           int(trace_domains), warmup_nodes)
        res = SearchResult()
        V = self.num_vars
        t_status = np.zeros(trace, np.int32)
        t_hash = np.zeros(trace, np.uint64)
        t_lo = np.zeros((trace, V) if trace_domains else (0, V), np.int32)
        t_hi = np.zeros((trace, V) if trace_domains else (0, V), np.int32)
        self._check(self._fn("search_run")(
            self._h, C.byref(cfg), C.byref(res), t_status.ctypes.data_as(_i32p), t_hash.ctypes.data_as(_u64p),
            t_lo.ctypes.data_as(_i32p), t_hi.ctypes.data_as(_i32p), trace))
        n = int(min(trace, res.num_nodes))
        tr = {"status": t_status[:n], "hash": t_hash[:n]}
        if trace_domains:
            tr["lo"], tr["hi"] = t_lo[:n], t_hi[:n]
        return res, tr


class _Counters:
    """Cumulative search counters with the field names of `SearchResult`."""

    def __init__(self):
        self.status = 0
        self.has_bb_value = 0
        self.bb_value = 0
        self.num_nodes = self.num_solution = self.num_failed_node = self.num_prune = 0
        self.propagations = self.iterations = 0
        self.seconds = self.kernel_seconds = 0.0


class PySearchHandle:
    """The resumable search of `pcp_search_open/step/close`, node by node over the store
    surface (restore / prop_alloc / consistency / domains / label).  Any `EngineBase` can use
    it; the device engine overrides `search_open` with the native driver."""

    def __init__(self, engine: "EngineBase", all_solutions: bool = False, node_limit: int = 0, warmup_nodes: int = 0,
                 bb_mode: int = 0, bb_var: int = 0):
        self.e = engine
        self.bb_mode = bb_mode
        self.bb_var = bb_var
        self.all_solutions = all_solutions
        self.node_limit = node_limit
        self.warmup = warmup_nodes
        self.stack = []
        self.started = False
        self.stopped = False
        self.exhausted_reported = False
        self.res = _Counters()

    def _enter_child(self) -> int:
        import time
        t0 = time.perf_counter()
        r = self.res
        if self.bb_mode and r.has_bb_value:  # branch_and_bound.rs:76-87
            if self.bb_mode == 1:
                self.e.prop_alloc(0, [[self.bb_var, 0], [-1, r.bb_value]])   # var < incumbent
            else:
                self.e.prop_alloc(0, [[-1, r.bb_value], [self.bb_var, 0]])   # var > incumbent
        st, stats = self.e.consistency()
        if r.num_nodes >= self.warmup:
            r.propagations += int(stats.propagations)
            r.iterations += int(stats.iterations)
            r.kernel_seconds += float(stats.kernel_ms) * 1e-3
        status = st
        if st == UNKNOWN:
            lo, hi = self.e.domains()
            size = hi.astype(np.int64) - lo.astype(np.int64) + 1
            var = int(np.argmin(np.where(size > 1, size, np.iinfo(np.int64).max)))
            val = int((int(lo[var]) + int(hi[var])) / 2)
            label = self.e.label()
            self.stack.append((label, var, val, 1))
            self.stack.append((label, var, val, 0))
        if st == TRUE and self.bb_mode:  # branch_and_bound.rs:89-92
            lo, _ = self.e.domains(self.bb_var, 1)
            r.has_bb_value, r.bb_value = 1, int(lo[0])
        r.num_nodes += 1
        stop = self.node_limit and r.num_nodes >= self.node_limit
        if stop:
            status = 2
        elif st == TRUE:
            r.num_solution += 1
        elif st == FALSE:
            r.num_failed_node += 1
        if r.num_nodes > self.warmup:
            r.seconds += time.perf_counter() - t0
        return status

    def step(self, max_nodes: int = 0) -> _Counters:
        r = self.res
        start = r.num_nodes
        if self.stopped:
            r.status = 2
            return r
        while True:
            if self.started and not self.stack:
                if self.all_solutions or self.exhausted_reported:
                    r.status = 2
                else:
                    self.exhausted_reported = True
                    r.status = -1
                return r
            if max_nodes and r.num_nodes - start >= max_nodes:
                r.status = 0
                return r
            if not self.started:
                self.started = True
            else:
                label, var, val, alt = self.stack.pop()
                self.e.restore(label)
                if alt == 0:
                    self.e.prop_alloc(0, [[var, 0], [-1, val + 1]])  # x <= val (binary_split.rs:46-51)
                else:
                    self.e.prop_alloc(0, [[-1, val], [var, 0]])      # x > val  (binary_split.rs:52-57)
            child = self._enter_child()
            if child == 2:
                self.stopped = True
                r.status = 2
                return r
            if child == 1 and not self.all_solutions:
                r.status = 1
                return r

    def set_incumbent(self, value: int) -> None:
        """pcp_search_set_incumbent: adopt a better incumbent found elsewhere."""
        assert self.bb_mode, "search opened without bb_mode"
        r = self.res
        if not r.has_bb_value or (value < r.bb_value if self.bb_mode == 1 else value > r.bb_value):
            r.has_bb_value, r.bb_value = 1, int(value)

    def close(self) -> None:
        pass


def _search_open(self, node_limit: int = 0, all_solutions: bool = False, warmup_nodes: int = 0, bb_mode: int = 0,
                 bb_var: int = 0, **_kw):
    return PySearchHandle(self, all_solutions=all_solutions, node_limit=node_limit, warmup_nodes=warmup_nodes,
                          bb_mode=bb_mode, bb_var=bb_var)


EngineBase.search_open = _search_open
