"""ctypes wrapper of the CPU ORACLE (oracle/libpcp_oracle.so) -- test infrastructure.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
import this module.  The product package (pcp_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from pcp_b200._capi import EngineBase, bind, _i32p

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpcp_oracle.so")

FAITHFUL, TUNED, FLAT = 0, 1, 2
SET = 4  # + SET: IntervalSet<i32> domains (libpcp's VStoreSet / FDSpace) instead of Interval<i32>


def build(force: bool = False) -> str:
    """Compile the oracle with g++ (oracle/Makefile)."""
    if force or not os.path.exists(_SO) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO)
            for f in ("pcp_oracle.hpp", "pcp_oracle_common.hpp", "pcp_oracle_body.hpp", "pcp_oracle_capi.cpp",
                  "pcp_oracle_capi_body.hpp", "Makefile")):
        subprocess.run(["make", "-C", _HERE, "-B", "libpcp_oracle.so"], check=True, capture_output=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        bind(_lib, "pcpo_")
        _lib.pcpo_engine_create.restype = C.c_int
        _lib.pcpo_engine_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        _lib.pcpo_test_propagation.restype = C.c_int
        _lib.pcpo_test_propagation.argtypes = [C.c_void_p, C.c_int32, _i32p, _i32p, _i32p, _i32p, _i32p]
        _lib.pcpo_prop_dependencies.restype = C.c_int
        _lib.pcpo_prop_dependencies.argtypes = [C.c_void_p, C.c_int32, _i32p, _i32p]
        _lib.pcpo_store_is_subsumed.restype = C.c_int
        _lib.pcpo_store_is_subsumed.argtypes = [C.c_void_p, _i32p]
        _lib.pcpo_reactor_new.restype = C.c_void_p
        _lib.pcpo_reactor_new.argtypes = [C.c_int32, C.c_int32]
        _lib.pcpo_reactor_free.argtypes = [C.c_void_p]
        for n in ("subscribe", "unsubscribe"):
            f = getattr(_lib, "pcpo_reactor_" + n)
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
        _lib.pcpo_reactor_react.restype = C.c_int
        _lib.pcpo_reactor_react.argtypes = [C.c_void_p, C.c_int32, C.c_int32, _i32p, _i32p]
        _lib.pcpo_reactor_is_empty.argtypes = [C.c_void_p]
        _lib.pcpo_fifo_new.restype = C.c_void_p
        _lib.pcpo_fifo_new.argtypes = [C.c_int32]
        _lib.pcpo_fifo_free.argtypes = [C.c_void_p]
        for n in ("schedule", "unschedule"):
            f = getattr(_lib, "pcpo_fifo_" + n)
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_int32]
        _lib.pcpo_fifo_pop.argtypes = [C.c_void_p]
        _lib.pcpo_fifo_is_empty.argtypes = [C.c_void_p]
        _lib.pcpo_event_new.restype = C.c_int
        _lib.pcpo_event_new.argtypes = [C.c_int32] * 4 + [_i32p]
    return _lib


class OracleEngine(EngineBase):
    """libpcp's Space (vstore + cstore) restated on the CPU."""

    _prefix = "pcpo_"

    def __init__(self, variant: int = TUNED):
        self._lib = lib()
        h = C.c_void_p()
        rc = self._lib.pcpo_engine_create(variant, C.byref(h))
        assert rc == 0
        self._h = h

    def test_propagation(self, prop: int):
        """propagators/mod.rs:110-129: (before, propagate_ok, delta[(var, event)], after)."""
        before, ok, after = C.c_int32(), C.c_int32(), C.c_int32()
        delta = np.zeros(2 * 4096, np.int32)
        n = C.c_int32(4096)
        self._check(self._lib.pcpo_test_propagation(self._h, prop, C.byref(before), C.byref(ok), C.byref(after),
                                                    delta.ctypes.data_as(_i32p), C.byref(n)))
        d = [(int(delta[2 * i]), int(delta[2 * i + 1])) for i in range(n.value)]
        return before.value, bool(ok.value), d, after.value

    def store_is_subsumed(self) -> int:
        """Subsumption for Store (propagation/store.rs:231-237)."""
        k = C.c_int32()
        self._check(self._lib.pcpo_store_is_subsumed(self._h, C.byref(k)))
        return k.value

    def dependencies(self, prop: int):
        deps = np.zeros(2 * 65536, np.int32)
        n = C.c_int32(65536)
        self._check(self._lib.pcpo_prop_dependencies(self._h, prop, deps.ctypes.data_as(_i32p), C.byref(n)))
        return [(int(deps[2 * i]), int(deps[2 * i + 1])) for i in range(n.value)]


class Reactor:
    """reactors/indexed_deps.rs:23-113."""

    def __init__(self, num_vars, num_events):
        self._l = lib()
        self._h = C.c_void_p(self._l.pcpo_reactor_new(num_vars, num_events))

    def subscribe(self, var, ev, prop):
        return self._l.pcpo_reactor_subscribe(self._h, var, ev, prop)

    def unsubscribe(self, var, ev, prop):
        return self._l.pcpo_reactor_unsubscribe(self._h, var, ev, prop)

    def react(self, var, ev):
        out = np.zeros(1024, np.int32)
        n = C.c_int32(1024)
        rc = self._l.pcpo_reactor_react(self._h, var, ev, out.ctypes.data_as(_i32p), C.byref(n))
        assert rc == 0
        return [int(x) for x in out[:n.value]]

    def is_empty(self):
        return bool(self._l.pcpo_reactor_is_empty(self._h))

    def __del__(self):
        self._l.pcpo_reactor_free(self._h)


class Fifo:
    """schedulers/relaxed_fifo.rs:27-71."""

    def __init__(self, capacity):
        self._l = lib()
        self._h = C.c_void_p(self._l.pcpo_fifo_new(capacity))

    def schedule(self, i):
        return self._l.pcpo_fifo_schedule(self._h, i)

    def unschedule(self, i):
        return self._l.pcpo_fifo_unschedule(self._h, i)

    def pop(self):
        r = self._l.pcpo_fifo_pop(self._h)
        return None if r < 0 else r

    def is_empty(self):
        return bool(self._l.pcpo_fifo_is_empty(self._h))

    def __del__(self):
        self._l.pcpo_fifo_free(self._h)


def event_new(little, big):
    ev = C.c_int32()
    rc = lib().pcpo_event_new(little[0], little[1], big[0], big[1], C.byref(ev))
    if rc != 0:
        raise AssertionError("contract violation")
    return None if ev.value < 0 else ev.value
