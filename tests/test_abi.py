"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol that
include/pcp_b200.h declares; without a GPU the engine refuses to start (no CPU fallback)."""
import os
import re

import pytest

import pcp_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "pcp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pcp_[a-z_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = pcp_b200.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 17
    assert sorted(pcp_b200.ABI_SYMBOLS) == declared
    for s in declared:
        assert hasattr(lib, s), s


def test_struct_layouts_match_header():
    import ctypes as C
    from pcp_b200 import _capi
    assert C.sizeof(_capi.Operand) == 8
    assert C.sizeof(_capi.Config) == 16
    assert C.sizeof(_capi.Stats) == 24
    assert C.sizeof(_capi.SearchConfig) == 40
    assert C.sizeof(_capi.SearchResult) == 80


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pcp_b200.PcpError):
        pcp_b200.Engine()


def test_product_never_imports_oracle():
    """Only tests/, smoke() and bench.py may touch oracle/ (it is the checker, not the product)."""
    pkg = os.path.join(ROOT, "pcp_b200")
    bad = re.compile(r"^\s*(from\s+oracle|import\s+oracle)|#include\s+\"[^\"]*oracle|libpcp_oracle|pcpo_[a-z]+\(", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not bad.search(text), f


@pytest.mark.gpu
def test_plain_c_client_solves_nqueens_8():
    """tests/abi_client.c (built by __graft_entry__.build()): a C program that only includes
    include/pcp_b200.h drives example/src/nqueens.rs at N = 8 to its 92 solutions -- host-driven
    node loop on Interval and IntervalSet domains, pcp_search_run, and two engines through
    pcp_consistency_batch -- and sees contract violations as PCP_ERR_INVALID."""
    import subprocess
    exe = os.path.join(ROOT, "tests", "abi_client")
    if not os.path.exists(exe):
        import __graft_entry__ as g
        g.build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "92 solutions" in r.stdout and "pcp_search_run 92" in r.stdout and "two engines 92" in r.stdout
