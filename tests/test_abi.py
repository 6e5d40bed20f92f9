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


def test_rust_shim_binds_every_symbol_and_constant():
    """rust_shim/src/ffi.rs (source only: no Rust toolchain here) declares every entry point of the
    header with the same arity, and the same constants."""
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "pcp_b200.h")).read(), flags=re.S)
    ffi = open(os.path.join(ROOT, "rust_shim", "src", "ffi.rs")).read()
    c_fns = {m.group(1): m.group(2) for m in re.finditer(r"\b(pcp_[a-z_]+)\s*\(([^;{]*?)\)\s*;", hdr)}
    r_fns = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (pcp_[a-z_]+)\(([^;]*?)\)(?:\s*->\s*[^;]+)?;", ffi, flags=re.S)}
    assert sorted(c_fns) == sorted(_declared_symbols())
    assert sorted(r_fns) == sorted(c_fns), sorted(set(c_fns) ^ set(r_fns))
    for name, args in c_fns.items():
        n_c = 0 if args.strip() in ("", "void") else args.count(",") + 1
        n_r = r_fns[name].count(":")
        assert n_c == n_r, name
    for m in re.finditer(r"#define\s+(PCP_[A-Z_0-9]+)\s+\(?(-?\d+)u?\)?\s", hdr):
        name, val = m.group(1), int(m.group(2))
        if name in ("PCP_NUM_KINDS", "PCP_F_MAX_NODES", "PCP_F_MAX_VARS"):
            continue
        r = re.search(r"pub const %s: \w+ = (-?\d+);" % name, ffi)
        assert r and int(r.group(1)) == val, name
    for m in re.finditer(r"\b(PCP_[A-Z_]+)\s*=\s*(\d+)", hdr):   # enum pcp_prop_kind
        r = re.search(r"pub const %s: i32 = (\d+);" % m.group(1), ffi)
        assert r and int(r.group(1)) == int(m.group(2)), m.group(1)


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
