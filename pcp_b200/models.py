"""Workload generators (model definitions as propagator descriptors).

Each generator returns a `Model`: initial domains plus batches of `(kind, operands)`
in the exact allocation order of the reference model it restates, so propagator
indices match libpcp's.  Operands are `(var, off)` pairs as in `pcp_operand`
(include/pcp_b200.h).

Reference models:
  * n-queens, example flavour: example/src/nqueens.rs:27-50 (2*C(n,2) diagonal XNeqY
    interleaved per pair, then join_distinct -> C(n,2) XNeqY, propagators/distinct.rs:26-45)
  * n-queens, test flavour: src/libpcp/search/mod.rs:64-91 (same diagonals, one n-ary Distinct)
  * chained_lt: src/libpcp/propagation/store.rs:362-384 (dead test)
  * all-interval and the random arithmetic CSP are constructed from reference
    primitives; they do not exist in the reference.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np

# enum pcp_prop_kind
X_LESS_Y = 0
X_NEQ_Y = 1
X_EQ_Y = 2
X_GREATER_Y_PLUS_Z = 3
X_LESS_Y_PLUS_Z = 4
X_EQ_Y_PLUS_Z = 5
DISTINCT = 6
DISJ2_X_EQ_Y_PLUS_Z = 7
X_EQ_Y_MUL_Z = 8
ALL_EQUAL = 9

KIND_NAMES = {
    X_LESS_Y: "XLessY", X_NEQ_Y: "XNeqY", X_EQ_Y: "XEqY",
    X_GREATER_Y_PLUS_Z: "XGreaterYPlusZ", X_LESS_Y_PLUS_Z: "XLessYPlusZ",
    X_EQ_Y_PLUS_Z: "XEqYPlusZ", DISTINCT: "Distinct", DISJ2_X_EQ_Y_PLUS_Z: "Disj2XEqYPlusZ",
    X_EQ_Y_MUL_Z: "XEqYMulZ", ALL_EQUAL: "AllEqual",
}
KIND_BY_NAME = {v: k for k, v in KIND_NAMES.items()}

VAR_CONSTANT = -1


def var_sum(sum_id: int) -> int:
    return -2 - sum_id


@dataclass
class Model:
    name: str
    lo: np.ndarray  # int32 [V]
    hi: np.ndarray  # int32 [V]
    # batches in allocation order: (kind, ops int32 [n_props, n_ops, 2])
    batches: List[Tuple[int, np.ndarray]] = field(default_factory=list)
    sums: List[np.ndarray] = field(default_factory=list)  # each int32 [n_terms, 2]

    @property
    def num_vars(self) -> int:
        return int(self.lo.shape[0])

    @property
    def num_props(self) -> int:
        return int(sum(b[1].shape[0] for b in self.batches))

    def add(self, kind: int, ops) -> None:
        a = np.ascontiguousarray(np.asarray(ops, dtype=np.int32))
        if a.ndim == 2:
            a = a[None, :, :]
        assert a.ndim == 3 and a.shape[2] == 2
        self.batches.append((kind, a))

    def load_into(self, engine) -> None:
        """Allocate variables, sums and propagators into an engine (device or oracle)."""
        engine.vars_alloc(self.lo, self.hi)
        for s in self.sums:
            engine.sum_alloc(s)
        for kind, ops in self.batches:
            engine.props_alloc(kind, ops)


def _queens_diagonals(n: int) -> np.ndarray:
    i, j = np.triu_indices(n, 1)  # row-major: i ascending, j ascending within i
    i = i.astype(np.int32)
    j = j.astype(np.int32)
    d = j - i
    ops = np.zeros((i.shape[0], 2, 2, 2), dtype=np.int32)  # [pair, which(+/-), operand, (var,off)]
    ops[:, 0, 0, 0] = i
    ops[:, 0, 1, 0] = j
    ops[:, 0, 1, 1] = d       # Xi != Xj + (j - i)   (nqueens.rs:41-43)
    ops[:, 1, 0, 0] = i
    ops[:, 1, 1, 0] = j
    ops[:, 1, 1, 1] = -d      # Xi != Xj - (j - i)   (nqueens.rs:45-47)
    return ops.reshape(-1, 2, 2)


def nqueens(n: int, flavour: str = "example") -> Model:
    """n-queens on Interval domains [1, n].

    flavour="example": example/src/nqueens.rs:27-50 -- P = 3*C(n,2) binary XNeqY.
    flavour="distinct": src/libpcp/search/mod.rs:64-91 -- 2*C(n,2) XNeqY + Distinct(queens).
    """
    m = Model(f"nqueens-{n}-{flavour}", np.full(n, 1, np.int32), np.full(n, n, np.int32))
    if n >= 2:
        m.add(X_NEQ_Y, _queens_diagonals(n))
    if flavour == "example":
        if n >= 2:
            i, j = np.triu_indices(n, 1)
            ops = np.zeros((i.shape[0], 2, 2), dtype=np.int32)
            ops[:, 0, 0] = i
            ops[:, 1, 0] = j
            m.add(X_NEQ_Y, ops)  # join_distinct (distinct.rs:40-44)
    elif flavour == "distinct":
        ops = np.zeros((1, n, 2), dtype=np.int32)
        ops[0, :, 0] = np.arange(n)
        m.add(DISTINCT, ops)
    else:
        raise ValueError(flavour)
    return m


def chained_lt(n: int, lo: int = 1, hi: int = 10) -> Model:
    """X1 < X2 < ... < Xn on [lo, hi] (propagation/store.rs:362-384)."""
    m = Model(f"chained-lt-{n}", np.full(n, lo, np.int32), np.full(n, hi, np.int32))
    if n >= 2:
        ops = np.zeros((n - 1, 2, 2), dtype=np.int32)
        ops[:, 0, 0] = np.arange(n - 1)
        ops[:, 1, 0] = np.arange(1, n)
        m.add(X_LESS_Y, ops)
    return m


def all_interval(n: int, symmetry_breaking: bool = False, decompose_distinct: bool = False) -> Model:
    """All-interval series of size n: s_i in [0,n-1], d_i in [1,n-1],
    Distinct(s), Distinct(d), d_i = |s_{i+1} - s_i| encoded with reference primitives as
    Disjunction[XEqYPlusZ(s_{i+1}, s_i, d_i), XEqYPlusZ(s_i, s_{i+1}, d_i)]
    (logic/disjunction.rs:29-33)."""
    V = 2 * n - 1
    lo = np.zeros(V, np.int32)
    hi = np.full(V, n - 1, np.int32)
    lo[n:] = 1
    m = Model(f"all-interval-{n}", lo, hi)
    s = np.arange(n, dtype=np.int32)
    d = np.arange(n, 2 * n - 1, dtype=np.int32)
    for group in (s, d):
        if decompose_distinct:
            i, j = np.triu_indices(group.shape[0], 1)
            ops = np.zeros((i.shape[0], 2, 2), dtype=np.int32)
            ops[:, 0, 0] = group[i]
            ops[:, 1, 0] = group[j]
            m.add(X_NEQ_Y, ops)
        else:
            ops = np.zeros((1, group.shape[0], 2), dtype=np.int32)
            ops[0, :, 0] = group
            m.add(DISTINCT, ops)
    ops = np.zeros((n - 1, 6, 2), dtype=np.int32)
    ops[:, 0, 0] = s[1:]
    ops[:, 1, 0] = s[:-1]
    ops[:, 2, 0] = d
    ops[:, 3, 0] = s[:-1]
    ops[:, 4, 0] = s[1:]
    ops[:, 5, 0] = d
    m.add(DISJ2_X_EQ_Y_PLUS_Z, ops)
    if symmetry_breaking and n >= 2:
        m.add(X_LESS_Y, np.array([[[0, 0], [1, 0]]], dtype=np.int32))
    return m


_MASK64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(seed: int, count: int) -> np.ndarray:
    """`count` outputs of splitmix64 started at `seed` (vectorised)."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, count + 1, dtype=np.uint64)
        z = (np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)) & _MASK64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _MASK64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _MASK64
        return z ^ (z >> np.uint64(31))


def random_arith_csp(num_vars: int = 100_000, num_props: int = 1_000_000,
                     seed: int = 0x5043503230323630, slack: int = 64, value_range: int = 1000) -> Model:
    """Planted-solution arithmetic CSP: every propagator is
    XEqYPlusZ((x, c), y, z) with c = s[y] + s[z] - s[x] over distinct x, y, z, so the hidden
    assignment s satisfies all of them and the fixpoint never fails.  All bounds stay
    below 2^12 in magnitude (no i32 overflow)."""
    assert num_vars >= 3
    r = _splitmix64(seed, 3 * num_vars + 3 * num_props)
    s = (r[:num_vars] % np.uint64(value_range)).astype(np.int64)
    a = (r[num_vars:2 * num_vars] % np.uint64(slack + 1)).astype(np.int64)
    b = (r[2 * num_vars:3 * num_vars] % np.uint64(slack + 1)).astype(np.int64)
    pr = r[3 * num_vars:].reshape(num_props, 3)
    x = (pr[:, 0] % np.uint64(num_vars)).astype(np.int64)
    y = (pr[:, 1] % np.uint64(num_vars - 1)).astype(np.int64)
    y = y + (y >= x)                               # distinct from x
    z = (pr[:, 2] % np.uint64(num_vars - 2)).astype(np.int64)
    lo2 = np.minimum(x, y)
    hi2 = np.maximum(x, y)
    z = z + (z >= lo2)
    z = z + (z >= hi2)                             # distinct from x and y
    c = s[y] + s[z] - s[x]
    m = Model(f"random-arith-{num_vars}-{num_props}", (s - a).astype(np.int32), (s + b).astype(np.int32))
    ops = np.zeros((num_props, 3, 2), dtype=np.int32)
    ops[:, 0, 0] = x
    ops[:, 0, 1] = c
    ops[:, 1, 0] = y
    ops[:, 2, 0] = z
    m.add(X_EQ_Y_PLUS_Z, ops)
    return m
