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

# formula-tree node tags (pcp_formula_alloc, include/pcp_b200.h)
F_CONJUNCTION, F_DISJUNCTION, F_BOOLEAN, F_BOOLEAN_NEG, F_NOT, F_LEAF = 1, 2, 3, 4, 5, 16


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
        return int(sum(1 if b[0] < 0 else b[1].shape[0] for b in self.batches))

    def add_formula(self, words) -> None:
        """One propagator given as a formula tree (prefix words of pcp_formula_alloc)."""
        self.batches.append((-1, np.ascontiguousarray(np.asarray(words, dtype=np.int32).reshape(-1))))

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
            if kind < 0:
                engine.formula_alloc(ops)
            else:
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


def f_leaf(kind: int, *ops) -> list:
    """A leaf propagator inside a formula tree: tag, then (var, off) per operand."""
    w = [F_LEAF + kind]
    for var, off in ops:
        w += [int(var), int(off)]
    return w


def f_not(f: list) -> list:
    return [F_NOT] + list(f)


def f_and(*fs) -> list:
    return [F_CONJUNCTION, len(fs)] + [w for f in fs for w in f]


def f_or(*fs) -> list:
    return [F_DISJUNCTION, len(fs)] + [w for f in fs for w in f]


def f_implication(f: list, g: list) -> list:
    """logic/mod.rs:30-35 (as written there: Disjunction[f, g.not()])."""
    return f_or(f, f_not(g))


def f_equivalence(f: list, g: list) -> list:
    """logic/mod.rs:37-45."""
    return f_and(f_implication(f, g), f_implication(g, f))


def cumulative(starts, durations, resources, capacity, constant: bool = False) -> Model:
    r"""`Cumulative::join` (propagators/cumulative.rs:59-114: the decomposition of Schutt et al.)
    on the instance the reference's tests build (cumulative.rs:161-232): tasks given as
    (lo, hi) domains; with `constant`, assigned domains become Constant views instead of
    variables.  Allocation order of the propagators as in the reference: per task j and other
    task i: equivalence(b_ij, s_i <= s_j /\ s_j < s_i + d_i), r_ij = b_ij * r_i; then
    c >= r_j + sum_i r_ij."""
    lo, hi = [], []

    def alloc(d):
        lo.append(int(d[0]))
        hi.append(int(d[1]))
        return (len(lo) - 1, 0)

    def create(d):
        if d[0] == d[1] and constant:
            return (VAR_CONSTANT, int(d[0]))
        return alloc(d)

    s_ = [create(d) for d in starts]
    d_ = [create(d) for d in durations]
    r_ = [create(d) for d in resources]
    r_ub = [int(d[1]) for d in resources]
    cap = alloc(capacity)
    tasks = len(s_)
    m = Model(f"cumulative-{tasks}", np.zeros(0, np.int32), np.zeros(0, np.int32))

    def plus(view, c):
        return (view[0], view[1] + c)

    if tasks == 1:
        m.add(X_LESS_Y, [r_[0], plus(cap, 1)])           # c >= r[0]: x_geq_y (cmp/mod.rs:44-52)
    else:
        for j in range(tasks):
            terms = []
            for i in range(tasks):
                if i == j:
                    continue
                conj = f_and(f_leaf(X_LESS_Y, s_[i], plus(s_[j], 1)),             # s[i] <= s[j]
                             f_leaf(X_LESS_Y_PLUS_Z, s_[j], s_[i], d_[i]))        # s[j] < s[i] + d[i]
                b = alloc((0, 1))                                                 # Boolean::new
                m.add_formula(f_equivalence([F_BOOLEAN, b[0], 0], conj))
                r = alloc((0, r_ub[i]))
                m.add(X_EQ_Y_MUL_Z, [r, b, r_[i]])                                # r = b * r[i]
                terms.append(r)
            m.sums.append(np.array(terms, np.int32))
            sid = len(m.sums) - 1
            m.add(X_GREATER_Y_PLUS_Z, [plus(cap, 1), r_[j], (var_sum(sid), 0)])   # c >= r[j] + sum
    m.lo = np.array(lo, np.int32)
    m.hi = np.array(hi, np.int32)
    return m
