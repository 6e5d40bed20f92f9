"""The CPU oracle instantiated at IntervalSet<i32> domains -- libpcp's VStoreSet / FDSpace
(variable/mod.rs:38, search/mod.rs:41-43), the instantiation example/src/nqueens.rs and every
search test of the reference run on -- against the known answers those tests hold, plus the set
semantics the restatement assumes for the un-vendored `intervallum` crate."""
import json
import os

import numpy as np
import pytest

from oracle.oracle_api import FAITHFUL, FLAT, SET, TUNED, OracleEngine
from pcp_b200 import models

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as f:
    GOLDEN = json.load(f)

ASSIGNMENT, BOUND, INNER = 0, 1, 2
SET_VARIANTS = [SET + FAITHFUL, SET + TUNED, SET + FLAT]


@pytest.mark.parametrize("variant", SET_VARIANTS)
def test_nqueens_all_solutions_on_interval_set(variant):
    """search/engine/all_solution.rs:67-98 as the reference runs it: FDSpace = IntervalSet."""
    for n, count in enumerate(GOLDEN["search"]["nqueens_all_solutions"]["counts"], start=1):
        if variant == SET + FAITHFUL and n > 8:
            continue
        for flavour in ("example", "distinct"):
            e = OracleEngine(variant)
            models.nqueens(n, flavour).load_into(e)
            res, _ = e.search(all_solutions=True)
            assert res.status == 2 and res.num_solution == count, (n, flavour)


def test_search_goldens_on_interval_set():
    """one_solution.rs:120-140, stop_node.rs:82-104, branch_and_bound.rs:112-138 on FDSpace."""
    for n, status in GOLDEN["search"]["nqueens_one_solution"]["status"].items():
        e = OracleEngine(SET + TUNED)
        models.nqueens(int(n), "distinct").load_into(e)
        res, _ = e.search()
        assert res.status == status, n
        if status == 1:
            lo, hi = e.domains()
            assert (lo == hi).all() and (e.domain_sizes() == 1).all()
    g = GOLDEN["search"]["stop_node"]
    e = OracleEngine(SET + TUNED)
    models.nqueens(g["n"], "distinct").load_into(e)
    res, _ = e.search(node_limit=g["limit"], all_solutions=True)
    assert res.status == g["status"] and res.num_nodes == g["num_nodes"]
    for mode, key in ((2, "maximize"), (1, "minimize")):
        e = OracleEngine(SET + TUNED)
        e.vars_alloc([0, 0], [10, 10])
        e.prop_alloc(models.X_LESS_Y, [[0, 0], [1, 0]])
        res, _ = e.search(all_solutions=True, bb_mode=mode, bb_var=0)
        assert res.status == 2 and res.has_bb_value and res.bb_value == GOLDEN["search"]["branch_and_bound"][key]


def test_distributors_and_selector_on_interval_set():
    """binary_split.rs:75-134 (`type Domain = IntervalSet<i32>`), enumerate.rs:71-77,
    first_smallest_var.rs:52-78."""
    g = GOLDEN["search"]["binary_split"]
    root = np.array(g["root"], np.int32)
    for var, children in g["children"].items():
        var = int(var)
        mid = int((root[var, 0] + root[var, 1]) / 2)
        for child, ops in zip(children, ([[var, 0], [-1, mid + 1]], [[-1, mid], [var, 0]])):
            e = OracleEngine(SET + FAITHFUL)
            e.vars_alloc(root[:, 0], root[:, 1])
            e.prop_alloc(models.X_LESS_Y, ops)
            assert e.consistency()[0] == 1
            lo, hi = e.domains()
            assert [int(lo[var]), int(hi[var])] == child
            assert int(e.domain_sizes()[var]) == child[1] - child[0] + 1
    # Enumerate + MinVal (enumerate.rs:71-77): x == min | x != min
    for var, children in {0: [[1, 1], [2, 10]], 1: [[2, 2], [3, 4]], 2: [[1, 1], [2, 2]]}.items():
        mn = int(root[var, 0])
        for child, kind in zip(children, (models.X_EQ_Y, models.X_NEQ_Y)):
            e = OracleEngine(SET + FAITHFUL)
            e.vars_alloc(root[:, 0], root[:, 1])
            e.prop_alloc(kind, [[var, 0], [-1, mn]])
            assert e.consistency()[0] == 1
            lo, hi = e.domains()
            assert [int(lo[var]), int(hi[var])] == child
    for case in GOLDEN["search"]["first_smallest_var"]["cases"]:
        d = np.array(case["vars"], np.int32)
        e = OracleEngine(SET + TUNED)
        e.vars_alloc(d[:, 0], d[:, 1])
        size = e.domain_sizes().astype(np.int64)
        cand = np.where(size > 1, size, np.iinfo(np.int64).max)
        assert int(np.argmin(cand)) == case["expect"]


def test_x_neq_y_removes_interior_values_and_raises_inner():
    """The difference between the two instantiations: on IntervalSet,
    `difference(&v)` removes an interior value (x_neq_y.rs:82-93) and the store classifies the
    change as Inner (events/mod.rs:62-63); on Interval the same vector is a no-op
    (x_neq_y.rs:128, test case 5)."""
    for variant, expect_delta, expect_size in ((SET + FAITHFUL, [(1, INNER)], 9), (SET + FLAT, [(1, INNER)], 9),
                                               (FAITHFUL, [], 10)):
        e = OracleEngine(variant)
        e.vars_alloc([5, 0], [5, 9])
        p = e.prop_alloc(models.X_NEQ_Y, [[0, 0], [1, 0]])
        before, ok, delta, after = e.test_propagation(p)
        assert (before, ok, delta) == (0, True, expect_delta)
        assert int(e.domain_sizes()[1]) == expect_size
        lo, hi = e.domains()
        assert (int(lo[1]), int(hi[1])) == (0, 9)
        bits = e.domain_bits(0, 1)
        assert int(bits[1, 0]) == (0x3FF & ~(1 << 5) if variant & SET else 0x3FF)
        # entailment: x = {5} and y without 5 are disjoint *as sets* -> XNeqY is entailed on
        # IntervalSet; on Interval the hulls still overlap
        assert after == (1 if variant & SET else 0)
    # bound values: Bound events, and an assignment when one value is left
    e = OracleEngine(SET + FAITHFUL)
    e.vars_alloc([3, 3, 4], [3, 5, 4])
    p0 = e.prop_alloc(models.X_NEQ_Y, [[0, 0], [1, 0]])
    p1 = e.prop_alloc(models.X_NEQ_Y, [[2, 0], [1, 0]])
    assert e.test_propagation(p0)[2] == [(1, BOUND)]
    assert e.test_propagation(p1)[2] == [(1, ASSIGNMENT)]
    lo, hi = e.domains()
    assert (int(lo[1]), int(hi[1])) == (5, 5)


def test_bounds_propagators_skip_holes():
    """shrink_left / shrink_right on a set land on the next value that is still there."""
    e = OracleEngine(SET + FAITHFUL)
    e.vars_alloc([0, 0], [9, 9])
    for v in (4, 5, 6):                                  # x loses 4, 5, 6
        e.prop_alloc(models.X_NEQ_Y, [[0, 0], [-1, v]])
    assert e.consistency()[0] == 1                       # x and {v} are disjoint sets now: all entailed
    assert int(e.domain_sizes()[0]) == 7
    e.prop_alloc(models.X_LESS_Y, [[-1, 3], [0, 0]])     # 3 < x  ->  x in {7, 8, 9}
    assert e.consistency()[0] == 1
    lo, hi = e.domains()
    assert (int(lo[0]), int(hi[0])) == (7, 9)
    e.prop_alloc(models.X_EQ_Y, [[0, 0], [1, 2]])        # x == y + 2  ->  y in {5, 6, 7}
    assert e.consistency()[0] == 0
    lo, hi = e.domains()
    assert (int(lo[1]), int(hi[1])) == (5, 7)
    assert int(e.domain_bits(0, 1)[1, 0]) == 0b11100000


def test_set_variants_agree_on_search_trace():
    """Faithful (reactor rebuilt per node), tuned and flat give the same nodes, statuses and
    domain contents (the trace hash covers every maximal run of every domain)."""
    rng = np.random.default_rng(11)
    mixed = models.Model("mixed-set", rng.integers(-5, 5, 12).astype(np.int32), rng.integers(6, 15, 12).astype(np.int32))
    for kind, n_ops in ((0, 2), (1, 2), (1, 2), (2, 2), (3, 3), (4, 3), (5, 3)):
        for _ in range(3):
            vs = rng.choice(12, n_ops, replace=False)
            mixed.add(kind, np.stack([vs, rng.integers(-3, 4, n_ops)], axis=1).astype(np.int32))
    mixed.add(models.ALL_EQUAL, [[0, 0], [5, 1], [9, -1]])
    mixed.add(models.DISTINCT, [[1, 0], [2, 0], [3, 0], [4, 2]])
    for model in (models.nqueens(11, "example"), models.nqueens(10, "distinct"), models.all_interval(7), mixed):
        traces = []
        for variant in SET_VARIANTS:
            e = OracleEngine(variant)
            model.load_into(e)
            res, tr = e.search(node_limit=400, all_solutions=True, trace=400, trace_domains=True)
            traces.append((res.num_nodes, res.num_solution, tr))
        for other in traces[1:]:
            assert traces[0][0] == other[0] and traces[0][1] == other[1]
            for k in ("status", "hash", "lo", "hi"):
                assert (traces[0][2][k] == other[2][k]).all(), (model.name, k)


def test_interval_set_prunes_more_than_interval():
    """Same model, same search: the IntervalSet instantiation visits fewer nodes (interior
    values go as soon as a queen is placed), the solution count is the same."""
    nodes = {}
    for variant in (FLAT, SET + FLAT):
        e = OracleEngine(variant)
        models.nqueens(10).load_into(e)
        res, _ = e.search(all_solutions=True)
        assert res.num_solution == 724
        nodes[variant] = res.num_nodes
    assert nodes[SET + FLAT] < nodes[FLAT]
