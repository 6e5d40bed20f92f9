"""The CPU oracle against every known-answer vector libpcp's own tests hold for the
propagation hot path.  This is what pins the oracle; the GPU is then checked
against the oracle (tests/test_gpu_*.py)."""
import json
import os

import numpy as np
import pytest

from oracle.oracle_api import FAITHFUL, FLAT, TUNED, Fifo, OracleEngine, Reactor, event_new
from pcp_b200 import models

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as f:
    GOLDEN = json.load(f)

ASSIGNMENT, BOUND, INNER = 0, 1, 2


def _engine_for(vec, variant=FAITHFUL):
    e = OracleEngine(variant)
    d = np.array(vec["domains"], np.int32)
    e.vars_alloc(d[:, 0], d[:, 1])
    p = e.prop_alloc(models.KIND_BY_NAME[vec["kind"]], vec["ops"])
    return e, p


@pytest.mark.parametrize("vec", GOLDEN["propagators"], ids=lambda v: v["name"])
@pytest.mark.parametrize("variant", [FAITHFUL, FLAT], ids=["faithful", "flat"])
def test_propagator_vector(vec, variant):
    """propagators/mod.rs:110-129: is_subsumed, propagate, exact delta, is_subsumed.  The FLAT
    variant (inline (var, off) descriptors -- the timed CPU baseline, the reference arm and the
    checker of the full-size GPU tests) replays every vector too."""
    e, p = _engine_for(vec, variant)
    before, ok, delta, after = e.test_propagation(p)
    assert before == vec["before"], vec["ref"]
    assert ok == vec["ok"], vec["ref"]
    if vec["ok"]:
        assert delta == [tuple(d) for d in vec["delta"]], vec["ref"]
    assert after == vec["after"], vec["ref"]
    if "domains_after" in vec:
        lo, hi = e.domains()
        assert [[int(a), int(b)] for a, b in zip(lo, hi)] == vec["domains_after"]


@pytest.mark.parametrize("vec", GOLDEN["propagators"], ids=lambda v: v["name"])
@pytest.mark.parametrize("variant", [FAITHFUL, TUNED, FLAT])
def test_propagator_vector_through_loop(vec, variant):
    """The same vectors through Store::consistency: status False iff propagate failed or the
    propagator is disentailed; otherwise the after-entailment decides True/Unknown once the
    propagator is at its own fixpoint (store.rs:166-189,250-256)."""
    e, p = _engine_for(vec, variant)
    status, stats = e.consistency()
    if not vec["ok"] or vec["after"] == -1:
        assert status == -1
        return
    # Re-running the propagator can only narrow further; a True after one pass stays True.
    if vec["after"] == 1:
        assert status == 1
        assert e.active()[p] == 0
    else:
        assert status in (0, 1)
    assert stats.propagations >= 1


@pytest.mark.parametrize("u", GOLDEN["store_updates"], ids=lambda u: f"{u['ref']}#{u['num']}")
def test_store_update(u):
    """variable/store.rs:151-166 + events/mod.rs:46-70 through MonotonicUpdate::update."""
    e = OracleEngine()
    e.vars_alloc([u["source"][0]], [u["source"][1]])
    ok = e.var_update(0, u["target"][0], u["target"][1])
    assert ok == u["ok"]
    lo, hi = e.domains()
    expect = u["target"] if u["ok"] else u["source"]
    assert [int(lo[0]), int(hi[0])] == expect
    if u["ok"]:
        ev = event_new(u["target"], u["source"])
        assert ([ev] if ev is not None else []) == u["events"]


@pytest.mark.parametrize("t", GOLDEN["store_intersections"], ids=lambda t: f"{t['x']}^{t['y']}")
def test_store_intersection(t):
    """variable/store.rs:509-526 == XEqY::propagate (x_eq_y.rs:102-107)."""
    e = OracleEngine()
    e.vars_alloc([t["x"][0], t["y"][0]], [t["x"][1], t["y"][1]])
    p = e.prop_alloc(models.X_EQ_Y, [[0, 0], [1, 0]])
    _, ok, delta, _ = e.test_propagation(p)
    assert ok == t["ok"]
    if ok:
        assert delta == [tuple(d) for d in t["delta"]]
        lo, hi = e.domains()
        assert [int(lo[0]), int(hi[0])] == t["target"] and [int(lo[1]), int(hi[1])] == t["target"]


def test_store_contract_violations():
    """variable/store.rs:330-365: empty alloc, non-monotone updates and bad indices panic."""
    from pcp_b200._capi import ContractViolation
    e = OracleEngine()
    with pytest.raises(ContractViolation):
        e.vars_alloc([1], [0])
    e.vars_alloc([0], [10])
    with pytest.raises(ContractViolation):
        e.var_update(0, 11, 11)
    with pytest.raises(ContractViolation):
        e.var_update(0, -5, 15)
    with pytest.raises(ContractViolation):
        e.var_update(3, 0, 0)
    assert e.var_update(0, 5, 5) is True
    assert e.var_update(0, 1, 0) is False  # empty_update (store.rs:330-336)
    with pytest.raises(ContractViolation):
        e.prop_alloc(models.X_LESS_Y, [[0, 0], [7, 0]])


def test_reactor_subscribe():
    """reactors/indexed_deps.rs:159-187."""
    r = Reactor(3, 3)
    assert r.is_empty()
    assert r.subscribe(0, ASSIGNMENT, 4) == 0
    assert not r.is_empty()
    assert r.react(0, ASSIGNMENT) == [4]
    assert r.react(0, INNER) == []
    r.subscribe(0, INNER, 5)
    assert r.react(0, INNER) == [5]
    assert r.react(0, ASSIGNMENT) == [4, 5]
    r.subscribe(1, BOUND, 6)
    r.subscribe(1, ASSIGNMENT, 7)
    assert r.react(1, INNER) == []
    assert r.react(1, BOUND) == [6]
    assert r.react(1, ASSIGNMENT) == [7, 6]
    assert r.react(2, ASSIGNMENT) == []
    r.subscribe(2, ASSIGNMENT, 8)
    assert r.react(1, INNER) == []


def test_reactor_unsubscribe_and_panics():
    """reactors/indexed_deps.rs:189-231."""
    r = Reactor(3, 3)
    r.subscribe(0, ASSIGNMENT, 4)
    r.subscribe(0, BOUND, 5)
    assert r.react(0, ASSIGNMENT) == [4, 5]
    assert r.unsubscribe(0, BOUND, 5) == 0
    assert r.react(0, ASSIGNMENT) == [4]
    r.unsubscribe(0, ASSIGNMENT, 4)
    assert r.react(0, ASSIGNMENT) == []
    r.subscribe(0, ASSIGNMENT, 4)
    assert r.react(0, ASSIGNMENT) == [4]
    r2 = Reactor(3, 3)
    assert r2.subscribe(0, ASSIGNMENT, 0) == 0
    assert r2.subscribe(0, ASSIGNMENT, 0) == -1   # subscribe_fail_test
    assert Reactor(3, 3).unsubscribe(0, ASSIGNMENT, 0) == -1  # unsubscribe_fail_test
    r3 = Reactor(3, 3)
    r3.subscribe(0, ASSIGNMENT, 0)
    assert r3.subscribe(0, BOUND, 0) == -1        # subscribe_two_events_fail_test


def test_relaxed_fifo():
    """schedulers/relaxed_fifo.rs:78-132."""
    def schedule_21(s):
        s.schedule(2)
        s.schedule(1)

    def pop_1(s):
        assert not s.is_empty()
        assert s.pop() == 1
        assert s.pop() is None
        assert s.is_empty()

    s = Fifo(3)
    schedule_21(s)
    assert s.pop() == 2
    pop_1(s)
    s.schedule(1)
    s.schedule(1)
    pop_1(s)

    s = Fifo(3)
    schedule_21(s)
    s.unschedule(1)
    assert s.pop() == 2
    assert s.pop() is None
    schedule_21(s)
    s.unschedule(2)
    pop_1(s)
    schedule_21(s)
    s.unschedule(2)
    s.unschedule(2)
    pop_1(s)
    assert Fifo(3).schedule(3) == -1
    assert Fifo(3).unschedule(3) == -1


@pytest.mark.parametrize("variant", [FAITHFUL, TUNED, FLAT])
def test_nqueens_all_solutions(variant):
    """search/engine/all_solution.rs:67-74 (counts are propagation-strength independent)."""
    g = GOLDEN["search"]["nqueens_all_solutions"]
    for n, count in enumerate(g["counts"], start=1):
        if variant == FAITHFUL and n > 8:
            continue
        for flavour in ("example", "distinct"):
            e = OracleEngine(variant)
            models.nqueens(n, flavour).load_into(e)
            res, _ = e.search(all_solutions=True)
            assert res.status == 2  # EndOfSearch
            assert res.num_solution == count, (n, flavour)


def test_nqueens_one_solution():
    """search/engine/one_solution.rs:120-128."""
    for n, status in GOLDEN["search"]["nqueens_one_solution"]["status"].items():
        e = OracleEngine(TUNED)
        models.nqueens(int(n), "distinct").load_into(e)
        res, _ = e.search()
        assert res.status == status, n
        if status == 1:
            lo, hi = e.domains()
            assert (lo == hi).all()
            q = lo.astype(int)
            assert len(set(q)) == len(q)
            assert len({q[i] + i for i in range(len(q))}) == len(q)
            assert len({q[i] - i for i in range(len(q))}) == len(q)


def test_stop_node():
    """search/stop_node.rs:82-104."""
    g = GOLDEN["search"]["stop_node"]
    e = OracleEngine(TUNED)
    models.nqueens(g["n"], "distinct").load_into(e)
    res, _ = e.search(node_limit=g["limit"], all_solutions=True)
    assert res.status == g["status"]
    assert res.num_nodes == g["num_nodes"]


@pytest.mark.parametrize("mode,key", [(2, "maximize"), (1, "minimize")])
def test_branch_and_bound(mode, key):
    """search/branch_and_bound.rs:112-138: x < y on [0,10]^2."""
    e = OracleEngine(TUNED)
    e.vars_alloc([0, 0], [10, 10])
    e.prop_alloc(models.X_LESS_Y, [[0, 0], [1, 0]])
    res, _ = e.search(all_solutions=True, bb_mode=mode, bb_var=0)
    assert res.status == 2
    assert res.has_bb_value and res.bb_value == GOLDEN["search"]["branch_and_bound"][key]


def test_binary_split_children():
    """search/branching/binary_split.rs:110-134 with MiddleVal: children domains, each True."""
    g = GOLDEN["search"]["binary_split"]
    root = np.array(g["root"], np.int32)
    for var, children in g["children"].items():
        var = int(var)
        mid = int((root[var, 0] + root[var, 1]) / 2)
        for child, ops in zip(children, ([[var, 0], [-1, mid + 1]], [[-1, mid], [var, 0]])):
            e = OracleEngine(FAITHFUL)
            e.vars_alloc(root[:, 0], root[:, 1])
            e.prop_alloc(models.X_LESS_Y, ops)  # x <= mid | x > mid (binary_split.rs:46-57)
            status, _ = e.consistency()
            assert status == 1
            lo, hi = e.domains()
            assert [int(lo[var]), int(hi[var])] == child


@pytest.mark.parametrize("variant", [FAITHFUL, TUNED, FLAT])
def test_chained_lt_and_root(variant):
    """propagation/store.rs:362-392 (dead tests): loop-level known answers on Interval."""
    for n, status in GOLDEN["search"]["chained_lt"]["status"].items():
        e = OracleEngine(variant)
        models.chained_lt(int(n)).load_into(e)
        st, _ = e.consistency()
        assert st == status, n
        lo, hi = e.domains()
        if int(n) == 9:
            assert list(lo) == list(range(1, 10)) and list(hi) == list(range(2, 11))
        if int(n) == 10:
            assert list(lo) == list(range(1, 11)) and list(hi) == list(range(1, 11))
    for n, status in GOLDEN["search"]["nqueens_root"]["status"].items():
        e = OracleEngine(variant)
        models.nqueens(int(n), "example").load_into(e)
        st, _ = e.consistency()
        assert st == status, n
        lo, hi = e.domains()
        assert (lo == 1).all() and (hi == int(n)).all()


def test_all_variants_agree_on_search_trace():
    """Same node sequence, statuses and domains from the three oracle variants (FLAT is what the
    full-size GPU tests and the CPU baseline use)."""
    rng = np.random.default_rng(7)
    mixed = models.Model("mixed", rng.integers(-5, 5, 14).astype(np.int32), rng.integers(6, 15, 14).astype(np.int32))
    for kind, n_ops in ((0, 2), (1, 2), (2, 2), (3, 3), (4, 3), (5, 3), (8, 3)):
        for _ in range(3):
            vs = rng.choice(14, n_ops, replace=False)
            mixed.add(kind, np.stack([vs, rng.integers(-3, 4, n_ops)], axis=1).astype(np.int32))
    mixed.add(models.ALL_EQUAL, [[0, 0], [5, 1], [9, -1]])
    mixed.add(models.DISTINCT, [[1, 0], [2, 0], [3, 0], [4, 2]])
    for model in (models.nqueens(12, "example"), models.nqueens(10, "distinct"), models.all_interval(7),
                  models.all_interval(6, decompose_distinct=True), mixed):
        traces = []
        for variant in (FAITHFUL, TUNED, FLAT):
            e = OracleEngine(variant)
            model.load_into(e)
            res, tr = e.search(node_limit=400, all_solutions=True, trace=400, trace_domains=True)
            traces.append((res.num_nodes, res.num_solution, tr))
        for other in traces[1:]:
            assert traces[0][0] == other[0] and traces[0][1] == other[1]
            for k in ("status", "hash", "lo", "hi"):
                assert (traces[0][2][k] == other[2][k]).all(), (model.name, k)


def test_label_restore():
    """Snapshot::label/restore of (vstore, cstore): kernel/restoration.rs:20-30,
    propagation/store.rs:312-323."""
    e = OracleEngine(TUNED)
    models.nqueens(6, "example").load_into(e)
    st, _ = e.consistency()
    assert st == 0
    l0 = e.label()
    lo0, hi0 = e.domains()
    n0 = e.num_props
    a0 = e.active().copy()
    e.prop_alloc(models.X_LESS_Y, [[0, 0], [-1, 2]])  # q0 <= 1
    st, _ = e.consistency()
    assert st in (0, -1)
    e.restore(l0)
    lo1, hi1 = e.domains()
    assert (lo0 == lo1).all() and (hi0 == hi1).all()
    assert e.num_props == n0 and (e.active() == a0).all()


def test_all_interval_small_solutions():
    """Constructed C3 model: every solution found is an all-interval series."""
    for n in (3, 4, 5, 6):
        e = OracleEngine(TUNED)
        models.all_interval(n).load_into(e)
        res, _ = e.search()
        assert res.status == 1
        lo, hi = e.domains()
        assert (lo == hi).all()
        s, d = lo[:n], lo[n:]
        assert sorted(s) == list(range(n))
        assert sorted(d) == list(range(1, n))
        assert all(abs(int(s[i + 1]) - int(s[i])) == int(d[i]) for i in range(n - 1))
    counts = {}
    for n in (3, 4, 5):
        e = OracleEngine(TUNED)
        models.all_interval(n).load_into(e)
        res, _ = e.search(all_solutions=True)
        counts[n] = res.num_solution
    # brute force count of all-interval series
    import itertools
    for n, c in counts.items():
        bf = sum(1 for p in itertools.permutations(range(n))
                 if len({abs(p[i + 1] - p[i]) for i in range(n - 1)}) == n - 1)
        assert c == bf, n
