"""Formula trees (logic/: Conjunction, Disjunction, Boolean, BooleanNeg, not(), implication,
equivalence) and Cumulative by decomposition (propagators/cumulative.rs:59-114): the oracle
against the reference's Cumulative fixtures -- the only live tests of Store::consistency in
libpcp (cumulative.rs:254-319) -- and the device against the oracle (GPU tests)."""
import json
import os

import numpy as np
import pytest

from pcp_b200 import models
from pcp_b200.models import F_BOOLEAN, F_BOOLEAN_NEG, f_and, f_equivalence, f_implication, f_leaf, f_not, f_or

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "cumulative_vectors.json")) as f:
    CUMULATIVE = json.load(f)["vectors"]


def _oracle(variant=0):
    from oracle.oracle_api import OracleEngine
    return OracleEngine(variant)


def _engine(**kw):
    from pcp_b200 import Engine
    return Engine(**kw)


def _model(v):
    return models.cumulative(v["starts"], v["durations"], v["resources"], v["capacity"], v["constant"])


@pytest.mark.parametrize("v", CUMULATIVE, ids=lambda v: v["name"])
@pytest.mark.parametrize("variant", [0, 1], ids=["faithful", "tuned"])
def test_cumulative_fixture_on_oracle(v, variant):
    """cumulative.rs:236-252: Store::is_subsumed before, consistency(), Store::is_subsumed after."""
    e = _oracle(variant)
    _model(v).load_into(e)
    assert e.store_is_subsumed() == v["before"], v["ref"]
    assert e.consistency()[0] == v["after"], v["ref"]
    assert e.store_is_subsumed() == v["after"], v["ref"]


def test_connectives_on_oracle():
    """Disjunction::propagate only acts when all children but one are disentailed
    (disjunction.rs:96-116); Boolean reads its 0/1 variable (boolean.rs:108-122); not() pushes to
    the leaves (De Morgan, cmp/mod.rs complements)."""
    e = _oracle()
    e.vars_alloc([0, 0, 0], [9, 9, 1])
    x, y, b = (0, 0), (1, 0), 2
    # b <-> x < y
    e.formula_alloc(f_equivalence([F_BOOLEAN, b, 0], f_leaf(models.X_LESS_Y, x, y)))
    assert e.consistency()[0] == 0
    e.prop_alloc(models.X_LESS_Y, [[-1, 6], [0, 0]])   # x > 6
    e.prop_alloc(models.X_LESS_Y, [[1, 0], [-1, 4]])   # y < 4   => x < y is disentailed => b = 0
    assert e.consistency()[0] == 1
    lo, hi = e.domains()
    assert (int(lo[2]), int(hi[2])) == (0, 0)
    e2 = _oracle()
    e2.vars_alloc([0, 0, 1], [9, 9, 1])                # b = 1 => x < y propagates
    e2.formula_alloc(f_equivalence([F_BOOLEAN, b, 0], f_leaf(models.X_LESS_Y, x, y)))
    assert e2.consistency()[0] == 0
    lo, hi = e2.domains()
    assert (int(hi[0]), int(lo[1])) == (8, 1)
    e3 = _oracle()
    e3.vars_alloc([0, 0, 0], [9, 9, 0])                # b = 0 => not (x < y) = x >= y propagates
    e3.formula_alloc(f_equivalence([F_BOOLEAN, b, 0], f_leaf(models.X_LESS_Y, x, y)))
    e3.prop_alloc(models.X_LESS_Y, [[0, 0], [-1, 3]])  # x < 3
    assert e3.consistency()[0] == 0
    lo, hi = e3.domains()
    assert int(hi[1]) == 2


def _random_tree(rng, V, bools, depth=0, inner=None):
    # one event class per tree: leaves that subscribe at Inner (XNeqY, XEqY) and leaves that subscribe
    # at Bound must not share a variable inside one propagator (indexed_deps.rs:69-77); now and then a
    # tree mixes them anyway (inner=None below) and must be rejected
    if depth == 0 and inner is None and rng.random() < 0.85:
        inner = bool(rng.random() < 0.35)
    r = rng.random()
    if depth >= 2 or r < 0.35:
        k = int(rng.integers(0, 7))
        if inner is True:
            k = int(rng.integers(1, 3))
        elif inner is False:
            k = int(rng.choice([0, 3, 4, 5, 6]))
        if k == 6:
            return [F_BOOLEAN if rng.random() < 0.5 else F_BOOLEAN_NEG, int(rng.choice(bools)), 0]
        kind = [models.X_LESS_Y, models.X_NEQ_Y, models.X_EQ_Y, models.X_GREATER_Y_PLUS_Z, models.X_LESS_Y_PLUS_Z,
                models.X_EQ_Y_PLUS_Z][k]
        n = 2 if k < 3 else 3
        vs = rng.choice(V, n, replace=False)
        ops = [(int(v), int(rng.integers(-3, 4))) for v in vs]
        if rng.random() < 0.2:
            ops[int(rng.integers(0, n))] = (-1, int(rng.integers(-2, 12)))
        leaf = f_leaf(kind, *ops)
        return f_not(leaf) if (rng.random() < 0.25 and kind != models.X_EQ_Y_PLUS_Z) else leaf
    kids = [_random_tree(rng, V, bools, depth + 1, inner) for _ in range(int(rng.integers(2, 4)))]
    if any(models.F_LEAF + models.X_EQ_Y_PLUS_Z in k for k in kids):
        return f_and(*kids) if r < 0.7 else f_or(*kids)
    t = f_and(*kids) if r < 0.7 else f_or(*kids)
    return f_not(t) if rng.random() < 0.2 else t


def random_tree_store(seed):
    rng = np.random.default_rng(seed)
    V = int(rng.integers(5, 9))
    nb = 3
    lo = np.concatenate([rng.integers(-3, 4, V), np.zeros(nb, np.int64)]).astype(np.int32)
    hi = np.concatenate([lo[:V] + rng.integers(2, 9, V), np.ones(nb, np.int64)]).astype(np.int32)
    m = models.Model(f"trees-{seed}", lo, hi)
    bools = list(range(V, V + nb))
    for _ in range(int(rng.integers(2, 7))):
        m.add_formula(_random_tree(rng, V, bools))
    for _ in range(int(rng.integers(1, 5))):
        vs = rng.choice(V, 2, replace=False)
        m.add(int(rng.integers(0, 3)), [[int(vs[0]), int(rng.integers(-2, 3))], [int(vs[1]), 0]])
    return m


def _subscribes_twice(m):
    """A tree whose leaves read one variable at two different events (Bound and Inner) panics in
    the reference when the store subscribes it (indexed_deps.rs:69-77): the faithful oracle
    raises at its first consistency()."""
    from pcp_b200 import ContractViolation
    probe = _oracle(0)
    m.load_into(probe)
    try:
        probe.search(node_limit=120, all_solutions=True)
    except ContractViolation as ex:
        # "subscribed": rejected at allocation by the device too.  Anything else is the other panic
        # formula trees can run into -- Conjunction::propagate reaching a Boolean whose variable is
        # already 0 (boolean.rs:126-135 updates with {1}, not a subset: variable/store.rs:153-156) --
        # where the reference has no defined result: such stores are left out.
        return "subscribed" in str(ex) or None
    return False


def test_random_trees_faithful_vs_tuned_oracle():
    checked = 0
    for seed in range(60):
        m = random_tree_store(seed)
        if _subscribes_twice(m) is not False:
            continue
        checked += 1
        a, b = _oracle(0), _oracle(1)
        m.load_into(a)
        m.load_into(b)
        ra, ta = a.search(node_limit=120, all_solutions=True, trace=120, trace_domains=True)
        rb, tb = b.search(node_limit=120, all_solutions=True, trace=120, trace_domains=True)
        assert ra.num_nodes == rb.num_nodes and (ta["status"] == tb["status"]).all() and (ta["hash"] == tb["hash"]).all()
    assert checked >= 15


# ---- device -------------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("v", CUMULATIVE, ids=lambda v: v["name"])
def test_cumulative_fixture_on_device(v):
    """The reference's Cumulative fixtures through the device fixpoint: same status as the
    reference expects, same domains and `active` set as the oracle."""
    dev, ora = _engine(), _oracle()
    m = _model(v)
    m.load_into(dev)
    m.load_into(ora)
    ds, _ = dev.consistency()
    assert ds == v["after"] == ora.consistency()[0], v["ref"]
    if ds != -1:
        dlo, dhi = dev.domains()
        olo, ohi = ora.domains()
        assert (dlo == olo).all() and (dhi == ohi).all()
        assert (dev.active() == ora.active()).all()


@pytest.mark.gpu
def test_random_trees_on_device():
    """Seeded stores of random formula trees (depth <= 3, negations, Boolean leaves, constants)
    next to plain binary propagators: fixpoint, then 80 search nodes, per-node bit-exact."""
    from pcp_b200 import ContractViolation
    checked = 0
    for seed in range(60):
        m = random_tree_store(seed)
        dev, ora = _engine(), _oracle(1)
        twice = _subscribes_twice(m)
        if twice is None:
            continue
        if twice:   # a panic in the reference: PCP_ERR_INVALID here, at allocation
            with pytest.raises(ContractViolation):
                m.load_into(dev)
            continue
        checked += 1
        m.load_into(dev)
        m.load_into(ora)
        rd, td = dev.search(node_limit=80, all_solutions=True, trace=80, trace_domains=True)
        ro, to = ora.search(node_limit=80, all_solutions=True, trace=80, trace_domains=True)
        assert rd.num_nodes == ro.num_nodes, seed
        assert (td["status"] == to["status"]).all(), seed
        assert (td["hash"] == to["hash"]).all(), seed
        ok = td["status"] != -1
        assert (td["lo"][ok] == to["lo"][ok]).all() and (td["hi"][ok] == to["hi"][ok]).all()
    assert checked >= 15


@pytest.mark.gpu
def test_cumulative_scheduling_search_on_device():
    """A small resource-constrained schedule searched to the end: 4 tasks with start windows, unit
    capacity steps; every node of the tree bit-exact against the oracle."""
    m = models.cumulative([(0, 6), (0, 6), (0, 6), (0, 6)], [(2, 2), (3, 3), (2, 2), (1, 1)],
                          [(2, 2), (1, 1), (2, 2), (1, 1)], (3, 3), constant=True)
    dev, ora = _engine(), _oracle(1)
    m.load_into(dev)
    m.load_into(ora)
    rd, td = dev.search(node_limit=400, all_solutions=True, trace=400, trace_domains=True)
    ro, to = ora.search(node_limit=400, all_solutions=True, trace=400, trace_domains=True)
    assert rd.num_nodes == ro.num_nodes and rd.num_solution == ro.num_solution
    assert (td["status"] == to["status"]).all() and (td["hash"] == to["hash"]).all()


@pytest.mark.gpu
def test_formula_limits_are_loud():
    from pcp_b200 import PcpError
    dev = _engine()
    dev.vars_alloc([0] * 20, [9] * 20)
    big = f_or(*[f_leaf(models.X_LESS_Y, (i, 0), (i + 1, 0)) for i in range(14)])
    with pytest.raises(PcpError):
        dev.formula_alloc(big)                       # more than PCP_F_MAX_VARS variables
    with pytest.raises(PcpError):
        dev.formula_alloc(f_not(f_leaf(models.X_EQ_Y_PLUS_Z, (0, 0), (1, 0), (2, 0))))   # unimplemented!() in the reference
    with pytest.raises(PcpError):
        dev.formula_alloc([F_BOOLEAN, 3, 0])         # not a 0/1 variable
    assert dev.num_props == 0
