"""GPU parity tests on IntervalSet<i32> domains (PCP_FLAG_INTERVAL_SET: libpcp's VStoreSet /
FDSpace, variable/mod.rs:38, search/mod.rs:41-43 -- the instantiation example/src/nqueens.rs
runs on): the `set` kernel variants through the C ABI against the oracle instantiated at
IntervalSet (tests/test_oracle_set.py pins that one).  Bit-exact after every fixpoint: status,
bounds, cardinalities, the full value sets and the `active` set."""
import json
import os

import numpy as np
import pytest

from pcp_b200 import models

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as f:
    GOLDEN = json.load(f)

UNSUPPORTED_ON_SETS = {"XEqYMulZ", "AllEqual"}


def _oracle(variant=1):
    from oracle.oracle_api import SET, OracleEngine
    return OracleEngine(SET + variant)


def _engine(**kw):
    from pcp_b200 import Engine
    return Engine(interval_set=True, **kw)


def _assert_same_state(dev, ora, check_active=True):
    dlo, dhi = dev.domains()
    olo, ohi = ora.domains()
    assert (dlo == olo).all() and (dhi == ohi).all()
    assert (dev.domain_sizes() == ora.domain_sizes()).all()
    base = int(min(dlo.min(), olo.min()))
    words = (int(max(dhi.max(), ohi.max())) - base) // 32 + 1
    assert (dev.domain_bits(base, words) == ora.domain_bits(base, words)).all()
    if check_active and dev.num_props:
        assert (dev.active() == ora.active()).all()


def _compare_search(model, node_limit, dev_kw=None, all_solutions=True, oracle_variant=2, **skw):
    dev, ora = _engine(**(dev_kw or {})), _oracle(oracle_variant)
    model.load_into(dev)
    model.load_into(ora)
    rd, td = dev.search(node_limit=node_limit, all_solutions=all_solutions, trace=node_limit or 100000,
                        trace_domains=True, **skw)
    ro, to = ora.search(node_limit=node_limit, all_solutions=all_solutions, trace=node_limit or 100000,
                        trace_domains=True, **skw)
    assert rd.num_nodes == ro.num_nodes
    assert rd.status == ro.status
    assert rd.num_solution == ro.num_solution and rd.num_failed_node == ro.num_failed_node
    assert (td["status"] == to["status"]).all()
    ok = td["status"] != -1
    assert (td["hash"] == to["hash"]).all()          # covers every run of values of every domain
    assert (td["lo"][ok] == to["lo"][ok]).all() and (td["hi"][ok] == to["hi"][ok]).all()
    return rd, ro, dev, ora


@pytest.mark.parametrize("vec", [v for v in GOLDEN["propagators"] if v["kind"] not in UNSUPPORTED_ON_SETS],
                         ids=lambda v: v["name"])
def test_reference_vector_on_sets(vec):
    """The reference's propagator vectors with the domains allocated as IntervalSet: the device
    fixpoint against the set oracle (on hole-free inputs the bound propagators behave as on
    Interval; XNeqY / Distinct remove interior values)."""
    d = np.array(vec["domains"], np.int32)
    kind = models.KIND_BY_NAME[vec["kind"]]
    dev, ora = _engine(), _oracle(0)
    for e in (dev, ora):
        e.vars_alloc(d[:, 0], d[:, 1])
        e.prop_alloc(kind, vec["ops"])
    ds, dstats = dev.consistency()
    os_, _ = ora.consistency()
    assert ds == os_, vec["ref"]
    if ds != -1:
        _assert_same_state(dev, ora)
        assert dstats.propagations >= 1


def test_x_neq_y_interior_value_and_entailment():
    dev, ora = _engine(), _oracle(0)
    for e in (dev, ora):
        e.vars_alloc([5, 0, 0], [5, 9, 9])
        e.prop_alloc(models.X_NEQ_Y, [[0, 0], [1, 0]])     # y loses the interior value 5
        e.prop_alloc(models.X_NEQ_Y, [[1, 0], [2, 3]])     # two unassigned views: nothing to do
    assert dev.consistency()[0] == ora.consistency()[0] == 0
    _assert_same_state(dev, ora)
    assert int(dev.domain_sizes()[1]) == 9 and int(dev.domain_bits(0, 1)[1, 0]) == 0x3FF & ~(1 << 5)
    assert list(dev.active()) == [0, 1]
    # holes make two unassigned variables disjoint as sets: entailed without any assignment
    dev, ora = _engine(), _oracle(0)
    for e in (dev, ora):
        e.vars_alloc([0, 0], [5, 5])
        for v in (1, 3, 5):
            e.prop_alloc(models.X_NEQ_Y, [[0, 0], [-1, v]])   # x in {0, 2, 4}
        for v in (0, 2, 4):
            e.prop_alloc(models.X_NEQ_Y, [[1, 0], [-1, v]])   # y in {1, 3, 5}
        e.prop_alloc(models.X_NEQ_Y, [[0, 0], [1, 0]])
    assert dev.consistency()[0] == ora.consistency()[0] == 1
    _assert_same_state(dev, ora)


def test_bounds_land_on_set_elements():
    dev, ora = _engine(), _oracle(0)
    for e in (dev, ora):
        e.vars_alloc([0, 0, 0], [40, 40, 40])
        for v in (4, 5, 6, 20, 33, 34):
            e.prop_alloc(models.X_NEQ_Y, [[0, 0], [-1, v]])
        e.prop_alloc(models.X_LESS_Y, [[-1, 3], [0, 0]])                 # 3 < x -> lo jumps to 7
        e.prop_alloc(models.X_LESS_Y, [[0, 0], [-1, 35]])                # x < 35 -> hi jumps to 32
        e.prop_alloc(models.X_EQ_Y, [[0, 0], [1, 2]])                    # x == y + 2 (set intersection)
        e.prop_alloc(models.X_EQ_Y_PLUS_Z, [[2, 0], [0, 0], [-1, 1]])    # z == x + 1 (bounds)
        e.prop_alloc(models.X_GREATER_Y_PLUS_Z, [[2, 0], [1, 0], [-1, 1]])
    assert dev.consistency()[0] == ora.consistency()[0]
    _assert_same_state(dev, ora)
    lo, hi = dev.domains()
    assert (int(lo[0]), int(hi[0])) == (7, 32) and int(dev.domain_sizes()[0]) == 25


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("flavour", ["example", "distinct"])
def test_nqueens_full_search_on_sets(n, flavour):
    """example/src/nqueens.rs as written (IntervalSet::new(1, n)) and the test flavour of
    search/mod.rs:64-91: every node of the full tree against the set oracle; solution counts
    against all_solution.rs:67-74."""
    rd, _, _, _ = _compare_search(models.nqueens(n, flavour), 0)
    assert rd.status == 2
    assert rd.num_solution == GOLDEN["search"]["nqueens_all_solutions"]["counts"][n - 1]


@pytest.mark.parametrize("n,limit", [(12, 500), (30, 300), (64, 200), (200, 120)])
def test_nqueens_node_budget_on_sets(n, limit):
    _compare_search(models.nqueens(n), limit)


def test_nqueens_1000_on_sets():
    """C2 on the domain the reference example really uses: V=1000, P=1,498,500, 128 KB of bit
    sets; the first 30 DFS nodes bit-exact (values sets included) against the set oracle."""
    rd, ro, dev, ora = _compare_search(models.nqueens(1000), 30)
    _assert_same_state(dev, ora)


@pytest.mark.parametrize("n", [5, 8, 12])
@pytest.mark.parametrize("decompose", [False, True])
def test_all_interval_on_sets(n, decompose):
    _compare_search(models.all_interval(n, decompose_distinct=decompose), 0 if n <= 8 else 300)


def test_random_mixed_store_on_sets():
    rng = np.random.default_rng(4321)
    for trial in range(30):
        V = int(rng.integers(3, 40))
        lo = rng.integers(-20, 20, V).astype(np.int32)
        hi = (lo + rng.integers(0, 25, V)).astype(np.int32)
        dev, ora = _engine(), _oracle(trial % 2)
        for e in (dev, ora):
            e.vars_alloc(lo, hi)
        for _ in range(int(rng.integers(1, 60))):
            kind = int(rng.choice([0, 1, 1, 2, 3, 4, 5, 6, 7]))
            n_ops = {0: 2, 1: 2, 2: 2, 3: 3, 4: 3, 5: 3, 6: int(rng.integers(1, min(V, 12) + 1)), 7: 6}[kind]
            if kind == 7:
                vs = np.concatenate([rng.choice(V, 3, replace=False), rng.choice(V, 3, replace=False)])
            else:
                vs = rng.choice(V, n_ops, replace=False)
            ops = np.stack([vs, rng.integers(-5, 6, n_ops)], axis=1).astype(np.int32)
            if rng.random() < 0.25:
                ops[int(rng.integers(0, n_ops))] = (-1, int(rng.integers(-20, 30)))
            for e in (dev, ora):
                e.prop_alloc(kind, ops)
        ds, _ = dev.consistency()
        os_, _ = ora.consistency()
        assert ds == os_, trial
        if ds != -1:
            _assert_same_state(dev, ora)


def test_random_mixed_search_on_sets():
    """Seeded stores of binary / ternary / Distinct propagators searched for 150 nodes with the
    Enumerate distributor (posts XEqY / XNeqY against constants: interior removals by posted
    propagators) and with BinarySplit."""
    rng = np.random.default_rng(99)
    for trial in range(6):
        V = 10
        m = models.Model(f"mixed-{trial}", rng.integers(-4, 4, V).astype(np.int32), rng.integers(6, 14, V).astype(np.int32))
        for kind, n_ops in ((0, 2), (1, 2), (1, 2), (1, 2), (2, 2), (3, 3), (4, 3), (5, 3)):
            vs = rng.choice(V, n_ops, replace=False)
            m.add(kind, np.stack([vs, rng.integers(-3, 4, n_ops)], axis=1).astype(np.int32))
        m.add(models.DISTINCT, np.stack([rng.choice(V, 5, replace=False), rng.integers(-1, 2, 5)], axis=1).astype(np.int32))
        _compare_search(m, 150, distributor=trial % 2, val_sel=trial % 2)


def test_label_restore_on_sets():
    dev, ora = _engine(), _oracle(1)
    m = models.nqueens(10)
    for e in (dev, ora):
        m.load_into(e)
        assert e.consistency()[0] == 0
    labels = [(dev.label(), ora.label())]
    for var, val in [(0, 3), (1, 7), (2, 1), (3, 9)]:
        for e in (dev, ora):
            e.prop_alloc(models.X_EQ_Y, [[var, 0], [-1, val]])
        ds, os_ = dev.consistency()[0], ora.consistency()[0]
        assert ds == os_
        if ds == -1:
            break
        _assert_same_state(dev, ora)
        labels.append((dev.label(), ora.label()))
    for dl, ol in reversed(labels):
        dev.restore(dl)
        ora.restore(ol)
        _assert_same_state(dev, ora)
        for e in (dev, ora):
            e.prop_alloc(models.X_NEQ_Y, [[4, 0], [-1, 4]])
        assert dev.consistency()[0] == ora.consistency()[0]
        _assert_same_state(dev, ora)
        dev.restore(dl)
        ora.restore(ol)


def test_var_update_and_unsupported_on_sets():
    from pcp_b200 import PcpError
    dev, ora = _engine(), _oracle(0)
    for e in (dev, ora):
        e.vars_alloc([0, 0], [9, 9])
        for v in (3, 4):
            e.prop_alloc(models.X_NEQ_Y, [[0, 0], [-1, v]])
        e.prop_alloc(models.X_LESS_Y, [[0, 0], [1, 0]])
        assert e.consistency()[0] == 0
        assert e.var_update(0, 3, 8) is True          # cur /\ [3, 8] = {5..8}
        assert e.consistency()[0] == 0
    _assert_same_state(dev, ora)
    assert int(dev.domains()[0][0]) == 5
    with pytest.raises(PcpError):
        dev.prop_alloc(models.X_EQ_Y_MUL_Z, [[0, 0], [1, 0], [-1, 2]])
    with pytest.raises(PcpError):
        dev.prop_alloc(models.ALL_EQUAL, [[0, 0], [1, 0]])
    s = dev.sum_alloc([[0, 0], [1, 0]])
    with pytest.raises(PcpError):
        dev.prop_alloc(models.X_LESS_Y, [[-2 - s, 0], [-1, 30]])


def test_incremental_mode_on_sets():
    _compare_search(models.nqueens(40), 250, dev_kw={"incremental": True})
    _compare_search(models.all_interval(9), 0, dev_kw={"incremental": True})
