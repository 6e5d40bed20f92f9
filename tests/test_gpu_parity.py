"""GPU parity tests: the device fixpoint (through the C ABI of include/pcp_b200.h) against
the CPU oracle on the same inputs -- bit-exact domains, status and `active` set after every
fixpoint.  Vectors are the reference's own (tests/golden/reference_vectors.json) plus
seeded workloads of the BASELINE configs at sizes the oracle finishes in seconds."""
import json
import os

import numpy as np
import pytest

from pcp_b200 import models

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as f:
    GOLDEN = json.load(f)


def _oracle(variant=1):
    from oracle.oracle_api import OracleEngine
    return OracleEngine(variant)


def _engine(**kw):
    from pcp_b200 import Engine
    return Engine(**kw)


def _assert_same_state(dev, ora, check_active=True):
    dlo, dhi = dev.domains()
    olo, ohi = ora.domains()
    assert (dlo == olo).all() and (dhi == ohi).all()
    if check_active and dev.num_props:
        assert (dev.active() == ora.active()).all()


@pytest.mark.parametrize("vec", GOLDEN["propagators"], ids=lambda v: v["name"])
def test_reference_vector(vec):
    """Every known-answer vector of the reference's propagator tests, run to the fixpoint
    through Consistency::consistency on the device, against the oracle and against the
    reference's expected entailment."""
    d = np.array(vec["domains"], np.int32)
    kind = models.KIND_BY_NAME[vec["kind"]]
    dev, ora = _engine(), _oracle(0)
    for e in (dev, ora):
        e.vars_alloc(d[:, 0], d[:, 1])
        e.prop_alloc(kind, vec["ops"])
    ds, dstats = dev.consistency()
    os_, _ = ora.consistency()
    assert ds == os_, vec["ref"]
    if not vec["ok"] or vec["after"] == -1:
        assert ds == -1
        return
    if vec["after"] == 1:
        assert ds == 1
    _assert_same_state(dev, ora)
    assert dstats.propagations >= 1
    if "domains_after" in vec:
        lo, hi = dev.domains()
        assert [[int(a), int(b)] for a, b in zip(lo, hi)] == vec["domains_after"]
    # single-pass expectations of the fixture that are also fixpoints: the narrowed
    # variables listed in `delta` did change
    lo, hi = dev.domains()
    for var, _ev in vec["delta"]:
        assert (int(lo[var]), int(hi[var])) != tuple(vec["domains"][var])


def test_chained_lt_and_nqueens_root():
    """propagation/store.rs:362-392."""
    for n, status in GOLDEN["search"]["chained_lt"]["status"].items():
        dev, ora = _engine(), _oracle()
        m = models.chained_lt(int(n))
        m.load_into(dev)
        m.load_into(ora)
        assert dev.consistency()[0] == status == ora.consistency()[0]
        if status != -1:
            _assert_same_state(dev, ora)
    for n, status in GOLDEN["search"]["nqueens_root"]["status"].items():
        dev = _engine()
        models.nqueens(int(n)).load_into(dev)
        assert dev.consistency()[0] == status
        lo, hi = dev.domains()
        assert (lo == 1).all() and (hi == int(n)).all()


def test_long_chain_many_iterations():
    """A 300-long x_i < x_{i+1} chain on [0, 299]: ~300 dependent propagation rounds, the
    worst case for the device loop (one variable moves per iteration)."""
    dev, ora = _engine(), _oracle()
    m = models.chained_lt(300, 0, 299)
    m.load_into(dev)
    m.load_into(ora)
    ds, st = dev.consistency()
    assert ds == ora.consistency()[0] == 1
    _assert_same_state(dev, ora)
    lo, hi = dev.domains()
    assert (lo == np.arange(300)).all() and (hi == np.arange(300)).all()
    dev2 = _engine()
    models.chained_lt(301, 0, 299).load_into(dev2)
    assert dev2.consistency()[0] == -1


@pytest.mark.parametrize("n", [10300, 10700, 13000])
def test_snapshot_capacity_boundary(n):
    """Stores just below / above the size whose domains still fit the CTA's shared-memory
    snapshot next to the TMA ring (both kernel variants, same results): a loose x_i < x_{i+1}
    chain over n variables (two propagation waves) and 40 search nodes on it."""
    m = models.chained_lt(n, 0, n + 5)
    dev, ora = _engine(), _oracle(2)
    m.load_into(dev)
    m.load_into(ora)
    assert dev.consistency()[0] == ora.consistency()[0] == 0
    _assert_same_state(dev, ora)
    _compare_search(m, 40, oracle_variant=2)


_ORACLE_TRACES = {}  # (model name, node limit) -> the oracle's (result, trace) of the last _compare_search


def _compare_search(model, node_limit, dev_kw=None, all_solutions=True, oracle_variant=1, **skw):
    dev, ora = _engine(**(dev_kw or {})), _oracle(oracle_variant)
    model.load_into(dev)
    model.load_into(ora)
    rd, td = dev.search(node_limit=node_limit, all_solutions=all_solutions, trace=node_limit or 100000,
                        trace_domains=True, **skw)
    ro, to = ora.search(node_limit=node_limit, all_solutions=all_solutions, trace=node_limit or 100000,
                        trace_domains=True, **skw)
    _ORACLE_TRACES.clear()
    _ORACLE_TRACES[(model.name, node_limit)] = (ro, to)
    assert rd.num_nodes == ro.num_nodes
    assert rd.status == ro.status
    assert rd.num_solution == ro.num_solution and rd.num_failed_node == ro.num_failed_node
    assert (td["status"] == to["status"]).all()
    ok = td["status"] != -1
    assert (td["hash"] == to["hash"]).all()
    assert (td["lo"][ok] == to["lo"][ok]).all() and (td["hi"][ok] == to["hi"][ok]).all()
    return rd, ro, dev, ora


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("flavour", ["example", "distinct"])
def test_nqueens_full_search(n, flavour):
    """C1 and the search goldens (all_solution.rs:67-74): every node of the full search tree
    bit-exact against the oracle; solution counts against the reference's table."""
    rd, _, _, _ = _compare_search(models.nqueens(n, flavour), 0)
    assert rd.status == 2
    assert rd.num_solution == GOLDEN["search"]["nqueens_all_solutions"]["counts"][n - 1]


@pytest.mark.parametrize("n,limit", [(12, 600), (30, 300), (64, 200), (200, 120)])
def test_nqueens_node_budget(n, limit):
    """C2 at reduced N: fixed DFS node budget (StopNode semantics), per-node bit-exact."""
    _compare_search(models.nqueens(n), limit)


def test_nqueens_1000_node_budget():
    """C2 at full size: V=1000, P=1,498,500; first 40 DFS nodes bit-exact vs the oracle."""
    rd, ro, dev, ora = _compare_search(models.nqueens(1000), 40)
    assert rd.propagations >= 40 * 1_498_500 * 0  # counts differ by schedule; just present
    _assert_same_state(dev, ora, check_active=True)


def test_nqueens_deep_crawl():
    """Deep DFS nodes on bounds-only domains: variables crawl across runs of assigned values
    (x_neq_y.rs:82-93 one step at a time in the reference; the device jumps) -- 1500 nodes of
    N=300 and 150 nodes of N=1000, per-node bit-exact incl. failed-node statuses."""
    _compare_search(models.nqueens(300), 1000, oracle_variant=2)   # 2 = the oracle's flat variant (same results, faster)
    _compare_search(models.nqueens(1000), 120, oracle_variant=2)


@pytest.mark.parametrize("n,nodes", [(60, 400), (250, 150)])
def test_host_driven_dfs_matches_oracle(n, nodes):
    """The node loop driven from the host through the store surface alone (restore, post the
    BinarySplit constraint, consistency, read domains, label) -- the path a libpcp host takes
    (search/branching/branch.rs:36-55) -- instead of the device-resident search."""
    from pcp_b200 import parallel
    dev, ora = _engine(), _oracle()
    m = models.nqueens(n)
    m.load_into(dev)
    m.load_into(ora)
    stack = []
    for k in range(nodes):
        if k:
            if not stack:
                break
            (dl, ol), d = stack.pop()
            dev.restore(dl)
            ora.restore(ol)
            parallel.post_decision(dev, d)
            parallel.post_decision(ora, d)
        ds, os_ = dev.consistency()[0], ora.consistency()[0]
        assert ds == os_, k
        if ds == -1:
            continue
        _assert_same_state(dev, ora)
        if ds == 0:
            lo, hi = dev.domains()
            var, val = parallel.select_branch(lo, hi)
            labels = (dev.label(), ora.label())
            stack.append((labels, (var, val, 1)))
            stack.append((labels, (var, val, 0)))


@pytest.mark.parametrize("n", [5, 8, 12, 30])
@pytest.mark.parametrize("decompose", [False, True])
def test_all_interval(n, decompose):
    """C3 at reduced N: Distinct-heavy model + 2-way disjunctions of XEqYPlusZ."""
    limit = 0 if n <= 8 else 300
    _compare_search(models.all_interval(n, decompose_distinct=decompose), limit)


def test_all_interval_500_node_budget():
    """C3 at full size (N=500, V=999)."""
    _compare_search(models.all_interval(500), 60)


@pytest.mark.parametrize("V,P,seed", [(50, 200, 1), (1000, 10_000, 2), (20_000, 200_000, 3)])
def test_random_arith_csp(V, P, seed):
    """C4 at reduced size: single fixpoint of XEqYPlusZ propagators, bit-exact lo/hi."""
    m = models.random_arith_csp(V, P, seed=seed)
    dev, ora = _engine(), _oracle()
    m.load_into(dev)
    m.load_into(ora)
    ds, st = dev.consistency()
    os_, _ = ora.consistency()
    assert ds == os_ and ds != -1
    _assert_same_state(dev, ora)
    assert st.iterations >= 1


def test_random_arith_csp_full_size():
    """C4 at full size: V=100,000, P=1,000,000 (L2-gather path, no shared-memory snapshot)."""
    m = models.random_arith_csp()
    dev, ora = _engine(), _oracle()
    m.load_into(dev)
    m.load_into(ora)
    ds, st = dev.consistency()
    os_, _ = ora.consistency()
    assert ds == os_ == 1  # bound propagation alone pins every variable to the planted solution
    _assert_same_state(dev, ora)
    lo, hi = dev.domains()
    assert (lo == hi).all()
    # idempotence: a second fixpoint changes nothing
    lo0, hi0 = dev.domains()
    ds2, st2 = dev.consistency()
    lo1, hi1 = dev.domains()
    assert ds2 == ds and (lo0 == lo1).all() and (hi0 == hi1).all()
    assert st2.iterations == 1


def test_random_mixed_store():
    """Seeded random stores mixing every device propagator kind, with constants and offsets."""
    rng = np.random.default_rng(1234)
    for trial in range(30):
        V = int(rng.integers(3, 40))
        lo = rng.integers(-20, 20, V).astype(np.int32)
        hi = (lo + rng.integers(0, 25, V)).astype(np.int32)
        dev, ora = _engine(), _oracle(trial % 2)
        for e in (dev, ora):
            e.vars_alloc(lo, hi)
        for _ in range(int(rng.integers(1, 60))):
            kind = int(rng.integers(0, 10))
            n_ops = {0: 2, 1: 2, 2: 2, 3: 3, 4: 3, 5: 3, 6: int(rng.integers(1, min(V, 12) + 1)), 7: 6, 8: 3,
                     9: int(rng.integers(1, min(V, 6) + 1))}[kind]
            if kind == 7:
                a = rng.choice(V, 3, replace=False)
                b = rng.choice(V, 3, replace=False)
                vs = np.concatenate([a, b])
            else:
                vs = rng.choice(V, n_ops, replace=False)
            ops = np.stack([vs, rng.integers(-5, 6, n_ops)], axis=1).astype(np.int32)
            if rng.random() < 0.2:  # a constant operand
                j = int(rng.integers(0, n_ops))
                ops[j] = (-1, int(rng.integers(-20, 30)))
            for e in (dev, ora):
                e.prop_alloc(kind, ops)
        ds, _ = dev.consistency()
        os_, _ = ora.consistency()
        assert ds == os_, trial
        if ds != -1:
            _assert_same_state(dev, ora)


def test_label_restore_and_tail():
    """Snapshot::label/restore through the ABI; branch constraints live in the tail."""
    dev, ora = _engine(), _oracle()
    m = models.nqueens(10)
    for e in (dev, ora):
        m.load_into(e)
        assert e.consistency()[0] == 0
    labels = [(dev.label(), ora.label())]
    for step, (var, val) in enumerate([(0, 3), (1, 7), (2, 1), (3, 9)]):
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
        assert dev.num_props == ora.num_props
        _assert_same_state(dev, ora)
        for e in (dev, ora):
            e.prop_alloc(models.X_LESS_Y, [[4, 0], [-1, 4]])
        assert dev.consistency()[0] == ora.consistency()[0]
        _assert_same_state(dev, ora)
        dev.restore(dl)
        ora.restore(ol)


@pytest.mark.parametrize("tail_limit", [1, 3, 50])
def test_small_tail_limit_rebuilds_csr(tail_limit):
    """CSR rebuilds in the middle of a search (tail limit reached) keep parity."""
    _compare_search(models.nqueens(9), 0, dev_kw={"tail_limit": tail_limit})


@pytest.mark.parametrize("model", [models.nqueens(9), models.nqueens(40), models.all_interval(9),
                                   models.nqueens(9, "distinct")], ids=lambda m: m.name)
def test_incremental_mode_same_fixpoints(model):
    """PCP_FLAG_INCREMENTAL: same per-node domains and statuses, fewer propagations."""
    limit = 0 if model.name.startswith("nqueens-9") else 250
    rd, ro, _, _ = _compare_search(model, limit, dev_kw={"incremental": True})
    dev_full = _engine()
    model.load_into(dev_full)
    rf, _ = dev_full.search(node_limit=limit, all_solutions=True)
    assert rf.num_nodes == rd.num_nodes
    assert rd.propagations <= rf.propagations


@pytest.mark.parametrize("model,nodes", [(models.nqueens(1000), 150), (models.all_interval(500), 60)],
                         ids=["c2-nqueens-1000", "c3-all-interval-500"])
def test_incremental_mode_full_size(model, nodes):
    """PCP_FLAG_INCREMENTAL at the sizes bench.py reports it for (C2, C3; C5 rides in
    test_nqueens_5000_full_size): every node's status and domains against the oracle."""
    rd, ro, _, _ = _compare_search(model, nodes, dev_kw={"incremental": True}, oracle_variant=2)
    assert rd.num_nodes == nodes


def test_bounds_outside_the_contract_are_rejected():
    """Device bound arithmetic is 32-bit: bounds, offsets, view ranges and worst-case Sum ranges
    must stay inside (-2^29, 2^29) (SURVEY 7 iv); anything else is PCP_ERR_INVALID at
    allocation time instead of a silent wrap-around."""
    from pcp_b200 import ContractViolation
    L = 1 << 29
    dev = _engine()
    for lo, hi in ((0, L), (-L, 0), (-(1 << 31), (1 << 31) - 1)):
        with pytest.raises(ContractViolation):
            dev.vars_alloc([lo], [hi])
    assert dev.num_vars == 0
    dev.vars_alloc([-(L - 1), 0, 0], [L - 1, 10, 10])
    with pytest.raises(ContractViolation):
        dev.prop_alloc(models.X_LESS_Y, [[1, L], [2, 0]])          # offset
    with pytest.raises(ContractViolation):
        dev.prop_alloc(models.X_LESS_Y, [[0, 5], [1, 0]])          # range of the view x0 + 5
    with pytest.raises(ContractViolation):
        dev.prop_alloc(models.X_LESS_Y, [[1, 0], [-1, -L]])        # constant
    with pytest.raises(ContractViolation):
        dev.sum_alloc([[0, 0], [1, 0]])                            # worst case of the sum
    assert dev.num_props == 0
    dev.prop_alloc(models.X_LESS_Y, [[1, 0], [0, 0]])              # at the limit: fine
    assert dev.consistency()[0] == 0
    lo, hi = dev.domains()
    assert int(lo[0]) == 1 and int(hi[1]) == 10


def test_store_calls_are_refused_while_a_device_search_is_open():
    """The device-resident search caches launch parameters: store entry points that would
    reallocate or change state underneath it return PCP_ERR_INVALID until pcp_search_close."""
    from pcp_b200 import ContractViolation
    dev = _engine()
    models.nqueens(12).load_into(dev)
    h = dev.search_open(all_solutions=True)
    assert h.step(5).status == 0
    for call in (lambda: dev.prop_alloc(models.X_LESS_Y, [[0, 0], [1, 0]]), lambda: dev.consistency(),
                 lambda: dev.label(), lambda: dev.restore(0), lambda: dev.vars_alloc([0], [1]),
                 lambda: dev.var_update(0, 1, 1)):
        with pytest.raises(ContractViolation):
            call()
    res = None
    for _ in range(100000):
        res = h.step(50)
        if res.status != 0:
            break
    h.close()
    assert res.num_solution == 14200
    assert dev.consistency()[0] in (0, 1, -1)  # usable again after close


@pytest.mark.parametrize("model,nodes", [(models.nqueens(64), 400), (models.nqueens(300), 300), (models.all_interval(12), 300),
                                         (models.nqueens(20, "distinct"), 400)], ids=lambda m: getattr(m, "name", str(m)))
def test_incremental_mode_host_driven_and_device_search(model, nodes):
    """PCP_FLAG_INCREMENTAL through both search paths: the device-resident search (a node below the
    root evaluates its posted constraint and the row of its variable in iteration 0) and the
    host-driven node loop -- the same per-node statuses and domains as the oracle's full fixpoints."""
    _compare_search(model, nodes, dev_kw={"incremental": True})
    _compare_search(model, nodes, dev_kw={"incremental": True, "host_search": True})


def test_var_update_and_contract_violations():
    from pcp_b200 import ContractViolation
    dev, ora = _engine(incremental=True), _oracle()
    for e in (dev, ora):
        models.chained_lt(6, 0, 20).load_into(e)
        assert e.consistency()[0] == 0
        assert e.var_update(2, 10, 12) is True
        assert e.consistency()[0] == 0
    _assert_same_state(dev, ora)
    assert dev.var_update(2, 5, 4) is False
    with pytest.raises(ContractViolation):
        dev.var_update(2, 0, 100)
    with pytest.raises(ContractViolation):
        dev.var_update(99, 0, 1)
    with pytest.raises(ContractViolation):
        dev.vars_alloc([3], [2])
    with pytest.raises(ContractViolation):
        dev.prop_alloc(models.X_LESS_Y, [[0, 0], [77, 0]])
    with pytest.raises(ContractViolation):
        dev.prop_alloc(models.X_LESS_Y, [[0, 0], [0, 1]])  # double subscription (indexed_deps.rs:69-77)
    with pytest.raises(ContractViolation):
        dev.restore(12345)
    n = dev.num_props
    with pytest.raises(ContractViolation):  # all-or-nothing batch
        dev.props_alloc(models.X_NEQ_Y, np.array([[[0, 0], [1, 0]], [[0, 0], [99, 0]]], np.int32))
    assert dev.num_props == n
    assert dev.consistency()[0] == ora.consistency()[0]
    _assert_same_state(dev, ora)


def test_branch_and_bound_and_stop_node():
    """search/branch_and_bound.rs:112-138 and search/stop_node.rs:82-104 through the device."""
    for mode, key in ((2, "maximize"), (1, "minimize")):
        dev = _engine()
        dev.vars_alloc([0, 0], [10, 10])
        dev.prop_alloc(models.X_LESS_Y, [[0, 0], [1, 0]])
        res, _ = dev.search(all_solutions=True, bb_mode=mode, bb_var=0)
        assert res.status == 2 and res.has_bb_value
        assert res.bb_value == GOLDEN["search"]["branch_and_bound"][key]
    g = GOLDEN["search"]["stop_node"]
    dev = _engine()
    models.nqueens(g["n"], "distinct").load_into(dev)
    res, _ = dev.search(node_limit=g["limit"], all_solutions=True)
    assert res.status == g["status"] and res.num_nodes == g["num_nodes"]


def test_size_independent_properties_nqueens_1000():
    """Full-size properties that need no oracle: the fixpoint is idempotent, monotone
    (domains only shrink along a branch), and restoring a label gives back the exact
    parent domains."""
    dev = _engine()
    models.nqueens(1000).load_into(dev)
    assert dev.consistency()[0] == 0
    root = dev.domains()
    l0 = dev.label()
    prev = root
    for var, val in [(0, 1), (1, 3), (2, 5), (3, 2), (4, 4)]:
        dev.prop_alloc(models.X_EQ_Y, [[var, 0], [-1, val]])
        st, stats = dev.consistency()
        assert st == 0
        cur = dev.domains()
        assert (cur[0] >= prev[0]).all() and (cur[1] <= prev[1]).all()
        st2, stats2 = dev.consistency()
        again = dev.domains()
        assert st2 == st and (again[0] == cur[0]).all() and (again[1] == cur[1]).all()
        assert stats2.iterations == 1
        prev = cur
    # queens 0..4 fixed: their values and diagonals are trimmed from the bounds of the others
    assert int(prev[0][5]) >= 1 and int(prev[0][0]) == int(prev[1][0]) == 1
    dev.restore(l0)
    back = dev.domains()
    assert (back[0] == root[0]).all() and (back[1] == root[1]).all()


def test_nqueens_5000_full_size():
    """C5 at full size: V=5000, P=37,492,500 (600 MB of descriptors, 300 MB compact stream, rows
    of 14,997 entries -- longer than the row staging area).  The first 32 DFS nodes bit-exact
    against the oracle's flat variant -- the descent passes nodes that need worklist iterations
    after the sweep (asserted) -- for the default engine and for PCP_FLAG_INCREMENTAL; then
    size-independent properties deeper in the tree: idempotence, monotonicity along a branch,
    exact restore."""
    m = models.nqueens(5000)
    rd, ro, dev, ora = _compare_search(m, 32, oracle_variant=2)
    assert rd.iterations > rd.num_nodes  # at least one node went through a worklist iteration
    del ora
    _, to = _ORACLE_TRACES[(m.name, 32)]
    inc = _engine(incremental=True)
    m.load_into(inc)
    ri, ti = inc.search(node_limit=32, all_solutions=True, trace=32, trace_domains=True)
    assert ri.num_nodes == 32 and (ti["status"] == to["status"]).all() and (ti["hash"] == to["hash"]).all()
    ok = ti["status"] != -1
    assert (ti["lo"][ok] == to["lo"][ok]).all() and (ti["hi"][ok] == to["hi"][ok]).all()
    assert ri.propagations < rd.propagations
    del inc
    dev2 = _engine()
    m.load_into(dev2)
    assert dev2.consistency()[0] == 0
    root = dev2.domains()
    l0 = dev2.label()
    prev = root
    for var, val in [(0, 1), (1, 3), (2, 5), (7, 2499), (8, 4999)]:
        dev2.prop_alloc(models.X_EQ_Y, [[var, 0], [-1, val]])
        st, _ = dev2.consistency()
        assert st == 0
        cur = dev2.domains()
        assert (cur[0] >= prev[0]).all() and (cur[1] <= prev[1]).all()
        st2, stats2 = dev2.consistency()
        again = dev2.domains()
        assert st2 == st and (again[0] == cur[0]).all() and (again[1] == cur[1]).all() and stats2.iterations == 1
        prev = cur
    # queen 0 sits on value 1, queen 8 on 4999+1-... : the bounds of the free queens moved off them
    assert int(prev[0][5]) >= 2 and int(prev[1][5]) <= 5000
    dev2.restore(l0)
    back = dev2.domains()
    assert (back[0] == root[0]).all() and (back[1] == root[1]).all()


def test_sum_views():
    """term/sum.rs:56-92: a Sum operand reads [sum lo, sum hi]; with more than one term its
    update never prunes (overlap test only, can fail); a single-term Sum delegates.  Random
    stores where cmp propagators take Sum operands, device vs the oracle's view trees."""
    from pcp_b200 import PcpError
    rng = np.random.default_rng(99)
    for trial in range(40):
        V = int(rng.integers(6, 30))
        lo = rng.integers(-10, 10, V).astype(np.int32)
        hi = (lo + rng.integers(0, 12, V)).astype(np.int32)
        dev, ora = _engine(), _oracle(1)
        for e in (dev, ora):
            e.vars_alloc(lo, hi)
        for _ in range(int(rng.integers(1, 25))):
            kind = int(rng.integers(0, 6))
            n_ops = 2 if kind < 3 else 3
            perm = rng.permutation(V)
            ops, used = [], 0
            for j in range(n_ops):
                if rng.random() < 0.4 and used + 4 <= V:  # a Sum operand over 1..3 fresh variables
                    k = int(rng.integers(1, 4))
                    terms = [[int(perm[used + t]), int(rng.integers(-3, 4))] for t in range(k)]
                    if rng.random() < 0.2:
                        terms.append([-1, int(rng.integers(-5, 6))])  # constant term
                    used += k
                    sid = [e.sum_alloc(terms) for e in (dev, ora)]
                    assert sid[0] == sid[1]
                    ops.append([-2 - sid[0], int(rng.integers(-4, 5))])
                else:
                    ops.append([int(perm[used]), int(rng.integers(-4, 5))])
                    used += 1
            for e in (dev, ora):
                e.prop_alloc(kind, ops)
        ds, _ = dev.consistency()
        os_, _ = ora.consistency()
        assert ds == os_, trial
        if ds != -1:
            _assert_same_state(dev, ora)
    # a Sum inside Distinct has no device lowering: loud error, not a CPU path
    dev = _engine()
    dev.vars_alloc([0, 0, 0], [5, 5, 5])
    s = dev.sum_alloc([[0, 0], [1, 0]])
    with pytest.raises(PcpError):
        dev.prop_alloc(models.DISTINCT, [[-2 - s, 0], [2, 0]])
    # the same variable directly and inside the sum: double subscription (indexed_deps.rs:69-77)
    from pcp_b200 import ContractViolation
    with pytest.raises(ContractViolation):
        dev.prop_alloc(models.X_LESS_Y, [[0, 0], [-2 - s, 0]])


def test_sum_capacity_constraint_search():
    """A small knapsack-like model with Sum operands through the search driver: per-node parity."""
    n = 8
    m = models.Model("sum-capacity", np.zeros(n, np.int32), np.full(n, 3, np.int32))
    m.sums.append(np.array([[i, 0] for i in range(n)], np.int32))
    m.sums.append(np.array([[i, 0] for i in range(0, n, 2)], np.int32))
    m.add(models.X_LESS_Y, [[-2, 0], [-1, 14]])            # sum(x) < 14
    m.add(models.X_LESS_Y, [[-1, 5], [-2, 0]])             # 5 < sum(x)
    m.add(models.X_LESS_Y_PLUS_Z, [[-3, 0], [1, 0], [3, 2]])  # sum(x_even) < x1 + x3 + 2
    ops = np.zeros((n - 1, 2, 2), np.int32)
    ops[:, 0, 0] = np.arange(n - 1)
    ops[:, 1, 0] = np.arange(1, n)
    ops[:, 1, 1] = 1
    m.add(models.X_LESS_Y, ops)                             # x_i < x_{i+1} + 1
    rd, _, _, _ = _compare_search(m, 400)
    assert rd.num_nodes == 71 and rd.num_solution == 3 and rd.num_failed_node == 33


@pytest.mark.parametrize("slice_nodes", [1, 5, 64])
def test_resumable_search_slices_match_one_shot(slice_nodes):
    """pcp_search_open/step/close in budget slices (the multi-GPU driver's pattern: the device
    search stops and resumes at arbitrary nodes, mixing barrier steps and fast descents) visits
    the same tree as one pcp_search_run: node, failure and solution counts (n-queens N=8: 92
    solutions, all_solution.rs:67-74)."""
    m = models.nqueens(8)
    one = _engine()
    m.load_into(one)
    r1, _ = one.search(all_solutions=True)
    dev = _engine()
    m.load_into(dev)
    h = dev.search_open(all_solutions=True)
    res = None
    for _ in range(100000):
        res = h.step(slice_nodes)
        if res.status != 0:
            break
    h.close()
    assert res is not None and res.status == r1.status == 2
    assert (res.num_nodes, res.num_solution, res.num_failed_node) == (r1.num_nodes, r1.num_solution, r1.num_failed_node)
    assert res.num_solution == 92


def test_sharded_search_single_rank_covers_the_tree():
    """pcp_b200.parallel on the device engine (world = 1): the frontier's subtrees together hold
    every solution of n-queens N=7 (40, all_solution.rs:67-74)."""
    from pcp_b200 import parallel
    dev = _engine()
    models.nqueens(7).load_into(dev)
    out = parallel.sharded_search(dev, 0, 1, node_budget=10**6, sync_every=50, parts_per_rank=6)
    assert out["solutions"] == 40 and out["subtrees"] == out["frontier"] >= 6


def _nccl_rank(rank, world, port, n, q):
    import os as _os, sys as _sys
    _sys.path.insert(0, ROOT)
    _os.environ["MASTER_ADDR"] = "127.0.0.1"
    _os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from pcp_b200 import Engine, parallel
    out = {}
    for key, kw in (("all", {}), ("bb", {"bb_mode": 1, "bb_var": 3})):
        e = Engine(device=rank, max_labels=1 << 12)
        models.nqueens(n).load_into(e)
        flag = parallel.StopFlag(torch.device("cuda", rank))
        out[key] = parallel.sharded_search(e, rank, world, node_budget=10**6, sync_every=32, stop_flag=flag, **kw)
        e.close()
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_search_two_gpus_nccl():
    """The N>1 path with the *device* engine under NCCL (world = 2, one GPU per rank): the two
    ranks' subtrees hold every solution of n-queens N=9 (352, all_solution.rs:67-74), the
    collectives match, and with BranchAndBound both ranks end with the incumbent one process
    finds alone (the all-reduced minimum of q_3)."""
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_rank, args=(r, 2, port, 9, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res[0]["all"]["solutions"] + res[1]["all"]["solutions"] == 352
    assert res[0]["all"]["exchanges"] == res[1]["all"]["exchanges"] >= 1
    solo = _oracle(2)
    models.nqueens(9).load_into(solo)
    ref, _ = solo.search(all_solutions=True, bb_mode=1, bb_var=3)
    assert res[0]["bb"]["incumbent"] == res[1]["bb"]["incumbent"] == ref.bb_value


@pytest.mark.parametrize("k,mode,fused", [(2, "interval", True), (5, "interval", True), (5, "interval", False),
                                          (4, "set", True), (6, "incremental", True), (3, "distinct", True)])
def test_consistency_batch_matches_lone_engines(k, mode, fused, monkeypatch):
    """pcp_consistency_batch: k engines, each on its own subtree of n-queens N=120, advance a
    host-driven DFS in lockstep (the k fixpoints of a round are launched together and run side
    by side on the GPU); after every round each engine's status, domains and `active` set equal
    those of the oracle replaying the same subtree alone.  Engines with the same kernel and geometry
    share ONE launch (pcp_fixpoint_batch_kernel, a group of CTAs per engine); PCP_BATCH=unfused keeps
    one launch per engine on its own stream.  Both, on Interval and IntervalSet domains, with the
    incremental flag, and with an n-ary Distinct in the store (the `full` kernel variant)."""
    import pcp_b200
    from pcp_b200 import parallel
    from oracle.oracle_api import SET, TUNED
    if not fused:
        monkeypatch.setenv("PCP_BATCH", "unfused")
    m = models.nqueens(40, "distinct") if mode == "distinct" else models.nqueens(120)
    kw = {"interval_set": True} if mode == "set" else ({"incremental": True} if mode == "incremental" else {})
    probe = _engine(**kw)
    m.load_into(probe)
    paths = parallel.expand_frontier(probe, parts=k)[:k]
    devs, oras, stacks = [], [], []
    for p in paths:
        d, o = _engine(**kw), _oracle(SET + TUNED if mode == "set" else 2)
        for e in (d, o):
            m.load_into(e)
            e.consistency()
            parallel.enter_subtree(e, e.label(), p)
        d.set_grid_limit(148 // k)
        devs.append(d)
        oras.append(o)
        stacks.append(None)  # None: the subtree's root is the next node
    for rnd in range(40):
        live = [i for i in range(k) if stacks[i] is None or stacks[i]]
        if not live:
            break
        for i in live:
            if stacks[i] is None:
                stacks[i] = []
                continue
            (dl, ol), dec = stacks[i].pop()
            devs[i].restore(dl)
            oras[i].restore(ol)
            parallel.post_decision(devs[i], dec)
            parallel.post_decision(oras[i], dec)
        sts, stats = pcp_b200.consistency_batch([devs[i] for i in live])
        assert sum(int(s.launches) for s in stats) == (1 if fused and len(live) > 1 else len(live))
        for i, st in zip(live, sts):
            assert st == oras[i].consistency()[0], (rnd, i)
            if st == -1:
                continue
            _assert_same_state(devs[i], oras[i])
            if st == 0:
                lo, hi = devs[i].domains()
                var, val = parallel.select_branch(lo, hi)
                labels = (devs[i].label(), oras[i].label())
                stacks[i].append((labels, (var, val, 1)))
                stacks[i].append((labels, (var, val, 0)))
        assert all(s.propagations > 0 for s in stats)


def test_batched_launch_at_c2_size():
    """The bench's configuration at full size: forks of one n-queens N=1000 engine, two CTAs each, one
    batched launch per round (compact descriptor stream, CTA 0 sweeping with the others); after every
    round each context's status and domains equal those of an oracle replaying the same subtree."""
    import pcp_b200
    from pcp_b200 import parallel
    k = 4
    m = models.nqueens(1000)
    first = _engine()
    m.load_into(first)
    paths = parallel.expand_frontier(first, parts=k)[:k]
    devs = [first] + [first.fork() for _ in range(k - 1)]
    oras, stacks = [], []
    for d, p in zip(devs, paths):
        o = _oracle(2)
        m.load_into(o)
        o.consistency()
        d.set_grid_limit(2)
        for e in (d, o):
            parallel.enter_subtree(e, e.label(), p)
        oras.append(o)
        stacks.append(None)
    for rnd in range(6):
        for i in range(k):
            if stacks[i] is None:
                stacks[i] = []
                continue
            assert stacks[i], "the first nodes of n-queens 1000 do not fail"
            (dl, ol), dec = stacks[i].pop()
            devs[i].restore(dl)
            oras[i].restore(ol)
            parallel.post_decision(devs[i], dec)
            parallel.post_decision(oras[i], dec)
        sts, stats = pcp_b200.consistency_batch(devs)
        assert sum(int(s.launches) for s in stats) == 1
        for i, st in enumerate(sts):
            assert st == oras[i].consistency()[0], (rnd, i)
            assert st == 0
            _assert_same_state(devs[i], oras[i])
            lo, hi = devs[i].domains()
            var, val = parallel.select_branch(lo, hi)
            labels = (devs[i].label(), oras[i].label())
            stacks[i].append((labels, (var, val, 1)))
            stacks[i].append((labels, (var, val, 0)))
    for d in devs[1:]:
        d.close()


def test_fork_shares_the_static_model_and_lets_go_of_it():
    """pcp_engine_fork: the child starts in the parent's state and shares its descriptors and
    reactor on the device; both then go their own way (own tail, domains, `active`, trail, labels).
    An engine that must rewrite its static part -- reactor rebuild when the tail limit is reached,
    restore to a label older than the shared prefix -- lets go of the shared block; the other keeps
    working on it.  Every engine against its own oracle after every step."""
    from pcp_b200 import parallel
    m = models.nqueens(40)
    parent, po = _engine(tail_limit=6), _oracle(2)
    early = None
    for e in (parent, po):
        e.vars_alloc(m.lo, m.hi)
        e.props_alloc(*m.batches[0][:1], m.batches[0][1][:10])      # ten propagators, then a label
    early = (parent.label(), po.label())
    for e in (parent, po):
        e.props_alloc(m.batches[0][0], m.batches[0][1][10:])
        e.props_alloc(*m.batches[1])
        assert e.consistency()[0] == 0
    child, co = parent.fork(), _oracle(2)
    m.load_into(co)
    assert co.consistency()[0] == 0
    _assert_same_state(child, co)
    _assert_same_state(parent, po)
    # different decisions on the two engines; the child posts past its tail limit (reactor rebuild -> own copy)
    for step in range(10):
        for (d, o, var) in ((parent, po, step), (child, co, 39 - step)):
            if d is parent and step >= 4:
                continue
            lo, hi = d.domains()
            dec = (var, int((int(lo[var]) + int(hi[var])) / 2), step & 1)
            parallel.post_decision(d, dec)
            parallel.post_decision(o, dec)
            st = d.consistency()[0]
            assert st == o.consistency()[0], step
            if st == -1:
                return
            _assert_same_state(d, o)
    # the parent goes back to a label older than the shared static part
    parent.restore(early[0])
    po.restore(early[1])
    assert parent.num_props == po.num_props == 10
    for e in (parent, po):
        e.prop_alloc(models.X_LESS_Y, [[0, 0], [1, 0]])
    assert parent.consistency()[0] == po.consistency()[0]
    _assert_same_state(parent, po)
    assert child.consistency()[0] == co.consistency()[0]
    _assert_same_state(child, co)
    # a search on a fork (device-resident: the tail of a shared model is the engine's own array)
    a, b = _engine(), _oracle(2)
    models.nqueens(30).load_into(a)
    models.nqueens(30).load_into(b)
    a.consistency()
    f = a.fork()
    a.close()                                            # the shared block outlives the parent
    rf, tf = f.search(node_limit=300, all_solutions=True, trace=300, trace_domains=True)
    rb, tb = b.search(node_limit=300, all_solutions=True, trace=300, trace_domains=True)
    assert rf.num_nodes == rb.num_nodes and (tf["status"] == tb["status"]).all() and (tf["hash"] == tb["hash"]).all()


@pytest.mark.parametrize("host_search,fused,kw", [(False, True, {}), (False, False, {}), (True, True, {}),
                                                  (False, True, {"interval_set": True}), (False, True, {"incremental": True}),
                                                  (False, True, {"flavour": "distinct"})],
                         ids=["device-search-one-launch", "device-search-threads", "host-pipelined", "device-search-set",
                              "device-search-incremental", "device-search-nary"])
def test_search_step_many_matches_lone_searches(host_search, fused, kw, monkeypatch):
    """pcp_search_step_many over 4 subtree contexts -- device-resident searches sharing one launch per
    slice (pcp_burst_batch_kernel: a group of CTAs per search), the same on one host thread and launch
    each (PCP_BATCH=unfused), host-driven searches pipelined over host threads: every context counts
    the nodes, failures and solutions of a lone search of its subtree, and together they hold every
    solution of n-queens N=9 (352, all_solution.rs:67-74)."""
    from pcp_b200 import Engine, parallel
    if not fused:
        monkeypatch.setenv("PCP_BATCH", "unfused")
    kw = dict(kw)
    m = models.nqueens(9, kw.pop("flavour", "example"))   # "distinct": n-ary Distinct, the `full` kernel variant
    ctx = parallel.SubtreeContexts(lambda: Engine(host_search=host_search, **kw), m, 4)
    ctx.open(all_solutions=True)
    res = None
    for _ in range(100000):
        res = ctx.step(37)
        if all(r.status != 0 for r in res):
            break
    assert sum(r.num_solution for r in res) == 352
    for path, r in zip(ctx.paths, res):
        from oracle.oracle_api import SET, TUNED
        lone = _oracle(SET + TUNED if kw.get("interval_set") else 2)
        m.load_into(lone)
        lone.consistency()
        parallel.enter_subtree(lone, lone.label(), path)
        ro, _ = lone.search(all_solutions=True)
        assert (r.num_nodes, r.num_solution, r.num_failed_node) == (ro.num_nodes, ro.num_solution, ro.num_failed_node)
    ctx.close()


def test_x_eq_y_mul_z_products_store():
    """XEqYMulZ (cmp/x_eq_y_mul_z.rs:68-116) at scale: p_i = a_i * b_i with the factors tied by
    chains of XLessY, signs mixed; fixpoint, status and `active` set against the oracle (a
    consistent and an inconsistent instance), then 60 search nodes."""
    rng = np.random.default_rng(77)
    n = 400
    lo = np.concatenate([rng.integers(-12, 4, 2 * n), np.full(n, -200)]).astype(np.int32)
    hi = np.concatenate([lo[:2 * n] + rng.integers(0, 14, 2 * n), np.full(n, 200)]).astype(np.int32)
    m = models.Model("products", lo, hi)
    ops = np.zeros((n, 3, 2), np.int32)
    ops[:, 0, 0] = 2 * n + np.arange(n)   # x = p_i
    ops[:, 1, 0] = np.arange(n)           # y = a_i
    ops[:, 2, 0] = n + np.arange(n)       # z = b_i
    m.add(models.X_EQ_Y_MUL_Z, ops)
    chain = np.zeros((n - 1, 2, 2), np.int32)
    chain[:, 0, 0] = 2 * n + np.arange(n - 1)
    chain[:, 1, 0] = 2 * n + np.arange(1, n)
    for slack, expect in ((150, 0), (40, -1)):   # p_i < p_{i+1} + slack: consistent / inconsistent
        mm = models.Model(m.name, m.lo, m.hi, list(m.batches))
        c = chain.copy()
        c[:, 1, 1] = slack
        mm.add(models.X_LESS_Y, c)
        dev, ora = _engine(), _oracle()
        mm.load_into(dev)
        mm.load_into(ora)
        assert dev.consistency()[0] == ora.consistency()[0] == expect
        if expect != -1:
            _assert_same_state(dev, ora)
        _compare_search(mm, 60)


def test_all_equal_classes_with_distinct_representatives():
    """AllEqual (all_equal.rs:47-103) at scale: 40 classes of 25 variables each, offsets on the
    operands, Distinct over one representative per class; root fixpoint, then 200 search nodes,
    per-node bit-exact against the oracle."""
    rng = np.random.default_rng(5)
    classes, size = 40, 25
    V = classes * size
    lo = rng.integers(0, 10, V).astype(np.int32)
    hi = (lo + rng.integers(30, 60, V)).astype(np.int32)
    m = models.Model("all-equal-classes", lo, hi)
    for c in range(classes):
        ops = np.stack([c * size + np.arange(size), rng.integers(-3, 4, size)], axis=1).astype(np.int32)
        m.add(models.ALL_EQUAL, ops)
    reps = np.stack([np.arange(classes) * size, np.zeros(classes, np.int64)], axis=1).astype(np.int32)
    m.add(models.DISTINCT, reps)
    dev, ora = _engine(), _oracle()
    m.load_into(dev)
    m.load_into(ora)
    assert dev.consistency()[0] == ora.consistency()[0] == 0
    _assert_same_state(dev, ora)
    _compare_search(m, 200)


def test_compact_and_wide_descriptor_streams(monkeypatch):
    """An all-XNeqY store over plain variables with 16-bit ids and offsets (n-queens, pairwise
    distinct) is swept from the compact 8-byte copy of its descriptors; PCP_NO_COMPACT keeps the
    16-byte stream (what every other store uses).  Both sweeps are checked node by node at sizes
    the oracle handles, and against each other."""
    for env in (None, "1"):
        if env:
            monkeypatch.setenv("PCP_NO_COMPACT", env)
        _compare_search(models.nqueens(200), 150)
        _compare_search(models.nqueens(64, "distinct"), 200)


def test_result_paths_agree():
    """pcp_consistency returns status + domains three ways -- zero-copy into the mapped mirror
    (default), copy + synchronise (timing enabled), header only + explicit copy (more than 8192
    variables) -- and pcp_stream hands out the stream the launches run on."""
    import torch
    for n in (300, 9000):   # 9000 variables: the domains are not part of the zero-copy result
        m = models.chained_lt(n, 0, n + 3)
        a, b = _engine(), _engine(timing=True)
        m.load_into(a)
        m.load_into(b)
        sa, _ = a.consistency()
        sb, stb = b.consistency()
        assert sa == sb == 0 and stb.kernel_ms > 0
        la, ha = a.domains()
        lb, hb = b.domains()
        assert (la == lb).all() and (ha == hb).all()
        assert (la == np.arange(n)).all() and (ha == np.arange(n) + 4).all()
        # work queued on the engine's stream is ordered before the next fixpoint
        s = torch.cuda.ExternalStream(a.cuda_stream(), device=torch.device("cuda", 0))
        with torch.cuda.stream(s):
            torch.zeros(1 << 20, device="cuda").add_(1)
        a.prop_alloc(models.X_LESS_Y, [[0, 0], [-1, 1]])      # x0 < 1
        assert a.consistency()[0] == 0
        la2, ha2 = a.domains()
        assert int(ha2[0]) == 0 and (la2 == la).all()
