"""The N>1 path on CPU: world_size-2 `gloo` process group, the oracle standing in for the
device engine.  Checks the host-side logic of pcp_b200.parallel: identical frontier on every
rank, disjoint round-robin slices, matching collective counts, stop-flag propagation."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, budget, stop_on_solution, q, bb_mode=0, bb_var=0):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle_api import FLAT, OracleEngine
    from pcp_b200 import models, parallel
    e = OracleEngine(FLAT)
    models.nqueens(n).load_into(e)
    paths = parallel.expand_frontier(e, parts=world * 4)
    flag = parallel.StopFlag()
    out = parallel.sharded_search(e, rank, world, node_budget=budget, sync_every=16, stop_flag=flag,
                                  stop_on_solution=stop_on_solution, bb_mode=bb_mode, bb_var=bb_var)
    lo, hi = e.domains()
    q.put((rank, [tuple(map(tuple, p)) for p in paths], out, bool((lo == hi).all())))
    dist.barrier()
    dist.destroy_process_group()


def _run(world, n, budget, stop_on_solution, bb_mode=0, bb_var=0):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, budget, stop_on_solution, q, bb_mode, bb_var))
             for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_sharded_search_fixed_budget():
    res = _run(2, 10, 64, False)
    (r0, paths0, out0, _), (r1, paths1, out1, _) = res
    assert paths0 == paths1 and len(paths0) >= 8          # same frontier on every rank
    assert out0["frontier"] == out1["frontier"] == len(paths0)
    assert out0["nodes"] == out1["nodes"] == 64            # weak scaling: fixed budget per rank
    assert out0["propagations"] > 0 and out1["propagations"] > 0
    assert not out0["stopped"] and not out1["stopped"]


def test_stop_flag_reaches_every_rank():
    res = _run(2, 8, 4000, True)
    outs = [r[2] for r in res]
    assert all(o["stopped"] for o in outs)                 # the finder's flag stops both ranks
    assert sum(o["solutions"] for o in outs) >= 1
    assert any(r[3] for r in res)                          # the finder holds a full assignment


@pytest.mark.parametrize("bb_mode", [1, 2])
def test_incumbent_exchange_branch_and_bound(bb_mode):
    """BranchAndBound across ranks (search/branch_and_bound.rs:76-92 + SURVEY 8e): every rank
    all-reduces the incumbent (min / max) after each round and adopts it, so both ranks end
    with the global optimum -- the value one process finds alone -- and the rank that did not
    find it itself still prunes with it (fewer nodes than the subtrees cost without exchange)."""
    sys.path.insert(0, ROOT)
    from oracle.oracle_api import FLAT, OracleEngine
    from pcp_b200 import models
    n, var = 9, 4
    solo = OracleEngine(FLAT)
    models.nqueens(n).load_into(solo)
    ref, _ = solo.search(all_solutions=True, bb_mode=bb_mode, bb_var=var)
    assert ref.has_bb_value
    res = _run(2, n, 10**6, False, bb_mode=bb_mode, bb_var=var)
    outs = [r[2] for r in res]
    assert [o["incumbent"] for o in outs] == [ref.bb_value, ref.bb_value]
    assert all(o["exchanges"] >= 2 for o in outs) and outs[0]["exchanges"] == outs[1]["exchanges"]
    plain = _run(2, n, 10**6, False)                      # the same subtrees searched exhaustively
    assert sum(o["nodes"] for o in outs) < sum(r[2]["nodes"] for r in plain)


def test_frontier_slices_partition_the_tree():
    """Single process: the frontier of the oracle covers the root exactly (the solution count of
    the subtrees adds up to the reference's golden count for n=6: 4)."""
    sys.path.insert(0, ROOT)
    from oracle.oracle_api import FLAT, OracleEngine
    from pcp_b200 import models, parallel
    e = OracleEngine(FLAT)
    models.nqueens(6).load_into(e)
    paths = parallel.expand_frontier(e, parts=6)
    root = e.label()
    total = 0
    for world_rank in range(3):
        for p in parallel.my_slice(paths, world_rank, 3):
            parallel.enter_subtree(e, root, p)
            res, _ = e.search(all_solutions=True)
            total += res.num_solution
    assert total == 4
