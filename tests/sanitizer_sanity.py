"""Run under compute-sanitizer (memcheck / racecheck): searches on the device against the oracle,
node by node -- profiles/r2_sanitizer.md.  Not collected by pytest."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # the repository root
import numpy as np
from pcp_b200 import Engine, models, parallel, search_step_many
from oracle.oracle_api import OracleEngine, SET, FLAT
def cmp(model, nodes, **kw):
    sets = kw.get("interval_set", False)
    d, o = Engine(**kw), OracleEngine(1 + (SET if sets else 0))
    model.load_into(d); model.load_into(o)
    rd, td = d.search(node_limit=nodes, all_solutions=True, trace=nodes, trace_domains=True)
    ro, to = o.search(node_limit=nodes, all_solutions=True, trace=nodes, trace_domains=True)
    assert rd.num_nodes == ro.num_nodes and (td["status"] == to["status"]).all() and (td["hash"] == to["hash"]).all(), model.name
    print("ok", model.name, kw, rd.num_nodes, flush=True)
cmp(models.nqueens(40), 60)
cmp(models.nqueens(40), 60, incremental=True)
cmp(models.nqueens(40), 60, host_search=True)
cmp(models.nqueens(40), 60, interval_set=True)
cmp(models.nqueens(40), 60, interval_set=True, host_search=True)
cmp(models.nqueens(40), 60, interval_set=True, incremental=True)
cmp(models.nqueens(12, "distinct"), 60, interval_set=True)
cmp(models.all_interval(10), 60)
cmp(models.all_interval(10), 60, interval_set=True)
cmp(models.cumulative([(0, 6)] * 4, [(2, 2), (3, 3), (2, 2), (1, 1)], [(2, 2), (1, 1), (2, 2), (1, 1)], (3, 3), constant=True), 60)
cmp(models.random_arith_csp(2000, 20000, seed=5), 3)
ctx = parallel.SubtreeContexts(lambda: Engine(host_search=True), models.nqueens(40), 4)
hs = ctx.open(all_solutions=True)
r = search_step_many(hs, 30)
print("contexts", [x.num_nodes for x in r]); ctx.close()
ctx = parallel.SubtreeContexts(lambda: Engine(incremental=True), models.nqueens(40), 4)
hs = ctx.open(all_solutions=True)
r = search_step_many(hs, 30)
print("contexts inc burst", [x.num_nodes for x in r]); ctx.close()
