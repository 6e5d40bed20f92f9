"""Run under compute-sanitizer (memcheck / racecheck): searches on the device against the oracle,
node by node -- profiles/r2_sanitizer.md.  Not collected by pytest."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # the repository root
import numpy as np
from pcp_b200 import Engine, models
from oracle.oracle_api import OracleEngine
def cmp(model, nodes, **kw):
    d, o = Engine(**kw), OracleEngine(1)
    model.load_into(d); model.load_into(o)
    rd, td = d.search(node_limit=nodes, all_solutions=True, trace=nodes, trace_domains=True)
    ro, to = o.search(node_limit=nodes, all_solutions=True, trace=nodes, trace_domains=True)
    n = min(len(td["status"]), len(to["status"]))
    bad = [i for i in range(n) if td["status"][i] != to["status"][i] or td["hash"][i] != to["hash"][i]]
    print(model.name, kw, "nodes", rd.num_nodes, ro.num_nodes, "first mismatches", bad[:5], flush=True)
    if bad:
        i = bad[0]
        print("  status dev/ora", td["status"][i], to["status"][i])
        if td["status"][i] != -1 and to["status"][i] != -1:
            print("  lo diff", np.nonzero(td["lo"][i] != to["lo"][i])[0], "hi diff", np.nonzero(td["hi"][i] != to["hi"][i])[0])
            print("  dev", list(zip(td["lo"][i], td["hi"][i])))
            print("  ora", list(zip(to["lo"][i], to["hi"][i])))
for kw in ({}, {"host_search": True}, {"incremental": True}):
    cmp(models.all_interval(10), 60, **kw)
    cmp(models.all_interval(10, decompose_distinct=True), 60, **kw)
cmp(models.nqueens(12, "distinct"), 60)
