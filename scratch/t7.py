import sys, time
sys.path.insert(0, '.')
import numpy as np
from oracle.oracle_api import OracleEngine, FLAT
from pcp_b200 import Engine, models
for n, limit in ((8, 0), (30, 300), (1000, 60)):
    m = models.nqueens(n)
    d = Engine(); o = OracleEngine(FLAT); m.load_into(d); m.load_into(o)
    rd, td = d.search(node_limit=limit, all_solutions=True, trace=limit or 5000, trace_domains=True)
    ro, to = o.search(node_limit=limit, all_solutions=True, trace=limit or 5000, trace_domains=True)
    ok = td['status'] != -1
    print(n, 'nodes', rd.num_nodes, ro.num_nodes, 'sol', rd.num_solution, ro.num_solution, 'status', rd.status, ro.status,
          'trace equal', bool((td['status']==to['status']).all() and (td['hash']==to['hash']).all() and (td['lo'][ok]==to['lo'][ok]).all()),
          'us/node', 1e6*rd.seconds/max(rd.num_nodes,1), 'kernel us/node', 1e6*rd.kernel_seconds/max(rd.num_nodes,1))
    dl, dh = d.domains(); ol, oh = o.domains()
    print('  final state equal', bool((dl==ol).all() and (dh==oh).all()), 'props', d.num_props, o.num_props, 'active eq', bool((d.active()==o.active()).all()))
m = models.nqueens(1000)
d = Engine(); m.load_into(d)
r,_ = d.search(node_limit=2010, all_solutions=True, warmup_nodes=10)
print('burst 2000 nodes: us/node', 1e6*r.seconds/2000, 'kernel us/node', 1e6*r.kernel_seconds/2000, 'props/s', r.propagations/r.seconds, 'iters/node', r.iterations/2000)
