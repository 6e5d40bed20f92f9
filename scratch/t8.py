import sys, time, os
sys.path.insert(0, '.')
from pcp_b200 import Engine, models
m = models.nqueens(1000)
for nodes in (210, 2010):
    e = Engine(timing=True); m.load_into(e)
    r,_ = e.search(node_limit=nodes, all_solutions=True, warmup_nodes=10)
    k = nodes-10
    print('burst' if not os.environ.get('PCP_NO_BURST') else 'per-node', nodes, 'us/node', round(1e6*r.seconds/k,2), 'kernel us/node', round(1e6*r.kernel_seconds/k,2), 'props/s %.3g' % (r.propagations/r.seconds), 'iters/node', r.iterations/k)
    e.close()
