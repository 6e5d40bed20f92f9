import sys, time
sys.path.insert(0, '.')
from pcp_b200 import Engine, models
m = models.nqueens(1000)
for timing in (False, True):
    e = Engine(timing=timing); m.load_into(e)
    r,_ = e.search(node_limit=210, all_solutions=True, warmup_nodes=10)
    print('timing', timing, 'nodes', r.num_nodes, 'sec', r.seconds, 'us/node', 1e6*r.seconds/200, 'kernel us/node', 1e6*r.kernel_seconds/200, 'iters', r.iterations/200)
    e.close()
