import sys
sys.path.insert(0, '.')
from pcp_b200 import Engine, models
m = models.nqueens(1000)
e = Engine(timing=False, host_search=True, max_labels=1 << 16); m.load_into(e)
r, _ = e.search(node_limit=410, all_solutions=True, warmup_nodes=10)
print('host_search us/node', round(1e6 * r.seconds / 400, 2))
e.close()
