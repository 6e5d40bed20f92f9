"""Host-driven node loop (PCP_FLAG_HOST_SEARCH) vs device search: wall-clock us/node."""
import sys
sys.path.insert(0, '.')
from pcp_b200 import Engine, models
m = models.nqueens(1000)
for hs in (True, False):
    for timing in (False, True):
        e = Engine(timing=timing, host_search=hs, max_labels=1 << 16); m.load_into(e)
        r, _ = e.search(node_limit=410, all_solutions=True, warmup_nodes=10)
        print('host_search' if hs else 'device_search', 'timing' if timing else 'no-timing', 'us/node', round(1e6 * r.seconds / 400, 2), 'iters/node', round(r.iterations / 400, 3))
        e.close()
