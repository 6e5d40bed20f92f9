import os, sys
os.environ["PCP_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pcp_b200 import models, Engine
m = models.nqueens(1000)
e = Engine(host_search=True)
m.load_into(e)
e.set_grid_limit(int(sys.argv[1]) if len(sys.argv) > 1 else 6)
h = e.search_open(all_solutions=True)
for i in range(5):
    print("== node", i, flush=True)
    h.step(1)
