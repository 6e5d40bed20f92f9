import os, sys, time
os.environ["PCP_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pcp_b200 import models, Engine
m = models.random_arith_csp()
e = Engine()
m.load_into(e)
root = e.label()
for i in range(3):
    e.restore(root)
    st, stats = e.consistency()
    print(i, st, stats, flush=True)
