import sys, time
sys.path.insert(0, '.')
from pcp_b200 import Engine, models
m = models.random_arith_csp()
e = Engine(timing=True); m.load_into(e)
root = e.label()
for i in range(4):
    e.restore(root)
    st, stats = e.consistency()
    print(i, st, 'kernel ms', round(stats.kernel_ms, 4), 'iters', stats.iterations, 'props', stats.propagations, file=sys.stderr)
