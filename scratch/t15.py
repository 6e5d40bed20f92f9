"""C3 (all-interval N=500) per-node phase trace."""
import sys
sys.path.insert(0, '.')
from pcp_b200 import Engine, models, parallel
m = models.all_interval(500)
e = Engine(timing=True); m.load_into(e)
stack = []; started = False
for n in range(30):
    if started:
        label, d = stack.pop(); e.restore(label); parallel.post_decision(e, d)
    started = True
    st, stats = e.consistency()
    print('NODE', n, 'st', st, 'kernel_us', round(stats.kernel_ms * 1e3, 1), 'iters', stats.iterations, 'props', stats.propagations, file=sys.stderr)
    if st == 0:
        lo, hi = e.domains(); var, val = parallel.select_branch(lo, hi); label = e.label()
        stack.append((label, (var, val, 1))); stack.append((label, (var, val, 0)))
