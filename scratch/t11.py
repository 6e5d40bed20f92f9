"""Iteration structure of deeper C2 nodes (PCP_TRACE=1 PCP_NO_BURST=1): which nodes need many iterations and why."""
import sys
sys.path.insert(0, '.')
from pcp_b200 import Engine, models, parallel
N = int(sys.argv[1]) if len(sys.argv) > 1 else 420
m = models.nqueens(1000)
e = Engine(timing=True); m.load_into(e)
stack = []; started = False
for n in range(N):
    if started:
        label, d = stack.pop(); e.restore(label); parallel.post_decision(e, d)
    started = True
    st, stats = e.consistency()
    print('NODE', n, 'st', st, 'kernel_us', round(stats.kernel_ms * 1e3, 1), 'iters', stats.iterations, 'props', stats.propagations, file=sys.stderr)
    if st == 0:
        lo, hi = e.domains(); var, val = parallel.select_branch(lo, hi); label = e.label()
        stack.append((label, (var, val, 1))); stack.append((label, (var, val, 0)))
