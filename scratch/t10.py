"""Phase traces: C4 single fixpoint, C2 nodes with an L2 flush before each (PCP_TRACE=1)."""
import sys
sys.path.insert(0, '.')
import torch
from pcp_b200 import Engine, models, parallel
which = sys.argv[1:] or ['c4', 'c2']
if 'c4' in which:
    m = models.random_arith_csp()
    e = Engine(timing=True); m.load_into(e)
    root = e.label()
    for i in range(3):
        e.restore(root); st, stats = e.consistency()
        print('C4 run', i, 'kernel ms', round(stats.kernel_ms, 4), 'iters', stats.iterations, 'props', stats.propagations, file=sys.stderr)
    e.close()
if 'c2' in which:
    m = models.nqueens(1000)
    e = Engine(timing=True); m.load_into(e)
    buf = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')
    stack = []; started = False
    for n in range(14):
        if started:
            label, d = stack.pop(); e.restore(label); parallel.post_decision(e, d)
        started = True
        if n >= 8:
            buf.add_(1); torch.cuda.synchronize()
        st, stats = e.consistency()
        print('C2 node', n, 'flushed' if n >= 8 else 'warm', 'kernel ms', round(stats.kernel_ms, 4), 'iters', stats.iterations, file=sys.stderr)
        if st == 0:
            lo, hi = e.domains(); var, val = parallel.select_branch(lo, hi); label = e.label()
            stack.append((label, (var, val, 1))); stack.append((label, (var, val, 0)))
