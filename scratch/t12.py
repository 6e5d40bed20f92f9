"""Per-node device time by iteration count, flushed (engine-stream flush) and warm."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from pcp_b200 import Engine, models, parallel
m = models.nqueens(1000)
for mode in ('flush', 'warm'):
    e = Engine(timing=True); m.load_into(e)
    buf = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')
    ext = torch.cuda.ExternalStream(e.cuda_stream(), device=torch.device('cuda', 0))
    torch.cuda.synchronize()
    stack = []; started = False; rec = []
    for n in range(210):
        if started:
            label, d = stack.pop(); e.restore(label); parallel.post_decision(e, d)
        started = True
        if mode == 'flush':
            with torch.cuda.stream(ext): buf.add_(1)
        st, stats = e.consistency()
        if n >= 10: rec.append((stats.iterations, stats.kernel_ms * 1e3, stats.propagations))
        if st == 0:
            lo, hi = e.domains(); var, val = parallel.select_branch(lo, hi); label = e.label()
            stack.append((label, (var, val, 1))); stack.append((label, (var, val, 0)))
    r = np.array(rec)
    print(mode, 'mean us', r[:, 1].mean().round(2), 'median', np.median(r[:, 1]).round(2))
    for it in sorted(set(r[:, 0].astype(int))):
        s = r[r[:, 0] == it]
        print('   iters', it, 'count', len(s), 'mean us', s[:, 1].mean().round(2), 'min', s[:, 1].min().round(2), 'max', s[:, 1].max().round(2))
    e.close()
