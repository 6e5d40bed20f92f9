import time, sys
sys.path.insert(0, '.')
from pcp_b200 import Engine, models, parallel
import numpy as np
e = Engine(timing=True); m = models.nqueens(1000); m.load_into(e)
T = {k: 0.0 for k in ('restore','alloc','cons','dom','label')}
stack=[]; started=False; kms=0; its=0
for n in range(60):
    if started:
        label,d = stack.pop()
        t=time.perf_counter(); e.restore(label); T['restore']+=time.perf_counter()-t
        t=time.perf_counter(); parallel.post_decision(e,d); T['alloc']+=time.perf_counter()-t
    started=True
    t=time.perf_counter(); st,stats=e.consistency(); dt=time.perf_counter()-t; T['cons']+=dt
    kms+=stats.kernel_ms; its+=stats.iterations
    if n<12 or n%10==0: print(n, st, 'cons ms', round(dt*1e3,3), 'kernel ms', round(stats.kernel_ms,4), 'iters', stats.iterations, 'props', stats.propagations)
    if st==0:
        t=time.perf_counter(); lo,hi=e.domains(); T['dom']+=time.perf_counter()-t
        var,val=parallel.select_branch(lo,hi)
        t=time.perf_counter(); label=e.label(); T['label']+=time.perf_counter()-t
        stack.append((label,(var,val,1))); stack.append((label,(var,val,0)))
print({k: round(v*1e3/60,4) for k,v in T.items()}, 'ms per node; kernel ms/node', kms/60, 'iters/node', its/60)
