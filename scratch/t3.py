import sys, os
sys.path.insert(0, '.')
from pcp_b200 import Engine, models, parallel
e = Engine(timing=True); m = models.nqueens(int(sys.argv[1]) if len(sys.argv)>1 else 1000); m.load_into(e)
stack=[]; started=False
for n in range(14):
    if started:
        label,d = stack.pop(); e.restore(label); parallel.post_decision(e,d)
    started=True
    st,stats=e.consistency()
    print(n, st, 'kernel ms', round(stats.kernel_ms,4), 'iters', stats.iterations, file=sys.stderr)
    if st==0:
        lo,hi=e.domains(); var,val=parallel.select_branch(lo,hi); label=e.label()
        stack.append((label,(var,val,1))); stack.append((label,(var,val,0)))
