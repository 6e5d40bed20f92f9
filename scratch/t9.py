"""Quick timing probe: C2 per-node + burst, C5 per-node, C4 single fixpoint."""
import sys, os, time
sys.path.insert(0, '.')
import numpy as np
from pcp_b200 import Engine, models, parallel
which = sys.argv[1:] or ['c2', 'c5', 'c4']
def per_node(m, nodes, tag):
    e = Engine(timing=True); m.load_into(e)
    stack=[]; started=False; kms=[]; its=0
    for n in range(nodes):
        if started:
            if not stack: break
            label,d = stack.pop(); e.restore(label); parallel.post_decision(e,d)
        started=True
        st,stats=e.consistency(); kms.append(stats.kernel_ms); its+=stats.iterations
        if st==0:
            lo,hi=e.domains(); var,val=parallel.select_branch(lo,hi); label=e.label()
            stack.append((label,(var,val,1))); stack.append((label,(var,val,0)))
    k=np.array(kms[3:])
    print(f'{tag} per-node: median {np.median(k)*1e3:.1f} us  mean {k.mean()*1e3:.1f} us  min {k.min()*1e3:.1f} us  iters/node {its/len(kms):.2f}')
    e.close()
if 'c2' in which:
    m = models.nqueens(1000)
    per_node(m, 60, 'C2')
    for nodes in (210, 2010):
        e = Engine(timing=True); m.load_into(e)
        r,_ = e.search(node_limit=nodes, all_solutions=True, warmup_nodes=10)
        k = nodes-10
        print('C2 burst', nodes, 'us/node', round(1e6*r.seconds/k,2), 'props/s %.3g' % (r.propagations/r.seconds), 'iters/node', round(r.iterations/k,3))
        e.close()
if 'c5' in which:
    per_node(models.nqueens(5000), 16, 'C5')
if 'c4' in which:
    m = models.random_arith_csp()
    e = Engine(timing=True); m.load_into(e)
    root = e.label(); ks=[]
    for i in range(8):
        e.restore(root); st,stats = e.consistency(); ks.append(stats.kernel_ms)
    print('C4 fixpoint ms', [round(x,3) for x in ks], 'iters', stats.iterations, 'props', stats.propagations)
    e.close()
