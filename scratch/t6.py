import sys; sys.path.insert(0,'.')
import numpy as np
from oracle.oracle_api import OracleEngine, TUNED
from pcp_b200 import models, Engine
n=8
m = models.Model("sum-capacity", np.zeros(n, np.int32), np.full(n, 3, np.int32))
m.sums.append(np.array([[i, 0] for i in range(n)], np.int32))
m.sums.append(np.array([[i, 0] for i in range(0, n, 2)], np.int32))
m.add(models.X_LESS_Y, [[-2, 0], [-1, 14]])
m.add(models.X_LESS_Y, [[-1, 5], [-2, 0]])
m.add(models.X_LESS_Y_PLUS_Z, [[-3, 0], [1, 0], [3, 2]])
ops = np.zeros((n - 1, 2, 2), np.int32); ops[:, 0, 0] = np.arange(n - 1); ops[:, 1, 0] = np.arange(1, n); ops[:, 1, 1] = 1
m.add(models.X_LESS_Y, ops)
d=Engine(); o=OracleEngine(TUNED); m.load_into(d); m.load_into(o)
rd,td=d.search(node_limit=400, all_solutions=True, trace=400, trace_domains=True)
ro,to=o.search(node_limit=400, all_solutions=True, trace=400, trace_domains=True)
print(rd.num_nodes, ro.num_nodes)
for i in range(min(len(td['status']), len(to['status']))):
    same = td['status'][i]==to['status'][i] and (td['status'][i]==-1 or ((td['lo'][i]==to['lo'][i]).all() and (td['hi'][i]==to['hi'][i]).all()))
    if not same:
        print('first diff at node', i, 'dev', td['status'][i], 'ora', to['status'][i])
        print(' dev', list(zip(td['lo'][i], td['hi'][i])))
        print(' ora', list(zip(to['lo'][i], to['hi'][i])))
        if i>0: print(' prev', td['status'][i-1], list(zip(to['lo'][i-1], to['hi'][i-1])))
        break
