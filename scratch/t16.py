import sys
sys.path.insert(0, '.')
import numpy as np
from pcp_b200 import Engine, models
from oracle.oracle_api import OracleEngine
rng = np.random.default_rng(77)
n = 400
lo = np.concatenate([rng.integers(-12, 4, 2 * n), np.full(n, -200)]).astype(np.int32)
hi = np.concatenate([lo[:2 * n] + rng.integers(0, 14, 2 * n), np.full(n, 200)]).astype(np.int32)
def build(with_chain):
    m = models.Model("products", lo, hi)
    ops = np.zeros((n, 3, 2), np.int32)
    ops[:, 0, 0] = 2 * n + np.arange(n); ops[:, 1, 0] = np.arange(n); ops[:, 2, 0] = n + np.arange(n)
    m.add(models.X_EQ_Y_MUL_Z, ops)
    if with_chain:
        chain = np.zeros((n - 1, 2, 2), np.int32)
        chain[:, 0, 0] = 2 * n + np.arange(n - 1); chain[:, 1, 0] = 2 * n + np.arange(1, n); chain[:, 1, 1] = 3
        m.add(models.X_LESS_Y, chain)
    return m
for with_chain in (False, True):
    m = build(with_chain)
    dev, ora = Engine(), OracleEngine(1)
    m.load_into(dev); m.load_into(ora)
    ds, st = dev.consistency(); os_, _ = ora.consistency()
    dlo, dhi = dev.domains(); olo, ohi = ora.domains()
    bad = np.nonzero((dlo != olo) | (dhi != ohi))[0]
    print('chain', with_chain, 'status', ds, os_, 'iters', st.iterations, 'mismatches', len(bad))
    for v in bad[:8]:
        i = v - 2 * n
        print('  var', v, 'dev', (dlo[v], dhi[v]), 'ora', (olo[v], ohi[v]), 'a', (dlo[i], dhi[i]) if i >= 0 else None, 'b', (dlo[n + i], dhi[n + i]) if i >= 0 else None)
