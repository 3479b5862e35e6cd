"""Where the reference's own call pattern (demos/case_study.py:284-293) spends its time through the drop-in functions:
taylor_first_order per model (device pass, D2H), the caller's strided assignments, one call per criterion (H2D of
mu / S, kernels, D2H).  Host wall clock, C2 shapes."""
import sys
import time

import numpy as np

sys.path.insert(0, '.')
import gpdoemd_b200 as g
from gpdoemd_b200 import synthetic

cfg = synthetic.CONFIGS['C2']
ctx = g.get_context(0)
models, ds = synthetic.build_models(g.GPModel, 2, cfg['M'], cfg['E'], cfg['D'], cfg['P'], cfg['n'], cfg['kernel'])
lo = np.max([d['xlo'] for d in ds], axis=0)
hi = np.min([d['xhi'] for d in ds], axis=0)
N, M, E = cfg['N'], cfg['M'], cfg['E']
X = synthetic.candidates(7, N, cfg['D'], lo, hi)
T = {}


def tick(name, t0):
    ctx.sync()
    T[name] = T.get(name, 0.0) + (time.perf_counter() - t0) * 1e3


for rep in range(4):
    if rep == 1:
        T.clear()
    mu = np.empty((N, M, E))
    S = np.empty((N, M, E, E))
    for i, mo in enumerate(models):
        t0 = time.perf_counter()
        m_i, S_i = g.taylor_first_order(mo, X)
        tick('taylor_first_order', t0)
        t0 = time.perf_counter()
        mu[:, i], S[:, i] = m_i, S_i
        tick('caller assignment mu[:, i], S[:, i] = ...', t0)
    for c in ('HR', 'BH'):
        t0 = time.perf_counter()
        dc = getattr(g, c)(mu, S, 0.01, None)
        tick('criterion ' + c, t0)
    t0 = time.perf_counter()
    j = int(np.argmax(dc))
    tick('np.argmax', t0)
tot = sum(T.values()) / 3
for k, v in T.items():
    print('%-45s %8.3f ms per step' % (k, v / 3))
print('%-45s %8.3f ms per step -> %.3g designs/s' % ('total', tot, N / (tot * 1e-3)))
t0 = time.perf_counter()
for _ in range(3):
    r = g.score_designs_multi(models, X, ['HR', 'BH'], noise_var=0.01)
ctx.sync()
print('fused score_designs_multi (pageable X, dc out)    %8.3f ms per step' % ((time.perf_counter() - t0) * 1e3 / 3))
