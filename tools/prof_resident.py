"""Where the resident drop-in pattern (taylor_multi + criteria) spends its time on C2 (host wall clock)."""
import sys
import time
import numpy as np
sys.path.insert(0, '.')
import gpdoemd_b200 as g
from gpdoemd_b200 import synthetic
cfg = synthetic.CONFIGS['C2']
ctx = g.get_context(0)
models, ds = synthetic.build_models(g.GPModel, 2, cfg['M'], cfg['E'], cfg['D'], cfg['P'], cfg['n'], cfg['kernel'])
lo = np.max([d['xlo'] for d in ds], axis=0); hi = np.min([d['xhi'] for d in ds], axis=0)
N = cfg['N']
X = ctx.pinned_empty((N, cfg['D']))
X[:] = synthetic.candidates(7, N, cfg['D'], lo, hi)
T = {}
def tick(name, t0):
    T[name] = T.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
for rep in range(6):
    if rep == 1:
        T.clear()
    t0 = time.perf_counter(); mu, S = g.taylor_multi(models, X, order=1); tick('taylor_multi', t0)
    for c in ('HR', 'BH'):
        t0 = time.perf_counter(); dc = getattr(g, c)(mu, S, 0.01, None); tick('criterion ' + c, t0)
        t0 = time.perf_counter(); j = int(np.argmax(dc)); tick('np.argmax', t0)
tot = sum(T.values()) / 5
for k, v in T.items():
    print('%-28s %8.3f ms per step' % (k, v / 5))
print('%-28s %8.3f ms per step -> %.3g designs/s' % ('total', tot, N / (tot * 1e-3)))
ctx.profile_enable(True)
mu, S = g.taylor_multi(models, X, order=1)
print({k: round(v['ms'], 3) for k, v in ctx.profile_read().items() if v['launches']})
