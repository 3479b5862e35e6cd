"""Host-side cost of one fused scoring call (Python prep + C planning + tiny GPU work)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import gpdoemd_b200 as g
from gpdoemd_b200 import synthetic, scoring

cfg = synthetic.CONFIGS['C2']
models, ds = synthetic.build_models(g.GPModel, 20250, cfg['M'], cfg['E'], cfg['D'], cfg['P'], cfg['n'], cfg['kernel'])
lo = np.max([d['xlo'] for d in ds], axis=0); hi = np.min([d['xhi'] for d in ds], axis=0)
ctx = g.get_context(0)
for N in (128, 100000):
    X_dev = ctx.dev_alloc(N * 3 * 8)
    ctx.fill_uniform(X_dev, N, 3, seed=1, lo=lo, hi=hi)
    f = lambda: g.score_designs_multi(models, None, ['HR', 'BH'], order=1, noise_var=0.01, return_dc=False, x_dev=X_dev, N=N)
    for _ in range(5): f()
    t0 = time.perf_counter()
    for _ in range(50): f()
    dt = (time.perf_counter() - t0) / 50
    t0 = time.perf_counter()
    for _ in range(50): scoring._prep(models, 'HR', 0.01, None)
    dp = (time.perf_counter() - t0) / 50
    ctx.profile_enable(True)
    for _ in range(10): f()
    prof = ctx.profile_read(); ctx.profile_enable(False)
    gpu = sum(v['ms'] for v in prof.values()) / 10
    print('N=%d: wall per call %.3f ms, python _prep %.3f ms, kernel time %.3f ms' % (N, dt * 1e3, dp * 1e3, gpu))
