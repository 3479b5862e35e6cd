"""Sweep of the tail-overlap head size (options head_periods / head_grid) on C2."""
import sys
import numpy as np
sys.path.insert(0, '.')
import gpdoemd_b200 as g
from gpdoemd_b200 import synthetic
ctx = g.get_context(0)
cfg = dict(synthetic.CONFIGS['C2'])
gms, ds = synthetic.build_models(g.GPModel, 2, cfg['M'], cfg['E'], cfg['D'], cfg['P'], cfg['n'], cfg['kernel'])
lo = np.max([d['xlo'] for d in ds], axis=0); hi = np.min([d['xhi'] for d in ds], axis=0)
N, D = cfg['N'], cfg['D']
X_dev = ctx.dev_alloc(N * D * 8)
ctx.fill_uniform(X_dev, N, D, seed=20252, lo=lo, hi=hi)
ref = None
for hp, hg in [(0, 0), (1, 32), (1, 48), (1, 64), (2, 64), (2, 96), (3, 96), (3, 128), (1, 0), (0, 0)]:
    ctx.set_option('head_periods', hp); ctx.set_option('head_grid', hg)
    for _ in range(3):
        r = g.score_designs_multi(gms, None, ['HR', 'BH'], order=1, noise_var=0.01, return_dc=False, x_dev=X_dev, N=N)
    ctx.profile_enable(True)
    tot = 0.0
    for _ in range(10):
        ctx.flush_l2(); ctx.timer_start()
        r = g.score_designs_multi(gms, None, ['HR', 'BH'], order=1, noise_var=0.01, return_dc=False, x_dev=X_dev, N=N)
        tot += ctx.timer_stop()
    prof = ctx.profile_read(); ctx.profile_enable(False)
    if ref is None:
        ref = r
    assert r['BH'][0] == ref['BH'][0] and r['BH'][1] == ref['BH'][1]
    print('head_periods=%d head_grid=%3d: %.3f ms per step  %s' % (hp, hg, tot / 10, {k: round(v['ms'] / 10, 3) for k, v in prof.items() if v['launches']}), flush=True)
