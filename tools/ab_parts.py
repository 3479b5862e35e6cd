"""A/B of the row-block parts of the first-order contraction (option tri_parts): results against whole items, step time."""
import sys
import numpy as np
sys.path.insert(0, '.')
import gpdoemd_b200 as g
from gpdoemd_b200 import synthetic
ctx = g.get_context(0)
ctx.set_scratch_bytes(64 << 30)
for name, N in (('C5', 370 * 128), ('C4', 148 * 128), ('C2', 100000)):
    cfg = dict(synthetic.CONFIGS[name])
    gms, ds = synthetic.build_models(g.GPModel, 2, cfg['M'], cfg['E'], cfg['D'], cfg['P'], cfg['n'], cfg['kernel'], build_on_device=cfg['n'] >= 2000)
    lo = np.max([d['xlo'] for d in ds], axis=0); hi = np.min([d['xhi'] for d in ds], axis=0)
    D = cfg['D']
    X_dev = ctx.dev_alloc(N * D * 8)
    ctx.fill_uniform(X_dev, N, D, seed=20252, lo=lo, hi=hi)
    dc_dev = ctx.dev_alloc(len(cfg['criteria']) * N * 8)
    ref = None
    for parts in ((0, 9, 4, 8, 0, 9) if name != 'C2' else (0, 9, 2, 4)):
        ctx.set_option('tri_parts', parts)
        for _ in range(2):
            r = g.score_designs_multi(gms, None, list(cfg['criteria']), order=1, noise_var=0.01, x_dev=X_dev, N=N, dc_dev=dc_dev)
        dc = ctx.dev_download(dc_dev, (len(cfg['criteria']), N))
        if ref is None:
            ref = dc
        ctx.profile_enable(True)
        tot = 0.0
        for _ in range(3):
            ctx.flush_l2(); ctx.timer_start()
            r = g.score_designs_multi(gms, None, list(cfg['criteria']), order=1, noise_var=0.01, return_dc=False, x_dev=X_dev, N=N)
            tot += ctx.timer_stop()
        prof = ctx.profile_read(); ctx.profile_enable(False)
        print('%s tri_parts=%2d: %.3f ms per step, tri %.3f, max rel diff vs whole items %.2e, argmax %s' % (
            name, parts, tot / 3, prof['tri']['ms'] / 3, float(np.max(np.abs(dc - ref)) / np.abs(ref).mean()), [r[c][0] for c in cfg['criteria']]), flush=True)
    ctx.set_option('tri_parts', 0)
    for m in gms:
        m._invalidate_device()
    ctx.dev_free(X_dev); ctx.dev_free(dc_dev)
