"""A/B of the fused evaluation + contraction kernel against the separate kernels: bit-level comparison of the criterion
vectors on ragged shapes, then step times on C2 (and optionally another config)."""
import sys
import time

import numpy as np

sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import gpdoemd_b200 as g
from gpdoemd_b200 import synthetic

ctx = g.get_context(0)


def models_for(M, E, D, P, n, kernel, nb=0, seed=5):
    gms, ds = synthetic.build_models(g.GPModel, seed, M, E, D, P, n, kernel, n_binary=nb, build_on_device=n >= 2000)
    lo = np.max([d['xlo'] for d in ds], axis=0)
    hi = np.min([d['xhi'] for d in ds], axis=0)
    return gms, lo, hi


worst = 0.0
for (M, E, D, P, n, kernel, nb, N) in [(2, 2, 3, 4, 500, 'RBF', 0, 128 * 163 + 17), (3, 1, 2, 2, 130, 'Matern52', 0, 700),
                                       (2, 2, 4, 3, 300, 'Matern32', 1, 1000), (2, 3, 1, 1, 37, 'Exponential', 0, 5),
                                       (2, 2, 3, 10, 260, 'RatQuad', 0, 400), (2, 1, 5, 6, 1000, 'RBF', 0, 3000),
                                       (2, 1, 3, 4, 2100, 'Matern52', 0, 128 * 150)]:
    gms, lo, hi = models_for(M, E, D, P, n, kernel, nb)
    X = synthetic.candidates(3, N, D, lo, hi, n_binary=nb)
    out = {}
    for flag in (0, 1):
        ctx.set_option('fused', flag)
        out[flag] = g.score_designs_multi(gms, X, ['HR', 'BH', 'JR'], order=1, noise_var=0.02)
        mu, s2 = gms[0].predict(X)
        out[flag]['pred'] = (mu, s2)
    for c in ('HR', 'BH', 'JR'):
        a, b = out[0][c][2], out[1][c][2]
        rel = float(np.max(np.abs(a - b)) / max(np.abs(a).mean(), 1e-300))
        worst = max(worst, rel)
        assert out[0][c][0] == out[1][c][0], (kernel, c, 'argmax differs')
    dm = float(np.max(np.abs(out[0]['pred'][0] - out[1]['pred'][0])))
    dv = float(np.max(np.abs(out[0]['pred'][1] - out[1]['pred'][1]) / np.abs(out[0]['pred'][1]).max()))
    print('%-11s M=%d E=%d D=%d P=%2d n=%4d nb=%d N=%6d: criteria max rel diff fused vs separate %.2e, mean %.1e, var %.1e'
          % (kernel, M, E, D, P, n, nb, N, rel, dm, dv), flush=True)
    for m in gms:
        m._invalidate_device()
print('worst relative difference', worst)

for name in sys.argv[1:] or ['C2']:
    cfg = dict(synthetic.CONFIGS[name])
    if name != 'C2':
        cfg['N'] = {'C4': 148 * 128, 'C5': 370 * 128}[name]
        ctx.set_scratch_bytes(64 << 30)
    gms, lo, hi = models_for(cfg['M'], cfg['E'], cfg['D'], cfg['P'], cfg['n'], cfg['kernel'], seed=2)
    N, D = cfg['N'], cfg['D']
    X_dev = ctx.dev_alloc(N * D * 8)
    ctx.fill_uniform(X_dev, N, D, seed=20252, lo=lo, hi=hi)
    for flag in (0, 1, 0, 1):
        ctx.set_option('fused', flag)
        for _ in range(3):
            r = g.score_designs_multi(gms, None, list(cfg['criteria']), order=1, noise_var=0.01, return_dc=False, x_dev=X_dev, N=N)
        ctx.profile_enable(True)
        tot = 0.0
        for _ in range(10):
            ctx.flush_l2()
            ctx.timer_start()
            r = g.score_designs_multi(gms, None, list(cfg['criteria']), order=1, noise_var=0.01, return_dc=False, x_dev=X_dev, N=N)
            tot += ctx.timer_stop()
        prof = ctx.profile_read()
        ctx.profile_enable(False)
        print('%s fused=%d: %.3f ms per step, families %s, argmax %s' % (
            name, flag, tot / 10, {k: round(v['ms'] / 10, 3) for k, v in prof.items() if v['launches']},
            {c: r[c][0] for c in cfg['criteria']}), flush=True)
    for m in gms:
        m._invalidate_device()
