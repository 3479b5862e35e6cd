"""Short driver for ncu captures: a few fused scoring steps of a BASELINE config on one GPU."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpdoemd_b200 as g
from gpdoemd_b200 import synthetic

ap = argparse.ArgumentParser()
ap.add_argument('--config', default='C2')
ap.add_argument('--candidates', type=int, default=None)
ap.add_argument('--steps', type=int, default=2)
ap.add_argument('--tri-variant', type=int, default=2)
ap.add_argument('--option', action='append', default=[], help='name=value for gpd_set_option')
a = ap.parse_args()
cfg = dict(synthetic.CONFIGS[a.config])
N = a.candidates or cfg['N']
ctx = g.get_context(0)
ctx.set_option('tri_variant', a.tri_variant)
for o in a.option:
    k, v = o.split('=')
    ctx.set_option(k, int(v))
models, ds = synthetic.build_models(g.GPModel, 2, cfg['M'], cfg['E'], cfg['D'], cfg['P'], cfg['n'], cfg['kernel'],
                                    build_on_device=cfg['n'] >= 2000)
if cfg['n'] >= 2000:
    ctx.set_scratch_bytes(64 << 30)
lo = np.max([d['xlo'] for d in ds], axis=0)
hi = np.min([d['xhi'] for d in ds], axis=0)
X_dev = ctx.dev_alloc(N * cfg['D'] * 8)
ctx.fill_uniform(X_dev, N, cfg['D'], seed=20252, lo=lo, hi=hi)
for _ in range(a.steps):
    r = g.score_designs_multi(models, None, list(cfg['criteria']), order=cfg['order'], noise_var=0.01, return_dc=False,
                              x_dev=X_dev, N=N)
print(r)
