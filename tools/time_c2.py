"""C2 step and per-family device times (A/B of library builds through GPDOEMD_B200_LIB)."""
import sys
import numpy as np
sys.path.insert(0, '.')
import gpdoemd_b200 as g
from gpdoemd_b200 import synthetic
ctx = g.get_context(0)
import os
for _o in filter(None, os.environ.get('GPD_OPTIONS', '').split(',')):     # e.g. GPD_OPTIONS=fused=2
    ctx.set_option(_o.split('=')[0], int(_o.split('=')[1]))
cfg = dict(synthetic.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else 'C2'])
if len(sys.argv) > 2:
    cfg['N'] = int(float(sys.argv[2]))
if len(sys.argv) > 3:
    cfg['kernel'] = sys.argv[3]          # kernel class override, e.g. RatQuad
if cfg['n'] >= 2000:
    ctx.set_scratch_bytes(64 << 30)
gms, ds = synthetic.build_models(g.GPModel, 2, cfg['M'], cfg['E'], cfg['D'], cfg['P'], cfg['n'], cfg['kernel'], build_on_device=cfg['n'] >= 2000)
lo = np.max([d['xlo'] for d in ds], axis=0); hi = np.min([d['xhi'] for d in ds], axis=0)
N, D = cfg['N'], cfg['D']
X_dev = ctx.dev_alloc(N * D * 8)
ctx.fill_uniform(X_dev, N, D, seed=20252, lo=lo, hi=hi)
crits = list(cfg['criteria'])
for _ in range(3):
    r = g.score_designs_multi(gms, None, crits, order=cfg['order'], noise_var=0.01, return_dc=False, x_dev=X_dev, N=N)
ctx.profile_enable(True)
tot = 0.0
K = 10 if cfg['n'] < 2000 else 3
for _ in range(K):
    ctx.flush_l2(); ctx.timer_start()
    r = g.score_designs_multi(gms, None, crits, order=cfg['order'], noise_var=0.01, return_dc=False, x_dev=X_dev, N=N)
    tot += ctx.timer_stop()
prof = ctx.profile_read()
print('%.3f ms per step  %s  argmax %s' % (tot / K, {k: round(v['ms'] / K, 3) for k, v in prof.items() if v['launches']}, [r[c][0] for c in crits]))
