import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpdoemd_b200 as g
from gpdoemd_b200 import synthetic
ctx = g.get_context(0)
models, ds = synthetic.build_models(g.GPModel, 2, 4, 2, 3, 4, 128, 'RBF')
lo = np.max([d['xlo'] for d in ds], axis=0); hi = np.min([d['xhi'] for d in ds], axis=0)
for N in (4096, 8192, 16384, 65536, 262144):
    X_dev = ctx.dev_alloc(N * 3 * 8)
    ctx.fill_uniform(X_dev, N, 3, seed=1, lo=lo, hi=hi)
    for _ in range(3):
        g.score_designs_multi(models, None, ['BH'], order=1, noise_var=0.01, return_dc=False, x_dev=X_dev, N=N)
    ctx.profile_enable(True)
    for _ in range(10):
        g.score_designs_multi(models, None, ['BH'], order=1, noise_var=0.01, return_dc=False, x_dev=X_dev, N=N)
    p = ctx.profile_read(); ctx.profile_enable(False)
    print(N, 'panel MB', N * 8 * 128 * 8 / 1e6, 'eval us/step', p['eval']['ms'] * 100, 'ns/cand', p['eval']['ms'] * 1e5 / N,
          'tri ns/cand', p['tri']['ms'] * 1e5 / N)
    ctx.dev_free(X_dev)
