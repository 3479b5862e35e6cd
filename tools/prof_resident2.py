import sys, time, ctypes as C
import numpy as np
sys.path.insert(0, '.')
import gpdoemd_b200 as g
from gpdoemd_b200 import synthetic, _lib
from gpdoemd_b200._lib import dptr, check
cfg = synthetic.CONFIGS['C2']
ctx = g.get_context(0)
models, ds = synthetic.build_models(g.GPModel, 2, cfg['M'], cfg['E'], cfg['D'], cfg['P'], cfg['n'], cfg['kernel'])
lo = np.max([d['xlo'] for d in ds], axis=0); hi = np.min([d['xhi'] for d in ds], axis=0)
N, M, E = cfg['N'], cfg['M'], cfg['E']
X = ctx.pinned_empty((N, cfg['D'])); X[:] = synthetic.candidates(7, N, cfg['D'], lo, hi)
handles = (C.c_void_p * M)(*[m.device_handle() for m in models])
mu = ctx.pinned_empty((N, M, E)); S = ctx.pinned_empty((N, M, E, E))
def run(with_out, keep):
    k = C.c_void_p()
    t0 = time.perf_counter()
    check(ctx.lib.gpd_marginals(ctx.handle, C.cast(handles, _lib.c_void_pp), M, dptr(X), N, 1, dptr(mu) if with_out else None,
                                dptr(S) if with_out else None, C.byref(k) if keep else None))
    dt = (time.perf_counter() - t0) * 1e3
    if keep: ctx.lib.gpd_resident_free(k)
    return dt
for name, a in (('outputs + keep', (True, True)), ('keep only (no D2H)', (False, True)), ('outputs only', (True, False))):
    for _ in range(3): run(*a)
    print('%-22s %.3f ms' % (name, np.mean([run(*a) for _ in range(8)])))
t0 = time.perf_counter()
for _ in range(8): r = g.score_designs_multi(models, X, ['HR', 'BH'], noise_var=0.01, return_dc=False)
print('fused score (no dc)     %.3f ms' % ((time.perf_counter() - t0) * 1e3 / 8))
t0 = time.perf_counter()
for _ in range(8): a, b = g.taylor_multi(models, X)
print('taylor_multi (python)   %.3f ms' % ((time.perf_counter() - t0) * 1e3 / 8))
