"""Development probe: print error magnitudes of every quantity vs the oracle (no asserts)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
import helpers
import gpdoemd_b200 as g
from gpdoemd_b200 import synthetic

ctx = g.get_context(0)
print('fp64 dmma peak TF/s', ctx.probe_fp64_peak())
for kernel in ['RBF', 'Matern52']:
    for n in (200, 500):
        oms, ds = helpers.oracle_models(11, M=1, E=2, D=3, P=3, n=n, kernel=kernel)
        gm = helpers.gpu_models(oms)[0]
        om = oms[0]
        X = synthetic.candidates(3, 300, 3, ds[0]['xlo'], ds[0]['xhi'])
        M0, S0 = om.predict(X)
        t = time.time()
        M1, S1 = gm.predict(X)
        print(kernel, n, 'predict time', time.time() - t)
        gp = om.gps[0][0]
        kss = gp.kern.kernx.variance * gp.kern.kernp.variance
        print('  mu  ', helpers.relerr(M1, M0, om.ystd), ' s2 ', helpers.relerr(S1, S0, kss * om.ystd ** 2))
        for e in range(2):
            pd = om.pmax - om.pmin
            print('  dmu ', helpers.relerr(gm.d_mu_d_p(e, X), om.d_mu_d_p(e, X), om.ystd[e] / pd),
                  ' d2mu', helpers.relerr(gm.d2_mu_d_p2(e, X), om.d2_mu_d_p2(e, X), om.ystd[e] / np.outer(pd, pd)),
                  ' ds2 ', helpers.relerr(gm.d_s2_d_p(e, X), om.d_s2_d_p(e, X), kss * om.ystd[e] ** 2 / pd),
                  ' d2s2', helpers.relerr(gm.d2_s2_d_p2(e, X), om.d2_s2_d_p2(e, X), kss * om.ystd[e] ** 2 / np.outer(pd, pd)))
for variant in (0, 1):
    ctx.set_option('tri_variant', variant)
    cfg = synthetic.CONFIGS['C2']
    t = time.time()
    gms, ds = synthetic.build_models(g.GPModel, 2, cfg['M'], cfg['E'], cfg['D'], cfg['P'], cfg['n'], cfg['kernel'])
    lo = np.max([d['xlo'] for d in ds], axis=0); hi = np.min([d['xhi'] for d in ds], axis=0)
    X = synthetic.candidates(7, cfg['N'], cfg['D'], lo, hi)
    g.score_designs(gms, X[:1000], order=1, criterion='BH', noise_var=0.01)
    print('setup s', time.time() - t)
    ctx.profile_enable(True)
    for it in range(3):
        ctx.timer_start()
        bi, bv, dc = g.score_designs(gms, X, order=1, criterion='BH', noise_var=0.01, return_dc=False)
        ms = ctx.timer_stop()
        print('variant', variant, 'C2 score ms', ms, 'cand/s', cfg['N'] / ms * 1e3, bi, bv)
    prof = ctx.profile_read()
    for k, v in prof.items():
        if v['launches']:
            print('   ', k, v, 'TF/s' , v['flops'] / max(v['ms'], 1e-9) / 1e9)
    ctx.profile_enable(False)
