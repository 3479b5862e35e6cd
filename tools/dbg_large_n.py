import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, helpers
from gpdoemd_b200 import synthetic
import gpdoemd_b200 as g
for kernel, n in (('Matern52', 4000), ('Matern52', 2000), ('Matern52', 1000), ('RBF', 2100)):
    for trim in (1, 0):
        g.get_context(0).set_option('tri_trim', trim)
        oms, ds = helpers.oracle_models(91, M=1, E=1, D=3, P=3, n=n, kernel=kernel)
        gm = helpers.gpu_models(oms)[0]; om = oms[0]
        X = synthetic.candidates(2, 200, 3, ds[0]['xlo'], ds[0]['xhi'])
        gp = om.gps[0][0]
        ystd = om.ystd; vscale = gp.kern.kernx.variance * gp.kern.kernp.variance * om.ystd ** 2
        pd = om.pmax - om.pmin
        M0, S0 = om.predict(X); M1, S1 = gm.predict(X)
        print(kernel, n, 'trim', trim, 'mu %.2e s2 %.2e' % (helpers.relerr(M1, M0, ystd), helpers.relerr(S1, S0, vscale)),
              'dmu %.2e ds2 %.2e d2s2 %.2e' % (helpers.relerr(gm.d_mu_d_p(0, X), om.d_mu_d_p(0, X), ystd[0] / pd),
              helpers.relerr(gm.d_s2_d_p(0, X), om.d_s2_d_p(0, X), vscale[0] / pd),
              helpers.relerr(gm.d2_s2_d_p2(0, X), om.d2_s2_d_p2(0, X), vscale[0] / np.outer(pd, pd))), flush=True)
        if trim == 0 and n <= 2000: break
