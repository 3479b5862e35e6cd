"""One-off exploration: many random configurations, fused score and every hook vs the oracle; prints failures."""
import sys, warnings; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
warnings.filterwarnings('ignore')
import numpy as np, helpers
import gpdoemd_b200 as g
from gpdoemd_b200 import synthetic
import os
for _o in filter(None, os.environ.get('GPD_OPTIONS', '').split(',')):     # e.g. GPD_OPTIONS=fused=1,tri_parts=4
    g.get_context(0).set_option(_o.split('=')[0], int(_o.split('=')[1]))
from oracle import criteria as ocrit, marginal as omarg
lo_s, hi_s = int(sys.argv[1]), int(sys.argv[2])
SPARSE = len(sys.argv) > 3 and sys.argv[3] == 'sparse'           # SparseGPModel surrogates (VarDTC posterior)
DEVICE_BUILD = len(sys.argv) > 3 and sys.argv[3] == 'device'     # product models built by gpd_gp_build instead of adopted
bad = 0
for seed in range(lo_s, hi_s):
    rng = np.random.default_rng(7000 + seed)
    kernel = ['RBF', 'Exponential', 'Matern32', 'Matern52', 'RatQuad'][seed % 5]
    D = int(rng.integers(1, 9)); nb = int(rng.integers(0, 3)) if D >= 3 else 0
    M, E = int(rng.integers(1, 6)), int(rng.integers(1, 9)); order = 1 + int(rng.integers(0, 2))
    P = [int(rng.integers(1, 11 if order == 1 else 5)) for _ in range(M)]
    n, N = int(rng.integers(5, 300)), int(rng.integers(1, 700))
    noise = float(rng.uniform(0.005, 0.05))
    try:
        if SPARSE:
            n = max(n, 40)
            oms, ds = helpers.oracle_sparse_models(900 + seed, M, E, D, P, n, kernel, num_inducing=int(rng.integers(8, 40)), n_binary=nb)
        else:
            oms, ds = helpers.oracle_models(900 + seed, M=M, E=E, D=D, P=P, n=n, kernel=kernel, n_binary=nb, gp_noise=1e-4)
        if DEVICE_BUILD:
            gms, _ = synthetic.build_models(g.GPModel, 900 + seed, M, E, D, P, n, kernel, n_binary=nb, gp_noise=1e-4, build_on_device=True)
        elif SPARSE:
            gms = [g.SparseGPModel.from_reference(m) for m in oms]
        else:
            gms = helpers.gpu_models(oms)
        lo = np.max([d['xlo'] for d in ds], axis=0); hi = np.min([d['xhi'] for d in ds], axis=0)
        X = synthetic.candidates(seed, N, D, lo, hi, n_binary=nb)
        marg = omarg.taylor_first_order if order == 1 else omarg.taylor_second_order
        mu = np.zeros((N, M, E)); S = np.zeros((N, M, E, E))
        for i, om in enumerate(oms): mu[:, i], S[:, i] = marg(om, X)
        errs = {}
        res = g.score_designs_multi(gms, X, ['HR', 'BH', 'BF', 'AW', 'JR'], order=order, noise_var=noise)
        for c in ['HR', 'BH', 'BF', 'AW', 'JR']:
            d0 = ocrit.CRITERIA[c](mu.copy(), S.copy(), noise, None)
            ok = np.isfinite(d0)
            sc = max(np.abs(d0[ok]).mean() if ok.any() else 1.0, 1.0 if c == 'JR' else 1e-300)
            errs[c] = float(np.max(np.abs(res[c][2][ok] - d0[ok])) / sc) if ok.any() else 0.0
        # hooks of one model
        om, gm = oms[0], gms[0]
        ystd = om.ystd; g0 = next(x for x in om.gps[0] if x is not None)
        vs = g0.kern.kernx.variance * g0.kern.kernp.variance * om.ystd ** 2; pd = om.pmax - om.pmin
        M0, S0 = om.predict(X); M1, S1 = gm.predict(X)
        errs['mu'] = helpers.relerr(M1, M0, ystd); errs['s2'] = helpers.relerr(S1, S0, vs)
        e = int(rng.integers(0, E))
        errs['dmu'] = helpers.relerr(gm.d_mu_d_p(e, X), om.d_mu_d_p(e, X), ystd[e] / pd)
        errs['ds2'] = helpers.relerr(gm.d_s2_d_p(e, X), om.d_s2_d_p(e, X), vs[e] / pd)
        if P[0] <= 6:
            errs['d2mu'] = helpers.relerr(gm.d2_mu_d_p2(e, X), om.d2_mu_d_p2(e, X), ystd[e] / np.outer(pd, pd))
            errs['d2s2'] = helpers.relerr(gm.d2_s2_d_p2(e, X), om.d2_s2_d_p2(e, X), vs[e] / np.outer(pd, pd))
        tol = {'mu': 1e-9, 's2': 1e-9, 'dmu': 1e-9, 'ds2': 1e-9, 'd2mu': 1e-9, 'd2s2': 1e-8}
        if DEVICE_BUILD or SPARSE:   # two backward-stable factorisations agree to cond * u; sparse: root of woodbury_inv
            tol = {k: 100 * v for k, v in tol.items()}
        fails = {k: v for k, v in errs.items() if not (v < tol.get(k, 1e-7))}
        tag = 'FAIL' if fails else 'ok'
        bad += bool(fails)
        print(tag, seed, kernel, 'D', D, 'nb', nb, 'M', M, 'E', E, 'P', P, 'n', n, 'N', N, 'order', order, fails if fails else '', flush=True)
    except Exception as ex:
        bad += 1
        print('EXC', seed, kernel, 'D', D, 'nb', nb, 'M', M, 'E', E, 'P', P, 'n', n, 'N', N, 'order', order, repr(ex)[:200], flush=True)
print('failures:', bad)
