import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, helpers
import gpdoemd_b200 as g
from gpdoemd_b200 import synthetic
from oracle import criteria as ocrit, marginal as omarg
for seed in [int(a) for a in sys.argv[1:]] or [2, 12, 14, 18, 22, 23]:
    rng = np.random.default_rng(1000 + seed)
    kernel = ['RBF', 'Exponential', 'Matern32', 'Matern52', 'RatQuad'][seed % 5]
    D = int(rng.integers(1, 7)); nb = int(rng.integers(0, 2)) if D >= 2 else 0
    M, E = int(rng.integers(1, 5)), int(rng.integers(1, 5)); order = 1 + int(rng.integers(0, 2))
    P = [int(rng.integers(1, 7 if order == 1 else 4)) for _ in range(M)]
    n, N = int(rng.integers(20, 161)), int(rng.integers(1, 301))
    crit = ['HR', 'BH', 'BF', 'AW', 'JR'][int(rng.integers(0, 5))]; noise = float(rng.uniform(0.005, 0.05))
    oms, ds = helpers.oracle_models(500 + seed, M=M, E=E, D=D, P=P, n=n, kernel=kernel, n_binary=nb, gp_noise=1e-4)
    gms = helpers.gpu_models(oms)
    lo = np.max([d['xlo'] for d in ds], axis=0); hi = np.min([d['xhi'] for d in ds], axis=0)
    X = synthetic.candidates(seed, N, D, lo, hi, n_binary=nb)
    marg = omarg.taylor_first_order if order == 1 else omarg.taylor_second_order
    gmarg = g.taylor_first_order if order == 1 else g.taylor_second_order
    mu = np.zeros((N, M, E)); S = np.zeros((N, M, E, E)); mu1 = mu.copy(); S1 = S.copy()
    for i, (om, gm) in enumerate(zip(oms, gms)):
        mu[:, i], S[:, i] = marg(om, X)
        mu1[:, i], S1[:, i] = gmarg(gm, X)
    print(seed, kernel, 'D', D, 'nb', nb, 'M', M, 'E', E, 'P', P, 'n', n, 'N', N, 'order', order, crit)
    print('   marginal mu err %.2e  S err %.2e' % (np.max(np.abs(mu1 - mu)) / np.abs(mu).max(), np.max(np.abs(S1 - S)) / np.abs(S).max()))
    for c in ['HR', 'BH', 'BF', 'AW', 'JR']:
        d0 = ocrit.CRITERIA[c](mu.copy(), S.copy(), noise, None)
        d_drop = getattr(g, c)(mu.copy(), S.copy(), noise, None)          # device criterion on oracle mu,S
        d_fused = g.score_designs(gms, X, order=order, criterion=c, noise_var=noise)[2]
        sc = np.abs(d0).mean()
        print('   %s: drop-in vs oracle %.2e, fused vs oracle %.2e' % (c, np.max(np.abs(d_drop - d0)) / sc, np.max(np.abs(d_fused - d0)) / sc))
