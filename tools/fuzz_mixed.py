"""One-off: rival models that differ in kernel class, training-set size AND dim_p, fused score vs the oracle."""
import sys, warnings; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
warnings.filterwarnings('ignore')
import numpy as np, helpers
import gpdoemd_b200 as g
from gpdoemd_b200 import synthetic
import os
for _o in filter(None, os.environ.get('GPD_OPTIONS', '').split(',')):     # e.g. GPD_OPTIONS=fused=1,tri_parts=4
    g.get_context(0).set_option(_o.split('=')[0], int(_o.split('=')[1]))
from oracle import criteria as ocrit, marginal as omarg
KS = ['RBF', 'Exponential', 'Matern32', 'Matern52', 'RatQuad']
bad = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(3000 + seed)
    D, E = int(rng.integers(1, 5)), int(rng.integers(1, 5))
    M = int(rng.integers(2, 5)); order = 1 + int(rng.integers(0, 2))
    N = int(rng.integers(50, 3000)); noise = 0.02
    oms, los, his = [], [], []
    desc = []
    for m in range(M):
        k = KS[int(rng.integers(0, 5))]; n = int(rng.integers(10, 400)); P = int(rng.integers(1, 5))
        o, d = helpers.oracle_models(4000 + 10 * seed + m, M=1, E=E, D=D, P=P, n=n, kernel=k, gp_noise=1e-4)
        oms.append(o[0]); los.append(d[0]['xlo']); his.append(d[0]['xhi']); desc.append((k, n, P))
    gms = helpers.gpu_models(oms)
    lo, hi = np.max(los, axis=0), np.min(his, axis=0)
    if np.any(hi <= lo):
        continue
    X = synthetic.candidates(seed, N, D, lo, hi)
    marg = omarg.taylor_first_order if order == 1 else omarg.taylor_second_order
    mu = np.zeros((N, M, E)); S = np.zeros((N, M, E, E))
    for i, om in enumerate(oms): mu[:, i], S[:, i] = marg(om, X)
    res = g.score_designs_multi(gms, X, ['HR', 'BH', 'BF', 'AW', 'JR'], order=order, noise_var=noise)
    errs = {}
    for c in res:
        d0 = ocrit.CRITERIA[c](mu.copy(), S.copy(), noise, None); ok = np.isfinite(d0)
        sc = max(np.abs(d0[ok]).mean(), 1.0 if c == 'JR' else 1e-300)
        errs[c] = float(np.max(np.abs(res[c][2][ok] - d0[ok])) / sc)
    fails = {k: v for k, v in errs.items() if not v < 1e-7}
    bad += bool(fails)
    print('FAIL' if fails else 'ok', seed, 'D', D, 'E', E, 'order', order, 'N', N, desc, fails if fails else '', flush=True)
print('failures:', bad)
