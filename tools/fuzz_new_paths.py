"""Random configurations through the entry points added in round 2: taylor_multi + criteria on resident marginals (vs the
per-model drop-in functions, bit for bit), gpd_marginal_dx (vs central differences of the oracle's taylor_first_order),
the sharded call with a one-rank communicator (vs gpd_score_multi) and the fused option (vs the separate kernels)."""
import sys, warnings; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
warnings.filterwarnings('ignore')
import numpy as np, helpers
import gpdoemd_b200 as g
from gpdoemd_b200 import synthetic
from oracle import marginal as omarg


def check_seed(seed, ctx, comm):
    """-> (info string, dict of failures) for one random configuration; AttributeError = a binary combination without GP."""
    rng = np.random.default_rng(9000 + seed)
    kernel = ['RBF', 'Exponential', 'Matern32', 'Matern52', 'RatQuad'][seed % 5]
    D = int(rng.integers(1, 9)); nb = int(rng.integers(0, 3)) if D >= 3 else 0
    M, E = int(rng.integers(1, 5)), int(rng.integers(1, 9)); order = 1 + int(rng.integers(0, 2))
    P = [int(rng.integers(1, 11 if order == 1 else 5)) for _ in range(M)]
    n, N = int(rng.integers(30, 300)), int(rng.integers(1, 600))
    info = '%d %s D %d nb %d M %d E %d P %s n %d N %d order %d' % (seed, kernel, D, nb, M, E, P, n, N, order)
    oms, ds = helpers.oracle_models(1900 + seed, M=M, E=E, D=D, P=P, n=n, kernel=kernel, n_binary=nb, gp_noise=1e-4)
    gms = helpers.gpu_models(oms)
    lo = np.max([d['xlo'] for d in ds], axis=0); hi = np.min([d['xhi'] for d in ds], axis=0)
    X = synthetic.candidates(seed, N, D, lo, hi, n_binary=nb)
    fails = {}
    # 1. taylor_multi == per-model functions, resident criteria == upload path
    f = g.taylor_first_order if order == 1 else g.taylor_second_order
    mu0 = np.empty((N, M, E)); S0 = np.empty((N, M, E, E))
    for i, m in enumerate(gms): mu0[:, i], S0[:, i] = f(m, X)
    mu, S = g.taylor_multi(gms, X, order=order)
    if not (np.array_equal(mu, mu0) and np.array_equal(S, S0)): fails['taylor_multi'] = float(np.max(np.abs(S - S0)))
    for c in ('HR', 'BH', 'BF', 'AW', 'JR'):
        a, b = getattr(g, c)(mu, S, 0.02), getattr(g, c)(mu0, S0, 0.02)
        if not np.array_equal(a, b, equal_nan=True): fails['resident ' + c] = float(np.nanmax(np.abs(a - b)))
    # 2. sharded (world 1) == unsharded; fused option == separate kernels (first order)
    r0 = g.score_designs_multi(gms, X, ['BH', 'AW'], order=order, noise_var=0.02, idx_offset=17)
    r1 = g.score_designs_multi(gms, X, ['BH', 'AW'], order=order, noise_var=0.02, idx_offset=17, comm=comm)
    if r0['BH'][0] != r1['BH'][0] or not np.array_equal(r0['AW'][2], r1['AW'][2]): fails['sharded'] = 1
    if order == 1:
        ctx.set_option('fused', 1)
        try:
            r2 = g.score_designs_multi(gms, X, ['BH', 'AW'], order=1, noise_var=0.02, idx_offset=17)
        finally:
            ctx.set_option('fused', 0)
        e = helpers.relerr(r2['BH'][2], r0['BH'][2], np.abs(r0['BH'][2]).mean())
        if not e < 1e-9: fails['fused'] = e
    # 3. design gradients of model 0 vs central differences of the oracle
    om, gm = oms[0], gms[0]
    Xs = X[:min(N, 24)]
    M1, S1, dM, dS = g.taylor_first_order_dx(gm, Xs)
    span = hi - lo
    for d in range(D - nb):
        h = 1e-6 * span[d]
        Xp, Xm = Xs.copy(), Xs.copy(); Xp[:, d] += h; Xm[:, d] -= h
        Mp, Sp = omarg.taylor_first_order(om, Xp); Mm, Sm = omarg.taylor_first_order(om, Xm)
        fdM, fdS = (Mp - Mm) / (2 * h), (Sp - Sm) / (2 * h)
        tol = 5e-6 if kernel != 'Exponential' else 1e-4
        eM = helpers.relerr(dM[:, :, d], fdM, max(np.abs(fdM).max(), 1e-300)); eS = helpers.relerr(dS[:, :, :, d], fdS, max(np.abs(fdS).max(), 1e-300))
        if not (eM < tol and eS < 20 * tol): fails['dx%d' % d] = (eM, eS)
    for d in range(D - nb, D):
        if np.any(dM[:, :, d] != 0): fails['binary dx'] = 1
    return info, fails


if __name__ == '__main__':
    ctx = g.get_context(0)
    comm = g.LibraryComm(ctx, g.LibraryComm.unique_id(), 0, 1)
    bad = 0
    for seed in range(int(sys.argv[1]), int(sys.argv[2])):
        try:
            info, fails = check_seed(seed, ctx, comm)
            bad += bool(fails)
            print('FAIL' if fails else 'ok', info, fails if fails else '', flush=True)
        except AttributeError:
            print('skip', seed, 'missing binary combination (the reference raises too)', flush=True)
        except Exception as ex:
            bad += 1
            print('EXC', seed, repr(ex)[:300], flush=True)
    print('failures:', bad)
