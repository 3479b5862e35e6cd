"""Design gradients.  Parity = central finite differences of the ORACLE:

* CPU: the chain rule through HR / BH / BF / AW / JR (design_gradients.criterion_gradient) against finite differences of
  the oracle's criteria on an analytic family mu(x), S(x); the L-BFGS refinement loop with an injected objective;
* GPU: gpd_marginal_dx (mu, S, dmu/dx, dS/dx of the first-order Taylor marginal) against central differences of the
  oracle's taylor_first_order at h = 1e-6, for every kernel class and with a binary design variable; the criterion
  gradient end to end; L-BFGS refinement reaches at least the compass search's optimum with fewer evaluations."""
import numpy as np
import pytest

import helpers
from gpdoemd_b200 import criterion_gradient, refine_designs, synthetic
from oracle import criteria as ocrit
from oracle import marginal as omarg


def _family(rng, K, M, E, D):
    A = rng.normal(size=(M, E, D))
    b = rng.normal(size=(M, E))
    B0 = np.array([np.eye(E) * 2.0 + 0.3 * (lambda Q: Q @ Q.T)(rng.normal(size=(E, E))) for _ in range(M)])
    Bd = 0.15 * rng.normal(size=(M, E, E, D))
    Bd = 0.5 * (Bd + Bd.transpose(0, 2, 1, 3))

    def f(X):
        mu = np.einsum('med,kd->kme', A, X) + b
        S = B0[None] + np.einsum('mabd,kd->kmab', Bd, X)
        return mu, S
    dmu = np.broadcast_to(A, (K, M, E, D)).copy()
    dS = np.broadcast_to(Bd, (K, M, E, E, D)).copy()
    return f, dmu, dS


@pytest.mark.parametrize('kind', ['HR', 'BH', 'BF', 'AW', 'JR'])
def test_chain_rule_against_finite_differences_of_the_oracle_criteria(kind):
    rng = np.random.default_rng(3)
    K, M, E, D = 7, 3, 2, 3
    fam, dmu, dS = _family(rng, K, M, E, D)
    X = rng.uniform(-1, 1, (K, D))
    noise = np.array([[0.05, 0.01], [0.01, 0.08]])
    pps = [0.2, 0.5, 0.3]
    mu, S = fam(X)
    f, g = criterion_gradient(kind, mu, S, dmu, dS, noise, pps)
    f0 = ocrit.CRITERIA[kind](mu, S.copy(), noise, pps)
    assert helpers.relerr(f, f0) < 1e-12
    h = 1e-6
    for d in range(D):
        Xp, Xm = X.copy(), X.copy()
        Xp[:, d] += h
        Xm[:, d] -= h
        mp, Sp = fam(Xp)
        mm, Sm = fam(Xm)
        fd = (ocrit.CRITERIA[kind](mp, Sp, noise, pps) - ocrit.CRITERIA[kind](mm, Sm, noise, pps)) / (2 * h)
        assert helpers.relerr(g[:, d], fd, np.abs(fd).max()) < 1e-7, (kind, d)


@pytest.mark.parametrize('shape', [(5, 5, 1, 2), (4, 4, 3, 3), (3, 2, 4, 2)])
def test_chain_rule_other_shapes(shape):
    """the demo shape (M = 5 rival models, one output) and larger E, every criterion, random prior probabilities"""
    K, M, E, D = shape
    rng = np.random.default_rng(K * M + E)
    fam, dmu, dS = _family(rng, K, M, E, D)
    X = rng.uniform(-1, 1, (K, D))
    noise, pps = 0.05 * np.eye(E), rng.uniform(0.1, 1, M)
    mu, S = fam(X)
    for kind in ('HR', 'BH', 'BF', 'AW', 'JR'):
        f, g = criterion_gradient(kind, mu, S, dmu, dS, noise, pps)
        assert helpers.relerr(f, ocrit.CRITERIA[kind](mu, S.copy(), noise, pps)) < 1e-12
        for d in range(D):
            h = 1e-6
            Xp, Xm = X.copy(), X.copy()
            Xp[:, d] += h
            Xm[:, d] -= h
            fd = (ocrit.CRITERIA[kind](*fam(Xp), noise, pps) - ocrit.CRITERIA[kind](*fam(Xm), noise, pps)) / (2 * h)
            assert helpers.relerr(g[:, d], fd, np.abs(fd).max()) < 1e-7, (kind, d)


def test_lbfgs_refinement_host_logic():
    opt = np.array([0.3, -0.2, 1.0])

    def scorer(X):
        return -np.sum((X - opt) ** 2, axis=1)

    def vg(X):
        return scorer(X), -2.0 * (X - opt)
    lo, hi = np.array([-1., -1., 0.]), np.array([1., 1., 1.])
    X0 = np.array([[0.9, 0.9, 1.0], [-0.5, 0.1, 0.0], [0.31, -0.21, 1.0]])
    X, f, evals = refine_designs(None, X0, bounds=(lo, hi), binary=(2,), iters=40, scorer=scorer, value_and_grad=vg)
    assert np.all(np.diff(f) <= 0) and np.all((X >= lo) & (X <= hi)) and set(X[:, 2]) == {0.0, 1.0}
    assert np.allclose(X[:, :2], opt[:2], atol=1e-6)
    _, _, evals_c = refine_designs(None, X0, bounds=(lo, hi), binary=(2,), iters=40, scorer=scorer, method='compass')
    assert evals < evals_c
    # bound-constrained optimum
    X, f, _ = refine_designs(None, np.array([[0.0, 0.0]]), bounds=([-1, -1], [0.2, 1]), iters=40,
                             scorer=lambda Z: -np.sum((Z - np.array([0.7, 0.1])) ** 2, axis=1),
                             value_and_grad=lambda Z: (-np.sum((Z - np.array([0.7, 0.1])) ** 2, axis=1),
                                                       -2 * (Z - np.array([0.7, 0.1]))))
    assert np.allclose(X[0], [0.2, 0.1], atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize('kernel,nb', [('RBF', 0), ('Matern52', 1), ('Matern32', 0), ('Exponential', 0), ('RatQuad', 0)])
def test_marginal_dx_against_central_differences_of_the_oracle(kernel, nb, gpu_ctx):
    import gpdoemd_b200 as g
    D = 3
    oms, ds = helpers.oracle_models(81, M=1, E=2, D=D, P=3, n=140, kernel=kernel, n_binary=nb, gp_noise=1e-4)
    om, gm = oms[0], helpers.gpu_models(oms)[0]
    X = synthetic.candidates(4, 150, D, ds[0]['xlo'], ds[0]['xhi'], n_binary=nb)
    M1, S1, dM, dS = g.taylor_first_order_dx(gm, X)
    M0, S0 = omarg.taylor_first_order(om, X)
    ystd = om.ystd
    sc = np.sqrt(np.einsum('nii->ni', S0))
    assert helpers.relerr(M1, M0, ystd) < 1e-9
    assert helpers.relerr(S1, S0, sc[:, :, None] * sc[:, None, :]) < 1e-9
    span = ds[0]['xhi'] - ds[0]['xlo']
    for d in range(D):
        if d >= D - nb:
            assert np.all(dM[:, :, d] == 0) and np.all(dS[:, :, :, d] == 0)      # binary design variable
            continue
        h = 1e-6 * span[d]
        Xp, Xm = X.copy(), X.copy()
        Xp[:, d] += h
        Xm[:, d] -= h
        Mp, Sp = omarg.taylor_first_order(om, Xp)
        Mm, Sm = omarg.taylor_first_order(om, Xm)
        fdM, fdS = (Mp - Mm) / (2 * h), (Sp - Sm) / (2 * h)
        # finite-difference limited: O(h^2 f''') + O(u f / h); the Exponential kernel has a cusp at training inputs
        tol = 2e-6 if kernel != 'Exponential' else 2e-5
        assert helpers.relerr(dM[:, :, d], fdM, np.abs(fdM).max()) < tol, (kernel, d)
        assert helpers.relerr(dS[:, :, :, d], fdS, np.abs(fdS).max()) < 10 * tol, (kernel, d)


@pytest.mark.gpu
@pytest.mark.parametrize('crit', ['HR', 'BH', 'BF', 'AW', 'JR'])
def test_criterion_gradient_end_to_end(crit, gpu_ctx):
    import gpdoemd_b200 as g
    oms, ds = helpers.oracle_models(83, M=3, E=2, D=2, P=2, n=120, kernel='Matern52', gp_noise=1e-4)
    gms = helpers.gpu_models(oms)
    lo = np.max([d['xlo'] for d in ds], axis=0)
    hi = np.min([d['xhi'] for d in ds], axis=0)
    X = synthetic.candidates(6, 40, 2, lo, hi)

    def oracle(Xc):
        mu = np.zeros((len(Xc), 3, 2))
        S = np.zeros((len(Xc), 3, 2, 2))
        for i, om in enumerate(oms):
            mu[:, i], S[:, i] = omarg.taylor_first_order(om, Xc)
        return ocrit.CRITERIA[crit](mu, S, 0.02, None)
    f, grad = g.criterion_with_gradient(gms, X, crit, noise_var=0.02)
    assert helpers.relerr(f, oracle(X), np.abs(oracle(X)).mean()) < 1e-7
    for d in range(2):
        h = 1e-6 * (hi[d] - lo[d])
        Xp, Xm = X.copy(), X.copy()
        Xp[:, d] += h
        Xm[:, d] -= h
        fd = (oracle(Xp) - oracle(Xm)) / (2 * h)
        assert helpers.relerr(grad[:, d], fd, np.abs(fd).max()) < 1e-5, (crit, d)


@pytest.mark.gpu
def test_lbfgs_refinement_on_the_device(gpu_ctx):
    import gpdoemd_b200 as g
    oms, ds = helpers.oracle_models(41, M=3, E=2, D=3, P=2, n=150, kernel='Matern52', n_binary=1)
    gms = helpers.gpu_models(oms)
    lo = np.max([d['xlo'] for d in ds], axis=0)
    hi = np.min([d['xhi'] for d in ds], axis=0)
    grid = helpers.demo_mesh(lo, hi, num=12)
    bi, bv, dc = g.score_designs(gms, grid, order=1, criterion='BH', noise_var=0.01)
    top = grid[np.argsort(-dc)[:8]]
    Xl, fl, el = refine_designs(gms, top, criterion='BH', noise_var=0.01, bounds=(lo, hi), binary=(2,), iters=30, method='lbfgs')
    Xc, fc, ec = refine_designs(gms, top, criterion='BH', noise_var=0.01, bounds=(lo, hi), binary=(2,), iters=30, method='compass')
    print('BH: grid %.6g  compass %.6g (%d evaluations)  lbfgs %.6g (%d evaluations)' % (bv, fc[0], ec, fl[0], el))
    assert fl[0] >= bv and fl[0] >= fc[0] * (1 - 1e-6)
    assert np.all((Xl >= lo - 1e-12) & (Xl <= hi + 1e-12)) and set(Xl[:, 2]) <= {0.0, 1.0}
    # JR: analytic gradient as well (default method), never worse than the start points and than the compass search
    bj, vj, dcj = g.score_designs(gms, grid, order=1, criterion='JR', noise_var=0.01)
    topj = grid[np.argsort(-dcj)[:4]]
    Xj, fj, ej = refine_designs(gms, topj, criterion='JR', noise_var=0.01, bounds=(lo, hi), binary=(2,), iters=30)
    Xjc, fjc, ejc = refine_designs(gms, topj, criterion='JR', noise_var=0.01, bounds=(lo, hi), binary=(2,), iters=30, method='compass')
    print('JR: grid %.6g  compass %.6g (%d evaluations)  lbfgs %.6g (%d evaluations)' % (vj, fjc[0], ejc, fj[0], ej))
    assert np.all(np.isfinite(fj)) and fj[0] >= vj and fj[0] >= fjc[0] * (1 - 1e-6)
