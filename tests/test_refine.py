"""Continuous design refinement: host logic of the batched compass search with an injected scorer
(CPU), and on the device through the fused scorer (GPU): the refined criterion never decreases, beats the grid
argmax, respects the bounds and the fixed binary columns, and matches the oracle's criterion at the refined points."""
import numpy as np
import pytest

import helpers
from gpdoemd_b200 import refine_designs, synthetic


def test_compass_search_host_logic():
    opt = np.array([0.3, -0.2, 1.0])

    def scorer(X):
        return -np.sum((X - opt) ** 2, axis=1)
    lo, hi = np.array([-1., -1., 0.]), np.array([1., 1., 1.])
    X0 = np.array([[0.9, 0.9, 1.0], [-0.5, 0.1, 0.0], [0.31, -0.21, 1.0]])
    X, f, evals = refine_designs(None, X0, bounds=(lo, hi), binary=(2,), iters=60, scorer=scorer)
    assert X.shape == (3, 3) and evals > 3
    assert np.all(np.diff(f) <= 0)                              # sorted by decreasing criterion
    assert np.all((X >= lo) & (X <= hi))
    assert set(X[:, 2]) == {0.0, 1.0}                           # binary column untouched
    top = X[X[:, 2] == 1.0]
    assert np.allclose(top[:, :2], opt[:2], atol=1e-4)          # both starts with b = 1 reach the optimum
    assert np.isclose(f[0], 0.0, atol=1e-7)
    # bound-constrained optimum: the maximiser lies outside the box
    X, f, _ = refine_designs(None, np.array([[0.0, 0.0]]), bounds=([-1, -1], [0.2, 1]), iters=60,
                             scorer=lambda Z: -np.sum((Z - np.array([0.7, 0.1])) ** 2, axis=1))
    assert np.allclose(X[0], [0.2, 0.1], atol=1e-4)
    # NaN polls never win
    X, f, _ = refine_designs(None, np.array([[0.0, 0.0]]), bounds=([-1, -1], [1, 1]), iters=5,
                             scorer=lambda Z: np.where(Z[:, 0] > 0, np.nan, -np.sum(Z ** 2, axis=1)))
    assert np.allclose(X[0], 0.0) and f[0] == 0.0


@pytest.mark.gpu
def test_refinement_improves_on_the_grid_argmax(gpu_ctx):
    import gpdoemd_b200 as g
    from oracle import criteria as ocrit
    from oracle import marginal as omarg
    oms, ds = helpers.oracle_models(41, M=3, E=2, D=3, P=2, n=150, kernel='Matern52', n_binary=1)
    gms = helpers.gpu_models(oms)
    lo = np.max([d['xlo'] for d in ds], axis=0)
    hi = np.min([d['xhi'] for d in ds], axis=0)
    grid = helpers.demo_mesh(lo, hi, num=12)                     # 12 x 12 x {0, 1}
    bi, bv, dc = g.score_designs(gms, grid, order=1, criterion='BH', noise_var=0.01)
    top = grid[np.argsort(-dc)[:8]]
    X, f, evals = refine_designs(gms, top, criterion='BH', order=1, noise_var=0.01, bounds=(lo, hi), binary=(2,), iters=30)
    assert f[0] >= bv and np.all(np.diff(f) <= 0)
    assert f[0] > bv * (1 + 1e-6)                                # a 12-point grid is not at the continuum optimum
    assert np.all((X >= lo - 1e-12) & (X <= hi + 1e-12)) and set(X[:, 2]) <= {0.0, 1.0}
    # the refined values are the oracle's criterion at the refined points
    mu = np.zeros((len(X), 3, 2))
    S = np.zeros((len(X), 3, 2, 2))
    for i, om in enumerate(oms):
        mu[:, i], S[:, i] = omarg.taylor_first_order(om, X)
    d0 = ocrit.BH(mu, S, 0.01, None)
    assert helpers.relerr(f, d0) < 1e-7
