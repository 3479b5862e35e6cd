"""CPU restatement of ``GPdoemd/discrimination_criteria.py:32-130`` (TEST INFRASTRUCTURE ONLY).

The reference evaluates ``scipy.stats.multivariate_normal.logpdf`` per (observation, model); here the same
Gaussian log-density is written out with a Cholesky factor (identical for symmetric positive definite S up to
rounding).  Pinned against the reference's own module run in the build container: tests/golden/obs_golden.npz
(tests/golden/make_obs_golden.py).
"""
import numpy as np
from scipy.stats import chi2 as scipy_chi2


def loglik(Y, M, S):
    """-> logpdf (n, N), maha (n, N)"""
    n, N, E = M.shape
    lp, mh = np.empty((n, N)), np.empty((n, N))
    for j in range(n):
        for i in range(N):
            L = np.linalg.cholesky(S[j, i])
            z = np.linalg.solve(L, Y[j] - M[j, i])
            mh[j, i] = z.dot(z)
            lp[j, i] = -0.5 * (E * np.log(2 * np.pi) + 2 * np.sum(np.log(np.diag(L))) + mh[j, i])
    return lp, mh


def _weights(scores):
    # discrimination_criteria.py:45-51,122-128
    return np.array([1. / np.sum([np.exp(min(max(s2 - s1, -1000), 100)) for s2 in scores]) for s1 in scores])


def gaussian_posterior(Y, M, S):
    return _weights(loglik(Y, M, S)[0].sum(axis=0))


def gaussian_posterior_update(y, M, S, prior):
    p = np.exp(loglik(y[None], M[None], S[None])[0][0]) * prior
    return p / p.sum()


def chi2(Y, M, S, D):
    n, N, E = M.shape
    S = np.diag(S) if S.ndim == 1 else S
    S = np.broadcast_to(S, (n, N, E, E)) if S.ndim == 2 else S
    maha = loglik(Y, M, S)[1].sum(axis=0)
    dof = n * E - np.asarray(D)
    if np.any(dof <= 0):
        raise RuntimeWarning('Degrees of freedom not greater than zero.')
    return 1. - scipy_chi2.cdf(maha, dof)


def akaike(Y, M, S, D):
    return _weights(2 * loglik(Y, M, S)[0].sum(axis=0) - 2 * np.asarray(D))
