"""Discrimination criteria on the observed data set -- SURVEY 8f row 3.

Same names, signatures, shapes and error behaviour as ``GPdoemd/discrimination_criteria.py:32-130``
(``gaussian_posterior``, ``gaussian_posterior_update``, ``chi2``, ``akaike``).  The per-(observation, model)
Gaussian log-densities and squared Mahalanobis distances -- the E x E inverses and determinants -- come from ONE
device call (``gpd_gauss_loglik``, the same register-resident inverse as the design criteria); the sums over the
handful of observations and the weight normalisation are host glue, in the reference's order of operations.
"""
import numpy as np
from scipy.stats import chi2 as scipy_chi2

from . import _lib
from ._lib import check, dptr
from .context import get_context


def _loglik(Y, M, S, device=None):
    """-> logpdf (n, N), maha (n, N) for Y (n, E), M (n, N, E), S (n, N, E, E)."""
    Y, M, S = _lib.as_f64(Y), _lib.as_f64(M), _lib.as_f64(S)
    n, N, E = M.shape
    assert Y.shape == (n, E) and S.shape == (n, N, E, E)
    ctx = get_context(device)
    logpdf, maha = np.empty((n, N)), np.empty((n, N))
    check(ctx.lib.gpd_gauss_loglik(ctx.handle, dptr(Y), dptr(M), dptr(S), n, N, E, dptr(logpdf), dptr(maha)))
    return logpdf, maha


def _weights(scores):
    """1 / sum_j exp(clip(s_j - s_i, -1000, 100))   (discrimination_criteria.py:45-51, :122-128)"""
    out = []
    for s1 in scores:
        acc = 0
        for s2 in scores:
            acc += np.exp(np.min((np.max((s2 - s1, -1000)), 100)))
        out.append(1. / acc)
    return np.array(out)


def gaussian_posterior(Y, M, S):
    """discrimination_criteria.py:32-52.  Y [n x E], M [n x N x E], S [n x N x E x E]  ->  [N]"""
    logpdf, _ = _loglik(Y, M, S)
    return _weights([np.sum(logpdf[:, i]) for i in range(logpdf.shape[1])])


def gaussian_posterior_update(y, M, S, prior):
    """discrimination_criteria.py:57-66.  y [E], M [N x E], S [N x E x E], prior [N]  ->  [N]"""
    logpdf, _ = _loglik(np.asarray(y)[None, :], np.asarray(M)[None], np.asarray(S)[None])
    pup = np.exp(logpdf[0]) * prior
    return pup / np.sum(pup)


def chi2(Y, M, S, D):
    """discrimination_criteria.py:72-103.  S is [E], [E x E] or [n x N x E x E]; D [N] parameters per model."""
    n, N, E = M.shape
    S = np.asarray(S, dtype=float)
    if S.ndim == 1:
        S = np.diag(S)
    if S.ndim == 2:
        S = np.broadcast_to(S, (n, N, E, E))
    _, maha = _loglik(Y, M, S)
    maha = np.sum(maha, axis=0)
    dof = n * E - np.asarray(D)
    if np.any(dof <= 0):
        raise RuntimeWarning('Degrees of freedom not greater than zero.')
    return 1. - scipy_chi2.cdf(maha, dof)


def akaike(Y, M, S, D):
    """discrimination_criteria.py:109-130.  -> Akaike weights [N]"""
    logpdf, _ = _loglik(Y, M, S)
    logL = [np.sum(logpdf[:, i]) for i in range(logpdf.shape[1])]
    aic = 2 * np.array(logL) - 2 * np.asarray(D)
    return _weights(aic)
