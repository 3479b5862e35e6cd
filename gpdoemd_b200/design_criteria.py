"""Drop-in design criteria ``HR, BH, BF, AW, JR`` (``GPdoemd/design_criteria.py:49-194``).

Signature ``f(mu, s2, noise_var=None, pps=None) -> (n,)`` and input normalisation exactly as
``_reshape`` (``design_criteria.py:202-238``); the per-design E x E inverses, determinants and pair sums
run in CUDA (csrc/kernels_crit.cu).

One deliberate difference: the reference's BH and BF add ``noise_var`` into the caller's ``s2`` in
place (``design_criteria.py:81,114`` on a reshape view).  Here inputs are never modified unless
``mutate_like_reference=True`` is passed, which reproduces that side effect for callers that depend on
it (the demo loop ``demos/case_study.py:289-290`` scores BF/AW/JR on the accumulated noise).
"""
import numpy as np

from . import _lib
from ._lib import check, dptr
from .context import get_context


def _reshape(mu, s2, noise_var=None, pps=None):
    if mu.ndim == 2:
        mu = np.expand_dims(mu, axis=2)
    n, M, E = mu.shape
    if noise_var is None:
        noise_var = np.zeros((E, E))
    if isinstance(noise_var, (int, float)):
        noise_var = np.array([noise_var])
    if noise_var.ndim == 1:
        tmp = np.eye(E)
        np.fill_diagonal(tmp, noise_var)
        noise_var = tmp
    assert noise_var.shape == (E, E)
    assert np.all(np.diag(noise_var) >= 0.)
    s2 = s2.reshape((n, M, E, E))
    if pps is None:
        pps = np.ones(M)
    if isinstance(pps, (list, tuple)):
        pps = np.array(pps)
    assert pps.shape == (M,) and np.all(pps >= 0)
    pps = pps / np.sum(pps)
    return mu, s2, noise_var, pps, n, M, E


def _run(kind, mu, s2, noise_var, pps, device=None):
    from .marginal import resident_pair
    res = resident_pair(mu, s2)           # untouched outputs of taylor_multi: score the device copy, no upload
    mu, s2v, noise_var, pps, n, M, E = _reshape(mu, s2, noise_var, pps)
    dc = np.empty(n)
    if res is not None:
        check(res.ctx.lib.gpd_criterion_resident(res.ctx.handle, _lib.CRITERION_IDS[kind], res.handle,
                                                 dptr(_lib.as_f64(noise_var)), dptr(_lib.as_f64(pps)), dptr(dc)))
        return dc, s2v, noise_var
    ctx = get_context(device)
    check(ctx.lib.gpd_criterion(ctx.handle, _lib.CRITERION_IDS[kind], dptr(_lib.as_f64(mu)), dptr(_lib.as_f64(s2v)),
                                n, M, E, dptr(_lib.as_f64(noise_var)), dptr(_lib.as_f64(pps)), dptr(dc)))
    return dc, s2v, noise_var


def HR(mu, s2, noise_var=None, pps=None):
    """Hunter & Reiner (1965): sum_{i<j} |mu_i - mu_j|^2   (design_criteria.py:49-63)."""
    return _run('HR', mu, s2, None, None)[0]


def BH(mu, s2, noise_var=None, pps=None, mutate_like_reference=False):
    """Box & Hill (1967), Prasad & Someswara Rao (1977) extension   (design_criteria.py:66-92)."""
    dc, s2v, nv = _run('BH', mu, s2, noise_var, pps)
    if mutate_like_reference:
        s2v += nv            # (read-only taylor_multi outputs raise here: they cannot be mutated and stay resident)
    return dc


def BF(mu, s2, noise_var=None, pps=None, mutate_like_reference=False):
    """Buzzi-Ferraris et al. (1990)   (design_criteria.py:95-123)."""
    dc, s2v, nv = _run('BF', mu, s2, noise_var, pps)
    if mutate_like_reference:
        s2v += nv
    return dc


def AW(mu, s2, noise_var=None, pps=None):
    """Michalik et al. (2010) Akaike-weights criterion   (design_criteria.py:126-144)."""
    return _run('AW', mu, s2, noise_var, pps)[0]


def JR(mu, s2, noise_var=None, pps=None):
    """Olofsson et al. (2019) quadratic Jensen-Renyi divergence   (design_criteria.py:147-194)."""
    return _run('JR', mu, s2, noise_var, pps)[0]
