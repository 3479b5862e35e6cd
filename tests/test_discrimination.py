"""Discrimination criteria on the observed set: the oracle restatement (CPU) and the product's
drop-in functions (GPU, gpd_gauss_loglik) against golden vectors produced by the reference's own
GPdoemd/discrimination_criteria.py (tests/golden/make_obs_golden.py)."""
import os

import numpy as np
import pytest

from oracle import discrimination as odisc

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'obs_golden.npz'))
CASES = range(5)
TOL = 1e-9        # weights / p-values: FP64, relative (floor 1e-300: losing models have weights ~ 1e-100)


def _check(mod, k):
    key = 'c%d_' % k
    Y, M, S, D = G[key + 'Y'], G[key + 'M'], G[key + 'S'], G[key + 'D']
    np.testing.assert_allclose(mod.gaussian_posterior(Y, M, S), G[key + 'gp'], rtol=TOL, atol=1e-300)
    np.testing.assert_allclose(mod.akaike(Y, M, S, D), G[key + 'aw'], rtol=TOL, atol=1e-300)
    np.testing.assert_allclose(mod.chi2(Y, M, S, D), G[key + 'chi2_full'], rtol=1e-7, atol=1e-14)
    np.testing.assert_allclose(mod.chi2(Y, M, S[0, 0], D), G[key + 'chi2_mat'], rtol=1e-7, atol=1e-14)
    np.testing.assert_allclose(mod.chi2(Y, M, np.diag(S[0, 0]).copy(), D), G[key + 'chi2_vec'], rtol=1e-7, atol=1e-14)
    np.testing.assert_allclose(mod.gaussian_posterior_update(Y[-1], M[-1], S[-1], G[key + 'prior']), G[key + 'gpu'],
                               rtol=TOL, atol=1e-300)


def _check_winner(mod):
    """tests/test_discrimination_criteria.py:31-100: model 0 wins; shapes and normalisation."""
    Y, M, S = G['w_Y'], G['w_M'], G['w_S']
    for P, ref in ((mod.gaussian_posterior(Y, M, S), G['w_gp']), (mod.akaike(Y, M, S, np.array([2] * 3)), G['w_aw'])):
        assert P.shape == (3,) and 0.999 < np.sum(P) < 1.001 and P[0] > 0.99
        np.testing.assert_allclose(P, ref, rtol=TOL, atol=1e-300)
    c = mod.chi2(Y, M, S, np.array([2] * 3))
    np.testing.assert_allclose(c, G['w_chi2'], rtol=1e-7, atol=1e-14)
    with pytest.raises(RuntimeWarning):
        mod.chi2(Y, M, S, np.array([100] * 3))


@pytest.mark.parametrize('k', CASES)
def test_oracle_vs_reference_golden(k):
    _check(odisc, k)


def test_oracle_winner_case():
    _check_winner(odisc)


@pytest.mark.gpu
@pytest.mark.parametrize('k', CASES)
def test_device_vs_reference_golden(k, gpu_ctx):
    from gpdoemd_b200 import discrimination_criteria as gdisc
    _check(gdisc, k)


@pytest.mark.gpu
def test_device_winner_case_and_loglik(gpu_ctx):
    from gpdoemd_b200 import discrimination_criteria as gdisc
    _check_winner(gdisc)
    key = 'c3_'                                   # E = 8: the largest register-resident inverse
    lp1, mh1 = gdisc._loglik(G[key + 'Y'], G[key + 'M'], G[key + 'S'])
    lp0, mh0 = odisc.loglik(G[key + 'Y'], G[key + 'M'], G[key + 'S'])
    np.testing.assert_allclose(lp1, lp0, rtol=1e-11)
    np.testing.assert_allclose(mh1, mh0, rtol=1e-11)
