"""Size limits the reference does not have: each fails LOUDLY with a message that names the limit;
nothing falls back to a slower or a CPU path.  The limits are listed in INTEGRATION.md section "Limits"."""
import numpy as np
import pytest

from gpdoemd_b200 import synthetic
from gpdoemd_b200._lib import GpdError

pytestmark = pytest.mark.gpu


def _models(M=1, E=1, D=2, P=2, n=40, kernel='Matern52'):
    import gpdoemd_b200 as g
    gms, ds = synthetic.build_models(g.GPModel, 3, M, E, D, P, n, kernel)
    lo = np.max([d['xlo'] for d in ds], axis=0)
    hi = np.min([d['xhi'] for d in ds], axis=0)
    return gms, synthetic.candidates(1, 50, D, lo, hi)


def test_more_than_eight_outputs(gpu_ctx):
    gms, X = _models(E=9)
    with pytest.raises(GpdError, match=r'num_outputs out of range \(1\.\.8\)'):
        gms[0].predict(X)


def test_criteria_with_more_than_eight_outputs(gpu_ctx):
    import gpdoemd_b200 as g
    mu, S = np.zeros((5, 2, 9)), np.tile(np.eye(9), (5, 2, 1, 1))
    with pytest.raises(GpdError, match=r'criteria support 1 <= E <= 8 outputs'):
        g.BH(mu, S, 0.1)


def test_more_than_sixteen_parameters(gpu_ctx):
    gms, X = _models(P=17)
    with pytest.raises(GpdError, match=r'dim_p out of range \(1\.\.16\)'):
        gms[0].predict(X)


def test_more_than_eight_continuous_design_dimensions(gpu_ctx):
    gms, X = _models(D=9)
    with pytest.raises(GpdError, match=r'x-kernel acts on 1\.\.8 non-binary design dimensions'):
        gms[0].predict(X)


def test_first_order_works_up_to_sixteen_parameters_second_order_up_to_ten(gpu_ctx):
    import gpdoemd_b200 as g
    gms, X = _models(M=2, P=16)
    bi, bv, dc = g.score_designs(gms, X, order=1, criterion='BH', noise_var=0.01)       # 17-column bucket
    assert np.all(np.isfinite(dc))
    gms, X = _models(M=2, P=11)
    with pytest.raises(GpdError, match=r'(P <= 10|at most 72 coefficient columns)'):
        g.score_designs(gms, X, order=2, criterion='BH', noise_var=0.01)
    with pytest.raises(GpdError, match=r'(P <= 10|at most 72 coefficient columns)'):
        gms[0].d2_s2_d_p2(0, X)
    gms, X = _models(M=2, P=10)
    bi, bv, dc = g.score_designs(gms, X, order=2, criterion='BH', noise_var=0.01)
    assert np.all(np.isfinite(dc))


def test_more_than_sixty_four_rival_models(gpu_ctx):
    import gpdoemd_b200 as g
    gms, X = _models(M=1)
    with pytest.raises(GpdError, match=r'number of models out of range \(1\.\.64\)'):
        g.score_designs(gms * 65, X, order=1, criterion='HR', noise_var=0.01)


def test_singular_covariance_raises_like_numpy(gpu_ctx):
    """np.linalg.inv raises LinAlgError on a singular S + noise (design_criteria.py:83); so does the device
    path (a zero pivot sets a flag, the entry point returns GPD_ERR_SINGULAR) instead of returning NaN as the argmax."""
    import gpdoemd_b200 as g
    mu = np.random.default_rng(0).normal(size=(6, 2, 2))
    S = np.tile(np.eye(2), (6, 2, 1, 1))
    S[3, 1] = 0.0                                   # singular with noise_var = None -> 0
    with pytest.raises(np.linalg.LinAlgError):
        g.BH(mu, S)
    assert np.all(np.isfinite(g.BH(mu, S, 0.1)))    # the flag is cleared: the next call is clean
    assert np.all(np.isfinite(g.HR(mu, S)))         # HR never inverts
