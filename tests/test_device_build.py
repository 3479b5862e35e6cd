"""GP-state build on the device (gpd_gp_build): covariance matrix, blocked Cholesky with GPy's
jitchol retries and alpha computed in device memory, against the oracle's LAPACK restatement of GPy's
ExactGaussianInference (oracle/gp.py) on the same training data and hyper-parameters.

A Cholesky factor is pinned by its BACKWARD error (|L L^T - Ky| <= c n u |Ky|), which does not depend on the
conditioning; L itself, alpha and the predictions agree with LAPACK's only up to cond(Ky) * u, so those tolerances
are stated per conditioning (noise 1e-2: cond ~ 1e4; the reference's default gp_noise_var 1e-6: cond ~ 1e8).
"""
import numpy as np
import pytest

import helpers
from gpdoemd_b200 import synthetic

U = 2.0 ** -53


def _pair(seed, n, kernel, noise, E=1, D=3, P=2, n_binary=0):
    """oracle model (host LAPACK) and the product model built on the device, same data and hyper-parameters"""
    from gpdoemd_b200 import GPModel
    oms, ds = helpers.oracle_models(seed, M=1, E=E, D=D, P=P, n=n, kernel=kernel, gp_noise=noise, n_binary=n_binary)
    gms, _ = synthetic.build_models(GPModel, seed, 1, E, D, P, n, kernel, n_binary=n_binary, gp_noise=noise,
                                    build_on_device=True)
    return oms[0], gms[0], ds[0]


def _K_diff(kern, Z):
    """the covariance in the difference form the device uses (GPy's expanded form differs by O(|z/l|^2 u))"""
    out = 1.0
    for k in (kern.kernx, kern.kernp):
        A = np.asarray(Z)[:, np.asarray(k.active_dims)] / k.lengthscale
        out = out * k.K_of_r(np.sqrt(((A[:, None, :] - A[None, :, :]) ** 2).sum(-1)))
    return out


def test_device_gp_is_lazy_and_has_no_cpu_fallback():
    """CPU side: constructing / re-parameterising a DeviceExactGP does no arithmetic; the posterior needs the GPU."""
    import gpdoemd_b200 as g
    from gpdoemd_b200 import kernels
    from gpdoemd_b200.gp_state import ProductKernel
    rng = np.random.default_rng(0)
    Z, Y = rng.uniform(size=(20, 3)), rng.normal(size=(20, 1))
    gp = g.DeviceExactGP(Z, Y, ProductKernel(kernels.RBF(2, [0, 1], 'kernx'), kernels.RBF(1, [2], 'kernp')))
    h = gp[:]
    assert h.shape == (2 + 1 + 1 + 1 + 1,)
    gp[:] = np.r_[1.5, 0.4, 0.6, 0.8, 0.7, 1e-3]
    assert gp.noise_var == 1e-3 and gp.kern.kernx.variance == 1.5
    assert gp._posterior is None and gp._dims() == (2, 1)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        with pytest.raises(Exception):
            gp.posterior


@pytest.mark.gpu
@pytest.mark.parametrize('kernel,n', [('RBF', 37), ('Matern52', 128), ('Matern32', 300), ('RBF', 500),
                                      ('Matern52', 1000), ('RatQuad', 129)])
def test_factor_backward_error_and_alpha_residual(kernel, n, gpu_ctx):
    om, gm, d = _pair(3, n, kernel, 1e-6)
    go, gg = om.gps[0][0], gm.gps[0][0]
    post = gg.posterior                      # device build, alpha and L downloaded
    assert gg.jitter == 0.0
    L, alpha = post.woodbury_chol, post.woodbury_vector[:, 0]
    Ky = _K_diff(go.kern, go.X)
    Ky[np.diag_indices(n)] += 1e-6 + 1e-8
    # backward error of the factorisation, entrywise against |L||L^T| (Higham 10.4: <= (n+1) u)
    bound = np.abs(L).dot(np.abs(L).T)
    assert np.max(np.abs(L.dot(L.T) - Ky) / bound) < 4 * (n + 1) * U
    assert np.all(np.triu(L, 1) == 0) and np.all(np.diag(L) > 0)
    # alpha: normwise backward error of the solve
    res = Ky.dot(alpha) - go.Y[:, 0]
    eta = np.linalg.norm(res) / (np.linalg.norm(Ky, 2) * np.linalg.norm(alpha) + np.linalg.norm(go.Y))
    assert eta < 50 * n * U
    # against LAPACK's factor / solution: equal up to cond(Ky) u
    cond = np.linalg.cond(Ky)
    L0, a0 = go.posterior.woodbury_chol, go.posterior.woodbury_vector[:, 0]
    eL, ea = np.linalg.norm(L - L0) / np.linalg.norm(L0), np.linalg.norm(alpha - a0) / np.linalg.norm(a0)
    print('n=%d %s: cond %.1e, backward %.1f (n+1)u, eta %.1f n u, |dL|/|L| %.1e, |dalpha|/|alpha| %.1e'
          % (n, kernel, cond, np.max(np.abs(L.dot(L.T) - Ky) / bound) / ((n + 1) * U), eta / (n * U), eL, ea))
    assert eL < 100 * cond * U
    assert ea < 100 * cond * U


@pytest.mark.gpu
@pytest.mark.parametrize('noise,tol_mean,tol_var', [(1e-2, 1e-10, 1e-10), (1e-6, 1e-9, 1e-9)])
def test_predictions_of_device_built_model(noise, tol_mean, tol_var, gpu_ctx):
    """Every hook of a model whose GPs were built on the device vs the oracle (host LAPACK inference).  With the
    reference's default gp noise 1e-6 the two alphas differ by up to cond * u (3e-9 relative at n = 500, both being
    backward stable), but the predictions stay inside the north-star tolerance (measured on B200: mean 1e-12,
    variance 3e-14 at noise 1e-6)."""
    om, gm, d = _pair(8, 400, 'Matern52', noise, E=2, D=3, P=3)
    X = synthetic.candidates(2, 500, 3, d['xlo'], d['xhi'])
    g0 = om.gps[0][0]
    ystd, vscale = om.ystd, g0.kern.kernx.variance * g0.kern.kernp.variance * om.ystd ** 2
    pd = om.pmax - om.pmin
    M0, S0 = om.predict(X)
    M1, S1 = gm.predict(X)
    print('noise %g: mean %.2e, var %.2e' % (noise, helpers.relerr(M1, M0, ystd), helpers.relerr(S1, S0, vscale)))
    assert helpers.relerr(M1, M0, ystd) < tol_mean
    assert helpers.relerr(S1, S0, vscale) < tol_var
    for e in range(2):
        assert helpers.relerr(gm.d_mu_d_p(e, X), om.d_mu_d_p(e, X), ystd[e] / pd) < tol_mean * 10
        assert helpers.relerr(gm.d2_s2_d_p2(e, X), om.d2_s2_d_p2(e, X), vscale[e] / np.outer(pd, pd)) < 1e-7


@pytest.mark.gpu
def test_device_build_with_binary_routing_and_reload(gpu_ctx):
    """gp_surrogate + gp_load_hyp on the device for a model with a binary design variable (one GP per output and
    combination, gp_model.py:104-124); loading new hyper-parameters rebuilds the device state."""
    om, gm, d = _pair(31, 260, 'Matern32', 1e-3, E=2, D=3, P=2, n_binary=1)
    X = synthetic.candidates(6, 301, 3, d['xlo'], d['xhi'], n_binary=1)
    M0, S0 = om.predict(X)
    M1, S1 = gm.predict(X)
    assert helpers.relerr(M1, M0, om.ystd) < 1e-8
    hyp2 = [[h * np.r_[np.full(len(h) - 1, 1.3), 1.0] for h in hs] for hs in d['hyp']]
    for m in (om, gm):
        m.hyp = hyp2
        m.gp_load_hyp()
    M0b, _ = om.predict(X)
    M1b, _ = gm.predict(X)
    assert helpers.relerr(M1b, M0b, om.ystd) < 1e-8
    assert helpers.relerr(M0b, M0, om.ystd) > 1e-4          # the reload really changed the surrogate


@pytest.mark.gpu
def test_jitchol_retry_on_the_device(gpu_ctx):
    """A covariance that is numerically indefinite at GPy's fixed 1e-8 (variance 1e10, long lengthscales): the
    device factorisation reports the bad pivot and retries with jitter mean(diag) * 1e-6 like GPy's jitchol."""
    import gpdoemd_b200 as g
    from gpdoemd_b200 import kernels
    from gpdoemd_b200.gp_state import ProductKernel
    rng = np.random.default_rng(4)
    n = 150
    Z, Y = rng.uniform(size=(n, 3)), rng.normal(size=(n, 1))
    kern = ProductKernel(kernels.RBF(2, [0, 1], 'kernx'), kernels.RBF(1, [2], 'kernp'))
    gp = g.DeviceExactGP(Z, Y, kern)
    gp[:] = np.r_[1e10, 5.0, 5.0, 1.0, 5.0, 0.0]
    post = gp.posterior
    diag_mean = 1e10 + 1e-8
    ratio = gp.jitter / (diag_mean * 1e-6)
    k = int(round(np.log10(ratio)))
    assert 0 <= k <= 4 and np.isclose(ratio, 10.0 ** k, rtol=1e-12)
    L = post.woodbury_chol
    Ky = _K_diff(kern, Z) + np.eye(n) * (1e-8 + gp.jitter)
    assert np.max(np.abs(L.dot(L.T) - Ky)) / 1e10 < 1e-12
    assert np.all(np.isfinite(post.woodbury_vector))


@pytest.mark.gpu
def test_build_and_loglik_error_paths(gpu_ctx):
    """C-ABI error behaviour of the new entry points: non-zero status + message, nothing computed, no fallback."""
    import ctypes as C
    from gpdoemd_b200 import kernels
    from gpdoemd_b200._lib import GpdError, check, dptr
    from gpdoemd_b200.gp_state import DeviceExactGP, GPState, ProductKernel
    lib = gpu_ctx.lib
    rng = np.random.default_rng(1)
    Z, Y = rng.uniform(size=(30, 3)), rng.normal(size=(30, 1))
    gp = DeviceExactGP(Z, Y, ProductKernel(kernels.Matern32(2, [0, 1], 'kernx'), kernels.Matern32(1, [2], 'kernp')))
    st = GPState.from_gp(gp, 2, 1, with_posterior=False)
    with pytest.raises(GpdError, match='negative noise'):
        st.build(gpu_ctx, Y[:, 0], -1.0)
    d = st._desc()
    d.factor_kind = 1                                                 # GPD_FACTOR_ROOT is not something to build
    h = C.c_void_p()
    y = np.ascontiguousarray(Y[:, 0])
    assert lib.gpd_gp_build(gpu_ctx.handle, C.byref(d), dptr(y), 1e-3, C.byref(h), None, None, None) != 0
    assert b'Cholesky' in lib.gpd_last_error() and not h.value
    with pytest.raises(AssertionError):
        st.upload(gpu_ctx)                                            # no posterior to upload
    st.ls_x[0] = -1.0
    with pytest.raises(GpdError, match='lengthscale'):
        st.build(gpu_ctx, Y[:, 0], 1e-3)
    # gauss_loglik: E out of range, empty observation set
    Yo, Mo, So = np.zeros((2, 9)), np.zeros((2, 3, 9)), np.tile(np.eye(9), (2, 3, 1, 1))
    lp, mh = np.empty((2, 3)), np.empty((2, 3))
    assert lib.gpd_gauss_loglik(gpu_ctx.handle, dptr(Yo), dptr(Mo), dptr(So), 2, 3, 9, dptr(lp), dptr(mh)) != 0
    assert b'1 <= E <= 8' in lib.gpd_last_error()
    check(lib.gpd_gauss_loglik(gpu_ctx.handle, None, None, None, 0, 3, 2, None, None))
    # a singular covariance: scipy's multivariate_normal.logpdf (discrimination_criteria.py:40) raises LinAlgError
    # ("singular matrix"); the device path reports the zero pivot the same way instead of returning inf / NaN
    from gpdoemd_b200 import discrimination_criteria as gdisc
    with pytest.raises(np.linalg.LinAlgError):
        gdisc._loglik(np.zeros((1, 2)), np.zeros((1, 1, 2)), np.zeros((1, 1, 2, 2)))
    lpz, _ = gdisc._loglik(np.zeros((1, 2)), np.zeros((1, 1, 2)), np.tile(np.eye(2), (1, 1, 1, 1)))
    assert np.isfinite(lpz[0, 0])


@pytest.mark.gpu
@pytest.mark.parametrize('n', [129, 300, 640, 1000, 2100])
def test_setup_variants_agree(n, gpu_ctx):
    """Tensor-pipe setup kernels (option setup_variant = 1: recursive-doubling inverse of the factor on DMMA, row-wise
    Cholesky block / panel kernels, multi-CTA substitutions for alpha) against the round-1 FMA kernels (= 0) on the same
    data: same factor up to rounding, same predictions.  n covers 2, 3, 5, 8 and 17 row blocks (pairs with a short or
    missing right half at every level of the recursion)."""
    from gpdoemd_b200 import GPModel
    res = {}
    try:
        for variant in (0, 1):
            gpu_ctx.set_option('setup_variant', variant)
            gms, ds = synthetic.build_models(GPModel, 5, 1, 1, 3, 2, n, 'Matern52', gp_noise=1e-4, build_on_device=True)
            gm = gms[0]
            X = synthetic.candidates(4, 300, 3, ds[0]['xlo'], ds[0]['xhi'])
            post = gm.gps[0][0].posterior
            M, S = gm.predict(X)
            res[variant] = (post.woodbury_chol.copy(), post.woodbury_vector[:, 0].copy(), M, S, gm.d2_s2_d_p2(0, X))
    finally:
        gpu_ctx.set_option('setup_variant', 1)
    L0, a0, M0, S0, H0 = res[0]
    L1, a1, M1, S1, H1 = res[1]
    assert np.all(np.triu(L1, 1) == 0) and np.all(np.diag(L1) > 0)
    # both are backward-stable factorisations of the same matrix: entrywise against |L||L^T|
    bound = np.abs(L0).dot(np.abs(L0).T)
    assert np.max(np.abs(L1.dot(L1.T) - L0.dot(L0.T)) / bound) < 8 * (n + 1) * U
    cond = np.linalg.cond(L0) ** 2
    assert np.linalg.norm(a1 - a0) / np.linalg.norm(a0) < 100 * cond * U
    g0 = gm.gps[0][0]
    ystd = float(np.ravel(gm.ystd)[0])
    vscale = float(g0.kern.kernx.variance * g0.kern.kernp.variance) * ystd ** 2     # k** in original units
    print('n=%d: cond %.1e, |dalpha| %.1e, mean %.1e, var %.1e, d2s2 %.1e' % (
        n, cond, np.linalg.norm(a1 - a0) / np.linalg.norm(a0), np.max(np.abs(M1 - M0)) / ystd,
        np.max(np.abs(S1 - S0)) / vscale, np.max(np.abs(H1 - H0)) / max(float(np.max(np.abs(H0))), 1e-300)))
    assert np.max(np.abs(M1 - M0)) < 1e-9 * ystd
    assert np.max(np.abs(S1 - S0)) < 1e-9 * vscale
    assert np.max(np.abs(H1 - H0)) < 1e-7 * max(float(np.max(np.abs(H0))), 1e-300)


@pytest.mark.gpu
def test_uploaded_factor_inverse_variants_agree(gpu_ctx):
    """gpd_gp_upload (host factor, the from_reference path) through both inversion kernels"""
    oms, ds = helpers.oracle_models(7, M=1, E=1, D=3, P=2, n=700, kernel='RBF', gp_noise=1e-4)
    X = synthetic.candidates(4, 257, 3, ds[0]['xlo'], ds[0]['xhi'])
    out = {}
    try:
        for variant in (0, 1):
            gpu_ctx.set_option('setup_variant', variant)
            gm = helpers.gpu_models(oms)[0]
            out[variant] = gm.predict(X)
    finally:
        gpu_ctx.set_option('setup_variant', 1)
    M0, S0 = oms[0].predict(X)
    for variant in (0, 1):
        assert helpers.relerr(out[variant][0], M0, oms[0].ystd) < 1e-9
    g0 = oms[0].gps[0][0]
    vscale = float(g0.kern.kernx.variance * g0.kern.kernp.variance) * float(np.ravel(oms[0].ystd)[0]) ** 2
    for variant in (0, 1):
        assert helpers.relerr(out[variant][1], S0, vscale) < 1e-9
    assert np.max(np.abs(out[1][1] - out[0][1])) < 1e-10 * vscale
