"""Independent third-party pin of the GP part of the oracle.

GPy (the library the reference calls) is not available in this image, so the oracle's restatement of GPy's exact GP
(oracle/gpy_kernels.py, oracle/gp.py) cannot be checked against GPy itself.  scikit-learn IS available and implements
the same mathematical objects independently: anisotropic RBF / Matern(1/2, 3/2, 5/2) kernels and the exact-GP
predictive mean and standard deviation (Rasmussen & Williams alg. 2.1).  A product kernel ``kernx * kernp`` on disjoint
column sets is expressed in scikit-learn (which has no active_dims) as a product of two kernels over ALL columns with
length-scale 1e150 on the columns a factor must ignore.

What this pins: the kernel definitions (K matrices), mu and sigma^2 of ``predict`` -- i.e. that the oracle computes
the textbook GP the reference asks GPy for.  What stays unpinned: GPy-specific conventions (SURVEY Appendix A items
marked with a warning sign: the 1e-8 jitter, gradients_XX at r = 0, RatQuad parameter order); the p-derivatives are
pinned by finite differences in tests/test_oracle_gp.py.
"""
import numpy as np
import pytest
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern

import helpers
from gpdoemd_b200 import synthetic

BIG = 1e150


def _sk_factor(name, ls_full):
    if name == 'RBF':
        return RBF(length_scale=ls_full)
    nu = {'Exponential': 0.5, 'Matern32': 1.5, 'Matern52': 2.5}[name]
    return Matern(length_scale=ls_full, nu=nu)


def _sk_kernel(name, gp, dim):
    kx, kp = gp.kern.kernx, gp.kern.kernp
    lx, lp = np.full(dim, BIG), np.full(dim, BIG)
    lx[np.asarray(kx.active_dims)] = kx.lengthscale
    lp[np.asarray(kp.active_dims)] = kp.lengthscale
    return ConstantKernel(kx.variance * kp.variance) * _sk_factor(name, lx) * _sk_factor(name, lp)


@pytest.mark.parametrize('kernel', ['RBF', 'Exponential', 'Matern32', 'Matern52'])
@pytest.mark.parametrize('noise', [1e-2, 1e-6])
def test_predict_matches_scikit_learn(kernel, noise):
    oms, ds = helpers.oracle_models(17, M=1, E=2, D=3, P=2, n=90, kernel=kernel, gp_noise=noise)
    om = oms[0]
    X = synthetic.candidates(5, 40, 3, ds[0]['xlo'], ds[0]['xhi'])
    Zs = np.c_[om.trans_x(X), np.tile(om.trans_p(om.pmean), (len(X), 1))]
    M0, S0 = om.predict(X)
    for e in range(2):
        gp = om.gps[e][0]
        k = _sk_kernel(kernel, gp, 5)
        # kernel matrix
        np.testing.assert_allclose(k(gp.X), gp.kern.K(gp.X), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(k(Zs, gp.X), gp.kern.K(Zs, gp.X), rtol=1e-9, atol=1e-12)
        # exact GP posterior with fixed hyper-parameters; GPy adds 1e-8 to the noise on the diagonal
        sk = GaussianProcessRegressor(kernel=k, alpha=noise + 1e-8, optimizer=None, normalize_y=False)
        sk.fit(gp.X, gp.Y[:, 0])
        mu, sd = sk.predict(Zs, return_std=True)
        mu_o = (M0[:, e] - om.ymean[e]) / om.ystd[e]          # back to the GP's standardised units
        s2_o = S0[:, e] / om.ystd[e] ** 2
        kss = gp.kern.kernx.variance * gp.kern.kernp.variance
        # conditioning: cond(K) ~ 1e4 (noise 1e-2) .. 1e8 (noise 1e-6); two backward-stable solves agree to cond * u
        tol = 1e-9 if noise > 1e-3 else 2e-6
        assert np.max(np.abs(mu - mu_o)) < tol * max(1.0, np.max(np.abs(mu_o)))
        assert np.max(np.abs(sd ** 2 - s2_o)) < tol * kss
