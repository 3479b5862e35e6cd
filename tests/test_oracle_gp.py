"""Internal-consistency pins of the GP part of the oracle (GPy itself is not available, so parity with
real GPy is UNPINNED -- see oracle/__init__.py): finite differences of the oracle's own mean/variance,
extended-precision recomputation, sigma^2 via L vs via K^-1, interpolation at the training inputs."""
import numpy as np
import pytest

import helpers
from gpdoemd_b200 import synthetic
from oracle import gpy_kernels

KERNELS = ['RBF', 'Exponential', 'Matern32', 'Matern52', 'RatQuad']


def _model(kernel, seed=3, n=60, E=2, D=2, P=3):
    oms, ds = helpers.oracle_models(seed, M=1, E=E, D=D, P=P, n=n, kernel=kernel)
    X = synthetic.candidates(1, 12, D, ds[0]['xlo'], ds[0]['xhi'])
    return oms[0], X


@pytest.mark.parametrize('kernel', KERNELS)
def test_derivative_hooks_vs_finite_differences(kernel):
    om, X = _model(kernel)
    p0 = om.pmean.copy()
    h = 1e-5 * (om.pmax - om.pmin)
    P = om.dim_p

    def at(p):
        om.pmean = p
        return om.predict(X)
    for e in range(om.num_outputs):
        om.pmean = p0
        dmu, ds2 = om.d_mu_d_p(e, X), om.d_s2_d_p(e, X)
        d2mu, d2s2 = om.d2_mu_d_p2(e, X), om.d2_s2_d_p2(e, X)
        for a in range(P):
            ea = np.zeros(P)
            ea[a] = h[a]
            Mp, Sp = at(p0 + ea)
            Mm, Sm = at(p0 - ea)
            fd_mu = (Mp[:, e] - Mm[:, e]) / (2 * h[a])
            fd_s2 = (Sp[:, e] - Sm[:, e]) / (2 * h[a])
            assert np.max(np.abs(fd_mu - dmu[:, a])) <= 1e-6 * max(1.0, np.max(np.abs(dmu[:, a])))
            assert np.max(np.abs(fd_s2 - ds2[:, a])) <= 1e-5 * max(1e-3, np.max(np.abs(ds2[:, a])))
            # second derivatives: central difference of the analytic first derivative
            om.pmean = p0 + ea
            gp_, sp_ = om.d_mu_d_p(e, X), om.d_s2_d_p(e, X)
            om.pmean = p0 - ea
            gm_, sm_ = om.d_mu_d_p(e, X), om.d_s2_d_p(e, X)
            fd2 = (gp_ - gm_) / (2 * h[a])
            fs2 = (sp_ - sm_) / (2 * h[a])
            assert np.max(np.abs(fd2 - d2mu[:, a, :])) <= 1e-5 * max(1.0, np.max(np.abs(d2mu)))
            assert np.max(np.abs(fs2 - d2s2[:, a, :])) <= 1e-4 * max(1e-2, np.max(np.abs(d2s2)))
        om.pmean = p0
        assert np.allclose(d2mu, d2mu.transpose(0, 2, 1), rtol=1e-10, atol=1e-12)
        assert np.allclose(d2s2, d2s2.transpose(0, 2, 1), rtol=1e-8, atol=1e-10)


def test_variance_via_cholesky_equals_explicit_inverse():
    om, X = _model('Matern52', n=80)
    gp = om.gps[0][0]
    Z = np.c_[om.trans_x(X), np.tile(om.trans_p(om.pmean), (len(X), 1))]
    K = gp.kern.K(gp.X, Z)
    mu, var = gp.predict_noiseless(Z)
    iK = gp.posterior.woodbury_inv
    var2 = gp.kern.Kdiag(Z) - np.einsum('in,ij,jn->n', K, iK, K)
    kss = gp.kern.kernx.variance * gp.kern.kernp.variance
    assert np.max(np.abs(var[:, 0] - var2)) < 1e-9 * kss
    assert np.allclose(mu[:, 0], K.T.dot(iK.dot(gp.Y))[:, 0], rtol=1e-7, atol=1e-9)


def test_interpolation_at_training_inputs():
    om, _ = _model('Matern52', n=50)
    Xtr = om.trans_x(om.Z[:, :om.dim_x], back=True)
    for i in (0, 7, 23):
        om.pmean = om.trans_p(om.Z[i, om.dim_x:], back=True)
        M, S = om.predict(Xtr[i:i + 1])
        ytrue = om.trans_y(om.Y[i:i + 1], back=True)
        assert np.allclose(M, ytrue, atol=1e-3 * om.ystd.max())
        assert np.all(S < 1e-4 * om.ystd ** 2) and np.all(S > -1e-9)


def test_extended_precision_truth():
    """80-bit recomputation of mean and variance on a 40-point GP: the float64 oracle agrees to c*u*cond."""
    om, X = _model('RBF', n=40, E=1)
    gp = om.gps[0][0]
    ld = np.longdouble
    Z = np.c_[om.trans_x(X), np.tile(om.trans_p(om.pmean), (len(X), 1))].astype(ld)
    Xt = gp.X.astype(ld)

    def kern(A, B):
        out = np.ones((len(A), len(B)), dtype=ld)
        for part in (gp.kern.kernx, gp.kern.kernp):
            a = A[:, part.active_dims] / part.lengthscale.astype(ld)
            b = B[:, part.active_dims] / part.lengthscale.astype(ld)
            r2 = ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1)
            out = out * ld(part.variance) * np.exp(-r2 / 2)
        return out
    Ky = kern(Xt, Xt) + np.eye(len(Xt), dtype=ld) * ld(gp.noise_var + 1e-8)
    n = len(Xt)
    L = np.zeros((n, n), dtype=ld)
    for j in range(n):                                  # long-double Cholesky
        L[j, j] = np.sqrt(Ky[j, j] - (L[j, :j] ** 2).sum())
        for i in range(j + 1, n):
            L[i, j] = (Ky[i, j] - (L[i, :j] * L[j, :j]).sum()) / L[j, j]

    def solve_lower(Lm, B):
        Xs = np.zeros_like(B)
        for i in range(n):
            Xs[i] = (B[i] - Lm[i, :i].dot(Xs[:i])) / Lm[i, i]
        return Xs
    y = gp.Y.astype(ld)
    w = solve_lower(L, y)
    Ks = kern(Xt, Z)
    V = solve_lower(L, Ks)
    mu_true = (V * w).sum(0)
    var_true = ld(gp.kern.kernx.variance * gp.kern.kernp.variance) - (V ** 2).sum(0)
    mu, var = gp.predict_noiseless(np.asarray(Z, dtype=float))
    kss = gp.kern.kernx.variance * gp.kern.kernp.variance
    cond = np.linalg.cond(gp.posterior.woodbury_chol) ** 2
    assert np.max(np.abs(mu[:, 0] - mu_true.astype(float))) < 50 * 2.2e-16 * cond * max(1.0, np.abs(mu).max())
    assert np.max(np.abs(var[:, 0] - var_true.astype(float))) < 50 * 2.2e-16 * cond * kss


def test_hyp_roundtrip_and_shapes():
    # tests/test_models/test_gp_model.py:61-137 of the reference: gps layout, hyp round trip, hook shapes
    om, X = _model('Matern32', E=2, D=2, P=3)
    assert len(om.gps) == 2 and len(om.gps[0]) == 1
    for e in range(2):
        assert np.array_equal(om.gps[e][0].get_params(), om.hyp[e][0])
    M, S = om.predict(X)
    assert M.shape == (12, 2) and S.shape == (12, 2)
    assert om.d_mu_d_p(0, X).shape == (12, 3) and om.d2_mu_d_p2(1, X).shape == (12, 3, 3)
    assert om.d_s2_d_p(0, X).shape == (12, 3) and om.d2_s2_d_p2(1, X).shape == (12, 3, 3)


def test_kernel_shapes_like_reference():
    # tests/test_kernels/test_stationary.py:12-94: dK2_drdr keeps the shape, r = 0 entries do not crash
    rng = np.random.default_rng(0)
    d, N = 3, 7
    X = rng.uniform(size=(N, d))
    for name, cls in gpy_kernels.KERNELS.items():
        k = cls(d, range(d), 'k')
        R = rng.uniform(size=(4, 5))
        R[2, 2] = 0.0
        assert k.dK2_drdr(R).shape == R.shape
        assert k.gradients_X(np.ones((5, N)), X[:5], X).shape == (5, d)
        g = k.gradients_XX(np.ones((5, N)), X[:5], X)
        assert g.shape == (5, N, d, d) and np.all(np.isfinite(g))
