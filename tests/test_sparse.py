"""SparseGPModel (GPdoemd/models/sparse_gp_model.py:35-74), the first "next" row of SURVEY 8f.

CPU: the oracle's VarDTC restatement with inducing inputs = all training inputs reproduces the exact GP;
the product's host setup reproduces the oracle's posterior; the triangular root T of woodbury_inv is exact.
GPU: every hook, both marginals and the fused scoring of sparse models against the oracle."""
import numpy as np
import pytest

import helpers
from gpdoemd_b200 import synthetic
from gpdoemd_b200.gp_state import GPState, lower_root_of
from oracle import criteria as ocrit
from oracle import gpy_kernels
from oracle import marginal as omarg
from oracle.model import OracleGPModel, OracleSparseGPModel


def test_oracle_sparse_with_all_inducing_equals_exact_gp():
    d = synthetic.model_data(5, 2, 3, 2, 80, gp_noise=1e-4)

    def mk(cls, **kw):
        m = cls({'name': 'm', 'dim_x': 3, 'dim_p': 2, 'num_outputs': 2})
        m.gp_surrogate(Z=d['Z'], Y=d['Y'], kern_x=gpy_kernels.Matern52, kern_p=gpy_kernels.Matern52, **kw)
        m.pmean, m.Sigma = d['pmean'], d['Sigma']
        return m
    ex = mk(OracleGPModel)
    ex.hyp = d['hyp']
    ex.gp_load_hyp()
    sp = mk(OracleSparseGPModel, num_inducing=80)
    sp.hyp = [[np.r_[ex.Z.ravel(), h] for h in hs] for hs in d['hyp']]
    sp.gp_load_hyp()
    X = synthetic.candidates(1, 25, 3, d['xlo'], d['xhi'])
    M0, S0 = ex.predict(X)
    M1, S1 = sp.predict(X)
    assert np.max(np.abs(M0 - M1)) < 1e-6 * ex.ystd.max()
    assert np.max(np.abs(S0 - S1)) < 1e-6 * S0.max()      # the 1e-8 Kmm jitter vs the exact model
    for e in range(2):
        assert np.allclose(ex.d_mu_d_p(e, X), sp.d_mu_d_p(e, X), rtol=0, atol=1e-5 * np.abs(ex.d_mu_d_p(e, X)).max())
        assert np.allclose(ex.d_s2_d_p(e, X), sp.d_s2_d_p(e, X), rtol=0, atol=1e-6 * np.abs(ex.d_s2_d_p(e, X)).max())
        assert np.allclose(ex.d2_s2_d_p2(e, X), sp.d2_s2_d_p2(e, X), rtol=0, atol=1e-6 * np.abs(ex.d2_s2_d_p2(e, X)).max())


def test_sparse_param_vector_layout_and_shapes():
    # tests/test_models/test_sparse_gp_model.py:61-92 of the reference: gps layout, hyp round trip, predict shapes
    oms, ds = helpers.oracle_sparse_models(3, M=1, E=2, D=2, P=2, n=14, kernel='RBF', num_inducing=10)
    om = oms[0]
    assert len(om.gps) == 2 and len(om.gps[0]) == 1 and om.gps[0][0].num_inducing == 10
    for e in range(2):
        v = om.gps[e][0].get_params()
        assert len(v) == 10 * 4 + (1 + 2) + (1 + 2) + 1 and np.array_equal(v, om.hyp[e][0])
    X = synthetic.candidates(1, 10, 2, ds[0]['xlo'], ds[0]['xhi'])
    M, S = om.predict(X)
    assert M.shape == (10, 2) and S.shape == (10, 2) and np.all(S > 0)


def test_host_sparse_setup_matches_oracle_and_root_is_exact():
    import gpdoemd_b200 as g
    oms, _ = helpers.oracle_sparse_models(4, M=1, E=1, D=3, P=2, n=120, kernel='Matern52', num_inducing=40)
    gms, _ = synthetic.build_models(g.SparseGPModel, 4, 1, 1, 3, 2, 120, 'Matern52', gp_noise=1e-4, num_inducing=40)
    go, gg = oms[0].gps[0][0], gms[0].gps[0][0]
    assert np.array_equal(go.Z, gg.Z)
    np.testing.assert_allclose(gg.posterior.woodbury_inv, go.posterior.woodbury_inv, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(gg.posterior.woodbury_vector, go.posterior.woodbury_vector, rtol=1e-9, atol=1e-9)
    st = GPState.from_gp(go, 3, 2)
    assert st.factor_kind == 1 and st.L.shape == (40, 40) and np.all(np.triu(st.L, 1) == 0)
    W = go.posterior.woodbury_inv
    assert np.max(np.abs(st.L.T.dot(st.L) - W)) < 1e-12 * np.abs(W).max()
    assert np.array_equal(st.Z, go.Z)
    # singular (rank-deficient) woodbury_inv still gets a root
    A = np.random.default_rng(0).normal(size=(8, 3))
    T = lower_root_of(A.dot(A.T))
    assert np.max(np.abs(T.T.dot(T) - A.dot(A.T))) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize('kernel,nb', [('RBF', 0), ('Matern52', 1)])
def test_sparse_hooks_on_gpu(kernel, nb, gpu_ctx):
    import gpdoemd_b200 as g
    oms, ds = helpers.oracle_sparse_models(7, M=1, E=2, D=3, P=3, n=400, kernel=kernel, num_inducing=150, n_binary=nb)
    om = oms[0]
    gm = g.SparseGPModel.from_reference(om)
    X = synthetic.candidates(3, 300, 3, ds[0]['xlo'], ds[0]['xhi'], n_binary=nb)
    gp0 = om.gps[0][0]
    kss = gp0.kern.kernx.variance * gp0.kern.kernp.variance
    ystd, vs, pd = om.ystd, kss * om.ystd ** 2, om.pmax - om.pmin
    M0, S0 = om.predict(X)
    M1, S1 = gm.predict(X)
    assert helpers.relerr(M1, M0, ystd) < 1e-9
    assert helpers.relerr(S1, S0, vs) < 1e-9
    for e in range(2):
        assert helpers.relerr(gm.d_mu_d_p(e, X), om.d_mu_d_p(e, X), ystd[e] / pd) < 1e-9
        assert helpers.relerr(gm.d2_mu_d_p2(e, X), om.d2_mu_d_p2(e, X), ystd[e] / np.outer(pd, pd)) < 1e-9
        assert helpers.relerr(gm.d_s2_d_p(e, X), om.d_s2_d_p(e, X), vs[e] / pd) < 1e-9
        assert helpers.relerr(gm.d2_s2_d_p2(e, X), om.d2_s2_d_p2(e, X), vs[e] / np.outer(pd, pd)) < 1e-8


@pytest.mark.gpu
def test_sparse_marginals_and_fused_scoring(gpu_ctx):
    import gpdoemd_b200 as g
    oms, ds = helpers.oracle_sparse_models(9, M=3, E=2, D=3, P=2, n=300, kernel='Matern32', num_inducing=100)
    gms = [g.SparseGPModel.from_reference(m) for m in oms]
    lo = np.max([d['xlo'] for d in ds], axis=0)
    hi = np.min([d['xhi'] for d in ds], axis=0)
    N = 500
    X = synthetic.candidates(2, N, 3, lo, hi)
    for order, fo, fg in ((1, omarg.taylor_first_order, g.taylor_first_order),
                          (2, omarg.taylor_second_order, g.taylor_second_order)):
        mu = np.zeros((N, 3, 2))
        S = np.zeros((N, 3, 2, 2))
        for i, om in enumerate(oms):
            mu[:, i], S[:, i] = fo(om, X)
            Mg, Sg = fg(gms[i], X)
            assert helpers.relerr(Mg, mu[:, i], om.ystd) < 1e-9
            sc = np.sqrt(np.einsum('nii->ni', S[:, i]))
            assert helpers.relerr(Sg, S[:, i], sc[:, :, None] * sc[:, None, :]) < 1e-8
        res = g.score_designs_multi(gms, X, ['BH', 'JR'], order=order, noise_var=0.01)
        for name in ('BH', 'JR'):
            d0 = ocrit.CRITERIA[name](mu, S, 0.01, None)
            assert helpers.relerr(res[name][2], d0, np.abs(d0).mean()) < 1e-7
            assert res[name][0] == int(np.argmax(d0))


@pytest.mark.gpu
def test_sparse_model_built_by_the_product(gpu_ctx):
    """SparseGPModel.gp_surrogate + hyp + gp_load_hyp without any oracle object, against the oracle's numbers."""
    import gpdoemd_b200 as g
    oms, ds = helpers.oracle_sparse_models(11, M=1, E=1, D=2, P=2, n=200, kernel='RBF', num_inducing=64)
    gms, _ = synthetic.build_models(g.SparseGPModel, 11, 1, 1, 2, 2, 200, 'RBF', gp_noise=1e-4, num_inducing=64)
    X = synthetic.candidates(5, 100, 2, ds[0]['xlo'], ds[0]['xhi'])
    M0, S0 = oms[0].predict(X)
    M1, S1 = gms[0].predict(X)
    gp0 = oms[0].gps[0][0]
    vs = gp0.kern.kernx.variance * gp0.kern.kernp.variance * oms[0].ystd ** 2
    assert helpers.relerr(M1, M0, oms[0].ystd) < 1e-8
    assert helpers.relerr(S1, S0, vs) < 1e-8          # against the prior variance: sigma^2 itself is ~1e-5 of it here
    assert gms[0].max_num_inducing == 500


@pytest.mark.gpu
def test_sparse_variance_floor_like_gpys_base_posterior(gpu_ctx):
    """GPy's base Posterior._raw_predict (SparseGPRegression) clips the predictive variance at 1e-15 in
    standardised units, PosteriorExact (exact GPs) does not.  The device applies the same floor to GPD_FACTOR_ROOT
    models only.  A sparse posterior whose woodbury_inv overshoots (k** - k^T W k < 0 at the inducing inputs) shows it;
    an exact model with the same numbers keeps the negative value."""
    import gpdoemd_b200 as g
    oms, ds = helpers.oracle_sparse_models(13, M=1, E=1, D=2, P=2, n=120, kernel='RBF', num_inducing=40)
    om = oms[0]
    gp = om.gps[0][0]
    gp.posterior.woodbury_inv = gp.posterior.woodbury_inv * (1.0 + 1e-3)      # overshoot: variances go slightly negative
    gm = g.SparseGPModel.from_reference(om)
    Zt = gp._predictive_variable                                                # the inducing inputs, transformed units
    om.pmean = om.trans_p(Zt[7, 2:], back=True)
    gm.pmean = om.pmean.copy()
    X = om.trans_x(Zt[:30, :2], back=True)
    M0, S0 = om.predict(X)
    M1, S1 = gm.predict(X)
    floor = 1e-15 * om.ystd ** 2
    assert np.any(S0 == floor) and np.all(S0 >= floor)          # the oracle (GPy's clip) bites on this input
    assert np.all(S1 >= floor)
    vs = gp.kern.kernx.variance * gp.kern.kernp.variance * om.ystd ** 2
    assert helpers.relerr(S1, S0, vs) < 1e-8 and np.array_equal(S1 == floor, S0 == floor)
