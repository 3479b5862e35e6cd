"""Committed GP golden vectors (tests/golden/gp_golden.npz, written by the oracle):
 * CPU: the oracle still reproduces them (drift guard);
 * GPU (-m gpu): the CUDA path, built through the product's own host setup (GPModel.gp_surrogate, no
   oracle object involved) and called through the ORIGINAL-unit C-ABI entry points, matches them."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
import make_gp_golden as mg  # noqa: E402

import helpers  # noqa: E402

G = np.load(os.path.join(HERE, 'golden', 'gp_golden.npz'))
CRITS = ('HR', 'BH', 'BF', 'AW', 'JR')


@pytest.mark.parametrize('case', list(mg.CASES))
def test_oracle_reproduces_gp_golden(case):
    oms, X = mg.build(case)
    for i, om in enumerate(oms):
        pm, ps = om.predict(X)
        np.testing.assert_allclose(pm, G['%s_m%d_mu' % (case, i)], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(ps, G['%s_m%d_s2' % (case, i)], rtol=1e-7, atol=1e-12)
        np.testing.assert_allclose(om.d_mu_d_p(0, X), G['%s_m%d_dmu' % (case, i)][:, 0], rtol=1e-8, atol=1e-11)


def _gpu_models(case):
    import gpdoemd_b200 as g
    seed, M, E, D, P, n, kernel, nb, N = mg.CASES[case]
    gms, ds = g.synthetic.build_models(g.GPModel, seed, M, E, D, P, n, kernel, n_binary=nb)
    lo = np.max([d['xlo'] for d in ds], axis=0)
    hi = np.min([d['xhi'] for d in ds], axis=0)
    X = g.synthetic.candidates(seed + 1, N, D, lo, hi, n_binary=nb)
    return gms, X


@pytest.mark.gpu
@pytest.mark.parametrize('case', list(mg.CASES))
def test_c_abi_original_units_vs_golden(case, gpu_ctx):
    """gpd_predict / gpd_dmu_dp / gpd_d2mu_dp2 / gpd_ds2_dp / gpd_d2s2_dp2 / gpd_marginal (original units)."""
    from gpdoemd_b200._lib import check, dptr
    gms, X = _gpu_models(case)
    lib = gpu_ctx.lib
    N = len(X)
    for i, gm in enumerate(gms):
        h = gm.device_handle()
        E, P = gm.num_outputs, gm.dim_p
        g0 = gm.gps[0][0]
        kss = g0.kern.kernx.variance * g0.kern.kernp.variance
        ystd, pd = gm.ystd, gm.pmax - gm.pmin
        vs = kss * ystd ** 2
        mu, s2 = np.empty((N, E)), np.empty((N, E))
        check(lib.gpd_predict(h, dptr(X), N, dptr(mu), dptr(s2)))
        assert helpers.relerr(mu, G['%s_m%d_mu' % (case, i)], ystd) < 1e-9
        assert helpers.relerr(s2, G['%s_m%d_s2' % (case, i)], vs) < 1e-9
        dmu = np.empty((N, E, P))
        check(lib.gpd_dmu_dp(h, dptr(X), N, dptr(dmu)))
        assert helpers.relerr(dmu, G['%s_m%d_dmu' % (case, i)], ystd[None, :, None] / pd) < 1e-9
        d2mu = np.empty((N, E, P, P))
        check(lib.gpd_d2mu_dp2(h, dptr(X), N, dptr(d2mu)))
        assert helpers.relerr(d2mu, G['%s_m%d_d2mu' % (case, i)], ystd[None, :, None, None] / np.outer(pd, pd)) < 1e-9
        ds2 = np.empty((N, E, P))
        check(lib.gpd_ds2_dp(h, dptr(X), N, dptr(ds2)))
        assert helpers.relerr(ds2, G['%s_m%d_ds2' % (case, i)], vs[None, :, None] / pd) < 1e-9
        d2s2 = np.empty((N, E, P, P))
        check(lib.gpd_d2s2_dp2(h, dptr(X), N, dptr(d2s2)))
        assert helpers.relerr(d2s2, G['%s_m%d_d2s2' % (case, i)], vs[None, :, None, None] / np.outer(pd, pd)) < 1e-8
        for order, km, ks in ((1, '_mu1', '_S1'), (2, '_mu2', '_S2')):
            M_, S_ = np.empty((N, E)), np.empty((N, E, E))
            check(lib.gpd_marginal(h, dptr(X), N, order, dptr(M_), dptr(S_)))
            Mg, Sg = G[case + km][:, i], G[case + ks][:, i]
            assert helpers.relerr(M_, Mg, ystd) < 1e-9
            sc = np.sqrt(np.einsum('nii->ni', Sg))
            assert helpers.relerr(S_, Sg, sc[:, :, None] * sc[:, None, :]) < (1e-9 if order == 1 else 1e-8)


@pytest.mark.gpu
@pytest.mark.parametrize('case', list(mg.CASES))
def test_fused_scoring_vs_golden(case, gpu_ctx):
    import gpdoemd_b200 as g
    gms, X = _gpu_models(case)
    E = gms[0].num_outputs
    noise = 0.01 * (1 + np.arange(E))
    for order in (1, 2):
        res = g.score_designs_multi(gms, X, list(CRITS), order=order, noise_var=noise)
        for name in CRITS:
            d0 = G['%s_dc%d_%s' % (case, order, name)]
            bi, bv, d1 = res[name]
            assert helpers.relerr(d1, d0, np.abs(d0).mean()) < 1e-7, (case, order, name)
            assert bi == int(np.argmax(d0)) and bv == d1[bi]
        # single-criterion entry point agrees bit for bit with the multi-criterion pass
        bi1, bv1, d11 = g.score_designs(gms, X, order=order, criterion='JR', noise_var=noise)
        assert np.array_equal(d11, res['JR'][2]) and bi1 == res['JR'][0]


@pytest.mark.gpu
def test_empty_and_error_paths(gpu_ctx):
    import gpdoemd_b200 as g
    from gpdoemd_b200._lib import GpdError, check, dptr
    gms, X = _gpu_models('m32')
    gm = gms[0]
    M0, S0 = gm.predict(X[:0])
    assert M0.shape == (0, 1) and S0.shape == (0, 1)
    lib = gpu_ctx.lib
    h = gm.device_handle()
    out = np.empty((4, 1))
    assert lib.gpd_predict(h, dptr(X[:4].copy()), -1, dptr(out), dptr(out)) != 0
    assert b'negative' in lib.gpd_last_error()
    with pytest.raises(GpdError):
        check(lib.gpd_marginal(h, dptr(X[:4].copy()), 4, 3, dptr(out), dptr(out)))        # order 3
    with pytest.raises(GpdError):
        g.score_designs(gms, X, criterion='BH', order=5)
    gm2 = g.GPModel.from_reference(gm)
    gm2.Sigma = None
    with pytest.raises(AssertionError):
        g.taylor_first_order(gm2, X)                                                      # Sigma missing
