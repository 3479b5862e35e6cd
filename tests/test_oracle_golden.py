"""The CPU oracle against golden vectors produced by the reference's own code
(tests/golden/make_golden.py run against /root/reference in the build container) and against the
reference's golden routing vectors (tests/test_utils.py:16,41,66)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
import make_golden  # noqa: E402  (StubModel + notebook inputs; its main() is not run)

from oracle import criteria as ocrit  # noqa: E402
from oracle import marginal as omarg  # noqa: E402
from oracle import routing, transform  # noqa: E402

G = np.load(os.path.join(HERE, 'golden', 'reference_golden.npz'))
CRITS = ('HR', 'BH', 'BF', 'AW', 'JR')


@pytest.mark.parametrize('name', CRITS)
def test_criteria_notebook_input(name):
    M, S, mv = make_golden.notebook_inputs()
    assert np.array_equal(M, G['nb_mu']) and np.array_equal(S, G['nb_s2'])
    d = ocrit.CRITERIA[name](M, S, mv)
    np.testing.assert_allclose(d, G['nb_' + name], rtol=1e-12, atol=0)


@pytest.mark.parametrize('shape', [(3, 2), (4, 1), (2, 3), (8, 8), (5, 4)])
@pytest.mark.parametrize('name', CRITS)
def test_criteria_random_inputs(shape, name):
    key = 'r%d_%d_' % shape
    mu, s2, noise, pps = G[key + 'mu'], G[key + 's2'], G[key + 'noise'], G[key + 'pps']
    s2c = s2.copy()
    d = ocrit.CRITERIA[name](mu, s2c, noise, pps)
    assert np.array_equal(s2c, s2)                      # the oracle never mutates (reference BH/BF do)
    np.testing.assert_allclose(d, G[key + name], rtol=1e-11, atol=0)
    assert int(np.argmax(d)) == int(np.argmax(G[key + name]))


@pytest.mark.parametrize('name', CRITS)
def test_criteria_planted_argmax_13(name):
    d = ocrit.CRITERIA[name](G['t13_mu'], G['t13_s2'].copy(), 0.1 * np.eye(2), np.ones(3) / 3)
    assert d.shape == (15,) and int(np.argmax(d)) == 13
    np.testing.assert_allclose(d, G['t13_' + name], rtol=1e-11)


def test_reference_mutation_quirk_is_documented():
    # reference BH leaves s2 + noise in the caller's array (design_criteria.py:81)
    np.testing.assert_allclose(G['quirk_BH_s2_after'], G['r3_2_s2'] + G['r3_2_noise'], rtol=1e-15)


def test_reshape_conventions():
    m = np.random.default_rng(0).normal(size=(15, 3))
    s = np.random.default_rng(1).uniform(0.1, 1, size=(15, 3))
    R = ocrit._reshape(m, s)
    assert R[0].shape == (15, 3, 1) and R[1].shape == (15, 3, 1, 1)
    assert R[2].shape == (1, 1) and abs(R[3].sum() - 1) < 1e-15
    R = ocrit._reshape(m, s, 0.1, [1, 1, 2])
    assert R[2][0, 0] == 0.1 and np.allclose(R[3], [0.25, 0.25, 0.5])
    with pytest.raises(AssertionError):
        ocrit._reshape(m, s, np.zeros((2, 2)))
    with pytest.raises(AssertionError):
        ocrit._reshape(m, s, 0.1, [1, -1, 1])


@pytest.mark.parametrize('EP', [(1, 1), (2, 3), (3, 5)])
def test_marginals_stub_model(EP):
    E, P = EP
    st = make_golden.StubModel(E, P, 7 * E + P)
    key = 'ty%d_%d_' % (E, P)
    X = G[key + 'X']
    M1, S1 = omarg.taylor_first_order(st, X)
    np.testing.assert_allclose(M1, G[key + 'M1'], rtol=1e-13)
    np.testing.assert_allclose(S1, G[key + 'S1'], rtol=1e-12, atol=1e-16)
    M2, S2 = omarg.taylor_second_order(st, X)
    np.testing.assert_allclose(M2, G[key + 'M2'], rtol=1e-12)
    np.testing.assert_allclose(S2, G[key + 'S2'], rtol=1e-12, atol=1e-16)
    assert S1.shape == (30, E, E) and np.all(np.einsum('nii->ni', S1) >= 1e-15)


def test_transforms():
    bt = transform.BoxTransform(G['tr_xmin'], G['tr_xmax'])
    X = G['tr_X']
    assert np.array_equal(bt(X), G['tr_box']) and np.array_equal(bt(X, back=True), G['tr_box_back'])
    assert np.array_equal(bt.var(X), G['tr_box_var'])
    assert np.array_equal(bt.cov(np.einsum('ni,nj->nij', X, X)), G['tr_box_cov'])
    mt = transform.MeanTransform(G['tr_xmin'], G['tr_xmax'])
    assert np.array_equal(mt(X), G['tr_mean']) and np.array_equal(mt(X, back=True), G['tr_mean_back'])
    pm, ps = mt.prediction(X, X ** 2, back=True)
    assert np.array_equal(pm, G['tr_pred_m']) and np.array_equal(ps, G['tr_pred_s'])
    np.testing.assert_allclose(bt(bt(X), back=True), X, rtol=1e-14)


def test_routing_reference_golden_vectors():
    # tests/test_utils.py:9-71 of the reference
    Z = np.c_[np.arange(3), np.ones(3), np.array([1, -1, 1])]
    r, j = routing.binary_dimensions(Z, [2])
    assert list(r) == [0, 1] and list(j) == [1, 0, 1]
    J = [0, 1, 3, 2, 2, 1, 0, 3, 3, 2, 0]
    for lo in (0, -1):
        B = np.array([[lo, lo], [1, lo], [1, 1], [lo, 1], [lo, 1], [1, lo], [lo, lo], [1, 1], [1, 1], [lo, 1], [lo, lo]])
        Z = np.c_[np.arange(11), B, np.ones(11)]
        r, j = routing.binary_dimensions(Z, [1, 2])
        assert list(r) == [0, 1, 2, 3] and list(j) == J
    r, j = routing.binary_dimensions(np.zeros((4, 3)), [])
    assert list(r) == [0] and np.all(j == 0)


@pytest.mark.parametrize('nb', [1, 2, 3])
def test_routing_vs_reference_run(nb):
    Z, J = G['route%d_Z' % nb], G['route%d_J' % nb]
    r, j = routing.binary_dimensions(Z, list(range(2, 2 + nb)))
    assert np.array_equal(j, J)
    # the product's vectorised host routing and the device weight table agree with the reference too
    from gpdoemd_b200.models import binary_dimensions
    r2, j2 = binary_dimensions(Z, list(range(2, 2 + nb)))
    assert list(r2) == list(range(2 ** nb)) and np.array_equal(j2, J)


@pytest.mark.parametrize('EP', [(1, 1), (2, 3), (3, 5)])
def test_laplace_approximation_vs_reference(EP):
    """gpdoemd_b200.laplace_approximation == GPdoemd/param_covar/laplace_approximation.py on a duck-typed model."""
    from gpdoemd_b200 import laplace_approximation
    E, P = EP
    st = make_golden.StubModel(E, P, 7 * E + P)
    X = G['la%d_%d_X' % (E, P)]
    st.meas_noise_var = 0.01 * (1 + np.arange(E))
    np.testing.assert_allclose(laplace_approximation(st, X), G['la%d_%d_diag' % (E, P)], rtol=1e-12)
    B = np.random.default_rng(8).normal(size=(E, E))
    st.meas_noise_var = 0.01 * np.eye(E) + 0.002 * B.dot(B.T)
    np.testing.assert_allclose(laplace_approximation(st, X), G['la%d_%d_full' % (E, P)], rtol=1e-11)
