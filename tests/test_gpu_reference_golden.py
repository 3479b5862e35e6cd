"""The CUDA path DIRECTLY against golden vectors produced by the reference's own code (tests/golden/make_golden.py run
on /root/reference): all five criteria on the deterministic notebook input, the seeded random inputs, the planted
argmax-13 fixture, the BH/BF mutation quirk, and both Taylor marginals on the closed-form duck-typed model."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
import make_golden  # noqa: E402

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(HERE, 'golden', 'reference_golden.npz'))
CRITS = ('HR', 'BH', 'BF', 'AW', 'JR')
TOL = 1e-7          # north star: 1e-7 on criteria (measured ~1e-12)


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.abs(b).mean(), 1e-300))


@pytest.mark.parametrize('name', CRITS)
def test_notebook_input(name, gpu_ctx):
    import gpdoemd_b200 as g
    d = getattr(g, name)(G['nb_mu'].copy(), G['nb_s2'].copy(), G['nb_noise'].copy())
    assert d.shape == G['nb_' + name].shape and _rel(d, G['nb_' + name]) < TOL
    assert int(np.argmax(d)) == int(np.argmax(G['nb_' + name]))


@pytest.mark.parametrize('shape', [(3, 2), (4, 1), (2, 3), (8, 8), (5, 4)])
@pytest.mark.parametrize('name', CRITS)
def test_seeded_random_inputs(shape, name, gpu_ctx):
    import gpdoemd_b200 as g
    key = 'r%d_%d_' % shape
    mu, s2, noise, pps = G[key + 'mu'], G[key + 's2'], G[key + 'noise'], G[key + 'pps']
    s2c = s2.copy()
    d = getattr(g, name)(mu.copy(), s2c, noise.copy(), pps.copy())
    assert np.array_equal(s2c, s2)                                   # inputs are never modified by default
    assert _rel(d, G[key + name]) < TOL and int(np.argmax(d)) == int(np.argmax(G[key + name]))


@pytest.mark.parametrize('name', CRITS)
def test_planted_argmax_13(name, gpu_ctx):
    import gpdoemd_b200 as g
    d = getattr(g, name)(G['t13_mu'].copy(), G['t13_s2'].copy(), 0.1 * np.eye(2), np.ones(3) / 3)
    assert d.shape == (15,) and int(np.argmax(d)) == 13 and _rel(d, G['t13_' + name]) < TOL


def test_mutation_quirk_on_request(gpu_ctx):
    import gpdoemd_b200 as g
    s2m = G['r3_2_s2'].copy()
    g.BH(G['r3_2_mu'].copy(), s2m, G['r3_2_noise'].copy(), G['r3_2_pps'].copy(), mutate_like_reference=True)
    np.testing.assert_allclose(s2m, G['quirk_BH_s2_after'], rtol=1e-15)


@pytest.mark.parametrize('EP', [(1, 1), (2, 3), (3, 5)])
def test_marginals_of_a_duck_typed_model(EP, gpu_ctx):
    import gpdoemd_b200 as g
    E, P = EP
    st = make_golden.StubModel(E, P, 7 * E + P)
    key = 'ty%d_%d_' % (E, P)
    X = G[key + 'X']
    M1, S1 = g.taylor_first_order(st, X)
    M2, S2 = g.taylor_second_order(st, X)
    for a, b in ((M1, G[key + 'M1']), (S1, G[key + 'S1']), (M2, G[key + 'M2']), (S2, G[key + 'S2'])):
        assert a.shape == b.shape
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12 * np.abs(b).max())
