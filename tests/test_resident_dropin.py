"""taylor_multi + criteria on resident marginals: same numbers as the per-model drop-in functions and as the fused
scorer; residency is dropped (never trusted) as soon as an array is not the untouched read-only original."""
import numpy as np
import pytest

import helpers
from gpdoemd_b200 import synthetic

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('order', [1, 2])
def test_taylor_multi_equals_the_per_model_functions(order, gpu_ctx):
    import gpdoemd_b200 as g
    from gpdoemd_b200.marginal import resident_pair
    oms, ds = helpers.oracle_models(51, M=3, E=2, D=3, P=[2, 3, 2], n=120, kernel='Matern52', n_binary=1)
    gms = helpers.gpu_models(oms)
    lo = np.max([d['xlo'] for d in ds], axis=0)
    hi = np.min([d['xhi'] for d in ds], axis=0)
    X = synthetic.candidates(4, 700, 3, lo, hi, n_binary=1)
    f = g.taylor_first_order if order == 1 else g.taylor_second_order
    mu0 = np.empty((700, 3, 2))
    S0 = np.empty((700, 3, 2, 2))
    for i, m in enumerate(gms):
        mu0[:, i], S0[:, i] = f(m, X)
    mu, S = g.taylor_multi(gms, X, order=order)
    assert np.array_equal(mu, mu0) and np.array_equal(S, S0)           # identical kernels, identical arithmetic
    assert resident_pair(mu, S) is not None and not mu.flags.writeable
    with pytest.raises(ValueError):
        mu[0, 0, 0] = 1.0
    launches = gpu_ctx.launch_count()
    for c in ('HR', 'BH', 'BF', 'AW', 'JR'):
        a = getattr(g, c)(mu, S, 0.02, [0.2, 0.5, 0.3])               # resident path
        b = getattr(g, c)(mu0, S0, 0.02, [0.2, 0.5, 0.3])             # upload path
        assert np.array_equal(a, b), c
    assert gpu_ctx.launch_count() > launches
    # anything derived is an ordinary host array: correct results through the upload path
    assert resident_pair(mu[:350], S[:350]) is None and resident_pair(mu.copy(), S) is None
    half = g.BH(mu[:350], S[:350], 0.02)
    assert np.array_equal(half, g.BH(mu0[:350], S0[:350], 0.02))
    mod = np.array(S)
    mod[3, 1] *= 2.0
    assert not np.array_equal(g.BH(mu, mod, 0.02), g.BH(mu, S, 0.02))
    # and the fused scorer agrees
    out = g.score_designs_multi(gms, X, ['BH', 'JR'], order=order, noise_var=0.02)
    assert np.allclose(out['BH'][2], g.BH(mu, S, 0.02), rtol=1e-12, atol=0)
    assert out['JR'][0] == int(np.argmax(g.JR(mu, S, 0.02)))


def test_resident_marginals_survive_other_device_work_and_die_with_the_arrays(gpu_ctx):
    import gc
    import gpdoemd_b200 as g
    oms, ds = helpers.oracle_models(52, M=2, E=1, D=2, P=2, n=60, kernel='RBF')
    gms = helpers.gpu_models(oms)
    X = synthetic.candidates(5, 300, 2, np.max([d['xlo'] for d in ds], axis=0), np.min([d['xhi'] for d in ds], axis=0))
    mu, S = g.taylor_multi(gms, X)
    a = g.BH(mu, S, 0.01)
    gms[0].predict(X)                                   # scratch reused by another entry point in between
    g.score_designs(gms, X, criterion='HR')
    assert np.array_equal(g.BH(mu, S, 0.01), a)
    owner = mu._gpd_resident
    del mu, S
    gc.collect()
    assert owner.handle is not None                     # still referenced here
    del owner
    gc.collect()
    mu2, S2 = g.taylor_multi(gms, X[:0])                # empty candidate set
    assert mu2.shape == (0, 2, 1) and S2.shape == (0, 2, 1, 1)


@pytest.mark.parametrize('seed', range(12))
def test_randomised_configurations_through_the_round2_entry_points(seed, gpu_ctx):
    """tools/fuzz_new_paths.py (330 seeds run clean on a B200 in round 2), 12 of them as a regression test: taylor_multi
    and the resident criteria bit-identical with the per-model / upload paths, the sharded call with a one-rank
    communicator, the fused option, and gpd_marginal_dx against central differences of the oracle."""
    import os
    import sys
    import gpdoemd_b200 as g
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
    import fuzz_new_paths
    comm = g.LibraryComm(gpu_ctx, g.LibraryComm.unique_id(), 0, 1)
    try:
        info, fails = fuzz_new_paths.check_seed(seed, gpu_ctx, comm)
    except AttributeError:
        pytest.skip('binary combination without a GP (the reference raises too)')
    finally:
        comm.close()
    assert not fails, (info, fails)
