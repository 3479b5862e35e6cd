"""A memcheck of our own for the scratch region (compute-sanitizer is closed on the GPU pool): with the debug option
"scratch_redzones" every sub-buffer carved out of the per-call scratch region -- right-hand-side panels, raw
contractions, hook outputs, criterion workspace, work-item lists, private panels of the fused kernel -- is followed by
a 256-byte red zone; after each call none of those bytes may have changed.  The matrix covers ragged sizes, every entry
point, both Taylor orders, the fused option, and the n = 4000 / P = 10 / E = 8 buckets of the BASELINE configs."""
import numpy as np
import pytest

import helpers
from gpdoemd_b200 import synthetic

pytestmark = pytest.mark.gpu


@pytest.fixture()
def redzones(gpu_ctx):
    gpu_ctx.set_option('scratch_redzones', 1)
    yield gpu_ctx
    gpu_ctx.set_option('scratch_redzones', 0)


def _clean(ctx, what):
    zones, bad = ctx.check_scratch()
    assert zones > 0, 'red zones are not active for ' + what
    assert bad == 0, '%s overwrote %d red-zone bytes' % (what, bad)


def test_the_detector_sees_a_damaged_byte(redzones):
    import gpdoemd_b200 as g
    oms, ds = helpers.oracle_models(3, M=1, E=1, D=2, P=2, n=40, kernel='RBF')
    gm = helpers.gpu_models(oms)[0]
    gm.predict(synthetic.candidates(1, 50, 2, ds[0]['xlo'], ds[0]['xhi']))
    _clean(redzones, 'predict')
    redzones.set_option('scratch_redzones', 2)           # self-test: damage one byte, then count
    assert redzones.check_scratch()[1] == 1


@pytest.mark.parametrize('case', [  # M, E, D, P, n, kernel, n_binary, N
    (2, 2, 3, 4, 500, 'RBF', 0, 128 * 9 + 17), (3, 1, 2, 2, 130, 'Matern52', 0, 1), (2, 2, 4, 3, 300, 'Matern32', 1, 129),
    (2, 3, 1, 1, 37, 'Exponential', 0, 257), (2, 8, 3, 10, 260, 'RatQuad', 0, 400), (2, 1, 5, 6, 1000, 'RBF', 0, 333),
    (1, 2, 8, 16, 127, 'Matern52', 2, 700), (3, 2, 3, 2, 2, 'RBF', 0, 5)], ids=lambda c: '%s-n%d-N%d' % (c[5], c[4], c[7]))
def test_every_entry_point_stays_inside_its_buffers(case, redzones):
    import gpdoemd_b200 as g
    M, E, D, P, n, kernel, nb, N = case
    oms, ds = helpers.oracle_models(17, M=M, E=E, D=D, P=P, n=n, kernel=kernel, n_binary=nb, gp_noise=1e-4)
    gms = helpers.gpu_models(oms)
    lo = np.max([d['xlo'] for d in ds], axis=0)
    hi = np.min([d['xhi'] for d in ds], axis=0)
    X = synthetic.candidates(3, N, D, lo, hi, n_binary=nb)
    ctx = redzones
    try:
        gms[0].predict(X)
    except AttributeError:
        pytest.skip('a binary combination without a GP (the reference raises too)')
    _clean(ctx, 'predict')
    gms[0].d_mu_d_p(0, X); _clean(ctx, 'd_mu_d_p')
    second = P <= 10
    if second:
        gms[0].d2_s2_d_p2(0, X); _clean(ctx, 'all hooks')
    mu1, S1 = g.taylor_first_order(gms[0], X); _clean(ctx, 'taylor_first_order')
    if second:
        g.taylor_second_order(gms[0], X); _clean(ctx, 'taylor_second_order')
    for order in ((1, 2) if second else (1,)):
        mu, S = g.taylor_multi(gms, X, order=order); _clean(ctx, 'taylor_multi')
        g.JR(mu, S, 0.02); _clean(ctx, 'resident criterion')
        g.BF(np.array(mu), np.array(S), 0.02); _clean(ctx, 'uploaded criterion')
        for fused in (0, 1):
            ctx.set_option('fused', fused)
            try:
                g.score_designs_multi(gms, X, ['HR', 'BH', 'BF', 'AW', 'JR'], order=order, noise_var=0.02)
            finally:
                ctx.set_option('fused', 0)
            _clean(ctx, 'score_designs_multi order %d fused %d' % (order, fused))
    g.taylor_first_order_dx(gms[0], X); _clean(ctx, 'taylor_first_order_dx')
    assert ctx.argmax(np.arange(5000.0))[1] == 4999       # (one plain buffer, no sub-allocation: no zones to check)


@pytest.mark.parametrize('order,P,E', [(1, 10, 8), (2, 4, 2)])
def test_large_training_set_buckets_stay_inside_their_buffers(order, P, E, redzones):
    """n = 4000 (32 row blocks, 96 padding rows), the 11-column bucket with E = 8 (C4) and second order (C3)."""
    import gpdoemd_b200 as g
    gms, ds = synthetic.build_models(g.GPModel, 2, 1, E, 3, P, 4000, 'Matern52', build_on_device=True)
    lo, hi = ds[0]['xlo'], ds[0]['xhi']
    X = synthetic.candidates(3, 300, 3, lo, hi)
    ctx = redzones
    ctx.set_scratch_bytes(8 << 30)
    try:
        g.score_designs_multi(gms, X, ['HR', 'BH', 'JR'], order=order, noise_var=0.02); _clean(ctx, 'score order %d' % order)
        if order == 1:
            ctx.set_option('fused', 1)
            g.score_designs_multi(gms, X, ['HR', 'BH', 'JR'], order=1, noise_var=0.02); _clean(ctx, 'fused score')
            ctx.set_option('fused', 0)
            g.taylor_first_order_dx(gms[0], X[:64]); _clean(ctx, 'marginal_dx')
        else:
            gms[0].d2_s2_d_p2(0, X); _clean(ctx, 'second variance derivative')
    finally:
        ctx.set_option('fused', 0)
        for m in gms:
            m._invalidate_device()
        ctx.set_scratch_bytes(6 << 30)
