"""Option "fused" (off by default): the first-order path through ONE persistent kernel that evaluates the kernel rows
into a private L2-resident panel and contracts them in place (csrc/fused_impl.cuh).  Same numbers as the separate
evaluation + triangular kernels up to the summation order of the mean contractions; the tail-wave items are cut into
row-block parts, so sizes around one and several waves of the 148-CTA grid are covered.  Variant 2
(csrc/fusedc_impl.cuh) keeps tri_kernel's consumer structure and adds a producer warpgroup that evaluates the CTA's next
item concurrently."""
import numpy as np
import pytest

import helpers
from gpdoemd_b200 import synthetic
from oracle import criteria as ocrit
from oracle import marginal as omarg

pytestmark = pytest.mark.gpu

CASES = [  # M, E, D, P, n, kernel, n_binary, N
    (2, 2, 3, 4, 500, 'RBF', 0, 128 * 163 + 17),        # > 4 waves + a cut tail
    (3, 1, 2, 2, 130, 'Matern52', 0, 700),              # less than one wave: every item is cut into parts
    (2, 2, 4, 3, 300, 'Matern32', 1, 1000),             # binary routing
    (2, 3, 1, 1, 37, 'Exponential', 0, 5),              # one row block, nothing to cut
    (2, 2, 3, 10, 260, 'RatQuad', 0, 400),              # 11-column bucket
    (2, 1, 5, 6, 1000, 'RBF', 0, 3000),                 # padded design dimensions, 8 row blocks
]


@pytest.mark.parametrize('variant', [1, 2], ids=['serial', 'concurrent'])
@pytest.mark.parametrize('case', CASES, ids=lambda c: '%s-n%d-N%d' % (c[5], c[4], c[7]))
def test_fused_kernel_matches_the_separate_kernels_and_the_oracle(case, variant, gpu_ctx):
    import gpdoemd_b200 as g
    M, E, D, P, n, kernel, nb, N = case
    oms, ds = helpers.oracle_models(7, M=M, E=E, D=D, P=P, n=n, kernel=kernel, n_binary=nb, gp_noise=1e-4)
    gms = helpers.gpu_models(oms)
    lo = np.max([d['xlo'] for d in ds], axis=0)
    hi = np.min([d['xhi'] for d in ds], axis=0)
    X = synthetic.candidates(3, N, D, lo, hi, n_binary=nb)
    out = {}
    try:
        for flag in (0, 1):
            gpu_ctx.set_option('fused', flag * variant)
            out[flag] = g.score_designs_multi(gms, X, ['HR', 'BH', 'JR'], order=1, noise_var=0.02)
            out[flag]['marg'] = g.taylor_first_order(gms[0], X)
    finally:
        gpu_ctx.set_option('fused', 0)
    for c in ('HR', 'BH', 'JR'):
        a, b = out[0][c][2], out[1][c][2]
        scale = max(np.abs(a).mean(), 1.0 if c == 'JR' else 1e-300)
        assert helpers.relerr(b, a, scale) < 1e-9, c
        assert out[0][c][0] == out[1][c][0] or abs(a[out[0][c][0]] - a[out[1][c][0]]) <= 1e-9 * scale
    ystd = oms[0].ystd
    assert helpers.relerr(out[1]['marg'][0], out[0]['marg'][0], ystd) < 1e-11
    assert np.array_equal(out[1]['marg'][1].shape, out[0]['marg'][1].shape)
    # against the oracle on a sample (the fused path stands on its own, not only relative to the other CUDA path)
    sel = np.unique(np.r_[0:min(N, 140), max(0, N - 140):N])
    mu = np.zeros((len(sel), M, E))
    S = np.zeros((len(sel), M, E, E))
    for i, om in enumerate(oms):
        mu[:, i], S[:, i] = omarg.taylor_first_order(om, X[sel])
    d0 = ocrit.BH(mu, S, 0.02, None)
    assert helpers.relerr(out[1]['BH'][2][sel], d0, np.abs(d0).mean()) < 1e-7
