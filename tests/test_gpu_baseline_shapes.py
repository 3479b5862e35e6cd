"""The fused scorer against the oracle on the BASELINE.json model shapes themselves.

C2 (n=500 RBF, P=4, first order, HR/BH), C3 (n=4000 Matern52, second order, AW/BF), C4 (E=8, P=10, n=4000, all five
criteria: the 11-column first-order bucket; M cut from 8 to 2 so that the oracle finishes in about a minute), C5 (E=4,
n=4000, BH), plus P=10 SECOND order (72 / 55-column buckets, gram_kernel<11>) on a smaller training set.  The GPU models
are built exactly as bench.py builds them (device-side GP build for n >= 2000); the oracle shares their posterior
(alpha, L) for n >= 2000 -- `oracle.model.adopt_exact_model` -- and refactorises on its own at n = 500.  Candidates:
421 = 3 full tiles + a ragged one (C3: 259 = 2 + a ragged one), so tile boundaries and the padded tail are inside
the sample."""
import numpy as np
import pytest

import helpers
from gpdoemd_b200 import synthetic
from oracle import criteria as ocrit
from oracle import gpy_kernels
from oracle import marginal as omarg
from oracle.model import OracleGPModel, adopt_exact_model

pytestmark = pytest.mark.gpu
TOL_CRIT = 1e-7


def _problem(cfg, M, adopt):
    import gpdoemd_b200 as g
    gms, ds = synthetic.build_models(g.GPModel, 2, M, cfg['E'], cfg['D'], cfg['P'], cfg['n'], cfg['kernel'],
                                     build_on_device=cfg['n'] >= 2000)
    if adopt:
        oms = [adopt_exact_model(m) for m in gms]
        for om in oms:
            om.want_var_jac = cfg['order'] == 2
    else:
        oms, _ = synthetic.build_models(OracleGPModel, 2, M, cfg['E'], cfg['D'], cfg['P'], cfg['n'], cfg['kernel'],
                                        kernel_lookup=gpy_kernels.KERNELS)
    lo = np.max([d['xlo'] for d in ds], axis=0)
    hi = np.min([d['xhi'] for d in ds], axis=0)
    return gms, oms, lo, hi


def _check(cfg, gms, oms, X, noise=0.01):
    import gpdoemd_b200 as g
    N, M, E = len(X), len(gms), cfg['E']
    marg = omarg.taylor_first_order if cfg['order'] == 1 else omarg.taylor_second_order
    mu = np.zeros((N, M, E))
    S = np.zeros((N, M, E, E))
    for i, om in enumerate(oms):
        mu[:, i], S[:, i] = marg(om, X)
    out = g.score_designs_multi(gms, X, list(cfg['criteria']), order=cfg['order'], noise_var=noise)
    for c in cfg['criteria']:
        d0 = ocrit.CRITERIA[c](mu, S.copy(), noise, None)
        bi, bv, d1 = out[c]
        scale = max(np.abs(d0).mean(), 1.0 if c == 'JR' else 1e-300)
        err = helpers.relerr(d1, d0, scale)
        print('%s order %d M=%d E=%d P=%d n=%d: %s rel err %.2e' % (cfg['kernel'], cfg['order'], M, E, cfg['P'], cfg['n'], c, err))
        assert err < TOL_CRIT, (c, err)
        j = int(np.argmax(d0))
        assert bi == j or abs(d0[bi] - d0[j]) <= 1e-9 * scale, c
    # and the marginals underneath, at the north-star 1e-9 (first order) through the drop-in functions on one model
    f_g = g.taylor_first_order if cfg['order'] == 1 else g.taylor_second_order
    M1, S1 = f_g(gms[0], X[:130])
    ystd = oms[0].ystd
    assert helpers.relerr(M1, mu[:130, 0], ystd) < 1e-9
    sc = np.sqrt(np.einsum('nii->ni', S[:130, 0]))
    assert helpers.relerr(S1, S[:130, 0], sc[:, :, None] * sc[:, None, :]) < (1e-9 if cfg['order'] == 1 else 1e-8)


@pytest.mark.parametrize('name', ['C2', 'C3', 'C4', 'C5'])
def test_fused_scorer_on_the_baseline_shapes(name, gpu_ctx):
    cfg = dict(synthetic.CONFIGS[name])
    M = 2 if name == 'C4' else cfg['M']
    if cfg['n'] >= 2000:
        gpu_ctx.set_scratch_bytes(16 << 30)
    gms, oms, lo, hi = _problem(cfg, M, adopt=cfg['n'] >= 2000)
    X = synthetic.candidates(11, 259 if name == 'C3' else 421, cfg['D'], lo, hi)   # C3: the second-order oracle is slow
    try:
        _check(cfg, gms, oms, X)
    finally:
        for m in gms:
            m._invalidate_device()
        gpu_ctx.set_scratch_bytes(6 << 30)


@pytest.mark.parametrize('kernel', ['RBF', 'Matern52'])
def test_second_order_with_ten_parameters(kernel, gpu_ctx):
    """P = 10 second order: 1 + 10 + 55 = 66 coefficient columns (72-column bucket), 55-column beta contraction,
    gram_kernel<11>, with E = 8 outputs and every criterion."""
    cfg = dict(M=2, E=8, D=3, P=10, n=300, kernel=kernel, order=2, criteria=('HR', 'BH', 'BF', 'AW', 'JR'))
    gms, oms, lo, hi = _problem(cfg, 2, adopt=False)
    X = synthetic.candidates(12, 300, 3, lo, hi)
    _check(cfg, gms, oms, X)


def test_first_order_eight_outputs_ten_parameters_all_criteria(gpu_ctx):
    """The C4 shape at full M = 8 (64 GPs, E = 8, P = 10, 11-column bucket) on a small training set."""
    cfg = dict(M=8, E=8, D=3, P=10, n=200, kernel='Matern52', order=1, criteria=('HR', 'BH', 'BF', 'AW', 'JR'))
    gms, oms, lo, hi = _problem(cfg, 8, adopt=False)
    X = synthetic.candidates(13, 421, 3, lo, hi)
    _check(cfg, gms, oms, X)
