"""Second mean derivative d2 mu / dp2 with p* close to a training parameter, against an 80-bit truth.

VERDICT r01: two fuzz configurations of the Exponential kernel missed the 1e-9 gate against the oracle (6e-9, and 1e-6
with a device-built factor).  kappa = exp(-r) has a cusp at r = 0, so its second p-derivative carries 1/r and 1/r^3
terms; GPy (restated by the oracle) computes r^2 in the EXPANDED form |a|^2 + |b|^2 - 2 a.b, whose absolute error u |a|^2
becomes a relative error u / r^2 of r.  These tests decide which side is off:

* CPU (`not gpu`): the oracle's error against the truth grows like u / r^2 for Exponential and stays at rounding level
  for Matern32 / Matern52 (their 1/r terms cancel analytically) -- a documented divergence OF THE ORACLE (= of GPy);
* GPU: the CUDA path (difference-form distances) holds 1e-9 against the truth for every kernel and distance, with an
  uploaded factor and with a device-built one (the truth then uses the device's own alpha: what remains between a
  device-built and a host-built model is the conditioning of alpha = K^-1 y, bounded separately)."""
import numpy as np
import pytest

import helpers
from gpdoemd_b200 import synthetic

KERNELS = ['Exponential', 'Matern32', 'Matern52']
DELTAS = [1e-3, 1e-5, 1e-7]


def _problem(kernel, delta):
    oms, ds = helpers.oracle_models(73, M=1, E=1, D=2, P=3, n=150, kernel=kernel, gp_noise=1e-4)
    om = oms[0]
    gp = om.gps[0][0]
    pst = gp._predictive_variable[17, 2:] + delta * np.array([1.0, -0.5, 0.3])     # p* next to training row 17
    om.pmean = om.trans_p(pst, back=True)
    X = synthetic.candidates(2, 60, 2, ds[0]['xlo'], ds[0]['xhi'])
    return om, om.trans_x(X)


@pytest.mark.parametrize('kernel', KERNELS)
def test_oracle_error_is_the_expanded_form_distance(kernel):
    errs = []
    for delta in DELTAS:
        om, Xt = _problem(kernel, delta)
        t, rmin = helpers.d2mu_truth_longdouble(om, 0, Xt)
        errs.append((rmin, helpers.relerr(om._d2_mu_d_p2(0, Xt), np.asarray(t, dtype=float))))
    if kernel == 'Exponential':
        # u / r^2 (with |a|^2 ~ 10): 1e-9 .. 1e-7 at r = 2e-3, per cent level at r = 2e-7 -- the 6e-9 / 1e-6 of the fuzz runs
        assert errs[0][1] < 1e-6 and errs[1][1] > 1e-7 and errs[2][1] > 1e-3, errs
        for rmin, e in errs:
            assert e < 1e3 * np.finfo(float).eps * 30.0 / rmin ** 2, (rmin, e)
    else:
        assert max(e for _, e in errs) < 1e-6, errs


@pytest.mark.gpu
@pytest.mark.parametrize('built', ['uploaded', 'device'])
@pytest.mark.parametrize('kernel', KERNELS)
def test_cuda_matches_the_truth_close_to_a_training_parameter(kernel, built, gpu_ctx):
    import gpdoemd_b200 as g
    for delta in DELTAS:
        om, Xt = _problem(kernel, delta)
        if built == 'uploaded':
            gm = g.GPModel.from_reference(om)
            alpha = None
        else:
            d = synthetic.model_data(73 * 1000, 1, 2, 3, 150, 0, 1e-4, kernel)
            gm = g.GPModel({'name': 'm', 'dim_x': 2, 'dim_p': 3, 'num_outputs': 1}, build_on_device=True)
            gm.gp_surrogate(Z=d['Z'], Y=d['Y'], kern_x=kernel, kern_p=kernel)
            gm.hyp = d['hyp']
            gm.gp_load_hyp()
            gm.Sigma = d['Sigma']
            alpha = gm.gps[0][0].posterior.woodbury_vector[:, 0]           # downloaded from the device build
            a_host = om.gps[0][0].posterior.woodbury_vector[:, 0]
            # device vs host alpha: both solve K alpha = y with cond(K) ~ 1e4 / noise; the gap is conditioning, not a bug
            assert np.max(np.abs(alpha - a_host)) <= 1e-7 * np.max(np.abs(a_host))
        gm.pmean = om.pmean.copy()
        t, rmin = helpers.d2mu_truth_longdouble(om, 0, Xt, alpha=alpha)
        t = np.asarray(t, dtype=float)
        got = gm._d2_mu_d_p2(0, Xt)
        err = helpers.relerr(got, t, 1.0)
        print('%s %s delta %.0e: r_min %.1e  max|truth| %.2e  CUDA vs 80-bit truth %.1e' % (kernel, built, delta, rmin,
                                                                                         np.abs(t).max(), err))
        assert err < 1e-9, (kernel, built, delta, err)


def _problem_1d(kernel):
    """The pattern the round-2 fuzz run flagged (tools/fuzz_explore.py seeds 6, 106, 126, 131): ONE model parameter.  In
    one dimension the 1/r terms of the second derivative cancel analytically, but only if r is the same number in both
    terms; with n = 285 training rows the closest one sits at r ~ 1e-3 and GPy's expanded-form r leaves 1e-9 behind."""
    oms, ds = helpers.oracle_models(1031, M=1, E=2, D=2, P=1, n=285, kernel=kernel, gp_noise=1e-4)
    om = oms[0]
    X = synthetic.candidates(131, 40, 2, ds[0]['xlo'], ds[0]['xhi'])
    return om, om.trans_x(X)


def test_one_parameter_oracle_gap_is_small_but_real():
    om, Xt = _problem_1d('Exponential')
    errs = []
    for e in range(2):
        t, rmin = helpers.d2mu_truth_longdouble(om, e, Xt)
        errs.append(helpers.relerr(om._d2_mu_d_p2(e, Xt), np.asarray(t, dtype=float), 1.0))
    assert max(errs) < 1e-6                       # rounding-level for a 1e-3 distance, yet it can exceed the 1e-9 gate


@pytest.mark.gpu
@pytest.mark.parametrize('kernel', KERNELS)
def test_cuda_matches_the_truth_with_one_parameter(kernel, gpu_ctx):
    import gpdoemd_b200 as g
    om, Xt = _problem_1d(kernel)
    gm = g.GPModel.from_reference(om)
    for e in range(2):
        t, rmin = helpers.d2mu_truth_longdouble(om, e, Xt)
        t = np.asarray(t, dtype=float)
        err = helpers.relerr(gm._d2_mu_d_p2(e, Xt), t, 1.0)
        print('%s P=1 output %d: r_min %.1e  CUDA vs truth %.1e  oracle vs truth %.1e' % (
            kernel, e, rmin, err, helpers.relerr(om._d2_mu_d_p2(e, Xt), t, 1.0)))
        assert err < 1e-9, (kernel, e, err)
