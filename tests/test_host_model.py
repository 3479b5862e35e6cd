"""Host-side logic of gpdoemd_b200.GPModel without a GPU: the reference's property validation, training-data
normalisation, assertion order and save/load (tests/test_models/test_surrogate_model.py:63-216,
tests/test_models/test_gp_model.py:61-137 of the reference), and the host exact-GP setup vs the oracle."""
import os

import numpy as np
import pytest

import helpers
import gpdoemd_b200 as g
from gpdoemd_b200 import synthetic
from gpdoemd_b200.gp_state import GPState


def _dict(**kw):
    d = {'name': 'm', 'call': lambda x, p: x, 'dim_x': 3, 'dim_p': 2, 'num_outputs': 2}
    d.update(kw)
    return d


def test_constructor_validation():
    m = g.GPModel(_dict())
    assert m.dim_b == 0 and m.non_binary_variables == [0, 1, 2]
    assert m.gp_noise_var == 1e-6 and np.array_equal(m.meas_noise_var, [1., 1.])
    with pytest.raises(AssertionError):
        g.GPModel(_dict(dim_x=0))
    with pytest.raises(AssertionError):
        g.GPModel(_dict(num_outputs=1.5))
    with pytest.raises(AssertionError):
        g.GPModel(_dict(name=3))
    with pytest.raises(AssertionError):
        g.GPModel(_dict(binary_variables=[3]))
    with pytest.raises(AssertionError):
        g.GPModel(_dict(binary_variables=[0.5]))
    with pytest.raises(ValueError):
        g.GPModel(_dict(binary_variables='a'))
    assert g.GPModel(_dict(binary_variables=1)).binary_variables == [1]
    with pytest.raises(AssertionError):
        g.GPModel(_dict(gp_noise_var=-1.0))
    with pytest.raises(AssertionError):
        g.GPModel(_dict(meas_noise_var=np.array([1., -1.])))


def test_pmean_sigma_properties():
    m = g.GPModel(_dict())
    assert m.pmean is None and m.Sigma is None
    with pytest.raises(AssertionError):
        m.pmean = [1, 2]
    with pytest.raises(AssertionError):
        m.pmean = np.zeros(3)
    m.pmean = np.array([1., 2.])
    with pytest.raises(AssertionError):
        m.Sigma = np.eye(3)
    S = np.eye(2)
    m.Sigma = S
    S[0, 0] = 5
    assert m.Sigma[0, 0] == 1           # the setter copies (model.py:153)
    del m.pmean
    assert m.pmean is None and np.array_equal(m._old_pmean, [1., 2.])


def test_training_data_normalisation():
    rng = np.random.default_rng(0)
    m = g.GPModel(_dict(binary_variables=[2]))
    Z = np.c_[rng.uniform(2, 5, (20, 2)), rng.integers(0, 2, 20), rng.uniform(-1, 1, (20, 2))]
    Y = rng.normal(3, 2, (20, 2))
    with pytest.raises(AssertionError):
        m.set_training_data(Z[:, :4], Y)
    m.set_training_data(Z, Y)
    assert m.Z.shape == (20, 5) and m.Z.min() >= 0 and m.Z.max() <= 1
    assert m.zmin[2] == 0 and m.zmax[2] == 1                 # binary columns keep 0/1 (surrogate_model.py:80-82)
    assert np.allclose(m.Y.mean(0), 0) and np.allclose(m.Y.std(0), 1)
    assert np.allclose(m.trans_z(m.Z, back=True), Z) and np.allclose(m.trans_y(m.Y, back=True), Y)
    assert m.X.shape == (20, 3) and m.P.shape == (20, 2) and len(m.xmin) == 3 and len(m.pmax) == 2
    m.pmean = np.array([0.1, -0.2])
    ZZ = m.get_Z(m.X[:4])
    assert ZZ.shape == (4, 5) and np.allclose(ZZ[:, 3:], m.trans_p(m.pmean))
    with pytest.raises(AssertionError):
        m.get_Z(m.X[:4], np.zeros((3, 2)))
    # the three-argument form (surrogate_model.py:159-165)
    m.set_training_data(Z[:, :3], Z[:, 3:], Y)
    assert m.Z.shape == (20, 5)


def test_predict_assertion_order():
    m = g.GPModel(_dict())
    X = np.zeros((2, 3))
    with pytest.raises(AssertionError):
        m.predict(X)                      # pmean missing
    m.pmean = np.zeros(2)
    with pytest.raises(AssertionError):
        m.predict(X)                      # Z / Y missing
    rng = np.random.default_rng(1)
    m.set_training_data(rng.uniform(size=(10, 5)), rng.normal(size=(10, 2)))
    with pytest.raises(AssertionError):
        m.predict(X)                      # hyp missing
    with pytest.raises(NotImplementedError):
        m.gp_optimise()


def test_gp_surrogate_layout_and_hyp_roundtrip(tmp_path):
    ms, ds = synthetic.build_models(g.GPModel, 4, M=1, E=2, D=3, P=2, n=40, kernel='Matern52', n_binary=1)
    m = ms[0]
    assert len(m.gps) == 2 and len(m.gps[0]) == 2           # E lists of 2^b GPs
    for e in range(2):
        for r in range(2):
            assert np.array_equal(m.gps[e][r][:], m.hyp[e][r])
    assert m.kern_p is m.kern_x                              # gp_model.py:60 quirk
    f = str(tmp_path / 'model')
    m.save(f)
    m2 = g.GPModel(_dict(dim_p=2, binary_variables=[2]))
    m2.load(f)
    assert np.allclose(m2.Z, m.Z, atol=1e-10) and np.allclose(m2.Y, m.Y, atol=1e-10)
    assert np.array_equal(m2.pmean, m.pmean) and np.array_equal(m2.Sigma, m.Sigma)
    m.clear_model()
    assert m.Z is None and m.Y is None and m.hyp is None and m.pmean is None and m.Sigma is None


@pytest.mark.parametrize('kernel', ['RBF', 'Exponential', 'Matern32', 'Matern52', 'RatQuad'])
def test_host_exact_gp_matches_oracle_state(kernel):
    """The host setup (ExactGP: K, Cholesky, alpha) reproduces the oracle's GPy restatement bit for bit,
    and GPState.from_gp reads the same arrays from either object."""
    oms, ds = helpers.oracle_models(6, M=1, E=1, D=3, P=2, n=50, kernel=kernel)
    gms, _ = synthetic.build_models(g.GPModel, 6, M=1, E=1, D=3, P=2, n=50, kernel=kernel)
    go, gg = oms[0].gps[0][0], gms[0].gps[0][0]
    assert np.array_equal(go.posterior.woodbury_chol, gg.posterior.woodbury_chol)
    assert np.array_equal(go.posterior.woodbury_vector, gg.posterior.woodbury_vector)
    so, sg = GPState.from_gp(go, 3, 2), GPState.from_gp(gg, 3, 2)
    for k in ('Z', 'alpha', 'L', 'ls_x', 'ls_p', 'x_active'):
        assert np.array_equal(getattr(so, k), getattr(sg, k)), k
    assert (so.kernel, so.var_x, so.var_p, so.power_x) == (sg.kernel, sg.var_x, sg.var_p, sg.power_x) and so.kernel == kernel


def test_from_reference_adopts_everything():
    oms, ds = helpers.oracle_models(8, M=1, E=2, D=3, P=2, n=30, kernel='RBF', n_binary=1)
    om = oms[0]
    m = g.GPModel.from_reference(om)
    assert m.binary_variables == om.binary_variables and m.dim_b == 1
    for k in ('xmin', 'xmax', 'pmin', 'pmax', 'ymean', 'ystd', 'pmean', 'Sigma'):
        assert np.array_equal(getattr(m, k), getattr(om, k)), k
    assert m.gps is om.gps and m.hyp is om.hyp


def test_default_device_follows_local_rank(monkeypatch):
    """Entry points without a model argument (criteria, assembly) must run on the rank's own GPU under torchrun:
    the library and torch share the CUDA runtime's current device (an 8-GPU bench once moved every rank to cuda:0)."""
    from gpdoemd_b200 import context
    monkeypatch.setattr(context, '_default_device', None)
    monkeypatch.delenv('LOCAL_RANK', raising=False)
    assert context.default_device() == 0
    monkeypatch.setenv('LOCAL_RANK', '5')
    assert context.default_device() == 5
    context.set_default_device(3)
    assert context.default_device() == 3


def test_state_checkpoint_roundtrip(tmp_path):
    """save_state / load_state: the exported posterior of every GP, transforms, pmean and Sigma survive a round trip
    bit for bit (CPU part; the restored model's predictions are compared on the GPU in test_gpu_parity)."""
    import gpdoemd_b200 as g
    from gpdoemd_b200 import synthetic
    from gpdoemd_b200.gp_state import GPState
    gms, ds = synthetic.build_models(g.GPModel, 5, 1, 2, 3, 2, 40, 'Matern32', n_binary=1)
    gm = gms[0]
    fn = str(tmp_path / 'm0_state')
    gm.save_state(fn)
    m2 = g.GPModel.load_state(fn)
    assert (m2.dim_x, m2.dim_p, m2.num_outputs, m2.binary_variables) == (3, 2, 2, [2])
    assert np.array_equal(m2.pmean, gm.pmean) and np.array_equal(m2.Sigma, gm.Sigma)
    assert np.array_equal(m2.zmin, gm.zmin) and np.array_equal(m2.ystd, gm.ystd)
    for e in range(2):
        for r in range(2):
            a, b = GPState.from_gp(gm.gps[e][r], 3, 2), m2.gps[e][r]
            assert isinstance(b, GPState) and a.kernel == b.kernel and a.factor_kind == b.factor_kind
            for k in ('x_active', 'ls_x', 'ls_p', 'Z', 'alpha', 'L'):
                assert np.array_equal(getattr(a, k), getattr(b, k))
            assert (a.var_x, a.var_p) == (b.var_x, b.var_p)


def test_missing_binary_combination_raises_like_the_reference():
    """gp_model.py:193-199: a design whose binary combination has no GP (no training row had it) makes the reference
    dereference None; the product raises the same AttributeError instead of returning zeros from the device."""
    import gpdoemd_b200 as g
    from gpdoemd_b200 import synthetic
    gms, ds = synthetic.build_models(g.GPModel, 7, 1, 1, 3, 2, 30, 'Matern32', n_binary=1)
    gm = gms[0]
    gm._gps = [[gm.gps[0][0], None]]                 # as if no training row had b = 1
    X = synthetic.candidates(1, 10, 3, ds[0]['xlo'], ds[0]['xhi'], n_binary=1)
    X[:, 2] = 0
    gm.check_routing(X)                              # only combination 0 occurs: fine
    X[3, 2] = 1
    with pytest.raises(AttributeError, match='predict_noiseless'):
        gm.check_routing(X)
    with pytest.raises(AttributeError):
        g.score_designs([gm], X)


def test_from_reference_reads_gpy_shaped_objects():
    """GPy is not importable here, so the adoption path is exercised on objects with GPy's SHAPES: hyper-parameters are
    1-element / d-element array subclasses (paramz `Param`), the kernel classes are subclasses of subclasses named like
    GPy's (GPdoemd.kernels.RBF(GPy.kern.RBF)), `active_dims` are integer arrays, the posterior holds an (n, 1)
    woodbury_vector.  GPState.from_gp must extract the same plain arrays as from the oracle's objects."""
    class Param(np.ndarray):                                    # paramz.Param is an ndarray subclass
        def __new__(cls, v):
            return np.atleast_1d(np.asarray(v, dtype=float)).view(cls)

    class Kern:                                                 # GPy.kern.src.kern.Kern
        pass

    class Stationary(Kern):                                     # GPy.kern.src.stationary.Stationary
        pass

    class Matern52(Stationary):                                 # GPy.kern.Matern52
        pass

    GpdoemdMatern52 = type('Matern52', (Matern52,), {})         # GPdoemd.kernels.Matern52(GPy.kern.Matern52)

    oms, ds = helpers.oracle_models(9, M=1, E=1, D=3, P=2, n=40, kernel='Matern52')
    go = oms[0].gps[0][0]

    def gpy_like(k):
        o = GpdoemdMatern52()
        o.variance, o.lengthscale = Param(k.variance), Param(k.lengthscale)
        o.active_dims = np.asarray(k.active_dims, dtype=np.int64)
        return o

    class Prod:
        pass

    class Posterior:
        pass

    class GPRegression:
        pass
    gp = GPRegression()
    gp.kern = Prod()
    gp.kern.kernx, gp.kern.kernp = gpy_like(go.kern.kernx), gpy_like(go.kern.kernp)
    gp.posterior = Posterior()
    gp.posterior.woodbury_chol = np.asfortranarray(go.posterior.woodbury_chol)         # LAPACK output order
    gp.posterior.woodbury_vector = go.posterior.woodbury_vector.copy()
    gp._predictive_variable = go._predictive_variable
    a, b = GPState.from_gp(gp, 3, 2), GPState.from_gp(go, 3, 2)
    assert a.kernel == b.kernel == 'Matern52' and (a.var_x, a.var_p) == (b.var_x, b.var_p)
    for k in ('Z', 'alpha', 'L', 'ls_x', 'ls_p', 'x_active'):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert a.L.flags['C_CONTIGUOUS'] and a.L.dtype == np.float64 and a.alpha.shape == (40,)
