"""``GPModel``: the reference's GP surrogate model class with the prediction part on the B200.

Same constructor dict, properties, asserts and method names as ``GPdoemd/models/model.py:31-160``,
``surrogate_model.py:32-262`` and ``gp_model.py:37-280``.  The five hooks ``_predict``, ``_d_mu_d_p``,
``_d2_mu_d_p2``, ``_d_s2_d_p``, ``_d2_s2_d_p2`` (transformed units, gp_model.py:184-280) are the device
boundary: they call ``gpd_hooks`` of the C ABI.  Hyper-parameter optimisation stays with the reference
(GPy); a trained reference model is adopted with ``GPModel.from_reference``.
"""
import ctypes as C
import pickle

import numpy as np

from . import _lib, kernels
from ._lib import check, dptr
from .context import get_context
from .gp_state import DeviceExactGP, ExactGP, GPState, ProductKernel, SparseGP


class BoxTransform:
    """transform.py:27-47"""

    def __init__(self, xmin, xmax):
        self.xmin, self.xmax = xmin, xmax
        self.diff = xmax - xmin

    def __call__(self, X, back=False):
        return self.xmin + X * self.diff if back else (X - self.xmin) / self.diff

    def var(self, X, back=False):
        return X * self.diff ** 2 if back else X / self.diff ** 2

    def cov(self, X, back=False):
        d = self.diff[:, None] * self.diff[None, :]
        return X * d if back else X / d


class MeanTransform:
    """transform.py:50-73"""

    def __init__(self, mean, std):
        self.mean, self.std = mean, std

    def __call__(self, X, back=False):
        return self.mean + X * self.std if back else (X - self.mean) / self.std

    def var(self, X, back=False):
        return X * self.std ** 2 if back else X / self.std ** 2

    def cov(self, X, back=False):
        d = self.std[:, None] * self.std[None, :]
        return X * d if back else X / d

    def prediction(self, M, S, back=False):
        Mt = self(M, back=back)
        return (Mt, self.var(S, back=back)) if S.ndim == 2 else (Mt, self.cov(S, back=back))


def binary_dimensions(Z, binary_variables):
    """utils.py:27-48, vectorised; combination index = b0 + 2 b1 for two binary columns
    (row order of meshgrid's default 'xy' indexing; tests/test_utils.py:16,41,66)."""
    lb = len(binary_variables)
    if lb == 0:
        return [0], np.zeros(Z.shape[0])
    lib = _lib.load()
    bits = np.asarray(Z)[:, binary_variables] > 0.5
    w = np.array([lib.gpd_binary_combo_weight(lb, v) for v in range(lb)])
    return range(2 ** lb), bits.dot(w)


class GPModel:
    def __init__(self, model_dict, device=0, build_on_device=False):
        # build_on_device: gp_surrogate / gp_load_hyp run the exact-GP inference (K, Cholesky, alpha) on the GPU
        # (gpd_gp_build) instead of host LAPACK + an n x n factor upload per GP
        self.build_on_device = bool(build_on_device)
        self.name = model_dict.get('name', 'model')
        self.call = model_dict.get('call', None)
        self.dim_x = model_dict['dim_x']
        self.dim_p = model_dict['dim_p']
        self.num_outputs = model_dict.get('num_outputs', 1)
        self.p_bounds = model_dict.get('p_bounds', [])
        self.meas_noise_var = model_dict.get('meas_noise_var', 1.)
        self.binary_variables = model_dict.get('binary_variables', [])
        self.gp_noise_var = model_dict.get('gp_noise_var', 1e-6)
        self._device = device
        self._dev_ctx = None          # Context the device handles live in
        self._dev_model = None        # gpd_model*
        self._dev_gps = []            # gpd_gp* handles
        self._dev_pstar = None
        self._dev_sigma = None
        self._cache = None

    # ---- model.py properties ---------------------------------------------------------------
    @property
    def name(self):
        return self._name

    @name.setter
    def name(self, value):
        assert isinstance(value, str), 'Model name has to be of type string'
        self._name = value

    @property
    def call(self):
        return self._call

    @call.setter
    def call(self, value):
        assert value is None or callable(value), 'Invalid call handle'
        self._call = value

    @property
    def num_outputs(self):
        return self._num_outputs

    @num_outputs.setter
    def num_outputs(self, value):
        assert isinstance(value, int) and value > 0, 'Invalid no. of outputs'
        self._num_outputs = value

    @property
    def meas_noise_var(self):
        return self._meas_noise_var

    @meas_noise_var.setter
    def meas_noise_var(self, value):
        if isinstance(value, (int, float)):
            value = np.array([value] * self.num_outputs)
        assert isinstance(value, np.ndarray), 'Variance has to be numpy array'
        if value.ndim == 1:
            assert value.shape == (self.num_outputs,), 'Incorrect shape'
            assert np.all(value > 0.), 'Variance must be +ve and >0'
        else:
            assert value.shape == (self.num_outputs, self.num_outputs), 'Incorrect shape'
        self._meas_noise_var = value

    @property
    def meas_noise_covar(self):
        return np.diag(self.meas_noise_var) if self.meas_noise_var.ndim == 1 else self.meas_noise_var

    @property
    def dim_x(self):
        return self._dim_x

    @dim_x.setter
    def dim_x(self, value):
        assert isinstance(value, int) and value > 0, 'Invalid x dimensionality'
        self._dim_x = value

    @property
    def dim_p(self):
        return self._dim_p

    @dim_p.setter
    def dim_p(self, value):
        assert isinstance(value, int) and value > 0, 'Invalid p dimensionality'
        self._dim_p = value

    @property
    def pmean(self):
        return getattr(self, '_pmean', None)

    @pmean.setter
    def pmean(self, value):
        if value is not None:
            assert isinstance(value, np.ndarray), 'pmean has to be numpy array'
            assert value.shape == (self.dim_p,), 'pmean has incorrect shape'
        self._pmean = value
        if not hasattr(self, '_old_pmean'):
            self._old_pmean = None

    @pmean.deleter
    def pmean(self):
        self._old_pmean = None if self.pmean is None else self.pmean.copy()
        self._pmean = None

    def param_estim(self, Xdata, Ydata, method, p_bounds=None):
        if (p_bounds is None) and (self.p_bounds is not None):
            p_bounds = np.asarray(self.p_bounds)
        self.pmean = method(self, Xdata, Ydata, p_bounds)

    @property
    def Sigma(self):
        return getattr(self, '_Sigma', None)

    @Sigma.setter
    def Sigma(self, value):
        if value is None:
            self._Sigma = None
        else:
            assert isinstance(value, np.ndarray), 'Sigma has to be numpy array'
            assert value.shape == (self.dim_p, self.dim_p), 'Incorrect shape'
            self._Sigma = value.copy()

    @Sigma.deleter
    def Sigma(self):
        self._Sigma = None

    # ---- surrogate_model.py properties -----------------------------------------------------
    @property
    def binary_variables(self):
        return self._binary_variables

    @binary_variables.setter
    def binary_variables(self, value):
        if isinstance(value, list):
            for v in value:
                assert isinstance(v, int), 'Value not integer'
                assert 0 <= v < self.dim_x, 'Value outside range'
            self._binary_variables = value
        elif isinstance(value, int):
            assert 0 <= value < self.dim_x, 'Value outside range'
            self._binary_variables = [value]
        else:
            raise ValueError('Binary variable must be list or integer')

    @property
    def non_binary_variables(self):
        return [i for i in range(self.dim_x) if i not in self.binary_variables]

    @property
    def dim_b(self):
        return len(self.binary_variables)

    @property
    def Z(self):
        return getattr(self, '_Z', None)

    @Z.setter
    def Z(self, value):
        if value is not None:
            assert value.shape[1] == self.dim_x + self.dim_p, 'Incorrect shape'
            self._zmin = np.min(value, axis=0)
            self._zmax = np.max(value, axis=0)
            for i in self.binary_variables:
                self._zmin[i] = 0
                self._zmax[i] = 1
            self.trans_z = BoxTransform(self.zmin, self.zmax)
            self.trans_x = BoxTransform(self.xmin, self.xmax)
            self.trans_p = BoxTransform(self.pmin, self.pmax)
            self._Z = self.trans_z(value)
            self._invalidate_device()      # the device model holds xmin/xmax/pmin/pmax

    @Z.deleter
    def Z(self):
        self._Z = None
        self._invalidate_device()

    zmin = property(lambda self: self._zmin)
    zmax = property(lambda self: self._zmax)
    X = property(lambda self: None if self.Z is None else self.Z[:, :self.dim_x])
    P = property(lambda self: None if self.Z is None else self.Z[:, self.dim_x:])
    xmin = property(lambda self: self.zmin[:self.dim_x])
    xmax = property(lambda self: self.zmax[:self.dim_x])
    pmin = property(lambda self: self.zmin[self.dim_x:])
    pmax = property(lambda self: self.zmax[self.dim_x:])

    def get_Z(self, X, P=None):
        if P is None:
            assert self.pmean is not None
            P = np.tile(self.trans_p(self.pmean), (len(X), 1))
        assert P.shape == (len(X), self.dim_p)
        return np.c_[X, P]

    @property
    def Y(self):
        return getattr(self, '_Y', None)

    @Y.setter
    def Y(self, value):
        if value is not None:
            assert value.shape[1] == self.num_outputs
            self._ymean = np.mean(value, axis=0)
            self._ystd = np.std(value, axis=0)
            self.trans_y = MeanTransform(self.ymean, self.ystd)
            self._Y = self.trans_y(value)
            self._invalidate_device()      # the device model holds ymean/ystd

    @Y.deleter
    def Y(self):
        self._Y = None
        self._invalidate_device()

    ymean = property(lambda self: self._ymean)
    ystd = property(lambda self: self._ystd)

    def set_training_data(self, Z, Y, _Y=None):
        if _Y is None:
            self.Z = Z
            self.Y = Y
        else:
            self.Z = np.c_[Z, Y]
            self.Y = _Y

    @property
    def hyp(self):
        return getattr(self, '_hyp', None)

    @hyp.setter
    def hyp(self, value):
        if value is not None:
            assert len(value) == self.num_outputs
            self._hyp = value
            self._invalidate_device()

    @hyp.deleter
    def hyp(self):
        self._hyp = None

    # ---- gp_model.py: kernels, surrogate construction ----------------------------------------
    @property
    def kern_x(self):
        return getattr(self, '_kern_x', None)

    @kern_x.setter
    def kern_x(self, value):
        if value is not None:
            kernels.kernel_name_of(value)          # raises for unsupported classes
            self._kern_x = value

    @property
    def kern_p(self):
        # gp_model.py:60 -- the reference's getter returns _kern_x: both parts use kern_x's class
        return None if not hasattr(self, '_kern_p') else self._kern_x

    @kern_p.setter
    def kern_p(self, value):
        if value is not None:
            kernels.kernel_name_of(value)
            self._kern_p = value

    def set_kernels(self, kern_x, kern_p):
        self.kern_x = kern_x
        self.kern_p = kern_p

    @property
    def gps(self):
        return getattr(self, '_gps', None)

    @gps.setter
    def gps(self, value):
        assert len(value) == self.num_outputs
        self._gps = value
        self._invalidate_device()

    @gps.deleter
    def gps(self):
        self._gps = None
        self._invalidate_device()

    @property
    def gp_noise_var(self):
        return self._gp_noise_var

    @gp_noise_var.setter
    def gp_noise_var(self, value):
        assert isinstance(value, (int, float)) and value > 0.
        self._gp_noise_var = value

    def gp_surrogate(self, Z=None, Y=None, kern_x=None, kern_p=None):
        self.set_training_data(Z, Y)
        assert self.Z is not None and self.Y is not None
        self.set_kernels(kern_x, kern_p)
        assert self.kern_x is not None and self.kern_p is not None
        host_cls = kernels.BY_NAME[kernels.kernel_name_of(self.kern_x)]
        R, J = binary_dimensions(self.Z, self.binary_variables)
        gps = []
        for e in range(self.num_outputs):
            gps.append([])
            for r in R:
                Jr = (J == r)
                if not np.any(Jr):
                    gps[e].append(None)
                    continue
                dim = self.dim_x + self.dim_p
                kernx = host_cls(self.dim_x - self.dim_b, self.non_binary_variables, 'kernx')
                kernp = host_cls(self.dim_p, range(self.dim_x, dim), 'kernp')
                if self.build_on_device:
                    gps[e].append(DeviceExactGP(self.Z[Jr], self.Y[np.ix_(Jr, [e])], ProductKernel(kernx, kernp),
                                                device=self._device))
                else:
                    gps[e].append(ExactGP(self.Z[Jr], self.Y[np.ix_(Jr, [e])], ProductKernel(kernx, kernp)))
        self.gps = gps

    def gp_load_hyp(self, index=None):
        if index is None:
            index = range(self.num_outputs)
        elif isinstance(index, int):
            index = [index]
        for e in index:
            for gp, hyp in zip(self.gps[e], self.hyp[e]):
                if gp is None:
                    continue
                gp[:] = hyp
        self._invalidate_device()

    def gp_optimise(self, index=None, max_lengthscale=10, **kwargs):
        raise NotImplementedError('GP hyper-parameter training stays on the reference CPU path (GPy); train '
                                  'with GPdoemd.models.GPModel.gp_optimise and adopt the result with '
                                  'GPModel.from_reference(model), or set .hyp and call gp_load_hyp()')

    gp_optimize = gp_optimise

    @classmethod
    def from_reference(cls, ref, device=0):
        """Adopt a trained reference ``GPdoemd.models.GPModel`` (or any object with its attributes)."""
        d = {'name': ref.name, 'call': getattr(ref, 'call', None), 'dim_x': ref.dim_x, 'dim_p': ref.dim_p,
             'num_outputs': ref.num_outputs, 'p_bounds': getattr(ref, 'p_bounds', []),
             'meas_noise_var': ref.meas_noise_var, 'binary_variables': list(ref.binary_variables),
             'gp_noise_var': getattr(ref, 'gp_noise_var', 1e-6)}
        m = cls(d, device=device)
        m._Z, m._zmin, m._zmax = np.array(ref.Z), np.array(ref.zmin, dtype=float), np.array(ref.zmax, dtype=float)
        m.trans_z = BoxTransform(m.zmin, m.zmax)
        m.trans_x = BoxTransform(m.xmin, m.xmax)
        m.trans_p = BoxTransform(m.pmin, m.pmax)
        m._Y, m._ymean, m._ystd = np.array(ref.Y), np.array(ref.ymean, dtype=float), np.array(ref.ystd, dtype=float)
        m.trans_y = MeanTransform(m.ymean, m.ystd)
        kx = getattr(ref, 'kern_x', None)
        if kx is not None:
            m._kern_x = m._kern_p = kx
        m._gps = ref.gps
        m._hyp = ref.hyp if ref.hyp is not None else [[None] * len(g) for g in ref.gps]
        if ref.pmean is not None:
            m.pmean = np.array(ref.pmean, dtype=float)
        if ref.Sigma is not None:
            m.Sigma = np.array(ref.Sigma, dtype=float)
        return m

    # ---- device state -------------------------------------------------------------------------
    def _invalidate_device(self):
        self._cache = None
        if getattr(self, '_dev_model', None) is None:
            return
        # the context the handles were created in; after Context.close() (gpd_destroy) its handles are gone with it
        # and must not be touched (gpd_model_free would dereference the deleted context)
        ctx = getattr(self, '_dev_ctx', None)
        if ctx is not None and ctx.handle:
            ctx.lib.gpd_model_free(self._dev_model)
            for h in self._dev_gps:
                ctx.lib.gpd_gp_free(h)
        self._dev_ctx = None
        self._dev_model, self._dev_gps = None, []
        self._dev_pstar = self._dev_sigma = None
        self._dev_key = None

    def __del__(self):
        try:
            self._invalidate_device()
        except Exception:
            pass

    def _ensure_device(self):
        if self._dev_model is not None:
            if self._dev_ctx.handle:
                return
            self._invalidate_device()          # the context was closed: rebuild in a fresh one
        assert self.gps is not None, 'no GP surrogate: call gp_surrogate() (and gp_load_hyp()) first'
        ctx = self._dev_ctx = get_context(self._device)
        E, R = self.num_outputs, 2 ** self.dim_b
        handles = (C.c_void_p * (E * R))()
        for e in range(E):
            assert len(self.gps[e]) == R
            for r, gp in enumerate(self.gps[e]):
                if gp is None:
                    handles[e * R + r] = None
                    continue
                if isinstance(gp, GPState):                       # restored from a state checkpoint
                    h = gp.upload(ctx)
                elif getattr(gp, 'built_on_device', False):
                    h = gp.build(ctx, self.dim_x, self.dim_p)
                else:
                    h = GPState.from_gp(gp, self.dim_x, self.dim_p).upload(ctx)
                self._dev_gps.append(h)
                handles[e * R + r] = h
        bv = np.asarray(self.binary_variables, dtype=np.int32)
        mh = C.c_void_p()
        f = _lib.as_f64
        xmin, xmax, pmin, pmax = f(self.xmin), f(self.xmax), f(self.pmin), f(self.pmax)
        ymean, ystd = f(self.ymean), f(self.ystd)
        try:
            check(ctx.lib.gpd_model_create(ctx.handle, E, self.dim_x, self.dim_p, len(bv),
                                           bv.ctypes.data_as(_lib.c_int32_p) if len(bv) else None,
                                           C.cast(handles, _lib.c_void_pp), dptr(xmin), dptr(xmax), dptr(pmin),
                                           dptr(pmax), dptr(ymean), dptr(ystd), C.byref(mh)))
        except Exception:
            for h in self._dev_gps:            # a refused model (size limit) must not leak its uploaded GPs
                ctx.lib.gpd_gp_free(h)
            self._dev_gps = []
            raise
        self._dev_model = mh

    def _sync_device(self, pstar=None):
        """Make the device model agree with pmean / Sigma (or an explicit transformed p*)."""
        self._ensure_device()
        if pstar is None:
            # fast path of a design loop: pmean / Sigma unchanged since the last call (compared by content: callers
            # may mutate the arrays in place)
            assert self.pmean is not None
            key = (self.pmean.tobytes(), None if self.Sigma is None else self.Sigma.tobytes())
            if key == getattr(self, '_dev_key', None):
                return
        key_after = None
        if pstar is None:
            pstar = self.trans_p(self.pmean)
            key_after = key
        pstar = _lib.as_f64(pstar)
        sig = None if self.Sigma is None else _lib.as_f64(self.Sigma)
        if sig is None and self._dev_sigma is not None:
            # `del model.Sigma` / `model.Sigma = None` after an upload: the device copy must not survive
            # (taylor_first_order.py:38 would fail on model.Sigma = None)
            self._invalidate_device()
            self._ensure_device()
        ctx = self._dev_ctx
        same_p = self._dev_pstar is not None and np.array_equal(self._dev_pstar, pstar)
        same_s = (sig is None) or (self._dev_sigma is not None and np.array_equal(self._dev_sigma, sig))
        if same_p and same_s:
            self._dev_key = key_after if key_after is not None else getattr(self, '_dev_key', None)
            return
        if not same_s:
            # Sigma travels with pmean in the C ABI; send the original-unit pmean that maps to pstar
            pm = _lib.as_f64(self.trans_p(pstar, back=True))
            check(ctx.lib.gpd_model_set_pmean_sigma(self._dev_model, dptr(pm), dptr(sig)))
            self._dev_sigma = sig.copy()
        check(ctx.lib.gpd_model_set_pstar(self._dev_model, dptr(pstar)))
        self._dev_pstar = pstar.copy()
        self._dev_key = key_after            # None after an explicit p* (hooks called with a foreign p)
        self._cache = None

    def device_handle(self):
        """gpd_model* with pmean/Sigma synchronised (used by the fused scoring path)."""
        assert self.Sigma is not None, 'model.Sigma is not set (taylor_first_order.py:38 reads it)'
        self._sync_device()
        return self._dev_model

    def replicate(self, device):
        """The same trained model bound to another GPU (GP state replicated: SURVEY 8e); host arrays are shared."""
        m = type(self).from_reference(self, device=device)
        m.build_on_device = self.build_on_device
        return m

    # ---- hooks (transformed units) = the device boundary ---------------------------------------
    _FLAG_KEYS = ('dmu', 'd2mu', 'ds2', 'd2s2')

    def check_routing(self, Xt):
        """The reference dereferences ``self.gps[e][r]`` for every binary combination r that occurs among the designs
        (gp_model.py:193-199) and dies with AttributeError when no training row had that combination (gps[e][r] is
        None, :111-113).  The device would leave such rows at zero, so the same error is raised here first."""
        if self.dim_b == 0 or all(gp is not None for row in self.gps for gp in row):
            return
        _, J = binary_dimensions(np.asarray(Xt), self.binary_variables)
        present = np.unique(J.astype(int))
        for e, row in enumerate(self.gps):
            for r in present:
                if row[r] is None:
                    raise AttributeError("'NoneType' object has no attribute 'predict_noiseless' (no GP for output %d, "
                                         "binary combination %d: no training row has it)" % (e, r))

    def _hooks(self, Xt, pstar, want):
        """All outputs at once for a batch of transformed designs; cached per (X, p*)."""
        Xt = _lib.as_f64(Xt)
        assert Xt.ndim == 2 and Xt.shape[1] == self.dim_x
        self.check_routing(Xt)
        self._sync_device(pstar)
        key = (Xt.shape, Xt.tobytes(), self._dev_pstar.tobytes())      # compared by content, not by hash
        c = self._cache
        if c is not None and c['key'] == key and all(k in c for k in want):
            return c
        have = set() if (c is None or c['key'] != key) else {k for k in self._FLAG_KEYS if k in c}
        want = set(want) | have
        N, E, P = len(Xt), self.num_outputs, self.dim_p
        out = {'key': key, 'mu': np.empty((N, E)), 's2': np.empty((N, E))}
        shapes = {'dmu': (N, E, P), 'd2mu': (N, E, P, P), 'ds2': (N, E, P), 'd2s2': (N, E, P, P)}
        for k in self._FLAG_KEYS:
            if k in want:
                out[k] = np.empty(shapes[k])
        ctx = self._dev_ctx
        check(ctx.lib.gpd_hooks(self._dev_model, dptr(Xt), N, 1, dptr(out['mu']), dptr(out['s2']),
                                dptr(out.get('dmu')), dptr(out.get('d2mu')), dptr(out.get('ds2')),
                                dptr(out.get('d2s2'))))
        self._cache = out
        return out

    def _predict(self, xnew, p):
        # the first p-derivative of the mean costs P extra FMAs per (candidate, training row) in the same
        # pass, so it is computed eagerly: the reference's predict + per-output d_mu_d_p loop
        # (taylor_first_order.py:29-33) then needs ONE device pass instead of two
        h = self._hooks(xnew, p, ('dmu',))
        return h['mu'].copy(), h['s2'].copy()

    def _pstar(self):
        assert self.pmean is not None
        return self.trans_p(self.pmean)

    def _d_mu_d_p(self, e, X):
        return self._hooks(X, self._pstar(), ('dmu',))['dmu'][:, e].copy()

    def _d2_mu_d_p2(self, e, X):
        return self._hooks(X, self._pstar(), ('dmu', 'd2mu'))['d2mu'][:, e].copy()

    def _d_s2_d_p(self, e, X):
        return self._hooks(X, self._pstar(), ('ds2',))['ds2'][:, e].copy()

    def _d2_s2_d_p2(self, e, X):
        return self._hooks(X, self._pstar(), ('dmu', 'd2mu', 'd2s2'))['d2s2'][:, e].copy()

    # ---- public wrappers, original units (surrogate_model.py:170-215) --------------------------
    def predict(self, xnew):
        assert self.pmean is not None
        assert self.Z is not None and self.Y is not None
        assert self.hyp is not None
        xt = self.trans_x(xnew)
        pt = self.trans_p(self.pmean)
        M, S = self._predict(xt, pt)
        return self.trans_y.prediction(M, S, back=True)

    def d_mu_d_p(self, e, X):
        der = self._d_mu_d_p(e, self.trans_x(X))
        return der * self.ystd[e] / (self.pmax - self.pmin)

    def d2_mu_d_p2(self, e, X):
        der = self._d2_mu_d_p2(e, self.trans_x(X))
        diff = self.pmax - self.pmin
        return der * self.ystd[e] / (diff[:, None] * diff[None, :])

    def d_s2_d_p(self, e, X):
        der = self._d_s2_d_p(e, self.trans_x(X))
        return der * self.ystd[e] ** 2 / (self.pmax - self.pmin)

    def d2_s2_d_p2(self, e, X):
        der = self._d2_s2_d_p2(e, self.trans_x(X))
        diff = self.pmax - self.pmin
        return der * self.ystd[e] ** 2 / (diff[:, None] * diff[None, :])

    # ---- clear / save / load (model.py:187-241, surrogate_model.py:238-261) ----------------------
    def clear_model(self):
        del self.Z
        del self.Y
        del self.hyp
        del self.gps
        del self.pmean
        del self.Sigma

    def _get_save_dict(self):
        back = (lambda v, t: None if v is None else t(v, back=True))
        return {'pmean': self.pmean, 'old_pmean': getattr(self, '_old_pmean', None), 'Sigma': self.Sigma,
                'hyp': self.hyp,
                'Z': None if self.Z is None else back(self.Z, self.trans_z),
                'Y': None if self.Y is None else back(self.Y, self.trans_y)}

    def _load_save_dict(self, d):
        self.pmean = d['pmean']
        self._old_pmean = d['old_pmean']
        self.Sigma = d['Sigma']
        self.Z = d['Z']
        self.Y = d['Y']
        self.hyp = d['hyp']

    # ---- state checkpoint: everything prediction needs, no GPy / refit on restore ---------------------------
    def save_state(self, filename):
        """Write transforms, pmean, Sigma and the exported posterior of every GP (Z, alpha, factor, hyper-parameters)
        to one ``.npz``.  Unlike ``save`` (model.py:198-241: hyp + raw Z, Y, the GPs are refitted on load) the
        restored model predicts immediately and bit-identically."""
        assert self.gps is not None and self.Z is not None and self.Y is not None
        d = {'meta_name': np.asarray(self.name), 'meta_dims': np.array([self.dim_x, self.dim_p, self.num_outputs]),
             'meta_binary': np.asarray(self.binary_variables, dtype=int), 'meas_noise_var': np.asarray(self.meas_noise_var),
             'gp_noise_var': np.asarray(self.gp_noise_var), 'zmin': self.zmin, 'zmax': self.zmax, 'ymean': self.ymean,
             'ystd': self.ystd, 'sparse': np.asarray(isinstance(self, SparseGPModel))}
        if self.pmean is not None:
            d['pmean'] = self.pmean
        if self.Sigma is not None:
            d['Sigma'] = self.Sigma
        for e, row in enumerate(self.gps):
            for r, gp in enumerate(row):
                if gp is None:
                    continue
                st = gp if isinstance(gp, GPState) else GPState.from_gp(gp, self.dim_x, self.dim_p)
                d.update(st.to_dict('gp%d_%d_' % (e, r)))
        np.savez(filename, **d)

    @classmethod
    def load_state(cls, filename, device=0):
        with np.load(filename if str(filename).endswith('.npz') else str(filename) + '.npz') as f:
            d = {k: f[k] for k in f.files}
        dim_x, dim_p, E = (int(v) for v in d['meta_dims'])
        klass = SparseGPModel if bool(d['sparse']) and cls is GPModel else cls
        mnv = d['meas_noise_var']
        m = klass({'name': str(d['meta_name']), 'dim_x': dim_x, 'dim_p': dim_p, 'num_outputs': E,
                   'meas_noise_var': float(mnv) if mnv.ndim == 0 else mnv,
                   'binary_variables': [int(v) for v in d['meta_binary']], 'gp_noise_var': float(d['gp_noise_var'])},
                  device=device)
        m._zmin, m._zmax = d['zmin'], d['zmax']
        m.trans_z = BoxTransform(m.zmin, m.zmax)
        m.trans_x = BoxTransform(m.xmin, m.xmax)
        m.trans_p = BoxTransform(m.pmin, m.pmax)
        m._ymean, m._ystd = d['ymean'], d['ystd']
        m.trans_y = MeanTransform(m.ymean, m.ystd)
        R = 2 ** m.dim_b
        gps = [[GPState.from_dict(d, 'gp%d_%d_' % (e, r)) if ('gp%d_%d_Z' % (e, r)) in d else None for r in range(R)]
               for e in range(E)]
        # the reference's predict() asserts Z, Y and hyp (surrogate_model.py:171-173): restored models carry the
        # inducing / training inputs of the first GP as a stand-in, hyper-parameters live inside the GP states
        first = next(g for row in gps for g in row if g is not None)
        m._Z, m._Y = first.Z, np.zeros((len(first.Z), E))
        m._gps = gps
        m._hyp = [[None] * R for _ in range(E)]
        if 'pmean' in d:
            m.pmean = d['pmean']
        if 'Sigma' in d:
            m.Sigma = d['Sigma']
        return m

    def save(self, filename):
        assert isinstance(filename, str)
        if not filename.endswith('.gpdoemd.model'):
            filename += '.gpdoemd.model'
        with open(filename, 'wb') as f:
            pickle.dump(self._get_save_dict(), f, pickle.HIGHEST_PROTOCOL)

    def load(self, filename):
        assert isinstance(filename, str)
        if not filename.endswith('.gpdoemd.model'):
            filename += '.gpdoemd.model'
        with open(filename, 'rb') as f:
            self._load_save_dict(pickle.load(f))


class SparseGPModel(GPModel):
    """``GPdoemd/models/sparse_gp_model.py:35-74``: GPModel whose surrogates are sparse GPs with at most
    ``max_num_inducing`` inducing inputs.  Every hook is inherited: on the device a sparse GP is the same
    computation as an exact one with the inducing inputs as "training rows", ``woodbury_vector`` as alpha and a
    triangular root T of ``woodbury_inv`` in place of L^-1 (``GPD_FACTOR_ROOT``)."""

    def __init__(self, model_dict, device=0, build_on_device=False):
        super().__init__(model_dict, device=device, build_on_device=False)   # sparse inference stays on the host
        self.max_num_inducing = model_dict.get('max_num_inducing', 500)

    def gp_surrogate(self, Z=None, Y=None, kern_x=None, kern_p=None, num_inducing=None):
        self.set_training_data(Z, Y)
        assert self.Z is not None and self.Y is not None
        self.set_kernels(kern_x, kern_p)
        assert self.kern_x is not None and self.kern_p is not None
        host_cls = kernels.BY_NAME[kernels.kernel_name_of(self.kern_x)]
        R, J = binary_dimensions(self.Z, self.binary_variables)
        numi = self.max_num_inducing if num_inducing is None else num_inducing
        gps = []
        for e in range(self.num_outputs):
            gps.append([])
            for r in R:
                Jr = (J == r)
                if not np.any(Jr):
                    gps[e].append(None)
                    continue
                dim = self.dim_x + self.dim_p
                kernx = host_cls(self.dim_x - self.dim_b, self.non_binary_variables, 'kernx')
                kernp = host_cls(self.dim_p, range(self.dim_x, dim), 'kernp')
                gps[e].append(SparseGP(self.Z[Jr], self.Y[np.ix_(Jr, [e])], ProductKernel(kernx, kernp),
                                       num_inducing=numi, rng=np.random.default_rng(1000 * e + r)))
        self.gps = gps
