"""GP state hand-off: what the device needs from a trained exact GP.

The reference trains ``GPy.models.GPRegression(Zr, Yr, kernx*kernp)`` objects on the CPU
(``GPdoemd/models/gp_model.py:96-181``).  Prediction then needs, per GP, exactly the attributes the
reference itself reads at ``gp_model.py:231-233,262-275``: ``gp._predictive_variable``,
``gp.posterior.woodbury_vector``, ``gp.posterior.woodbury_chol``, ``gp.kern.kernx/kernp.{variance,
lengthscale, active_dims[, power]}``.  ``GPState.from_gp`` reads those from any object exposing them
(a real GPy model, or ``ExactGP`` below).  ``ExactGP`` is the host-side exact-inference setup used when
a surrogate is built from ``(Z, Y, hyp)`` without GPy: one Cholesky per GP with LAPACK -- setup, not
the per-candidate path.
"""
import ctypes as C

import numpy as np
from scipy.linalg import lapack

from . import _lib, kernels
from ._lib import check, dptr


class ProductKernel:
    """``kernx * kernp`` (gp_model.py:123)."""

    def __init__(self, kernx, kernp):
        self.kernx, self.kernp = kernx, kernp

    def K(self, X, X2=None):
        return self.kernx.K(X, X2) * self.kernp.K(X, X2)

    def get_params(self):
        return np.r_[self.kernx.get_params(), self.kernp.get_params()]

    def set_params(self, v):
        k = self.kernx.set_params(v)
        return k + self.kernp.set_params(v[k:])


class _Posterior:
    def __init__(self, L, alpha):
        self.woodbury_chol = L
        self.woodbury_vector = alpha


class ExactGP:
    """Exact Gaussian-likelihood GP regression, zero mean (what GPy's GPRegression computes).

    Ky = K + (noise + 1e-8) I, L = chol(Ky), alpha = Ky^-1 y   (GPy ExactGaussianInference).
    ``gp[:]`` reads / writes the flat hyper-parameter vector in GPy's order
    ``[kernx.variance, kernx.lengthscale.., kernp.variance, kernp.lengthscale.., noise]`` (gp_model.py:142,180).
    """

    def __init__(self, Z, Y, kern, noise_var=1.0):
        self.X = np.ascontiguousarray(Z, dtype=float)
        self.Y = np.ascontiguousarray(Y, dtype=float).reshape(len(self.X), 1)
        self.kern = kern
        self.noise_var = float(noise_var)
        self._predictive_variable = self.X
        self.posterior = None
        self.update()

    def __getitem__(self, key):
        assert key == slice(None)
        return np.r_[self.kern.get_params(), self.noise_var]

    def __setitem__(self, key, v):
        assert key == slice(None)
        v = np.asarray(v, dtype=float)
        k = self.kern.set_params(v)
        assert k + 1 == len(v), 'hyper-parameter vector has the wrong length'
        self.noise_var = float(v[k])
        self.update()

    def update(self):
        Ky = self.kern.K(self.X)
        Ky[np.diag_indices_from(Ky)] += self.noise_var + 1e-8
        L, info = lapack.dpotrf(Ky, lower=1)
        jitter = np.diag(Ky).mean() * 1e-6
        tries = 0
        while info != 0:                      # GPy jitchol
            assert tries < 5, 'covariance not positive definite, even with jitter'
            L, info = lapack.dpotrf(Ky + np.eye(len(Ky)) * jitter, lower=1)
            jitter *= 10
            tries += 1
        L = np.tril(L)
        alpha, info = lapack.dpotrs(L, self.Y, lower=1)
        assert info == 0
        self.posterior = _Posterior(L, alpha)


class DeviceExactGP:
    """Exact GP regression whose inference runs on the B200 (``gpd_gp_build``): covariance matrix,
    blocked Cholesky with GPy's jitchol retries and alpha are formed in device memory, so neither a host LAPACK
    factorisation nor an n x n upload happens per (re)training.  Same surface as ``ExactGP`` / GPy's GPRegression as
    far as the reference reads it (``gp[:]``, ``kern``, ``_predictive_variable``, ``posterior``); ``posterior`` is
    downloaded lazily (one extra device build) because the hot path never needs it on the host."""
    built_on_device = True

    def __init__(self, Z, Y, kern, noise_var=1.0, device=0):
        self.X = np.ascontiguousarray(Z, dtype=float)
        self.Y = np.ascontiguousarray(Y, dtype=float).reshape(len(self.X), 1)
        self.kern = kern
        self.noise_var = float(noise_var)
        self._predictive_variable = self.X
        self._device = device
        self._posterior = None
        self.jitter = None            # jitter the last device factorisation needed (0.0 = none)

    def __getitem__(self, key):
        assert key == slice(None)
        return np.r_[self.kern.get_params(), self.noise_var]

    def __setitem__(self, key, v):
        assert key == slice(None)
        v = np.asarray(v, dtype=float)
        k = self.kern.set_params(v)
        assert k + 1 == len(v), 'hyper-parameter vector has the wrong length'
        self.noise_var = float(v[k])
        self._posterior = None

    def _dims(self):
        dim_p = len(np.asarray(self.kern.kernp.active_dims))
        return self.X.shape[1] - dim_p, dim_p

    def build(self, ctx, dim_x=None, dim_p=None, want_factor=False):
        """Run the inference on the device -> gpd_gp* (owned by the caller)."""
        if dim_x is None:
            dim_x, dim_p = self._dims()
        st = GPState.from_gp(self, dim_x, dim_p, with_posterior=False)
        h, alpha, L, self.jitter = st.build(ctx, self.Y[:, 0], self.noise_var, want_factor=want_factor)
        if want_factor:
            self._posterior = _Posterior(np.tril(L), alpha[:, None])
        return h

    @property
    def posterior(self):
        if self._posterior is None:
            from .context import get_context
            ctx = get_context(self._device)
            ctx.lib.gpd_gp_free(self.build(ctx, want_factor=True))
        return self._posterior


class SparseGP(ExactGP):
    """Sparse GP regression with inducing inputs (what GPy's SparseGPRegression / VarDTC computes): the host
    setup behind ``SparseGPModel`` (GPdoemd/models/sparse_gp_model.py:72).  ``gp[:]`` = inducing inputs
    (row-major), kernel parameters, noise.  The posterior is (woodbury_vector, woodbury_inv) on the inducing
    inputs, which are ``_predictive_variable`` exactly as the reference's hooks expect."""
    is_sparse = True

    def __init__(self, X, Y, kern, num_inducing=10, Z=None, rng=None, noise_var=1.0):
        X = np.ascontiguousarray(X, dtype=float)
        if Z is None:
            rng = np.random.default_rng(0) if rng is None else rng
            Z = X[rng.permutation(len(X))[:min(num_inducing, len(X))]].copy()
        self.Z = np.ascontiguousarray(Z, dtype=float)
        self.num_inducing = len(self.Z)
        super().__init__(X, Y, kern, noise_var)

    def __getitem__(self, key):
        assert key == slice(None)
        return np.r_[self.Z.ravel(), self.kern.get_params(), self.noise_var]

    def __setitem__(self, key, v):
        assert key == slice(None)
        v = np.asarray(v, dtype=float)
        nz = self.Z.size
        self.Z = v[:nz].reshape(self.Z.shape).copy()
        k = self.kern.set_params(v[nz:])
        assert nz + k + 1 == len(v), 'hyper-parameter vector has the wrong length'
        self.noise_var = float(v[nz + k])
        self.update()

    @staticmethod
    def _chol(A):
        L, info = lapack.dpotrf(np.ascontiguousarray(A), lower=1)
        jitter, tries = np.diag(A).mean() * 1e-6, 0
        while info != 0:
            assert tries < 5, 'matrix not positive definite, even with jitter'
            L, info = lapack.dpotrf(A + np.eye(len(A)) * jitter, lower=1)
            jitter *= 10
            tries += 1
        return np.tril(L)

    def update(self):
        m = len(self.Z)
        beta = 1.0 / max(self.noise_var, 1e-6)
        Kmm = self.kern.K(self.Z)
        Kmm[np.diag_indices(m)] += 1e-8
        Lm = self._chol(Kmm)
        psi1 = self.kern.K(self.X, self.Z)
        half = psi1 * np.sqrt(beta)
        LmInv, _ = lapack.dtrtri(Lm, lower=1)
        A = LmInv.dot(half.T.dot(half).dot(LmInv.T))
        LB = self._chol(np.eye(m) + A)
        rhs = psi1.T.dot(beta * self.Y)
        t, _ = lapack.dtrtrs(Lm, rhs, lower=1, trans=0)
        t, _ = lapack.dtrtrs(LB, t, lower=1, trans=0)
        t, _ = lapack.dtrtrs(LB, t, lower=1, trans=1)
        wv, _ = lapack.dtrtrs(Lm, t, lower=1, trans=1)
        Bi, _ = lapack.dpotri(LB, lower=1)
        Bi = -(np.tril(Bi) + np.tril(Bi, -1).T)
        Bi[np.diag_indices(m)] += 1.0
        t, _ = lapack.dtrtrs(Lm, Bi, lower=1, trans=1)
        W, _ = lapack.dtrtrs(Lm, t.T, lower=1, trans=1)
        W = W.T
        self.posterior = _SparsePosterior(0.5 * (W + W.T), wv)
        self._predictive_variable = self.Z


class _SparsePosterior:
    def __init__(self, woodbury_inv, woodbury_vector):
        self.woodbury_inv = woodbury_inv
        self.woodbury_vector = woodbury_vector


def lower_root_of(W):
    """Lower-triangular T with W = T^T T for a symmetric positive (semi-)definite W: the Cholesky factor of the
    index-reversed matrix, reversed back.  |T k|^2 = k^T W k, so T plays the role L^-1 plays for an exact GP."""
    n = len(W)
    Wr = np.ascontiguousarray(W[::-1, ::-1])
    scale = max(np.diag(W).mean(), 1e-300)
    jitter, info = 0.0, 1
    for _ in range(8):
        G, info = lapack.dpotrf(Wr + np.eye(n) * jitter, lower=1)
        if info == 0:
            break
        jitter = scale * 1e-14 if jitter == 0.0 else jitter * 100
    if info != 0:
        lam = np.linalg.eigvalsh(0.5 * (W + W.T))
        raise AssertionError('woodbury_inv is not positive semi-definite (eigenvalues %.3g .. %.3g): the sparse posterior '
                             'is numerically meaningless, K(Z, Z) of the inducing inputs is too ill-conditioned for its '
                             'hyper-parameters (the reference would return k** - k^T W k of this indefinite W, possibly '
                             'negative variances)' % (lam[0], lam[-1]))
    return np.ascontiguousarray(np.tril(G)[::-1, ::-1].T)


class GPState:
    """Plain-array snapshot of one trained GP in the reference's transformed units."""

    def __init__(self, kernel, dim_x, dim_p, x_active, var_x, ls_x, var_p, ls_p, Z, alpha=None, L=None,
                 power_x=2.0, power_p=2.0, factor_kind=0):
        self.kernel = kernel
        self.dim_x, self.dim_p = int(dim_x), int(dim_p)
        self.x_active = np.ascontiguousarray(x_active, dtype=np.int32)
        self.var_x, self.var_p = float(var_x), float(var_p)
        self.ls_x = _lib.as_f64(np.ravel(ls_x))
        self.ls_p = _lib.as_f64(np.ravel(ls_p))
        self.power_x, self.power_p = float(power_x), float(power_p)
        self.factor_kind = int(factor_kind)   # 0: L = chol(K + noise); 1: L = T with woodbury_inv = T^T T
        self.Z = _lib.as_f64(Z)
        self.alpha = None if alpha is None else _lib.as_f64(np.ravel(alpha))
        self.L = None if L is None else _lib.as_f64(L)
        n = len(self.Z)
        assert self.Z.shape == (n, self.dim_x + self.dim_p)
        assert (self.alpha is None) == (self.L is None)
        assert self.alpha is None or (self.alpha.shape == (n,) and self.L.shape == (n, n))
        assert len(self.ls_x) == len(self.x_active) and len(self.ls_p) == self.dim_p

    @classmethod
    def from_gp(cls, gp, dim_x, dim_p, with_posterior=True):
        kx, kp = gp.kern.kernx, gp.kern.kernp
        name = kernels.kernel_name_of(kx)
        assert kernels.kernel_name_of(kp) == name, 'both product parts share one class (gp_model.py:60)'
        p_dims = np.asarray(kp.active_dims, dtype=int)
        assert np.array_equal(p_dims, np.arange(dim_x, dim_x + dim_p)), 'kernp acts on columns D..D+P-1'
        sparse = bool(getattr(gp, 'is_sparse', False)) or 'Sparse' in type(gp).__name__
        if not with_posterior:   # training set + hyper-parameters only: the posterior is formed by gpd_gp_build
            assert not sparse, 'the device build covers exact GPs'
            factor, kind, alpha = None, 0, None
        elif sparse:     # sparse_gp_model.py: posterior lives on the inducing inputs, only woodbury_inv is meaningful
            factor, kind = lower_root_of(np.asarray(gp.posterior.woodbury_inv, dtype=float)), 1
        else:
            factor, kind = np.asarray(gp.posterior.woodbury_chol), 0
        if with_posterior:
            alpha = np.asarray(gp.posterior.woodbury_vector)[:, 0]
        return cls(name, dim_x, dim_p, np.asarray(kx.active_dims, dtype=int),
                   float(np.ravel(kx.variance)[0]), np.ravel(kx.lengthscale),
                   float(np.ravel(kp.variance)[0]), np.ravel(kp.lengthscale),
                   np.asarray(gp._predictive_variable), alpha,
                   factor, factor_kind=kind,
                   power_x=float(np.ravel(getattr(kx, 'power', 2.0))[0]),
                   power_p=float(np.ravel(getattr(kp, 'power', 2.0))[0]))

    # ---- checkpoint of the exported state (the (Z, alpha, L, theta) snapshot is all prediction needs) ----
    _ARRAYS = ('x_active', 'ls_x', 'ls_p', 'Z', 'alpha', 'L')
    _SCALARS = ('kernel', 'dim_x', 'dim_p', 'var_x', 'var_p', 'power_x', 'power_p', 'factor_kind')

    def to_dict(self, prefix=''):
        assert self.alpha is not None, 'nothing to save: the posterior lives on the device only'
        d = {prefix + k: getattr(self, k) for k in self._ARRAYS}
        d.update({prefix + k: np.asarray(getattr(self, k)) for k in self._SCALARS})
        return d

    @classmethod
    def from_dict(cls, d, prefix=''):
        g = lambda k: d[prefix + k]
        return cls(str(g('kernel')), int(g('dim_x')), int(g('dim_p')), g('x_active'), float(g('var_x')), g('ls_x'),
                   float(g('var_p')), g('ls_p'), g('Z'), g('alpha'), g('L'), power_x=float(g('power_x')),
                   power_p=float(g('power_p')), factor_kind=int(g('factor_kind')))

    def _desc(self):
        d = _lib.GpDesc()
        d.kernel = _lib.KERNEL_IDS[self.kernel]
        d.n = len(self.Z)
        d.dim_x, d.dim_p = self.dim_x, self.dim_p
        d.n_x_active = len(self.x_active)
        d.x_active = self.x_active.ctypes.data_as(_lib.c_int32_p)
        d.var_x, d.ls_x, d.power_x = self.var_x, dptr(self.ls_x), self.power_x
        d.var_p, d.ls_p, d.power_p = self.var_p, dptr(self.ls_p), self.power_p
        d.Z, d.alpha, d.L = dptr(self.Z), dptr(self.alpha), dptr(self.L)
        d.factor_kind = self.factor_kind
        return d

    def upload(self, ctx):
        """-> opaque device handle (gpd_gp*); the factor is inverted and packed on the device."""
        assert self.alpha is not None, 'no posterior: use build()'
        d = self._desc()
        h = C.c_void_p()
        check(ctx.lib.gpd_gp_upload(ctx.handle, C.byref(d), C.byref(h)))
        return h

    def build(self, ctx, y, noise_var, want_factor=False):
        """Exact-GP inference on the device (gpd_gp_build) -> (gpd_gp*, alpha, L or None, jitter)."""
        d = self._desc()
        n = len(self.Z)
        y = _lib.as_f64(np.ravel(y))
        assert y.shape == (n,)
        alpha = np.empty(n)
        L = np.empty((n, n)) if want_factor else None
        jit = C.c_double(0.0)
        h = C.c_void_p()
        check(ctx.lib.gpd_gp_build(ctx.handle, C.byref(d), dptr(y), float(noise_var), C.byref(h), dptr(alpha), dptr(L),
                                   C.byref(jit)))
        return h, alpha, L, jit.value
