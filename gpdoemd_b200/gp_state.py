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


class GPState:
    """Plain-array snapshot of one trained GP in the reference's transformed units."""

    def __init__(self, kernel, dim_x, dim_p, x_active, var_x, ls_x, var_p, ls_p, Z, alpha, L,
                 power_x=2.0, power_p=2.0):
        self.kernel = kernel
        self.dim_x, self.dim_p = int(dim_x), int(dim_p)
        self.x_active = np.ascontiguousarray(x_active, dtype=np.int32)
        self.var_x, self.var_p = float(var_x), float(var_p)
        self.ls_x = _lib.as_f64(np.ravel(ls_x))
        self.ls_p = _lib.as_f64(np.ravel(ls_p))
        self.power_x, self.power_p = float(power_x), float(power_p)
        self.Z = _lib.as_f64(Z)
        self.alpha = _lib.as_f64(np.ravel(alpha))
        self.L = _lib.as_f64(L)
        n = len(self.Z)
        assert self.Z.shape == (n, self.dim_x + self.dim_p)
        assert self.alpha.shape == (n,) and self.L.shape == (n, n)
        assert len(self.ls_x) == len(self.x_active) and len(self.ls_p) == self.dim_p

    @classmethod
    def from_gp(cls, gp, dim_x, dim_p):
        kx, kp = gp.kern.kernx, gp.kern.kernp
        name = kernels.kernel_name_of(kx)
        assert kernels.kernel_name_of(kp) == name, 'both product parts share one class (gp_model.py:60)'
        p_dims = np.asarray(kp.active_dims, dtype=int)
        assert np.array_equal(p_dims, np.arange(dim_x, dim_x + dim_p)), 'kernp acts on columns D..D+P-1'
        return cls(name, dim_x, dim_p, np.asarray(kx.active_dims, dtype=int),
                   float(np.ravel(kx.variance)[0]), np.ravel(kx.lengthscale),
                   float(np.ravel(kp.variance)[0]), np.ravel(kp.lengthscale),
                   np.asarray(gp._predictive_variable), np.asarray(gp.posterior.woodbury_vector)[:, 0],
                   np.asarray(gp.posterior.woodbury_chol),
                   power_x=float(np.ravel(getattr(kx, 'power', 2.0))[0]),
                   power_p=float(np.ravel(getattr(kp, 'power', 2.0))[0]))

    def upload(self, ctx):
        """-> opaque device handle (gpd_gp*); the factor is inverted and packed on the device."""
        d = _lib.GpDesc()
        d.kernel = _lib.KERNEL_IDS[self.kernel]
        d.n = len(self.Z)
        d.dim_x, d.dim_p = self.dim_x, self.dim_p
        d.n_x_active = len(self.x_active)
        d.x_active = self.x_active.ctypes.data_as(_lib.c_int32_p)
        d.var_x, d.ls_x, d.power_x = self.var_x, dptr(self.ls_x), self.power_x
        d.var_p, d.ls_p, d.power_p = self.var_p, dptr(self.ls_p), self.power_p
        d.Z, d.alpha, d.L = dptr(self.Z), dptr(self.alpha), dptr(self.L)
        h = C.c_void_p()
        check(ctx.lib.gpd_gp_upload(ctx.handle, C.byref(d), C.byref(h)))
        return h
