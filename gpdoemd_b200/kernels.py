"""Host-side kernel classes with the reference's constructor ``(d, drange, name)``.

Mirrors ``GPdoemd/kernels/stationary.py:33-99`` (RBF, Exponential, Matern32, Matern52, Cosine, RatQuad:
ARD stationary kernels on ``active_dims = drange`` with ``variance``, ``lengthscale`` and the radial
derivatives ``dK_dr``/``dK2_drdr``).  The reference subclasses GPy, which is not a dependency here;
these classes only carry hyper-parameters to the device and build the training covariance once on
the host when a surrogate is (re)fitted from ``(Z, Y, hyp)`` -- the per-candidate evaluation of
K(x*, X) and its derivatives runs in CUDA (csrc/kernels_eval.cu).
"""
import numpy as np


class Stationary:
    kernel_name = None

    def __init__(self, d, drange, name):
        self.input_dim = int(d)
        self.active_dims = np.asarray(list(drange), dtype=int)
        assert len(self.active_dims) == self.input_dim
        self.name = name
        self.variance = 1.0
        self.lengthscale = np.ones(self.input_dim)

    # flat hyper-parameter order of GPy: variance, lengthscale[...] (then power for RatQuad)
    def get_params(self):
        return np.r_[self.variance, self.lengthscale]

    def set_params(self, v):
        v = np.asarray(v, dtype=float)
        self.variance = float(v[0])
        self.lengthscale = np.array(v[1:1 + self.input_dim], dtype=float)
        return 1 + self.input_dim

    def _r(self, X, X2=None):
        A = np.asarray(X, dtype=float)[:, self.active_dims] / self.lengthscale
        B = A if X2 is None else np.asarray(X2, dtype=float)[:, self.active_dims] / self.lengthscale
        # same expanded form as GPy's _unscaled_dist, so a host refit reproduces GPy's factor
        r2 = -2.0 * A.dot(B.T) + (np.sum(A * A, 1)[:, None] + np.sum(B * B, 1)[None, :])
        if X2 is None:
            np.fill_diagonal(r2, 0.0)
        return np.sqrt(np.clip(r2, 0, np.inf))

    def K(self, X, X2=None):
        return self.K_of_r(self._r(X, X2))

    def Kdiag(self, X):
        return np.full(len(X), self.variance)


class RBF(Stationary):
    kernel_name = 'RBF'

    def K_of_r(self, r):
        return self.variance * np.exp(-0.5 * np.square(r))

    def dK_dr(self, r):
        return -r * self.K_of_r(r)

    def dK2_drdr(self, r):
        return (np.square(r) - 1.0) * self.K_of_r(r)


class Exponential(Stationary):
    kernel_name = 'Exponential'

    def K_of_r(self, r):
        return self.variance * np.exp(-r)

    def dK_dr(self, r):
        return -self.K_of_r(r)

    def dK2_drdr(self, r):
        return self.K_of_r(r)


class Matern32(Stationary):
    kernel_name = 'Matern32'

    def K_of_r(self, r):
        a = np.sqrt(3.0) * r
        return self.variance * (1.0 + a) * np.exp(-a)

    def dK_dr(self, r):
        return -3.0 * self.variance * r * np.exp(-np.sqrt(3.0) * r)

    def dK2_drdr(self, r):
        a = np.sqrt(3.0) * r
        return 3.0 * self.variance * (a - 1.0) * np.exp(-a)


class Matern52(Stationary):
    kernel_name = 'Matern52'

    def K_of_r(self, r):
        a = np.sqrt(5.0) * r
        return self.variance * (1.0 + a + 5.0 / 3.0 * np.square(r)) * np.exp(-a)

    def dK_dr(self, r):
        a = np.sqrt(5.0) * r
        return -5.0 / 3.0 * self.variance * r * (1.0 + a) * np.exp(-a)

    def dK2_drdr(self, r):
        a = np.sqrt(5.0) * r
        return 5.0 / 3.0 * self.variance * (a * a - a - 1.0) * np.exp(-a)


class Cosine(Stationary):
    kernel_name = 'Cosine'

    def K_of_r(self, r):
        return self.variance * np.cos(r)

    def dK_dr(self, r):
        return -self.variance * np.sin(r)

    def dK2_drdr(self, r):
        return -self.K_of_r(r)


class RatQuad(Stationary):
    kernel_name = 'RatQuad'

    def __init__(self, d, drange, name):
        super().__init__(d, drange, name)
        self.power = 2.0

    def get_params(self):
        return np.r_[self.variance, self.lengthscale, self.power]

    def set_params(self, v):
        k = super().set_params(v)
        self.power = float(v[k])
        return k + 1

    def K_of_r(self, r):
        return self.variance * np.exp(-self.power * np.log1p(0.5 * np.square(r)))

    def dK_dr(self, r):
        return -self.variance * self.power * r * np.exp(-(self.power + 1.0) * np.log1p(0.5 * np.square(r)))

    def dK2_drdr(self, r):
        r2 = np.square(r)
        u = np.log1p(0.5 * r2)
        return (-self.variance * self.power * np.exp(-(self.power + 1.0) * u)
                + self.variance * self.power * (self.power + 1.0) * r2 * np.exp(-(self.power + 2.0) * u))


BY_NAME = {k.kernel_name: k for k in (RBF, Exponential, Matern32, Matern52, Cosine, RatQuad)}


def kernel_name_of(obj):
    """Kernel id name of a class / instance (these classes or GPy/GPdoemd kernel classes) or a name string."""
    if isinstance(obj, str):
        if obj in BY_NAME:
            return obj
        raise ValueError('unsupported kernel %r (supported: %s)' % (obj, ', '.join(BY_NAME)))
    cls = obj if isinstance(obj, type) else type(obj)
    for c in cls.__mro__:
        if c.__name__ in BY_NAME:
            return c.__name__
    raise ValueError('unsupported kernel class %r (supported: %s)' % (cls.__name__, ', '.join(BY_NAME)))
