"""Restatement of the GPy ARD stationary kernels as used by GPdoemd (TEST INFRASTRUCTURE ONLY).

Follows ``GPdoemd/kernels/stationary.py:33-99`` (six thin subclasses of ``GPy.kern.*`` with ctor
``(d, drange, name)``, ``ARD=True`` and an added ``dK2_drdr``) and the GPy 1.9.x base class
``GPy/kern/src/stationary.py`` (``_scaled_dist``, ``_unscaled_dist``, ``K``, ``Kdiag``,
``gradients_X`` (pure-numpy branch), ``gradients_XX``) plus the ``active_dims`` slicing of
``GPy/kern/src/kernel_slice_operations.py``.  See SURVEY.md Appendix A.2.
"""
import numpy as np


class Stationary:
    """k(x, x') = variance * kappa(r),  r = sqrt(sum_q (x_q - x'_q)^2 / lengthscale_q^2)."""

    def __init__(self, d, drange, name, variance=1.0, lengthscale=None):
        # GPdoemd/kernels/stationary.py:34-36 -> GPy Stationary.__init__(input_dim, active_dims, ARD=True)
        self.input_dim = int(d)
        self.active_dims = np.asarray(list(drange), dtype=int)
        assert len(self.active_dims) == self.input_dim
        self.name = name
        self.variance = float(variance)
        self.lengthscale = (np.ones(self.input_dim) if lengthscale is None
                            else np.asarray(lengthscale, dtype=float).reshape(self.input_dim))

    # ---- radial profile (overridden per kernel) -------------------------------------------
    def K_of_r(self, r):
        raise NotImplementedError

    def dK_dr(self, r):
        raise NotImplementedError

    def dK2_drdr(self, r):
        raise NotImplementedError

    # ---- hyper-parameter vector in GPy's flat order [variance, lengthscale...] ------------
    def get_params(self):
        return np.r_[self.variance, self.lengthscale]

    def set_params(self, v):
        v = np.asarray(v, dtype=float)
        self.variance = float(v[0])
        self.lengthscale = v[1:1 + self.input_dim].copy()
        return 1 + self.input_dim

    num_params = property(lambda self: 1 + self.input_dim)

    # ---- distances: GPy Stationary._unscaled_dist / _scaled_dist ---------------------------
    def _slice(self, X):
        return None if X is None else np.asarray(X)[:, self.active_dims]

    @staticmethod
    def _unscaled_dist(X, X2=None):
        # GPy computes r^2 in the EXPANDED form |a|^2 + |b|^2 - 2 a.b and clips at 0.
        if X2 is None:
            Xsq = np.sum(np.square(X), 1)
            r2 = -2.0 * np.dot(X, X.T) + (Xsq[:, None] + Xsq[None, :])
            np.fill_diagonal(r2, 0.0)
            r2 = np.clip(r2, 0, np.inf)
            return np.sqrt(r2)
        X1sq = np.sum(np.square(X), 1)
        X2sq = np.sum(np.square(X2), 1)
        r2 = -2.0 * np.dot(X, X2.T) + (X1sq[:, None] + X2sq[None, :])
        r2 = np.clip(r2, 0, np.inf)
        return np.sqrt(r2)

    def _scaled_dist_sliced(self, Xs, X2s=None):
        # ARD branch of GPy Stationary._scaled_dist
        if X2s is not None:
            X2s = X2s / self.lengthscale
        return self._unscaled_dist(Xs / self.lengthscale, X2s)

    def _scaled_dist(self, X, X2=None):
        return self._scaled_dist_sliced(self._slice(X), self._slice(X2))

    def _inv_dist_sliced(self, Xs, X2s=None):
        dist = self._scaled_dist_sliced(Xs, X2s).copy()
        with np.errstate(divide='ignore'):
            return 1.0 / np.where(dist != 0.0, dist, np.inf)

    # ---- covariance ------------------------------------------------------------------------
    def K(self, X, X2=None):
        return self.K_of_r(self._scaled_dist(X, X2))

    def Kdiag(self, X):
        return np.full(len(X), self.variance)

    # ---- gradients w.r.t. the first argument (GPy _gradients_X_pure) -----------------------
    def gradients_X(self, dL_dK, X, X2=None):
        """sum_j dL_dK[i,j] * dk(X_i, X2_j)/dX_i  -> (N, full input dim), zeros off active dims."""
        X = np.asarray(X)
        Xs, X2s = self._slice(X), self._slice(X2)
        invdist = self._inv_dist_sliced(Xs, X2s)
        dL_dr = self.dK_dr(self._scaled_dist_sliced(Xs, X2s)) * dL_dK
        tmp = invdist * dL_dr
        if X2s is None:
            tmp = tmp + tmp.T
            X2s = Xs
        grad = np.empty(Xs.shape, dtype=np.float64)
        for q in range(self.input_dim):
            np.sum(tmp * (Xs[:, q][:, None] - X2s[:, q][None, :]), axis=1, out=grad[:, q])
        grad = grad / self.lengthscale ** 2
        out = np.zeros(X.shape)
        out[:, self.active_dims] = grad
        return out

    # ---- mixed second derivative (GPy Stationary.gradients_XX, cov=True) -------------------
    def gradients_XX(self, dL_dK, X, X2=None):
        """dL_dK[i,j] * d^2 k / dX_i dX2_j  = MINUS d^2 k / dX_i dX_i  -> (N, M, Qfull, Qfull).

        At r == 0 GPy substitutes (dK/dr)/r := -variance for every kernel.
        """
        X = np.asarray(X)
        Xs, X2s = self._slice(X), self._slice(X2)
        invdist = self._inv_dist_sliced(Xs, X2s)
        invdist2 = invdist ** 2
        r = self._scaled_dist_sliced(Xs, X2s)
        tmp1 = self.dK_dr(r) * invdist
        tmp2 = self.dK2_drdr(r) * invdist2
        l2 = np.ones(Xs.shape[1]) * self.lengthscale ** 2
        if X2s is None:
            X2s = Xs
            tmp1 = tmp1 - np.eye(Xs.shape[0]) * self.variance
        else:
            tmp1 = tmp1.copy()
            tmp1[invdist2 == 0.0] -= self.variance
        dist = Xs[:, None, :] - X2s[None, :, :]
        dist = dist[:, :, :, None] * dist[:, :, None, :]
        eye = np.eye(X2s.shape[1])
        grad = (((dL_dK * (tmp1 * invdist2 - tmp2))[:, :, None, None] * dist) / l2[None, None, :, None]
                - (dL_dK * tmp1)[:, :, None, None] * eye) / l2[None, None, None, :]
        Q = X.shape[1]
        out = np.zeros((X.shape[0], X2s.shape[0], Q, Q))
        out[np.ix_(range(X.shape[0]), range(X2s.shape[0]), self.active_dims, self.active_dims)] = grad
        return out


class RBF(Stationary):                       # stationary.py:33-36 ; GPy rbf.py
    kernel_id = 0

    def K_of_r(self, r):
        return self.variance * np.exp(-0.5 * r ** 2)

    def dK_dr(self, r):
        return -r * self.K_of_r(r)

    def dK2_drdr(self, r):
        return (r ** 2 - 1) * self.K_of_r(r)


class Exponential(Stationary):               # stationary.py:41-47 ; GPy stationary.py Exponential
    kernel_id = 1

    def K_of_r(self, r):
        return self.variance * np.exp(-r)

    def dK_dr(self, r):
        return -self.K_of_r(r)

    def dK2_drdr(self, r):
        return self.K_of_r(r)


class Matern32(Stationary):                  # stationary.py:52-59
    kernel_id = 2

    def K_of_r(self, r):
        return self.variance * (1.0 + np.sqrt(3.0) * r) * np.exp(-np.sqrt(3.0) * r)

    def dK_dr(self, r):
        return -3.0 * self.variance * r * np.exp(-np.sqrt(3.0) * r)

    def dK2_drdr(self, r):
        ar = np.sqrt(3.0) * r
        return 3.0 * self.variance * (ar - 1) * np.exp(-ar)


class Matern52(Stationary):                  # stationary.py:64-71
    kernel_id = 3

    def K_of_r(self, r):
        return self.variance * (1 + np.sqrt(5.0) * r + 5.0 / 3 * r ** 2) * np.exp(-np.sqrt(5.0) * r)

    def dK_dr(self, r):
        return self.variance * (10.0 / 3 * r - 5.0 * r - 5.0 * np.sqrt(5.0) / 3 * r ** 2) * np.exp(-np.sqrt(5.0) * r)

    def dK2_drdr(self, r):
        ar = np.sqrt(5) * r
        return 5.0 / 3.0 * self.variance * (ar ** 2 - ar - 1) * np.exp(-ar)


class Cosine(Stationary):                    # stationary.py:76-82
    kernel_id = 4

    def K_of_r(self, r):
        return self.variance * np.cos(r)

    def dK_dr(self, r):
        return -self.variance * np.sin(r)

    def dK2_drdr(self, r):
        return -self.K_of_r(r)


class RatQuad(Stationary):                   # stationary.py:87-99 ; GPy RatQuad(power=2.)
    kernel_id = 5

    def __init__(self, d, drange, name, variance=1.0, lengthscale=None, power=2.0):
        super().__init__(d, drange, name, variance, lengthscale)
        self.power = float(power)

    def get_params(self):
        # paramz link order in GPy RatQuad: variance, lengthscale (Stationary), then power
        return np.r_[self.variance, self.lengthscale, self.power]

    def set_params(self, v):
        k = super().set_params(v)
        self.power = float(v[k])
        return k + 1

    num_params = property(lambda self: 2 + self.input_dim)

    def K_of_r(self, r):
        r2 = np.square(r)
        return self.variance * np.exp(-self.power * np.log1p(r2 / 2.0))

    def dK_dr(self, r):
        r2 = np.square(r)
        return -self.variance * self.power * r * np.exp(-(self.power + 1) * np.log1p(r2 / 2.0))

    def dK2_drdr(self, r):
        r2 = np.square(r)
        lp = np.log1p(r2 / 2.0)
        a = (self.power + 1) * lp
        dr1 = -self.variance * self.power * np.exp(-a)
        a = a + lp
        dr2 = self.variance * self.power * (self.power + 1) * r2 * np.exp(-a)
        return dr1 + dr2


KERNELS = {c.__name__: c for c in (RBF, Exponential, Matern32, Matern52, Cosine, RatQuad)}


class Prod:
    """GPy ``kernx * kernp`` (``gp_model.py:123``): elementwise product of the two parts."""

    def __init__(self, kernx, kernp):
        self.kernx, self.kernp = kernx, kernp
        self.parts = [kernx, kernp]

    def K(self, X, X2=None):
        return self.kernx.K(X, X2) * self.kernp.K(X, X2)

    def Kdiag(self, X):
        return self.kernx.Kdiag(X) * self.kernp.Kdiag(X)

    def gradients_X(self, dL_dK, X, X2=None):
        # GPy Prod.gradients_X, two-part branch
        return (self.kernx.gradients_X(dL_dK * self.kernp.K(X, X2), X, X2)
                + self.kernp.gradients_X(dL_dK * self.kernx.K(X, X2), X, X2))

    def get_params(self):
        return np.r_[self.kernx.get_params(), self.kernp.get_params()]

    def set_params(self, v):
        k = self.kernx.set_params(v)
        return k + self.kernp.set_params(v[k:])
