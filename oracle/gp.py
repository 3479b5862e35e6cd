"""Restatement of the GPy exact-GP regression pieces the reference calls (TEST INFRASTRUCTURE ONLY).

Reference call sites (GPdoemd/models/gp_model.py): ``GPRegression(Zr, Yr, kernx*kernp)`` :123,
``gp[:] = hyp`` :142, ``predict_noiseless`` :199, ``predictive_gradients`` :217,248,
``posterior.woodbury_vector`` :232, ``posterior.woodbury_inv`` :263, ``_predictive_variable``
:231,262.  GPy algorithms restated (GPy 1.9.x): ``ExactGaussianInference.inference``
(Ky = K + (noise + 1e-8) I, ``pdinv``/``jitchol``, ``dpotrs``), ``PosteriorExact._raw_predict``
(``dtrtrs`` against the Cholesky factor), ``GP.predictive_gradients``.  SURVEY.md Appendix A.3-A.5.
"""
import numpy as np
from scipy.linalg import lapack

from .gpy_kernels import Prod


class Posterior:
    def __init__(self, L, alpha):
        self.woodbury_chol = L
        self.woodbury_vector = alpha
        self._woodbury_inv = None

    @property
    def woodbury_inv(self):
        # GPy pdinv: Ai, _ = dpotri(L, lower=1); symmetrify(Ai)   (computed lazily, like GPy)
        if self._woodbury_inv is None:
            Ai, info = lapack.dpotri(self.woodbury_chol, lower=1)
            assert info == 0
            Ai = np.tril(Ai)
            self._woodbury_inv = Ai + np.tril(Ai, -1).T
        return self._woodbury_inv


class GPRegression:
    """Exact GP with Gaussian likelihood, zero mean, no normaliser (GPy defaults)."""

    def __init__(self, Z, Y, kern, noise_var=1.0):
        assert isinstance(kern, Prod)
        self.X = np.ascontiguousarray(Z, dtype=float)
        self.Y = np.ascontiguousarray(Y, dtype=float).reshape(len(Z), -1)
        self.kern = kern
        self.noise_var = float(noise_var)        # GPy Gaussian likelihood default variance = 1.0
        self._predictive_variable = self.X
        self.output_dim = self.Y.shape[1]
        self.posterior = None
        self.update()

    # flat parameter vector, GPy order: kernx.variance, kernx.lengthscale, kernp.variance,
    # kernp.lengthscale, Gaussian_noise.variance  (gp_model.py:142,180)
    def get_params(self):
        return np.r_[self.kern.get_params(), self.noise_var]

    def set_params(self, v):
        v = np.asarray(v, dtype=float)
        k = self.kern.set_params(v)
        self.noise_var = float(v[k])
        assert k + 1 == len(v)
        self.update()

    def update(self):
        # ExactGaussianInference.inference
        K = self.kern.K(self.X)
        Ky = K.copy()
        Ky[np.diag_indices_from(Ky)] += self.noise_var + 1e-8
        L, info = lapack.dpotrf(np.ascontiguousarray(Ky), lower=1)
        if info != 0:   # GPy jitchol: add diag.mean()*1e-6*10^k jitter until it factorises
            jitter = np.diag(Ky).mean() * 1e-6
            for _ in range(5):
                L, info = lapack.dpotrf(Ky + np.eye(len(Ky)) * jitter, lower=1)
                if info == 0:
                    break
                jitter *= 10
            assert info == 0, 'not positive definite, even with jitter'
        L = np.tril(L)
        alpha, info = lapack.dpotrs(L, self.Y, lower=1)
        assert info == 0
        self.posterior = Posterior(L, alpha)

    def predict_noiseless(self, Xnew, chunk=4096):
        """PosteriorExact._raw_predict, full_cov=False, no likelihood noise -> (N,1), (N,1)."""
        Xnew = np.asarray(Xnew, dtype=float)
        N = len(Xnew)
        mu = np.empty((N, self.output_dim))
        var = np.empty((N, 1))
        L = self.posterior.woodbury_chol
        for s in range(0, N, chunk):
            Xc = Xnew[s:s + chunk]
            Kx = self.kern.K(self._predictive_variable, Xc)
            mu[s:s + chunk] = np.dot(Kx.T, self.posterior.woodbury_vector)
            Kxx = self.kern.Kdiag(Xc)
            tmp, info = lapack.dtrtrs(L, Kx, lower=1)
            var[s:s + chunk, 0] = Kxx - np.square(tmp).sum(0)
        return mu, var

    @classmethod
    def from_posterior(cls, Z, Y, kern, noise_var, L, alpha):
        """Adopt an already computed posterior (woodbury_chol, woodbury_vector) instead of refactorising: lets the
        n = 4000 parity cases feed the SAME (alpha, L) to the oracle and to the CUDA path, as GPModel.from_reference
        does in the other direction.  Hyper-parameters must already be set on ``kern``."""
        self = cls.__new__(cls)
        self.X = np.ascontiguousarray(Z, dtype=float)
        self.Y = np.ascontiguousarray(Y, dtype=float).reshape(len(Z), -1)
        self.kern = kern
        self.noise_var = float(noise_var)
        self._predictive_variable = self.X
        self.output_dim = self.Y.shape[1]
        self.posterior = Posterior(np.asarray(L, dtype=float), np.asarray(alpha, dtype=float).reshape(len(Z), -1))
        return self

    def predictive_gradients(self, Xnew, chunk=2048, want_var=True):
        """GP.predictive_gradients -> mean_jac (N,Q,output_dim), var_jac (N,Q).

        want_var=False skips the variance Jacobian: GPy always computes it (it needs woodbury_inv = K^-1, an n^3
        dpotri per GP) and the reference's d_mu_d_p discards it (gp_model.py:217); the large-n parity cases that
        only need the mean Jacobian switch it off to stay within seconds."""
        Xnew = np.asarray(Xnew, dtype=float)
        N, Q = Xnew.shape
        mean_jac = np.empty((N, Q, self.output_dim))
        var_jac = np.empty((N, Q))
        kern, Xt = self.kern, self._predictive_variable
        for s in range(0, N, chunk):
            Xc = Xnew[s:s + chunk]
            for i in range(self.output_dim):
                mean_jac[s:s + chunk, :, i] = kern.gradients_X(
                    self.posterior.woodbury_vector[:, i:i + 1].T, Xc, Xt)
            if not want_var:
                var_jac[s:s + chunk] = np.nan
                continue
            # d k(x*,x*)/dx* term: stationary => zero (GPy evaluates kern.gradients_X(eye, Xnew))
            dv_dX = np.zeros(Xc.shape)
            alpha = -2.0 * np.dot(kern.K(Xc, Xt), self.posterior.woodbury_inv)
            dv_dX += kern.gradients_X(alpha, Xc, Xt)
            var_jac[s:s + chunk] = dv_dX
        return mean_jac, var_jac
