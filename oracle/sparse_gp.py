"""Restatement of GPy's SparseGPRegression as used by GPdoemd's SparseGPModel (TEST INFRASTRUCTURE ONLY).

Reference call site: ``GPdoemd/models/sparse_gp_model.py:72``  ``SparseGPRegression(Zr, Yr, kernx*kernp,
num_inducing=numi)``; all prediction / derivative hooks are inherited from GPModel (gp_model.py:184-280) and
read ``gp._predictive_variable`` (= the inducing inputs), ``gp.posterior.woodbury_vector`` and
``gp.posterior.woodbury_inv``.

GPy algorithms restated (GPy 1.9.x): ``SparseGPRegression.__init__`` (inducing inputs = a random subset of
the training inputs), ``VarDTC.inference`` (``const_jitter = 1e-8``), ``Posterior._raw_predict`` with a 2-D
``woodbury_inv`` (var = Kxx - sum((W Kx) * Kx)), ``GP.predictive_gradients``.  Flat parameter order of
``gp[:]``: inducing inputs (row-major), kernel parameters, Gaussian noise variance.
PARITY WITH REAL GPy IS UNPINNED (GPy not installed); pinned by: inducing inputs = all training inputs
reproduces the exact GP (tests/test_oracle_sparse.py).
"""
import numpy as np
from scipy.linalg import lapack

from .gp import GPRegression


class SparsePosterior:
    def __init__(self, woodbury_inv, woodbury_vector, Lm):
        self.woodbury_inv = woodbury_inv
        self.woodbury_vector = woodbury_vector
        self.K_chol = Lm


class SparseGPRegression(GPRegression):
    is_sparse = True

    def __init__(self, X, Y, kern, num_inducing=10, Z=None, rng=None, noise_var=1.0):
        X = np.ascontiguousarray(X, dtype=float)
        if Z is None:
            rng = np.random.default_rng(0) if rng is None else rng      # GPy: np.random.permutation (unseeded)
            i = rng.permutation(len(X))[:min(num_inducing, len(X))]
            Z = X[i].copy()
        self.Z = np.ascontiguousarray(Z, dtype=float)
        self.num_inducing = len(self.Z)
        super().__init__(X, Y, kern, noise_var)

    def get_params(self):
        return np.r_[self.Z.ravel(), self.kern.get_params(), self.noise_var]

    def set_params(self, v):
        v = np.asarray(v, dtype=float)
        nz = self.Z.size
        self.Z = v[:nz].reshape(self.Z.shape).copy()
        k = self.kern.set_params(v[nz:])
        self.noise_var = float(v[nz + k])
        assert nz + k + 1 == len(v)
        self.update()

    def update(self):
        # VarDTC.inference, homoscedastic noise, uncertain_inputs = False
        Xd, Y, Zi = self.X, self.Y, self.Z
        m = len(Zi)
        beta = 1.0 / np.fmax(self.noise_var, 1e-6)
        Kmm = self.kern.K(Zi).copy()
        Kmm[np.diag_indices(m)] += 1e-8
        Lm = _jitchol(Kmm)
        psi1 = self.kern.K(Xd, Zi)
        tmp = psi1 * np.sqrt(beta)
        psi2_beta = tmp.T.dot(tmp)
        LmInv, _ = lapack.dtrtri(Lm, lower=1)
        A = LmInv.dot(psi2_beta.dot(LmInv.T))
        B = np.eye(m) + A
        LB = _jitchol(B)
        psi1Vf = psi1.T.dot(beta * Y)
        t1, _ = lapack.dtrtrs(Lm, psi1Vf, lower=1, trans=0)
        t2, _ = lapack.dtrtrs(LB, t1, lower=1, trans=0)
        t3, _ = lapack.dtrtrs(LB, t2, lower=1, trans=1)
        wv, _ = lapack.dtrtrs(Lm, t3, lower=1, trans=1)
        Bi, _ = lapack.dpotri(LB, lower=1)
        Bi = np.tril(Bi) + np.tril(Bi, -1).T
        Bi = -Bi
        Bi[np.diag_indices(m)] += 1.0                      # I - B^-1
        # backsub_both_sides(Lm, Bi, transpose='left') = Lm^-T Bi Lm^-1
        t4, _ = lapack.dtrtrs(Lm, Bi, lower=1, trans=1)
        W, _ = lapack.dtrtrs(Lm, t4.T, lower=1, trans=1)
        W = W.T
        self.posterior = SparsePosterior(0.5 * (W + W.T), wv, Lm)
        self._predictive_variable = self.Z

    def predict_noiseless(self, Xnew, chunk=4096):
        Xnew = np.asarray(Xnew, dtype=float)
        N = len(Xnew)
        mu = np.empty((N, self.output_dim))
        var = np.empty((N, 1))
        W = self.posterior.woodbury_inv
        for s in range(0, N, chunk):
            Xc = Xnew[s:s + chunk]
            Kx = self.kern.K(self._predictive_variable, Xc)
            mu[s:s + chunk] = np.dot(Kx.T, self.posterior.woodbury_vector)
            var[s:s + chunk, 0] = self.kern.Kdiag(Xc) - np.sum(np.dot(W.T, Kx) * Kx, 0)
        # GPy's base Posterior._raw_predict (the class SparseGPRegression uses) ends its full_cov=False branch with
        # `var = np.clip(var, 1e-15, np.inf)`; PosteriorExact (exact GPRegression, gp.py) does not.  Restated from
        # memory of GPy 1.9.x like the rest of this module: parity with real GPy is unpinned.
        return mu, np.clip(var, 1e-15, np.inf)


def _jitchol(A):
    L, info = lapack.dpotrf(np.ascontiguousarray(A), lower=1)
    jitter = np.diag(A).mean() * 1e-6
    tries = 0
    while info != 0:
        assert tries < 5, 'not positive definite, even with jitter'
        L, info = lapack.dpotrf(A + np.eye(len(A)) * jitter, lower=1)
        jitter *= 10
        tries += 1
    return np.tril(L)
