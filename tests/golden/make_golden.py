"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE'S OWN CODE.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

The reference modules used here import without GPy: GPdoemd/design_criteria.py, GPdoemd/marginal/*,
GPdoemd/transform.py (loaded by file path so that GPdoemd/__init__.py, which imports GPy-dependent
models, is never executed).  GPdoemd/utils.py needs the numpy>=2 shim `list(map(...))`,
applied by patching np.vstack for the call only -- the reference source is not modified.
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = '/root/reference'
OUT = os.path.dirname(os.path.abspath(__file__))


def load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def notebook_inputs():
    """demos/design_criteria.ipynb cells 3-9: two analytic 2-output models on linspace(-2, 2, 250)."""
    def M1(x):
        mu = np.array([0.1 * x + 0.1 * x ** 2, -0.1 * x + 0.1 * np.sin(x ** 2)])
        s2 = np.array([[0.7 * (0.01 + mu[0] ** 2), 0], [0, 0.7 * (0.01 + mu[1] ** 2)]])
        return mu, s2

    def M2(x):
        mu = np.array([0.1 * x + 0.1 * x ** 2 + 0.2 * np.sin(0.2 * x ** 3), -0.1 * x + 0.1 * np.cos(x ** 2)])
        s2 = np.array([[0.7 * (0.01 + mu[0] ** 2), 0], [0, 0.7 * (0.01 + mu[1] ** 2)]])
        return mu, s2
    X = np.linspace(-2., 2., 250)
    M = np.zeros((250, 2, 2))
    S = np.zeros((250, 2, 2, 2))
    for i, x in enumerate(X):
        M[i, 0], S[i, 0] = M1(x)
        M[i, 1], S[i, 1] = M2(x)
    return M, S, np.array([0.04, 0.01])


class StubModel:
    """Closed-form duck-typed model in the spirit of tests/test_marginal/test_taylor_approximation.py:17-51."""

    def __init__(self, E, P, seed):
        rng = np.random.default_rng(seed)
        self.num_outputs, self.dim_p = E, P
        A = rng.normal(size=(P, P))
        self.Sigma = A.dot(A.T) / P * 0.1
        self.W = rng.normal(size=(E, P, 3))
        self.Q = rng.normal(size=(E, P, P, 3))
        self.Q = self.Q + self.Q.transpose(0, 2, 1, 3)
        self.R = rng.normal(size=(E, P, P, 3))
        self.R = self.R + self.R.transpose(0, 2, 1, 3)

    def predict(self, X):
        f = np.c_[np.sin(X[:, 0]), np.cos(X[:, 1]), X[:, 0] * X[:, 1]]
        M = np.stack([f.dot(np.arange(1, 4) * (e + 1.0)) for e in range(self.num_outputs)], axis=1)
        S = np.stack([0.05 + 0.01 * (e + 1) * f[:, 0] ** 2 for e in range(self.num_outputs)], axis=1)
        return M, S

    def _f(self, X):
        return np.c_[np.sin(X[:, 0]), np.cos(X[:, 1]), X[:, 0] * X[:, 1]]

    def d_mu_d_p(self, e, X):
        return self._f(X).dot(self.W[e].T)

    def d2_mu_d_p2(self, e, X):
        return np.einsum('nk,abk->nab', self._f(X), self.Q[e])

    def d2_s2_d_p2(self, e, X):
        return 0.01 * np.einsum('nk,abk->nab', self._f(X), self.R[e])


def main():
    dc = load('ref_design_criteria', 'GPdoemd/design_criteria.py')
    # the marginal modules use only numpy
    t1 = load('ref_taylor_first_order', 'GPdoemd/marginal/taylor_first_order.py')
    t2 = load('ref_taylor_second_order', 'GPdoemd/marginal/taylor_second_order.py')
    tr = load('ref_transform', 'GPdoemd/transform.py')
    ut = load('ref_utils', 'GPdoemd/utils.py')
    la = load('ref_laplace', 'GPdoemd/param_covar/laplace_approximation.py')
    out = {}

    # ---- criteria: notebook input (deterministic) ---------------------------------------------
    M, S, mv = notebook_inputs()
    for name in ('HR', 'BH', 'BF', 'AW', 'JR'):
        out['nb_' + name] = getattr(dc, name)(M.copy(), S.copy(), mv.copy())
    out['nb_mu'], out['nb_s2'], out['nb_noise'] = M, S, mv

    # ---- criteria: seeded random inputs, several (M, E), full covariances, pps, 2-D noise --------
    for (nM, E) in ((3, 2), (4, 1), (2, 3), (8, 8), (5, 4)):
        rng = np.random.default_rng(1000 + 10 * nM + E)
        N = 40
        mu = rng.normal(size=(N, nM, E))
        A = rng.normal(size=(N, nM, E, E))
        s2 = 0.05 * np.einsum('nmij,nmkj->nmik', A, A) + 0.01 * np.eye(E)
        B = rng.normal(size=(E, E))
        noise = 0.02 * np.eye(E) + 0.002 * B.dot(B.T)
        pps = rng.uniform(0.5, 1.5, nM)
        key = 'r%d_%d_' % (nM, E)
        out[key + 'mu'], out[key + 's2'], out[key + 'noise'], out[key + 'pps'] = mu, s2, noise, pps
        for name in ('HR', 'BH', 'BF', 'AW', 'JR'):
            out[key + name] = getattr(dc, name)(mu.copy(), s2.copy(), noise.copy(), pps.copy())

    # ---- criteria: the reference test's construction with the planted argmax 13 --------------------
    rng = np.random.default_rng(12345)
    N, nM, E = 15, 3, 2
    mu = rng.standard_normal((N, nM, E))
    s2 = 0.1 * rng.random((N, nM, E, E))
    s2 += s2.transpose(0, 1, 3, 2) + np.array([[np.eye(E)] * nM] * N)
    mu[13] = (1 + np.arange(nM * E).reshape((nM, E))) * 10
    s2[13] = np.array([0.001 * np.eye(E)] * nM)
    out['t13_mu'], out['t13_s2'] = mu, s2
    for name in ('HR', 'BH', 'BF', 'AW', 'JR'):
        d = getattr(dc, name)(mu.copy(), s2.copy(), 0.1 * np.eye(E), np.ones(nM) / nM)
        assert np.argmax(d) == 13
        out['t13_' + name] = d

    # ---- the BH / BF in-place mutation quirk (design_criteria.py:81,114) ------------------------------
    s2m = out['r3_2_s2'].copy()
    dc.BH(out['r3_2_mu'].copy(), s2m, out['r3_2_noise'].copy(), out['r3_2_pps'].copy())
    out['quirk_BH_s2_after'] = s2m

    # ---- Taylor marginals with the closed-form stub ---------------------------------------------------
    for (E, P) in ((1, 1), (2, 3), (3, 5)):
        st = StubModel(E, P, 7 * E + P)
        X = np.random.default_rng(5).uniform(-1, 1, (30, 2))
        M1_, S1_ = t1.taylor_first_order(st, X)
        M2_, S2_ = t2.taylor_second_order(st, X)
        key = 'ty%d_%d_' % (E, P)
        out[key + 'X'], out[key + 'M1'], out[key + 'S1'], out[key + 'M2'], out[key + 'S2'] = X, M1_, S1_, M2_, S2_

    # ---- Laplace approximation of the parameter covariance (param_covar/laplace_approximation.py) -------------
    for (E, P) in ((1, 1), (2, 3), (3, 5)):
        st = StubModel(E, P, 7 * E + P)
        X = np.random.default_rng(6).uniform(-1, 1, (12, 2))
        st.meas_noise_var = 0.01 * (1 + np.arange(E))
        out['la%d_%d_diag' % (E, P)] = la.laplace_approximation(st, X)
        B = np.random.default_rng(8).normal(size=(E, E))
        st.meas_noise_var = 0.01 * np.eye(E) + 0.002 * B.dot(B.T)
        out['la%d_%d_full' % (E, P)] = la.laplace_approximation(st, X)
        out['la%d_%d_X' % (E, P)] = X

    # ---- transforms ---------------------------------------------------------------------------------------
    rng = np.random.default_rng(9)
    xmin, xmax = rng.uniform(-2, 0, 3), rng.uniform(1, 4, 3)
    X = rng.uniform(-2, 4, (10, 3))
    bt = tr.BoxTransform(xmin, xmax)
    out['tr_xmin'], out['tr_xmax'], out['tr_X'] = xmin, xmax, X
    out['tr_box'], out['tr_box_back'], out['tr_box_var'] = bt(X), bt(X, back=True), bt.var(X)
    out['tr_box_cov'] = bt.cov(np.einsum('ni,nj->nij', X, X))
    mt = tr.MeanTransform(xmin, xmax)
    out['tr_mean'], out['tr_mean_back'] = mt(X), mt(X, back=True)
    pm, ps = mt.prediction(X, X ** 2, back=True)
    out['tr_pred_m'], out['tr_pred_s'] = pm, ps

    # ---- routing (utils.py) with the numpy>=2 shim ----------------------------------------------------------
    real_vstack = np.vstack
    ut.np.vstack = lambda a: real_vstack(list(a))
    try:
        rng = np.random.default_rng(3)
        for nb in (1, 2, 3):
            Z = np.c_[rng.uniform(0, 1, (25, 2)), rng.integers(0, 2, (25, nb)).astype(float)]
            R, J = ut.binary_dimensions(Z, list(range(2, 2 + nb)))
            out['route%d_Z' % nb], out['route%d_J' % nb] = Z, np.asarray(J)
            assert list(R) == list(range(2 ** nb))
    finally:
        ut.np.vstack = real_vstack

    np.savez_compressed(os.path.join(OUT, 'reference_golden.npz'), **out)
    print('wrote', os.path.join(OUT, 'reference_golden.npz'), len(out), 'arrays')


if __name__ == '__main__':
    main()
