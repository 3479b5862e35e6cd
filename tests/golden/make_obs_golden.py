"""Golden vectors for the discrimination criteria on the observed set, produced by RUNNING THE REFERENCE'S OWN
``GPdoemd/discrimination_criteria.py`` (numpy + scipy only; loaded by file path so that GPdoemd/__init__.py, which
imports GPy-dependent models, is never executed).  Build container only:

    python tests/golden/make_obs_golden.py     ->  tests/golden/obs_golden.npz
"""
import importlib.util
import os

import numpy as np

REF = '/root/reference'
OUT = os.path.dirname(os.path.abspath(__file__))


def load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def case(seed, n, N, E):
    rng = np.random.default_rng(seed)
    Y = rng.normal(size=(n, E))
    M = Y[:, None, :] + 0.3 * rng.normal(size=(n, N, E)) * rng.uniform(0.2, 2.0, (1, N, 1))
    A = rng.normal(size=(n, N, E, E))
    S = 0.05 * np.einsum('nmij,nmkj->nmik', A, A) + 0.02 * np.eye(E)
    D = rng.integers(1, 4, N)
    return Y, M, S, D


def main():
    dc = load('ref_discrimination', 'GPdoemd/discrimination_criteria.py')
    out = {}
    for k, (n, N, E) in enumerate(((10, 3, 2), (6, 4, 1), (25, 5, 3), (12, 2, 8), (4, 6, 4))):
        Y, M, S, D = case(900 + k, n, N, E)
        key = 'c%d_' % k
        out[key + 'Y'], out[key + 'M'], out[key + 'S'], out[key + 'D'] = Y, M, S, D
        out[key + 'gp'] = dc.gaussian_posterior(Y, M, S)
        out[key + 'aw'] = dc.akaike(Y, M, S, D)
        out[key + 'chi2_full'] = dc.chi2(Y, M, S, D)
        out[key + 'chi2_mat'] = dc.chi2(Y, M, S[0, 0], D)
        out[key + 'chi2_vec'] = dc.chi2(Y, M, np.diag(S[0, 0]).copy(), D)
        prior = np.random.default_rng(k).uniform(0.5, 1.5, N)
        prior /= prior.sum()
        out[key + 'prior'] = prior
        out[key + 'gpu'] = dc.gaussian_posterior_update(Y[-1], M[-1], S[-1], prior)
    # the reference test's own constructions (tests/test_discrimination_criteria.py:12-24)
    n, N, E = 10, 3, 2
    Y = np.ones((n, E))
    mu = np.zeros((n, N, E)) + 0.1
    mu[:, 0] = 1.
    s2 = 0.05 * np.array([[np.eye(E)] * N] * n)
    out['w_Y'], out['w_M'], out['w_S'] = Y, mu, s2
    out['w_gp'] = dc.gaussian_posterior(Y, mu, s2)
    out['w_aw'] = dc.akaike(Y, mu, s2, np.array([2] * N))
    out['w_chi2'] = dc.chi2(Y, mu, s2, np.array([2] * N))
    np.savez_compressed(os.path.join(OUT, 'obs_golden.npz'), **out)
    print('wrote obs_golden.npz with %d arrays' % len(out))


if __name__ == '__main__':
    main()
