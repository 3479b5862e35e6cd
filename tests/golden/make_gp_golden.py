"""Golden vectors for the GP part, produced by the CPU oracle (GPy is not available, so these pin the
CUDA path to the oracle's numbers at the time of generation; see oracle/__init__.py for what pins the oracle).

    python tests/golden/make_gp_golden.py        # writes tests/golden/gp_golden.npz

Inputs are regenerated from seeds by gpdoemd_b200.synthetic; only outputs are stored."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

CASES = {
    # name: (seed, M, E, D, P, n, kernel, n_binary, N)
    'rbf':     (101, 2, 2, 3, 3, 150, 'RBF', 0, 64),
    'm52bin':  (102, 2, 2, 3, 2, 160, 'Matern52', 1, 64),
    'm32':     (103, 3, 1, 2, 4, 130, 'Matern32', 0, 48),
}


def build(case):
    import helpers
    from gpdoemd_b200 import synthetic
    seed, M, E, D, P, n, kernel, nb, N = CASES[case]
    oms, ds = helpers.oracle_models(seed, M=M, E=E, D=D, P=P, n=n, kernel=kernel, n_binary=nb)
    lo = np.max([d['xlo'] for d in ds], axis=0)
    hi = np.min([d['xhi'] for d in ds], axis=0)
    X = synthetic.candidates(seed + 1, N, D, lo, hi, n_binary=nb)
    return oms, X


def main():
    from oracle import criteria as ocrit
    from oracle import marginal as omarg
    out = {}
    for case in CASES:
        oms, X = build(case)
        M, E = len(oms), oms[0].num_outputs
        N = len(X)
        mu1 = np.zeros((N, M, E)); S1 = np.zeros((N, M, E, E))
        mu2 = np.zeros((N, M, E)); S2 = np.zeros((N, M, E, E))
        for i, om in enumerate(oms):
            pm, ps = om.predict(X)
            out['%s_m%d_mu' % (case, i)], out['%s_m%d_s2' % (case, i)] = pm, ps
            out['%s_m%d_dmu' % (case, i)] = np.stack([om.d_mu_d_p(e, X) for e in range(E)], 1)
            out['%s_m%d_d2mu' % (case, i)] = np.stack([om.d2_mu_d_p2(e, X) for e in range(E)], 1)
            out['%s_m%d_ds2' % (case, i)] = np.stack([om.d_s2_d_p(e, X) for e in range(E)], 1)
            out['%s_m%d_d2s2' % (case, i)] = np.stack([om.d2_s2_d_p2(e, X) for e in range(E)], 1)
            mu1[:, i], S1[:, i] = omarg.taylor_first_order(om, X)
            mu2[:, i], S2[:, i] = omarg.taylor_second_order(om, X)
        out[case + '_mu1'], out[case + '_S1'], out[case + '_mu2'], out[case + '_S2'] = mu1, S1, mu2, S2
        noise = 0.01 * (1 + np.arange(E))
        for name in ('HR', 'BH', 'BF', 'AW', 'JR'):
            out['%s_dc1_%s' % (case, name)] = ocrit.CRITERIA[name](mu1, S1, noise, None)
            out['%s_dc2_%s' % (case, name)] = ocrit.CRITERIA[name](mu2, S2, noise, None)
    np.savez_compressed(os.path.join(HERE, 'gp_golden.npz'), **out)
    print('wrote gp_golden.npz with', len(out), 'arrays')


if __name__ == '__main__':
    main()
