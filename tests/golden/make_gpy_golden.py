"""Golden vectors from the REAL reference GP path (needs GPy + paramz): GPdoemd.models.GPModel trained by GPy, every
hook output dumped to tests/golden/gpy_golden.npz.  Run with the reference on PYTHONPATH, e.g.

    PYTHONPATH=/root/reference python tests/golden/make_gpy_golden.py

It pins what the oracle restates from memory of GPy 1.9.x (the items marked "re-verify": the 1e-8
jitter, gradients_XX at r = 0, the flat hyper-parameter order, variance clipping).  tests/test_gpy_golden.py compares
the oracle AND (with -m gpu) the CUDA path through GPModel.from_reference against the file when it exists.

STATUS (round 2): GPy is not installed in this image, is not in /opt/wheelhouse, and neither this container nor the
GPU boxes have network access -- see profiles/r02_gpy_attempt.log for the committed attempt (tools/try_gpy.sh).  The
file gpy_golden.npz therefore does not exist and GP parity stays "unpinned against real GPy" (DESIGN.md section 2)."""
import os
import sys

import numpy as np


def main():
    import GPy  # noqa: F401  (fails here: see the module docstring)
    from GPdoemd.kernels import Exponential, Matern32, Matern52, RBF, RatQuad
    from GPdoemd.models import GPModel
    rng = np.random.default_rng(20250)
    out = {}
    for kname, kern in (('RBF', RBF), ('Exponential', Exponential), ('Matern32', Matern32), ('Matern52', Matern52),
                        ('RatQuad', RatQuad)):
        D, P, E, n = 2, 2, 2, 60
        Z = rng.uniform([-1, 0, 0.5, -2], [2, 3, 1.5, 1], (n, D + P))
        Y = np.c_[np.sin(Z.sum(1)), np.cos(Z[:, 0] * Z[:, 2]) + 0.1 * Z[:, 3]]
        m = GPModel({'name': kname, 'call': lambda x, p: None, 'dim_x': D, 'dim_p': P, 'num_outputs': E})
        m.gp_surrogate(Z=Z, Y=Y, kern_x=kern, kern_p=kern)
        m.gp_optimise(max_iters=200)
        m.pmean = np.array([1.0, -0.5])
        m.Sigma = np.array([[0.02, 0.003], [0.003, 0.05]])
        X = rng.uniform([-1, 0], [2, 3], (25, D))
        X[0] = Z[3, :D]                                        # a training input: r = 0 in the x-kernel
        pre = kname + '_'
        out[pre + 'Z'], out[pre + 'Y'], out[pre + 'X'] = Z, Y, X
        out[pre + 'pmean'], out[pre + 'Sigma'] = m.pmean, m.Sigma
        out[pre + 'hyp'] = np.array([[np.asarray(h) for h in row] for row in m.hyp])
        for e in range(E):
            gp = m.gps[e][0]
            out[pre + 'alpha%d' % e] = np.asarray(gp.posterior.woodbury_vector)
            out[pre + 'L%d' % e] = np.asarray(gp.posterior.woodbury_chol)
            out[pre + 'dmu%d' % e] = m.d_mu_d_p(e, X)
            out[pre + 'd2mu%d' % e] = m.d2_mu_d_p2(e, X)
            out[pre + 'ds2%d' % e] = m.d_s2_d_p(e, X)
            out[pre + 'd2s2%d' % e] = m.d2_s2_d_p2(e, X)
        out[pre + 'M'], out[pre + 'S'] = m.predict(X)
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'gpy_golden.npz')
    np.savez_compressed(dst, **out)
    print('wrote', dst)


if __name__ == '__main__':
    sys.exit(main())
