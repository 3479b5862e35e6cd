"""The design step of GPdoemd's demos/case_study.py:279-293 on gpdoemd_b200, three ways.

    python examples/design_loop.py            (needs a B200; synthetic rival models, no GPy)

1. drop-in functions, exactly the reference's loop:   taylor_first_order per model + BH(mu, s2, noise)
2. the whole loop in one device pass:                 score_designs_multi
3. continuous refinement of the best grid designs:    refine_designs
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpdoemd_b200 as g
from gpdoemd_b200 import synthetic


def main():
    M, E, D, P, n = 3, 2, 3, 2, 300
    # rival models with GP surrogates: training data + hyper-parameters -> K, Cholesky, alpha on the device
    models, ds = synthetic.build_models(g.GPModel, 11, M, E, D, P, n, 'Matern52', n_binary=1, build_on_device=True)
    lo = np.max([d['xlo'] for d in ds], axis=0)
    hi = np.min([d['xhi'] for d in ds], axis=0)
    grid = np.meshgrid(np.linspace(lo[0], hi[0], 40), np.linspace(lo[1], hi[1], 40), [0, 1])
    Xtest = np.vstack([np.ravel(a) for a in grid]).T                      # 40 x 40 x {0, 1} candidate designs
    measvar = 0.01

    # 1. the reference's loop with the drop-in functions
    mu = np.zeros((len(Xtest), M, E))
    s2 = np.zeros((len(Xtest), M, E, E))
    for i, model in enumerate(models):
        mu[:, i], s2[:, i] = g.taylor_first_order(model, Xtest)
    dc = g.BH(mu, s2, measvar)
    print('drop-in  : next design', Xtest[np.argmax(dc)], 'BH = %.6g' % dc.max())

    # 2. fused: prediction, marginal, every criterion and the argmax in one pass
    res = g.score_designs_multi(models, Xtest, ['HR', 'BH', 'BF', 'AW', 'JR'], order=1, noise_var=measvar)
    for name, (idx, val, _) in res.items():
        print('fused    : %s argmax %5d  value %.6g' % (name, idx, val))
    assert res['BH'][0] == int(np.argmax(dc))

    # 3. refine the eight best grid points off the grid (the binary column stays fixed)
    top = Xtest[np.argsort(-res['BH'][2])[:8]]
    Xr, fr, evals = g.refine_designs(models, top, 'BH', noise_var=measvar, bounds=(lo, hi), binary=(2,))
    print('refined  : design', Xr[0], 'BH = %.6g after %d extra evaluations' % (fr[0], evals))


if __name__ == '__main__':
    main()
