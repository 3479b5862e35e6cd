"""Synthetic rival-model problems of the BASELINE shapes: seeded training data and
hyper-parameters for M models x E outputs, plus candidate designs.  Data only -- the same arrays feed
the CUDA path, the oracle and the CPU baseline."""
import numpy as np

CONFIGS = {
    # name: M, E, D, P, n, kernel, order, criteria, N
    'C2': dict(M=4, E=2, D=3, P=4, n=500, kernel='RBF', order=1, criteria=('HR', 'BH'), N=100000),
    'C3': dict(M=4, E=2, D=3, P=4, n=4000, kernel='Matern52', order=2, criteria=('AW', 'BF'), N=1000000),
    'C4': dict(M=8, E=8, D=3, P=10, n=4000, kernel='Matern52', order=1, criteria=('HR', 'BH', 'BF', 'AW', 'JR'),
               N=1000000),
    'C5': dict(M=4, E=4, D=3, P=4, n=4000, kernel='Matern52', order=1, criteria=('BH',), N=100000000),
    # SparseGPModel surrogates (reference default max_num_inducing = 500) on the C2 shapes
    'S2': dict(M=4, E=2, D=3, P=4, n=4000, inducing=500, kernel='RBF', order=1, criteria=('HR', 'BH'), N=100000),
}


def model_data(seed, E, D, P, n, n_binary=0, gp_noise=1e-6, kernel=None):
    """Training set and hyper-parameters of one model.

    Z (n, D+P) original units, Y (n, E) smooth deterministic responses (a surrogate is trained on
    simulator output, not on noisy data), hyp[e][r] in GPy's flat order, pmean, Sigma, noise.
    """
    rng = np.random.default_rng(seed)
    xlo, xhi = rng.uniform(-2, 0, D), rng.uniform(1, 3, D)
    plo, phi = rng.uniform(-1, 0, P), rng.uniform(0.5, 2, P)
    Z = np.c_[rng.uniform(xlo, xhi, (n, D)), rng.uniform(plo, phi, (n, P))]
    binary = list(range(D - n_binary, D))
    for b in binary:
        Z[:, b] = rng.integers(0, 2, n)
    U = (Z - np.r_[xlo, plo]) / (np.r_[xhi, phi] - np.r_[xlo, plo])
    Y = np.empty((n, E))
    for e in range(E):
        w = rng.uniform(0.5, 1.5, D + P)
        Y[:, e] = (1.0 + e) * np.sin(3.0 * U.dot(w) / np.sqrt(D + P) + 0.7 * e) + 0.3 * U.dot(rng.normal(size=D + P))
    R = 2 ** n_binary
    Dx = D - n_binary
    # GPy flat order: kernx.variance, kernx.lengthscale[, kernx.power], kernp.variance, kernp.lengthscale[, kernp.power], noise
    pw = (lambda: [rng.uniform(1.0, 3.0)]) if kernel == 'RatQuad' else (lambda: [])
    hyp = [[np.r_[rng.uniform(0.5, 2.0), rng.uniform(0.3, 2.0, Dx), pw(), rng.uniform(0.5, 2.0), rng.uniform(0.3, 2.0, P), pw(),
                  gp_noise] for _ in range(R)] for _ in range(E)]
    pmean = plo + rng.uniform(0.25, 0.75, P) * (phi - plo)
    A = rng.normal(size=(P, P))
    Sigma = A.dot(A.T) / P * 1e-2 * np.outer(phi - plo, phi - plo)
    return dict(Z=Z, Y=Y, hyp=hyp, pmean=pmean, Sigma=Sigma, binary=binary, xlo=xlo, xhi=xhi)


def candidates(seed, N, D, xlo, xhi, n_binary=0):
    rng = np.random.default_rng(seed)
    X = rng.uniform(xlo, xhi, (N, D))
    for b in range(D - n_binary, D):
        X[:, b] = rng.integers(0, 2, N)
    return X


def build_models(cls, seed, M, E, D, P, n, kernel, n_binary=0, kernel_lookup=None, gp_noise=1e-6,
                 num_inducing=None, **kw):
    """Instantiate ``cls`` (gpdoemd_b200.GPModel or an API-compatible class) for M rival models."""
    models, datas = [], []
    for m in range(M):
        Pm = P[m] if isinstance(P, (list, tuple)) else P          # rival models may differ in dim_p
        d = model_data(seed * 1000 + m, E, D, Pm, n, n_binary, gp_noise, kernel)
        mo = cls({'name': 'm%d' % m, 'call': (lambda x, p: None), 'dim_x': D, 'dim_p': Pm, 'num_outputs': E,
                  'meas_noise_var': 0.01, 'binary_variables': d['binary']}, **kw)
        k = kernel if kernel_lookup is None else kernel_lookup[kernel]
        if num_inducing is None:
            mo.gp_surrogate(Z=d['Z'], Y=d['Y'], kern_x=k, kern_p=k)
            mo.hyp = d['hyp']
        else:
            # sparse surrogates (cls = a SparseGPModel class): gp[:] starts with the inducing inputs
            mo.gp_surrogate(Z=d['Z'], Y=d['Y'], kern_x=k, kern_p=k, num_inducing=num_inducing)
            mo.hyp = [[None if gp is None else np.r_[np.ravel(gp.Z), h] for gp, h in zip(gps, hs)]
                      for gps, hs in zip(mo.gps, d['hyp'])]
        mo.gp_load_hyp()
        mo.pmean = d['pmean']
        mo.Sigma = d['Sigma']
        models.append(mo)
        datas.append(d)
    return models, datas
