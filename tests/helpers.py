"""Shared builders for the parity tests: the SAME training data, hyper-parameters, alpha and Cholesky
factor go to the oracle (CPU) and to the CUDA path (gpdoemd_b200.GPModel.from_reference)."""
import numpy as np

from gpdoemd_b200 import synthetic
from oracle import gpy_kernels
from oracle.model import OracleGPModel, OracleSparseGPModel


def oracle_models(seed, M, E, D, P, n, kernel, n_binary=0, gp_noise=1e-6):
    return synthetic.build_models(OracleGPModel, seed, M, E, D, P, n, kernel, n_binary,
                                  kernel_lookup=gpy_kernels.KERNELS, gp_noise=gp_noise)


def oracle_sparse_models(seed, M, E, D, P, n, kernel, num_inducing, n_binary=0, gp_noise=1e-4):
    return synthetic.build_models(OracleSparseGPModel, seed, M, E, D, P, n, kernel, n_binary,
                                  kernel_lookup=gpy_kernels.KERNELS, gp_noise=gp_noise, num_inducing=num_inducing)


def gpu_models(oracle_ms):
    from gpdoemd_b200 import GPModel
    return [GPModel.from_reference(m) for m in oracle_ms]


def demo_mesh(xlo, xhi, num=20):
    """demos/case_study.py:203-207: 20 x 20 x {0, 1} mesh, third design variable binary."""
    g = np.meshgrid(np.linspace(xlo[0], xhi[0], num), np.linspace(xlo[1], xhi[1], num), [0, 1])
    return np.vstack([np.ravel(a) for a in g]).T


def relerr(a, b, scale=None):
    """max |a-b| / max(|b|, scale): relative error with a floor at the quantity's natural scale."""
    a, b = np.asarray(a), np.asarray(b)
    den = np.abs(b) if scale is None else np.maximum(np.abs(b), scale)
    den = np.where(den == 0, 1.0, den)
    return float(np.max(np.abs(a - b) / den))


def variance_derivatives_factor_form(om, e, Xt):
    """d s2 / d p and d2 s2 / d p2 (transformed units, no binary variables) recomputed in FP64 through triangular
    solves against the Cholesky factor instead of GPy's explicit K^-1 (gp_model.py:263, GPy predictive_gradients).
    Same mathematics as oracle.model.OracleGPModel._d_s2_d_p / _d2_s2_d_p2, but backward stable: the explicit
    inverse loses cond(K) * u in the products k^T K^-1 dk, which is what limits the GPy-form oracle at large n."""
    from scipy.linalg import solve_triangular
    gp = om.gps[e][0]
    gpX, L = gp._predictive_variable, gp.posterior.woodbury_chol
    dx = om.dim_x
    Z = om.get_Z(Xt)
    kp = gp.kern.kernp
    kx = gp.kern.kernx.K(Z, gpX)
    Zs, Xs = Z[:, kp.active_dims], gpX[:, kp.active_dims]
    rr = kp._scaled_dist_sliced(Zs, Xs)
    w = kp._inv_dist_sliced(Zs, Xs) * (kp.dK_dr(rr) * kx)
    dk = w[:, :, None] * (Zs[:, None, :] - Xs[None, :, :]) / kp.lengthscale ** 2          # (N, n, P)
    k = gp.kern.K(Z, gpX)
    v = solve_triangular(L, k.T, lower=True)                                               # (n, N)
    W = np.stack([solve_triangular(L, dk[:, :, a].T, lower=True) for a in range(dk.shape[2])], axis=2)   # (n, N, P)
    ds2 = -2.0 * np.einsum('in,ina->na', v, W)
    beta = solve_triangular(L, v, lower=True, trans='T').T                                  # (N, n) = k K^-1
    ddk = np.sum(kp.gradients_XX(-kx * beta, Z, gpX), axis=1)[:, dx:, dx:]
    d2s2 = -2.0 * (ddk + np.einsum('ina,inb->nab', W, W))
    return ds2, d2s2


def d2mu_truth_longdouble(om, e, Xt, alpha=None):
    """80-bit truth of the second mean derivative d2 mu_e / dp_a dp_b (transformed units, no binary variables) for the
    kernels with an exponential profile, from the SAME float64 inputs (training inputs, hyper-parameters, alpha):
    difference-form distances, every operation in np.longdouble.  SURVEY Appendix A.2 / A.5:
        d2mu_ab = sum_i alpha_i kx_i sigma_p^2 [ (k''/r^2 - k'/r^3) d_a d_b / (l_a^2 l_b^2) + (k'/r) delta_ab / l_a^2 ]."""
    ld = np.longdouble
    gp = om.gps[e][0]
    kx, kp = gp.kern.kernx, gp.kern.kernp
    name = type(kx).__name__
    a = (gp.posterior.woodbury_vector[:, 0] if alpha is None else np.ravel(alpha)).astype(ld)
    Xtr = gp._predictive_variable.astype(ld)
    pstar = om.trans_p(om.pmean).astype(ld)
    Zx = np.asarray(Xt, dtype=float).astype(ld)

    def profile(r):
        if name == 'Exponential':
            k = np.exp(-r)
            return k, -k, k
        if name == 'Matern32':
            s3 = np.sqrt(ld(3))
            ex = np.exp(-s3 * r)
            return (1 + s3 * r) * ex, -3 * r * ex, 3 * (s3 * r - 1) * ex
        if name == 'Matern52':
            s5 = np.sqrt(ld(5))
            ex = np.exp(-s5 * r)
            return ((1 + s5 * r + ld(5) / 3 * r * r) * ex, -(ld(5) / 3) * r * (1 + s5 * r) * ex,
                    (ld(5) / 3) * (5 * r * r - s5 * r - 1) * ex)
        if name == 'RBF':
            k = np.exp(-r * r / 2)
            return k, -r * k, (r * r - 1) * k
        raise ValueError(name)
    lx = kx.lengthscale.astype(ld)
    dxs = (Zx[:, None, kx.active_dims] - Xtr[None, :, kx.active_dims]) / lx
    rx = np.sqrt((dxs ** 2).sum(-1))
    kxv = ld(kx.variance) * profile(rx)[0]                                   # (N, n)
    lp = kp.lengthscale.astype(ld)
    d = pstar[None, :] - Xtr[:, kp.active_dims]                              # (n, P)
    r = np.sqrt(((d / lp) ** 2).sum(-1))
    assert np.all(r > 0), 'p* coincides with a training parameter: GPy special-cases r = 0'
    k0, k1, k2 = profile(r)
    P = len(lp)
    H = np.zeros((len(r), P, P), dtype=ld)
    c1 = k2 / r ** 2 - k1 / r ** 3
    c2 = k1 / r
    for u in range(P):
        for v in range(P):
            H[:, u, v] = c1 * d[:, u] * d[:, v] / (lp[u] ** 2 * lp[v] ** 2)
        H[:, u, u] += c2 / lp[u] ** 2
    H *= ld(kp.variance)
    w = kxv * a[None, :]                                                      # (N, n)
    return np.einsum('ni,iuv->nuv', w, H), float(r.min())
