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
