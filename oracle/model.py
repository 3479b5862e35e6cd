"""Oracle GP surrogate model with the reference's class API (TEST INFRASTRUCTURE ONLY).

Restates the prediction/derivative part of ``GPdoemd/models/gp_model.py:96-143,184-280`` on top of
``GPdoemd/models/surrogate_model.py:70-217`` (training-data normalisation, public wrappers and their
unit conversions) and the attribute contract of ``GPdoemd/models/model.py:32-160``.  GP hyper-parameter
*optimisation* (``gp_optimise``, :149-181) is GPy/L-BFGS work that stays on the reference; here
hyper-parameters are supplied (``hyp``), exactly as after ``gp_load_hyp`` (:128-143).
"""
import numpy as np

from . import gpy_kernels
from .gp import GPRegression
from .sparse_gp import SparseGPRegression
from .gpy_kernels import Prod
from .routing import binary_dimensions
from .transform import BoxTransform, MeanTransform


class OracleGPModel:
    def __init__(self, model_dict):
        # model.py:32-42, surrogate_model.py:36, gp_model.py:41
        self.name = model_dict.get('name', 'model')
        self.call = model_dict.get('call', None)
        self.dim_x = int(model_dict['dim_x'])
        self.dim_p = int(model_dict['dim_p'])
        self.num_outputs = int(model_dict.get('num_outputs', 1))
        self.p_bounds = model_dict.get('p_bounds', [])
        mv = model_dict.get('meas_noise_var', 1.0)
        self.meas_noise_var = np.array([mv] * self.num_outputs) if isinstance(mv, (int, float)) else mv
        bv = model_dict.get('binary_variables', [])
        self.binary_variables = [bv] if isinstance(bv, int) else list(bv)
        self.gp_noise_var = model_dict.get('gp_noise_var', 1e-6)
        self.pmean = None
        self.Sigma = None
        self.Z = self.Y = self.hyp = self.gps = None
        self.kern_x = None
        self.want_var_jac = True      # False: d_mu_d_p skips GPy's (discarded) variance Jacobian, see gp.py

    dim_b = property(lambda self: len(self.binary_variables))
    non_binary_variables = property(
        lambda self: [i for i in range(self.dim_x) if i not in self.binary_variables])

    # ---- surrogate_model.py:70-165 : training data in normalised units ---------------------
    def set_training_data(self, Z, Y):
        Z = np.asarray(Z, dtype=float)
        Y = np.asarray(Y, dtype=float)
        assert Z.shape[1] == self.dim_x + self.dim_p and Y.shape[1] == self.num_outputs
        self.zmin, self.zmax = np.min(Z, axis=0), np.max(Z, axis=0)
        for i in self.binary_variables:
            self.zmin[i], self.zmax[i] = 0, 1
        self.xmin, self.xmax = self.zmin[:self.dim_x], self.zmax[:self.dim_x]
        self.pmin, self.pmax = self.zmin[self.dim_x:], self.zmax[self.dim_x:]
        self.trans_z = BoxTransform(self.zmin, self.zmax)
        self.trans_x = BoxTransform(self.xmin, self.xmax)
        self.trans_p = BoxTransform(self.pmin, self.pmax)
        self.Z = self.trans_z(Z)
        self.ymean, self.ystd = np.mean(Y, axis=0), np.std(Y, axis=0)
        self.trans_y = MeanTransform(self.ymean, self.ystd)
        self.Y = self.trans_y(Y)

    def get_Z(self, X, P=None):                      # surrogate_model.py:102-107
        if P is None:
            assert self.pmean is not None
            P = np.tile(self.trans_p(self.pmean), (len(X), 1))
        assert P.shape == (len(X), self.dim_p)
        return np.c_[X, P]

    # ---- gp_model.py:96-143 ------------------------------------------------------------------
    def gp_surrogate(self, Z=None, Y=None, kern_x=None, kern_p=None):
        if Z is not None:
            self.set_training_data(Z, Y)
        assert self.Z is not None and self.Y is not None
        kx = gpy_kernels.KERNELS[kern_x] if isinstance(kern_x, str) else kern_x
        assert kx is not None and kern_p is not None
        # quirk 1 (gp_model.py:60): the kern_p getter returns _kern_x, so BOTH parts use kern_x's class
        self.kern_x = kx
        R, J = binary_dimensions(self.Z, self.binary_variables)
        dim = self.dim_x + self.dim_p
        gps = []
        for e in range(self.num_outputs):
            gps.append([])
            for r in R:
                Jr = (J == r)
                if not np.any(Jr):
                    gps[e].append(None)
                    continue
                kernx = kx(self.dim_x - self.dim_b, self.non_binary_variables, 'kernx')
                kernp = kx(self.dim_p, range(self.dim_x, dim), 'kernp')
                gps[e].append(GPRegression(self.Z[Jr], self.Y[np.ix_(Jr, [e])], Prod(kernx, kernp)))
        self.gps = gps

    def gp_load_hyp(self, index=None):
        index = range(self.num_outputs) if index is None else ([index] if isinstance(index, int) else index)
        for e in index:
            for gp, hyp in zip(self.gps[e], self.hyp[e]):
                if gp is not None:
                    gp.set_params(hyp)

    # ---- surrogate_model.py:170-215 : public wrappers (original units) ----------------------
    def predict(self, xnew):
        assert self.pmean is not None
        assert self.Z is not None and self.Y is not None
        assert self.hyp is not None
        M, S = self._predict(self.trans_x(xnew), self.trans_p(self.pmean))
        return self.trans_y.prediction(M, S, back=True)

    def d_mu_d_p(self, e, X):
        return self._d_mu_d_p(e, self.trans_x(X)) * self.ystd[e] / (self.pmax - self.pmin)

    def d2_mu_d_p2(self, e, X):
        diff = self.pmax - self.pmin
        return self._d2_mu_d_p2(e, self.trans_x(X)) * self.ystd[e] / (diff[:, None] * diff[None, :])

    def d_s2_d_p(self, e, X):
        return self._d_s2_d_p(e, self.trans_x(X)) * self.ystd[e] ** 2 / (self.pmax - self.pmin)

    def d2_s2_d_p2(self, e, X):
        diff = self.pmax - self.pmin
        return self._d2_s2_d_p2(e, self.trans_x(X)) * self.ystd[e] ** 2 / (diff[:, None] * diff[None, :])

    # ---- gp_model.py:184-280 : hooks in normalised units --------------------------------------
    def _groups(self, Zrows):
        R, J = binary_dimensions(Zrows, self.binary_variables)
        for r in R:
            Jr = (J == r)
            if np.any(Jr):
                yield r, Jr

    def _predict(self, xnew, p):
        znew = np.c_[xnew, np.tile(p, (len(xnew), 1))]
        n = len(znew)
        M = np.zeros((n, self.num_outputs))
        S = np.zeros((n, self.num_outputs))
        for r, Jr in self._groups(znew):
            for e in range(self.num_outputs):
                m, s = self.gps[e][r].predict_noiseless(znew[Jr])
                M[Jr, e], S[Jr, e] = m[:, 0], s[:, 0]
        return M, S

    def _d_mu_d_p(self, e, X):
        dmu = np.zeros((len(X), self.dim_p))
        for r, Jr in self._groups(X):
            Z = self.get_Z(X[Jr])
            dmu[Jr] = self.gps[e][r].predictive_gradients(Z, want_var=self.want_var_jac)[0][:, self.dim_x:, 0]
        return dmu

    def _d_s2_d_p(self, e, X):
        ds2 = np.zeros((len(X), self.dim_p))
        for r, Jr in self._groups(X):
            Z = self.get_Z(X[Jr])
            ds2[Jr] = self.gps[e][r].predictive_gradients(Z)[1][:, self.dim_x:]
        return ds2

    def _d2_mu_d_p2(self, e, X, chunk=256):
        dx = self.dim_x
        ddmu = np.zeros((len(X), self.dim_p, self.dim_p))
        for r, Jr in self._groups(X):
            gp = self.gps[e][r]
            gpX = gp._predictive_variable
            idx = np.flatnonzero(Jr)
            for s in range(0, len(idx), chunk):
                I = idx[s:s + chunk]
                Z = self.get_Z(X[I])
                tmp = -gp.kern.kernx.K(Z, gpX) * gp.posterior.woodbury_vector.T
                dt = np.sum(gp.kern.kernp.gradients_XX(tmp, Z, gpX), axis=1)
                ddmu[I] = dt[:, dx:, dx:]
        return ddmu

    def _d2_s2_d_p2(self, e, X, chunk=128):
        dx = self.dim_x
        dds2 = np.zeros((len(X), self.dim_p, self.dim_p))
        for r, Jr in self._groups(X):
            gp = self.gps[e][r]
            gpX = gp._predictive_variable
            iK = gp.posterior.woodbury_inv
            idx = np.flatnonzero(Jr)
            for s in range(0, len(idx), chunk):
                I = idx[s:s + chunk]
                Z = self.get_Z(X[I])
                kx = gp.kern.kernx.K(Z, gpX)
                # d k / d p : per training point i the reference calls kernp.gradients_X(kx[:,[i]], Z, gpX[[i]])
                # (gp_model.py:268-270); the same numbers for all i at once:
                kp = gp.kern.kernp
                Zs, Xs = Z[:, kp.active_dims], gpX[:, kp.active_dims]
                rr = kp._scaled_dist_sliced(Zs, Xs)
                w = kp._inv_dist_sliced(Zs, Xs) * (kp.dK_dr(rr) * kx)
                dk = w[:, :, None] * (Zs[:, None, :] - Xs[None, :, :]) / kp.lengthscale ** 2
                dkiKdk = np.einsum('ink,inm->ikm', np.einsum('ijk,jn->ink', dk, iK), dk)   # :271
                # d^2 k / d p^2
                kiK = np.matmul(gp.kern.K(Z, gpX), iK)
                ddk = np.sum(kp.gradients_XX(-kx * kiK, Z, gpX), axis=1)[:, dx:, dx:]
                dds2[I] = -2.0 * (ddk + dkiKdk)
        return dds2


def adopt_exact_model(model):
    """OracleGPModel that shares transforms, hyper-parameters and the posterior (alpha, L) of an exact-GP model object
    with the reference's attribute surface (a gpdoemd_b200.GPModel, whose GPs expose ``posterior.woodbury_chol`` /
    ``woodbury_vector`` -- downloaded from the device for device-built GPs): nothing is refactorised, so both sides
    of a parity check start from the same (alpha, L), like the reference and a GPModel.from_reference copy."""
    om = OracleGPModel({'name': model.name, 'dim_x': model.dim_x, 'dim_p': model.dim_p, 'num_outputs': model.num_outputs,
                        'binary_variables': list(model.binary_variables), 'meas_noise_var': model.meas_noise_var})
    om.Z, om.Y = np.array(model.Z), np.array(model.Y)
    om.zmin, om.zmax = np.array(model.zmin, dtype=float), np.array(model.zmax, dtype=float)
    om.xmin, om.xmax = om.zmin[:om.dim_x], om.zmax[:om.dim_x]
    om.pmin, om.pmax = om.zmin[om.dim_x:], om.zmax[om.dim_x:]
    om.trans_z, om.trans_x, om.trans_p = BoxTransform(om.zmin, om.zmax), BoxTransform(om.xmin, om.xmax), BoxTransform(om.pmin, om.pmax)
    om.ymean, om.ystd = np.array(model.ymean, dtype=float), np.array(model.ystd, dtype=float)
    om.trans_y = MeanTransform(om.ymean, om.ystd)
    from gpdoemd_b200 import kernels as host_kernels          # class names only (data, no computation)
    kx = gpy_kernels.KERNELS[host_kernels.kernel_name_of(model.kern_x)]
    om.kern_x = kx
    dim = om.dim_x + om.dim_p
    gps = []
    for e, row in enumerate(model.gps):
        gps.append([])
        for gp in row:
            if gp is None:
                gps[e].append(None)
                continue
            kernx = kx(om.dim_x - om.dim_b, om.non_binary_variables, 'kernx')
            kernp = kx(om.dim_p, range(om.dim_x, dim), 'kernp')
            kern = Prod(kernx, kernp)
            v = np.asarray(gp[:], dtype=float)
            k = kern.set_params(v)
            assert k + 1 == len(v)
            post = gp.posterior
            gps[e].append(GPRegression.from_posterior(gp._predictive_variable, gp.Y, kern, v[k], post.woodbury_chol,
                                                      post.woodbury_vector))
    om.gps = gps
    om.hyp = model.hyp
    om.pmean = None if model.pmean is None else np.array(model.pmean, dtype=float)
    om.Sigma = None if model.Sigma is None else np.array(model.Sigma, dtype=float)
    return om


class OracleSparseGPModel(OracleGPModel):
    """GPdoemd/models/sparse_gp_model.py:35-74: same hooks, GPy SparseGPRegression per output / binary combination."""

    def __init__(self, model_dict):
        super().__init__(model_dict)
        self.max_num_inducing = model_dict.get('max_num_inducing', 500)

    def gp_surrogate(self, Z=None, Y=None, kern_x=None, kern_p=None, num_inducing=None):
        if Z is not None:
            self.set_training_data(Z, Y)
        assert self.Z is not None and self.Y is not None
        kx = gpy_kernels.KERNELS[kern_x] if isinstance(kern_x, str) else kern_x
        assert kx is not None and kern_p is not None
        self.kern_x = kx
        R, J = binary_dimensions(self.Z, self.binary_variables)
        dim = self.dim_x + self.dim_p
        numi = self.max_num_inducing if num_inducing is None else num_inducing
        gps = []
        for e in range(self.num_outputs):
            gps.append([])
            for r in R:
                Jr = (J == r)
                if not np.any(Jr):
                    gps[e].append(None)
                    continue
                kernx = kx(self.dim_x - self.dim_b, self.non_binary_variables, 'kernx')
                kernp = kx(self.dim_p, range(self.dim_x, dim), 'kernp')
                gps[e].append(SparseGPRegression(self.Z[Jr], self.Y[np.ix_(Jr, [e])], Prod(kernx, kernp),
                                                 num_inducing=numi, rng=np.random.default_rng(1000 * e + r)))
        self.gps = gps
