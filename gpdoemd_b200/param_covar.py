"""``laplace_approximation(model, Xdata)`` -- SURVEY 8f row 3.

Same signature and result as ``GPdoemd/param_covar/laplace_approximation.py:27-61``:
``Sigma = (sum_n grad f  Sigma_exp^-1  grad f^T)^-1`` with ``grad f = model.d_mu_d_p`` at the observed
designs.  For a ``gpdoemd_b200.GPModel`` the derivatives of all E outputs come from ONE device pass
(``gpd_hooks``); the D x D accumulation over the handful of observed points is host glue, exactly like the
reference (which also works unchanged on a ``gpdoemd_b200.GPModel`` thanks to duck typing)."""
import numpy as np


def laplace_approximation(model, Xdata):
    E, D = model.num_outputs, model.dim_p
    meas_noise_var = model.meas_noise_var
    if isinstance(meas_noise_var, (int, float)):
        meas_noise_var = np.array([meas_noise_var] * E)
    imeasvar = np.diag(1. / meas_noise_var) if meas_noise_var.ndim == 1 else np.linalg.inv(meas_noise_var)
    dmu = []
    for e in range(E):                       # cached after the first call for a device model
        d = model.d_mu_d_p(e, Xdata)
        assert d.shape == (len(Xdata), D)
        dmu.append(d)
    iA = np.zeros((D, D))
    for e1 in range(E):
        iA += imeasvar[e1, e1] * np.matmul(dmu[e1].T, dmu[e1])
        if meas_noise_var.ndim == 1:
            continue
        for e2 in range(e1 + 1, E):
            if imeasvar[e1, e2] == 0.:
                continue
            iA += imeasvar[e1, e2] * np.matmul(dmu[e1].T, dmu[e2])
            iA += imeasvar[e2, e1] * np.matmul(dmu[e2].T, dmu[e1])
    return np.linalg.inv(iA)
