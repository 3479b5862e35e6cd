"""Drop-in ``taylor_first_order`` / ``taylor_second_order``.

Same signature and results as ``GPdoemd/marginal/taylor_first_order.py:27-43`` and
``taylor_second_order.py:27-60``:  ``(model, xnew) -> (M (N,E), S (N,E,E))``.

* a ``gpdoemd_b200.GPModel`` goes through ``gpd_marginal``: prediction, derivatives and the assembly
  S = diag(s2) + J Sigma J^T (+ second-order terms) run in one device pass;
* any other duck-typed model (``num_outputs, dim_p, Sigma, predict, d_mu_d_p[, d2_mu_d_p2,
  d2_s2_d_p2]`` -- the protocol of tests/test_marginal/test_taylor_approximation.py:17-51) supplies its
  own hook outputs and only the assembly runs on the device (``gpd_marginal_assemble``).
"""
import numpy as np

from . import _lib
from ._lib import check, dptr
from .context import get_context
from .models import GPModel


def _device_marginal(model, xnew, order):
    assert model.pmean is not None
    assert model.Z is not None and model.Y is not None
    assert model.hyp is not None
    assert model.Sigma is not None, 'model.Sigma is not set'
    X = _lib.as_f64(xnew)
    assert X.ndim == 2 and X.shape[1] == model.dim_x
    N, E = len(X), model.num_outputs
    model.check_routing(X)
    M = np.empty((N, E))
    S = np.empty((N, E, E))
    ctx = get_context(model._device)
    check(ctx.lib.gpd_marginal(model.device_handle(), dptr(X), N, order, dptr(M), dptr(S)))
    return M, S


def _assemble(model, xnew, order, device=None):
    N, E, D = len(xnew), model.num_outputs, model.dim_p
    M, s2 = model.predict(xnew)
    M, s2 = _lib.as_f64(M).copy(), _lib.as_f64(s2)
    assert M.shape == (N, E) and s2.shape == (N, E)
    dmu = np.empty((N, E, D))
    ddmu = dds2 = None
    if order == 2:
        ddmu = np.empty((N, E, D, D))
        dds2 = np.empty((N, E, D, D))
    for e in range(E):
        dmu[:, e] = model.d_mu_d_p(e, xnew)
        if order == 2:
            ddmu[:, e] = model.d2_mu_d_p2(e, xnew)
            dds2[:, e] = model.d2_s2_d_p2(e, xnew)
    Sigma = _lib.as_f64(model.Sigma)
    assert Sigma.shape == (D, D)
    S = np.empty((N, E, E))
    ctx = get_context(device)
    check(ctx.lib.gpd_marginal_assemble(ctx.handle, order, N, E, D, dptr(M), dptr(s2), dptr(dmu), dptr(ddmu),
                                        dptr(dds2), dptr(Sigma), dptr(S)))
    return M, S


def taylor_first_order(model, xnew):
    if isinstance(model, GPModel):
        return _device_marginal(model, xnew, 1)
    return _assemble(model, xnew, 1)


def taylor_second_order(model, xnew):
    if isinstance(model, GPModel):
        return _device_marginal(model, xnew, 2)
    return _assemble(model, xnew, 2)
