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
    ctx = get_context(model._device)
    M = ctx.pooled_pinned((N, E))           # page-locked results: the D2H copies run at full PCIe rate
    S = ctx.pooled_pinned((N, E, E))
    check(ctx.lib.gpd_marginal(model.device_handle(), dptr(X), N, order, dptr(M), dptr(S)))
    return M, S


class ResidentArray(np.ndarray):
    """Host array (read-only, page-locked) whose exact content also lives in device memory (``gpd_resident``).

    ``taylor_multi`` returns mu and S as this type; the criteria recognise the pair and score the device copy, so
    nothing N-sized crosses PCIe between the marginals and ``HR .. JR``.  Soundness rests on immutability: the arrays
    are not writeable, and every derived array (slice, reshape, copy, arithmetic result) is an ordinary host array --
    ``__array_finalize__`` never propagates the device handle -- which takes the upload path like any other input."""

    def __array_finalize__(self, obj):
        self._gpd_resident = None


class _Resident:
    """Owner of one gpd_resident handle (freed with the last array that refers to it)."""

    def __init__(self, ctx, handle, N, M, E):
        self.ctx, self.handle, self.shape_mu, self.shape_S = ctx, handle, (N, M, E), (N, M, E, E)

    def __del__(self):
        try:
            if self.handle and self.ctx.handle:
                self.ctx.lib.gpd_resident_free(self.handle)
        except Exception:
            pass
        self.handle = None


def resident_pair(mu, s2):
    """-> the _Resident behind (mu, s2) if both are the untouched, still read-only outputs of ONE taylor_multi call."""
    r = getattr(mu, '_gpd_resident', None)
    if r is None or getattr(s2, '_gpd_resident', None) is not r or not r.handle or not r.ctx.handle:
        return None
    if mu.flags.writeable or s2.flags.writeable or mu.shape != r.shape_mu or s2.shape != r.shape_S:
        return None
    return r


def taylor_multi(models, xnew, order=1):
    """``taylor_first_order`` / ``taylor_second_order`` for ALL rival models in one device pass:

        mu, s2 = taylor_multi(models, X, order)        # (N, M, E), (N, M, E, E)

    replaces the reference's loop ``for i, M in enumerate(models): mu[:, i], s2[:, i] = taylor_*_order(M, X)``
    (demos/case_study.py:284-288) -- same numbers, but X is uploaded once, the M models share one pass, and the
    returned arrays keep a device copy that ``HR, BH, BF, AW, JR`` consume directly (see ResidentArray)."""
    import ctypes as C
    assert len(models) >= 1 and all(isinstance(m, GPModel) for m in models), 'taylor_multi takes gpdoemd_b200 GP models'
    assert order in (1, 2)
    X = _lib.as_f64(xnew)
    M, E = len(models), models[0].num_outputs
    assert X.ndim == 2 and X.shape[1] == models[0].dim_x
    N = len(X)
    for m in models:
        assert m.pmean is not None and m.Z is not None and m.Y is not None and m.hyp is not None
        assert m.Sigma is not None, 'model.Sigma is not set'
        m.check_routing(X)
    ctx = get_context(models[0]._device)
    handles = (C.c_void_p * M)(*[m.device_handle() for m in models])
    mu = ctx.pooled_pinned((N, M, E))
    s2 = ctx.pooled_pinned((N, M, E, E))
    keep = C.c_void_p()
    check(ctx.lib.gpd_marginals(ctx.handle, C.cast(handles, _lib.c_void_pp), M, dptr(X), N, order, dptr(mu), dptr(s2),
                                C.byref(keep)))
    if N == 0:
        return mu, s2
    owner = _Resident(ctx, keep, N, M, E)
    mu, s2 = mu.view(ResidentArray), s2.view(ResidentArray)
    mu._gpd_resident = s2._gpd_resident = owner
    mu.flags.writeable = False
    s2.flags.writeable = False
    return mu, s2


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


def taylor_first_order_dx(model, xnew):
    """``taylor_first_order`` and its derivative with respect to the design:
    ``(model, xnew) -> M (N,E), S (N,E,E), dM_dx (N,E,D), dS_dx (N,E,E,D)`` in original units (the
    reference has no counterpart).  Columns of binary design variables hold zero."""
    assert isinstance(model, GPModel), 'design gradients need a gpdoemd_b200 GP model'
    assert model.pmean is not None and model.Sigma is not None, 'pmean / Sigma are not set'
    X = _lib.as_f64(xnew)
    assert X.ndim == 2 and X.shape[1] == model.dim_x
    N, E, D = len(X), model.num_outputs, model.dim_x
    model.check_routing(X)
    M, S = np.empty((N, E)), np.empty((N, E, E))
    dM, dS = np.empty((N, E, D)), np.empty((N, E, E, D))
    ctx = get_context(model._device)
    check(ctx.lib.gpd_marginal_dx(model.device_handle(), dptr(X), N, dptr(M), dptr(S), dptr(dM), dptr(dS)))
    return M, S, dM, dS


def taylor_first_order(model, xnew):
    if isinstance(model, GPModel):
        return _device_marginal(model, xnew, 1)
    return _assemble(model, xnew, 1)


def taylor_second_order(model, xnew):
    if isinstance(model, GPModel):
        return _device_marginal(model, xnew, 2)
    return _assemble(model, xnew, 2)
