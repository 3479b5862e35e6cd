"""Fused design scoring: the loop of ``demos/case_study.py:279-293`` in one device pass.

``score_designs(models, X, order, criterion, noise_var, pps)`` is equivalent to

    for i, M in enumerate(models): mu[:, i], s2[:, i] = taylor_{first,second}_order(M, X)
    dc = CRITERION(mu, s2, noise_var, pps);  best = np.argmax(dc)

with nothing N-sized leaving the GPU unless ``return_dc`` is set.  ``ShardedScorer`` splits the candidates
over the ranks of a ``torch.distributed`` process group (one process per GPU, GP factors replicated,
contiguous blocks) and combines the per-rank (max, index) pairs with one all-gather; ties go to the
lowest global index, like ``np.argmax`` on the unsharded vector.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, dptr
from .context import default_device, get_context
from .design_criteria import _reshape


_NORM_CACHE = {}


def _norm_key(v):
    if v is None or isinstance(v, (int, float)):
        return v
    a = np.asarray(v, dtype=float)
    return (a.shape, a.tobytes())


def _prep(models, criterion, noise_var, pps):
    M, E = len(models), models[0].num_outputs
    # noise / prior normalisation of design_criteria._reshape (:202-238), memoised: a design loop passes the same
    # measurement noise and priors on every call
    key = (M, E, _norm_key(noise_var), _norm_key(pps))
    hit = _NORM_CACHE.get(key)
    if hit is None:
        _, _, noise, pps_n, _, _, _ = _reshape(np.zeros((1, M, E)), np.zeros((1, M, E, E)), noise_var, pps)
        if len(_NORM_CACHE) > 64:
            _NORM_CACHE.clear()
        hit = _NORM_CACHE[key] = (_lib.as_f64(noise).copy(), _lib.as_f64(pps_n).copy())
    handles = (C.c_void_p * M)(*[m.device_handle() for m in models])
    return _lib.CRITERION_IDS[criterion], hit[0], hit[1], handles


def score_designs_multi(models, X, criteria, order=1, noise_var=None, pps=None, return_dc=True, idx_offset=0,
                        x_dev=None, N=None, dc_dev=None, dc_out=None):
    """Several criteria in one pass over the candidates -> {name: (best_index, best_value, dc or None)}.

    X is a host (N, dim_x) array, or pass ``x_dev`` (device pointer) and ``N`` for resident candidates
    (then ``dc_dev`` may be a device pointer to [len(criteria)][N] doubles).  ``dc_out``: optional preallocated
    C-contiguous float64 array (len(criteria), N) for the criterion vectors, e.g. page-locked memory from
    ``Context.pinned_empty`` (a pageable destination makes the 8 N bytes per criterion a staged copy).
    """
    dev = models[0]._device
    ctx = get_context(dev)
    _, noise, pps_n, handles = _prep(models, criteria[0], noise_var, pps)
    kinds = np.array([_lib.CRITERION_IDS[c] for c in criteria], dtype=np.int32)
    nk = len(kinds)
    bv = np.zeros(nk)
    bi = (C.c_int64 * nk)()
    if x_dev is None:
        X = _lib.as_f64(X)
        N = len(X)
        for m in models:
            m.check_routing(X)            # binary columns are 0 / 1 in both unit systems
        if return_dc and dc_out is not None:
            assert dc_out.shape == (nk, N) and dc_out.dtype == np.float64 and dc_out.flags['C_CONTIGUOUS']
            dc = dc_out
        else:
            dc = np.empty((nk, N)) if return_dc else None
        xp, x_on_dev = X.ctypes.data_as(C.c_void_p), 0
        dcp, dc_on_dev = (dc.ctypes.data_as(C.c_void_p) if return_dc else None), 0
    else:
        dc = None
        xp, x_on_dev = x_dev, 1
        dcp, dc_on_dev = dc_dev, 1
    check(ctx.lib.gpd_score_multi(ctx.handle, C.cast(handles, _lib.c_void_pp), len(models), xp, x_on_dev, int(N), order,
                                  kinds.ctypes.data_as(_lib.c_int32_p), nk, dptr(noise), dptr(pps_n), dcp, dc_on_dev,
                                  dptr(bv), bi, int(idx_offset)))
    return {c: (int(bi[k]), float(bv[k]), None if dc is None else dc[k]) for k, c in enumerate(criteria)}


def score_designs(models, X, order=1, criterion='BH', noise_var=None, pps=None, return_dc=True, idx_offset=0):
    """-> (best_index, best_value, dc or None).  X: (N, dim_x) host array in original units."""
    X = _lib.as_f64(X)
    N = len(X)
    for m in models:
        m.check_routing(X)
    dev = models[0]._device
    ctx = get_context(dev)
    kind, noise, pps_n, handles = _prep(models, criterion, noise_var, pps)
    dc = np.empty(N) if return_dc else None
    bv, bi = C.c_double(), C.c_int64()
    check(ctx.lib.gpd_score(ctx.handle, C.cast(handles, _lib.c_void_pp), len(models), dptr(X), N, order, kind,
                            dptr(noise), dptr(pps_n), dptr(dc), C.byref(bv), C.byref(bi), int(idx_offset)))
    return int(bi.value), bv.value, dc


def score_designs_dev(models, X_dev, N, order=1, criterion='BH', noise_var=None, pps=None, dc_dev=None,
                      idx_offset=0):
    """Same with candidates (and optionally the criterion vector) resident in device memory."""
    dev = models[0]._device
    ctx = get_context(dev)
    kind, noise, pps_n, handles = _prep(models, criterion, noise_var, pps)
    bv, bi = C.c_double(), C.c_int64()
    check(ctx.lib.gpd_score_dev(ctx.handle, C.cast(handles, _lib.c_void_pp), len(models), X_dev, int(N), order, kind,
                                dptr(noise), dptr(pps_n), dc_dev, C.byref(bv), C.byref(bi), int(idx_offset)))
    return int(bi.value), bv.value


def refine_designs(models, X_start, criterion='BH', order=1, noise_var=None, pps=None, bounds=None, binary=(),
                   step=None, iters=25, shrink=0.5, tol=1e-6, scorer=None):
    """Continuous refinement of the best grid designs: batched compass (pattern) search.

    The reference only evaluates a fixed candidate set and takes ``np.argmax`` (demos/case_study.py:279-293).  Here
    every start point (e.g. the top-k designs of the grid) is improved by polling the 2 * D axis neighbours at the
    current step length; ALL polls of ALL start points form one candidate batch per iteration and go through the
    fused device scorer (``gpd_score``), so one iteration costs one pass of the hot path over K * 2D designs.  A point
    moves to its best improving neighbour, otherwise its step shrinks.  Derivative-free on purpose: the criterion is
    evaluated by exactly the kernels the parity tests cover, and the criterion value of every point is monotone
    non-decreasing.

    X_start (K, D) original units; bounds (2, D) [lo; hi] (default: the models' common training box); ``binary``
    design columns are held fixed; step (D,) initial step (default 5 % of the box); returns
    ``(X (K, D), dc (K,), evaluations)`` sorted by decreasing criterion.
    ``scorer(X) -> dc`` replaces the device scorer (host-logic tests)."""
    X = np.array(X_start, dtype=float, ndmin=2)
    K, D = X.shape
    if bounds is None:
        lo = np.max([m.xmin for m in models], axis=0)
        hi = np.min([m.xmax for m in models], axis=0)
    else:
        lo, hi = np.asarray(bounds[0], dtype=float), np.asarray(bounds[1], dtype=float)
    assert lo.shape == (D,) and hi.shape == (D,) and np.all(hi >= lo)
    free = np.array([d for d in range(D) if d not in set(binary)], dtype=int)
    assert len(free) > 0, 'no continuous design dimension to refine'
    h0 = 0.05 * (hi - lo) if step is None else np.broadcast_to(np.asarray(step, dtype=float), (D,))
    H = np.tile(h0, (K, 1))                       # per-point step lengths
    if scorer is None:
        def scorer(Xb):
            return score_designs(models, Xb, order=order, criterion=criterion, noise_var=noise_var, pps=pps)[2]
    X = np.clip(X, lo, hi)
    f = np.array(scorer(X), dtype=float)
    evals = K
    dirs = np.concatenate([np.eye(D)[free], -np.eye(D)[free]])          # (2F, D)
    for _ in range(iters):
        active = np.any(H[:, free] > tol * np.maximum(hi - lo, 1e-300)[free], axis=1)
        if not np.any(active):
            break
        ia = np.nonzero(active)[0]
        polls = np.clip(X[ia, None, :] + dirs[None, :, :] * H[ia, None, :], lo, hi)      # (A, 2F, D)
        fp = np.asarray(scorer(polls.reshape(-1, D)), dtype=float).reshape(len(ia), len(dirs))
        evals += fp.size
        fp = np.where(np.isnan(fp), -np.inf, fp)
        j = np.argmax(fp, axis=1)
        best = fp[np.arange(len(ia)), j]
        better = best > f[ia]
        mv = ia[better]
        X[mv] = polls[better, j[better]]
        f[mv] = best[better]
        H[ia[~better]] *= shrink
    o = np.argsort(-f, kind='stable')
    return X[o], f[o], evals


def shard_bounds(N, rank, world):
    """Contiguous block [lo, hi) of rank: ceil(N/world) candidates each."""
    per = -(-N // world)
    lo = min(N, rank * per)
    return lo, min(N, lo + per)


def combine_best(vals, idxs):
    """np.argmax semantics over per-shard winners: NaN wins, ties -> lowest global index."""
    best = None
    for v, i in zip(vals, idxs):
        if i < 0:
            continue
        if best is None:
            best = (v, i)
            continue
        bv, bi = best
        nv, nb = np.isnan(v), np.isnan(bv)
        if nv or nb:
            better = nv and (not nb or i < bi)
        else:
            better = v > bv or (v == bv and i < bi)
        if better:
            best = (v, i)
    return best


class ShardedScorer:
    """Data-parallel scoring over a torch.distributed group (NCCL on GPUs, gloo in CPU tests).

    ``local_score(lo, hi) -> (best_value, best_global_index)`` is the per-rank work; the only
    collective is an all-gather of one (float64, int64) pair per rank.
    """

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def argmax_allgather(self, best_val, best_idx, device=None):
        import torch
        dist = self.dist
        backend = dist.get_backend(self.group)
        # the rank's own GPU, not "current device": the library and torch share the CUDA runtime's current device
        dev = device if device is not None else ('cuda:%d' % default_device() if backend == 'nccl' else 'cpu')
        v = torch.tensor([best_val], dtype=torch.float64, device=dev)
        i = torch.tensor([best_idx], dtype=torch.int64, device=dev)
        vs = [torch.empty_like(v) for _ in range(self.world)]
        is_ = [torch.empty_like(i) for _ in range(self.world)]
        dist.all_gather(vs, v, group=self.group)
        dist.all_gather(is_, i, group=self.group)
        vals = [float(t.item()) for t in vs]
        idxs = [int(t.item()) for t in is_]
        return combine_best(vals, idxs)

    def score(self, N, local_score):
        lo, hi = shard_bounds(N, self.rank, self.world)
        if hi > lo:
            v, i = local_score(lo, hi)
        else:
            v, i = 0.0, -1
        return self.argmax_allgather(v, i)
