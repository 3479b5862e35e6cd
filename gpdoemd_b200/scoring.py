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
                        x_dev=None, N=None, dc_dev=None, dc_out=None, comm=None):
    """Several criteria in one pass over the candidates -> {name: (best_index, best_value, dc or None)}.

    ``comm`` (a ``LibraryComm``): X / x_dev is THIS rank's contiguous block of a sharded candidate set (it may be
    empty) starting at global index ``idx_offset``; the call then goes through ``gpd_score_sharded`` and the returned
    best index / value are the global winners after the all-gather (identical on every rank).

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
    if comm is not None:
        assert comm.ctx is ctx, 'communicator and models live on different devices'
        check(ctx.lib.gpd_score_sharded(ctx.handle, comm.handle, C.cast(handles, _lib.c_void_pp), len(models), xp, x_on_dev,
                                        int(N), order, kinds.ctypes.data_as(_lib.c_int32_p), nk, dptr(noise), dptr(pps_n),
                                        dcp, dc_on_dev, dptr(bv), bi, int(idx_offset)))
    else:
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


def criterion_with_gradient(models, X, criterion='BH', noise_var=None, pps=None):
    """Criterion value and its gradient with respect to the design at the rows of X (first-order Taylor marginals):
    ``-> f (K,), g (K, D)``.  Device: ``gpd_marginal_dx`` per rival model; host: the chain rule through the criterion
    (design_gradients.py).  HR, BH, BF, AW, JR."""
    from .design_gradients import criterion_gradient
    from .marginal import taylor_first_order_dx
    X = np.array(X, dtype=float, ndmin=2)
    K, D = X.shape
    M, E = len(models), models[0].num_outputs
    mu, S = np.empty((K, M, E)), np.empty((K, M, E, E))
    dmu, dS = np.empty((K, M, E, D)), np.empty((K, M, E, E, D))
    for i, m in enumerate(models):
        mu[:, i], S[:, i], dmu[:, i], dS[:, i] = taylor_first_order_dx(m, X)
    return criterion_gradient(criterion, mu, S, dmu, dS, noise_var, pps)


def _refine_lbfgs(models, X, lo, hi, free, criterion, noise_var, pps, iters, value_and_grad=None):
    """Joint bound-constrained L-BFGS over all start points: the objective sum_k f(x_k) is separable, so one
    iteration = ONE batched device pass over the K current points (values and gradients)."""
    from scipy.optimize import minimize
    K, D = X.shape
    if value_and_grad is None:
        def value_and_grad(Xb):
            return criterion_with_gradient(models, Xb, criterion, noise_var, pps)
    scale = np.maximum(hi - lo, 1e-300)[free]                    # optimise in box units
    evals = [0]

    def fg(z):
        Xb = X.copy()
        Xb[:, free] = lo[free] + z.reshape(K, len(free)) * scale
        f, g = value_and_grad(Xb)
        evals[0] += K
        f = np.where(np.isfinite(f), f, -1e300)
        g = np.where(np.isfinite(g), g, 0.0)
        return -float(np.sum(f)), -(g[:, free] * scale).ravel()
    z0 = ((X[:, free] - lo[free]) / scale).ravel()
    res = minimize(fg, z0, jac=True, method='L-BFGS-B', bounds=[(0.0, 1.0)] * len(z0),
                   options={'maxiter': int(iters), 'maxfun': int(4 * iters) + 4, 'ftol': 1e-15, 'gtol': 1e-10})
    Xr = X.copy()
    Xr[:, free] = lo[free] + np.clip(res.x.reshape(K, len(free)), 0.0, 1.0) * scale
    return Xr, evals[0]


def refine_designs(models, X_start, criterion='BH', order=1, noise_var=None, pps=None, bounds=None, binary=(),
                   step=None, iters=25, shrink=0.5, tol=1e-6, scorer=None, method='auto', value_and_grad=None):
    """Continuous refinement of the best grid designs.

    method='lbfgs' (default for first order, all five criteria): bound-constrained L-BFGS on the analytic design gradient
    (``criterion_with_gradient``: x-derivatives of mu and S on the device, chain rule through the criterion on the
    host), all start points in one batch per iteration; the result of every point is then re-scored by the fused
    scorer and kept only if it improves on its start, and a short compass search polishes it.  method='compass'
    (second order): the batched derivative-free pattern search described below.

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
    if scorer is None:
        def scorer(Xb):
            return score_designs(models, Xb, order=order, criterion=criterion, noise_var=noise_var, pps=pps)[2]
    X = np.clip(X, lo, hi)
    f = np.array(scorer(X), dtype=float)
    evals = K
    from .design_gradients import SUPPORTED
    if method == 'auto':
        method = 'lbfgs' if (order == 1 and criterion in SUPPORTED and (models is not None or value_and_grad is not None)) else 'compass'
    assert method in ('lbfgs', 'compass')
    if method == 'lbfgs':
        assert order == 1 and criterion in SUPPORTED, 'analytic design gradients: first order, %s' % (SUPPORTED,)
        Xr, ne = _refine_lbfgs(models, X, lo, hi, free, criterion, noise_var, pps, iters, value_and_grad)
        fr = np.array(scorer(Xr), dtype=float)
        evals += ne + K
        better = fr > f                                        # NaN never wins; monotone per start point
        X[better], f[better] = Xr[better], fr[better]
        iters = min(iters, 6)                                  # short derivative-free polish (bound corners, kinks of the clip)
        H0 = 0.002 * (hi - lo)
        step = H0 if step is None else np.minimum(np.broadcast_to(np.asarray(step, dtype=float), (D,)), H0)
    h0 = 0.05 * (hi - lo) if step is None else np.broadcast_to(np.asarray(step, dtype=float), (D,))
    H = np.tile(h0, (K, 1))                       # per-point step lengths
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


class LibraryComm:
    """One NCCL communicator owned by the C library and bound to a device context (``gpd_comm_init``).  ``unique_id()``
    on one rank, the 128 bytes to every rank over any channel, then ``LibraryComm(ctx, uid, rank, world)`` on all of
    them concurrently (threads of one process or one process per GPU)."""

    def __init__(self, ctx, uid, rank, world):
        assert len(uid) == _lib.NCCL_ID_BYTES
        self.ctx, self.rank, self.world = ctx, int(rank), int(world)
        h = C.c_void_p()
        check(ctx.lib.gpd_comm_init(ctx.handle, bytes(uid), self.rank, self.world, C.byref(h)))
        self.handle = h

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(_lib.NCCL_ID_BYTES)
        check(_lib.load().gpd_nccl_unique_id(buf))
        return buf.raw

    def close(self):
        if self.handle:
            if self.ctx.handle:                      # a closed context took the stream the communicator used with it
                self.ctx.lib.gpd_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pack_best(vals, idxs):
    """(float64 value, int64 index) pairs as ONE int64 vector [bits(v0), i0, bits(v1), i1, ...] (one collective)."""
    out = np.empty(2 * len(vals), dtype=np.int64)
    out[0::2] = np.asarray(vals, dtype=np.float64).view(np.int64)
    out[1::2] = np.asarray(idxs, dtype=np.int64)
    return out


def unpack_best(packed):
    packed = np.ascontiguousarray(packed, dtype=np.int64)
    return packed[0::2].view(np.float64).copy(), packed[1::2].copy()


class ShardedScorer:
    """Data-parallel scoring over a torch.distributed group: one process per GPU, candidates in contiguous blocks,
    GP factors replicated.  The only collective is ONE all-gather of the per-rank (float64 best value, int64 best
    global index) pairs of all criteria:

    * NCCL groups: the C library owns a communicator of its own (bootstrapped over the torch group) and
      ``score_designs_multi`` runs ``gpd_score_sharded`` -- scoring, ncclAllGather and the np.argmax fold are
      enqueued on the library stream, one host synchronisation per call;
    * other backends (gloo in the CPU tests): ``argmax_allgather`` packs the pairs into one int64 tensor.
    """

    def __init__(self, group=None, device=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.backend = dist.get_backend(group)
        self.device = default_device() if device is None else int(device)
        self.comm = None
        if self.backend == 'nccl':
            import torch
            ctx = get_context(self.device)
            dev = torch.device('cuda', self.device)
            uid = LibraryComm.unique_id() if self.rank == 0 else bytes(_lib.NCCL_ID_BYTES)
            t = torch.frombuffer(bytearray(uid), dtype=torch.uint8).to(dev)
            dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            self.comm = LibraryComm(ctx, bytes(t.cpu().numpy().tobytes()), self.rank, self.world)

    def close(self):
        if self.comm is not None:
            self.comm.close()
            self.comm = None

    def score_designs_multi(self, models, X_local, criteria, idx_offset, **kw):
        """This rank's block through the fused scorer + the in-stream all-gather -> global winners per criterion."""
        assert self.comm is not None, 'gpd_score_sharded needs an NCCL process group (GPU ranks)'
        return score_designs_multi(models, X_local, criteria, idx_offset=idx_offset, comm=self.comm, **kw)

    def argmax_allgather(self, best_val, best_idx, device=None):
        """Fold per-rank winners of one or several criteria: scalars -> (value, index); sequences -> list of pairs."""
        import torch
        scalar = np.ndim(best_val) == 0
        vals, idxs = np.atleast_1d(best_val), np.atleast_1d(best_idx)
        dev = device if device is not None else ('cuda:%d' % self.device if self.backend == 'nccl' else 'cpu')
        mine = torch.from_numpy(pack_best(vals, idxs)).to(dev)
        gathered = [torch.empty_like(mine) for _ in range(self.world)]
        self.dist.all_gather(gathered, mine, group=self.group)
        allp = torch.stack(gathered).cpu().numpy()                  # (world, 2 nk): one device -> host copy
        out = []
        for k in range(len(vals)):
            v, i = unpack_best(allp[:, 2 * k:2 * k + 2].ravel())
            out.append(combine_best(v.tolist(), i.tolist()))
        return out[0] if scalar else out

    def score(self, N, local_score):
        lo, hi = shard_bounds(N, self.rank, self.world)
        if hi > lo:
            v, i = local_score(lo, hi)
        else:
            v, i = 0.0, -1
        return self.argmax_allgather(v, i)


class MultiDeviceScorer:
    """The reference's single-process design loop (demos/case_study.py:279-293) over SEVERAL GPUs of one box: one
    context, one replica of every model's GP state and one host thread per device; candidates are cut into contiguous
    blocks, each thread runs ``gpd_score_sharded`` on its block, and the NCCL all-gather inside that call makes every
    thread return the global winners (ties -> lowest global index, like ``np.argmax`` on the whole grid)."""

    def __init__(self, models, devices):
        from concurrent.futures import ThreadPoolExecutor
        self.devices = [int(d) for d in devices]
        assert len(self.devices) >= 1 and len(set(self.devices)) == len(self.devices)
        self.models = list(models)
        W = len(self.devices)
        self.pool = ThreadPoolExecutor(W)
        self.replicas = [[m if m._device == d else m.replicate(d) for m in self.models] for d in self.devices]
        uid = LibraryComm.unique_id()
        # ncclCommInitRank is collective: every rank's thread must be inside it at the same time
        self.comms = list(self.pool.map(lambda r: LibraryComm(get_context(self.devices[r]), uid, r, W), range(W)))

    def close(self):
        for c in self.comms:
            c.close()
        self.comms = []
        self.pool.shutdown()

    def score_designs_multi(self, X, criteria, order=1, noise_var=None, pps=None, return_dc=True):
        """-> {name: (best_index, best_value, dc or None)} over the whole X (host array, original units)."""
        X = _lib.as_f64(X)
        N, W = len(X), len(self.devices)
        for reps in self.replicas:                       # pmean / Sigma follow the caller's models
            for src, rep in zip(self.models, reps):
                if rep is not src:
                    rep.pmean, rep.Sigma = src.pmean, src.Sigma

        def work(r):
            lo, hi = shard_bounds(N, r, W)
            return score_designs_multi(self.replicas[r], X[lo:hi], criteria, order=order, noise_var=noise_var, pps=pps,
                                       return_dc=return_dc, idx_offset=lo, comm=self.comms[r])
        parts = list(self.pool.map(work, range(W)))
        out = {}
        for c in criteria:
            assert all(p[c][0] == parts[0][c][0] for p in parts), 'ranks disagree on the global argmax'
            dc = np.concatenate([p[c][2] for p in parts]) if return_dc else None
            out[c] = (parts[0][c][0], parts[0][c][1], dc)
        return out
