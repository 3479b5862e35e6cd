"""More than one GPU (skipped on a one-GPU box): the single-process MultiDeviceScorer (one context + one host thread per
device, gpd_score_sharded with the library's own NCCL communicator) returns exactly what one GPU returns on the whole
candidate set -- same criterion vector bit for bit, same argmax with the lowest-index tie-break across shards."""
import numpy as np
import pytest

import helpers
from gpdoemd_b200 import synthetic

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


def test_library_reports_its_nccl(gpu_ctx):
    import ctypes as C
    v = C.c_int32()
    assert gpu_ctx.lib.gpd_nccl_version(C.byref(v)) == 0 and v.value >= 22000


def test_sharded_call_with_a_single_rank_communicator(gpu_ctx):
    """world = 1: the all-gather degenerates, the result must equal gpd_score_multi (also with an EMPTY shard)."""
    import gpdoemd_b200 as g
    oms, ds = helpers.oracle_models(41, M=2, E=2, D=3, P=2, n=90, kernel='Matern52')
    gms = helpers.gpu_models(oms)
    X = synthetic.candidates(5, 300, 3, np.max([d['xlo'] for d in ds], axis=0), np.min([d['xhi'] for d in ds], axis=0))
    comm = g.LibraryComm(gpu_ctx, g.LibraryComm.unique_id(), 0, 1)
    try:
        a = g.score_designs_multi(gms, X, ['BH', 'JR'], noise_var=0.02, idx_offset=1000)
        b = g.score_designs_multi(gms, X, ['BH', 'JR'], noise_var=0.02, idx_offset=1000, comm=comm)
        for c in ('BH', 'JR'):
            assert a[c][0] == b[c][0] and a[c][1] == b[c][1] and np.array_equal(a[c][2], b[c][2])
        e = g.score_designs_multi(gms, X[:0], ['BH', 'JR'], noise_var=0.02, idx_offset=0, comm=comm)
        assert e['BH'][0] == -1 and e['BH'][2].shape == (0,)
    finally:
        comm.close()


@pytest.mark.parametrize('N', [1, 1000, 128 * 40 + 3])
def test_multi_device_scorer_equals_one_gpu(N, gpu_ctx):
    if _ndev() < 2:
        pytest.skip('needs at least 2 GPUs')
    import gpdoemd_b200 as g
    oms, ds = helpers.oracle_models(43, M=3, E=2, D=3, P=3, n=150, kernel='RBF')
    gms = helpers.gpu_models(oms)
    X = synthetic.candidates(6, N, 3, np.max([d['xlo'] for d in ds], axis=0), np.min([d['xhi'] for d in ds], axis=0))
    if N >= 1000:
        X[N - 7] = X[5]                      # an exact tie across shards: the lower global index must win
    one = g.score_designs_multi(gms, X, ['HR', 'BH', 'AW'], noise_var=0.01)
    sc = g.MultiDeviceScorer(gms, devices=list(range(min(_ndev(), 4))))
    try:
        for _ in range(2):
            many = sc.score_designs_multi(X, ['HR', 'BH', 'AW'], noise_var=0.01)
            for c in ('HR', 'BH', 'AW'):
                assert many[c][0] == one[c][0] == int(np.argmax(one[c][2])) and many[c][1] == one[c][1]
                assert np.array_equal(many[c][2], one[c][2])
    finally:
        sc.close()


def test_adopting_a_communicator_the_caller_owns(gpu_ctx):
    """gpd_comm_adopt: a host that already has an ncclComm_t (here created with NCCL's own C API through ctypes) hands it
    to the library; gpd_comm_destroy must leave it alive."""
    import ctypes as C
    import gpdoemd_b200 as g
    from gpdoemd_b200 import _lib
    from gpdoemd_b200._lib import check
    path = _lib._preload_nccl() or 'libnccl.so.2'
    nccl = C.CDLL(path)

    class UniqueId(C.Structure):
        _fields_ = [('internal', C.c_char * 128)]
    uid = UniqueId()
    assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
    import torch
    torch.cuda.set_device(gpu_ctx.device)
    comm = C.c_void_p()
    nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
    assert nccl.ncclCommInitRank(C.byref(comm), 1, uid, 0) == 0
    oms, ds = helpers.oracle_models(44, M=2, E=2, D=2, P=2, n=70, kernel='Matern32')
    gms = helpers.gpu_models(oms)
    X = synthetic.candidates(7, 200, 2, np.max([d['xlo'] for d in ds], axis=0), np.min([d['xhi'] for d in ds], axis=0))
    h = C.c_void_p()
    check(gpu_ctx.lib.gpd_comm_adopt(gpu_ctx.handle, comm, 0, 1, C.byref(h)))

    class Adopted:                        # the duck type score_designs_multi(comm=...) needs
        ctx, handle = gpu_ctx, h
    a = g.score_designs_multi(gms, X, ['BH'], noise_var=0.02)
    b = g.score_designs_multi(gms, X, ['BH'], noise_var=0.02, comm=Adopted)
    assert a['BH'][0] == b['BH'][0] and np.array_equal(a['BH'][2], b['BH'][2])
    gpu_ctx.lib.gpd_comm_destroy(h)
    nccl.ncclCommDestroy.argtypes = [C.c_void_p]
    assert nccl.ncclCommDestroy(comm) == 0          # still ours to destroy
