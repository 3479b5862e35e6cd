"""N > 1 path on CPU: two gloo ranks shard the candidates, score their blocks (with the oracle standing in
for the per-rank device work) and combine (max, index) pairs with the all-gather used on NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gpdoemd_b200.scoring import ShardedScorer, combine_best, pack_best, shard_bounds, unpack_best


def test_shard_bounds_cover_everything():
    for N in (1, 7, 100, 101, 10 ** 8):
        for W in (1, 2, 3, 8):
            blocks = [shard_bounds(N, r, W) for r in range(W)]
            assert blocks[0][0] == 0 and blocks[-1][1] == N
            for (a, b), (c, d) in zip(blocks, blocks[1:]):
                assert b == c and a <= b
            assert max(b - a for a, b in blocks) == -(-N // W)


def test_combine_best_numpy_argmax_semantics():
    assert combine_best([1.0, 3.0, 3.0], [5, 9, 7]) == (3.0, 7)          # tie -> lowest global index
    assert combine_best([1.0, 2.0], [4, -1]) == (1.0, 4)                # empty shard (-1) is ignored
    v, i = combine_best([1.0, float('nan'), 5.0], [0, 10, 20])
    assert np.isnan(v) and i == 10                                      # NaN wins like np.argmax
    v, i = combine_best([float('nan'), float('nan')], [30, 10])
    assert i == 10


def test_pack_best_round_trips_bit_exactly():
    vals = [1.5, float('nan'), -0.0, 1e-310, float('inf')]
    idxs = [0, 7, -1, 2 ** 40, 10 ** 8 - 1]
    v, i = unpack_best(pack_best(vals, idxs))
    assert np.array_equal(np.asarray(vals).view(np.int64), v.view(np.int64)) and list(i) == idxs


def _worker(rank, world, port, N, seed, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        dc = np.random.default_rng(seed).normal(size=N)       # every rank can regenerate the full vector
        dc[[N // 3, N - 2]] = dc.max() + 1.0                   # a tie across the two shards
        sc = ShardedScorer()

        def local(lo, hi):
            j = int(np.argmax(dc[lo:hi]))
            return float(dc[lo + j]), lo + j
        v, i = sc.score(N, local)
        # several criteria in ONE packed all-gather (what a fused multi-criterion pass produces per rank)
        lo, hi = shard_bounds(N, rank, world)
        dcs = [dc, -dc, np.where(np.arange(N) == N // 2, np.nan, dc)]
        if hi > lo:
            mine = [(float(d[lo + int(np.argmax(d[lo:hi]))]), lo + int(np.argmax(d[lo:hi]))) for d in dcs]
        else:
            mine = [(0.0, -1)] * len(dcs)
        multi = sc.argmax_allgather([m[0] for m in mine], [m[1] for m in mine])
        assert [m[1] for m in multi] == [int(np.argmax(d)) for d in dcs]
        out[rank] = (v, i, int(np.argmax(dc)), float(dc.max()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('N', [1, 10, 1001])
def test_two_rank_argmax_allgather(N):
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.get_context('spawn').Manager()      # no fork() of this multi-threaded test process
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, N, 5, out), nprocs=2, join=True)
    assert len(out) == 2
    for r in range(2):
        v, i, ref_i, ref_v = out[r]
        assert i == ref_i and v == ref_v
