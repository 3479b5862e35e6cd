"""Host restatement of the tile recursion behind csrc/kernels_inv.cu (no GPU): the off-diagonal tiles of L^-1 by
recursive doubling over pairs of ts-tile diagonal blocks, with the intermediate T = L21 X11 parked in the mirrored
strictly-upper tiles of X and the k ranges the kernel uses (MODE 1: c .. end of the left half, MODE 2: start of the
right half .. r).  Guards the index arithmetic for block counts that are not powers of two; the CUDA kernels are held
against the round-1 substitution kernels in tests/test_device_build.py::test_setup_variants_agree."""
import numpy as np
import pytest


def tile_recursion(L, n, B):
    nblk = (n + B - 1) // B
    n_pad = nblk * B
    Lp = np.eye(n_pad)
    Lp[:n, :n] = L
    X = np.zeros((n_pad, n_pad))
    for J in range(nblk):                                   # diag_inv_kernel
        s = slice(J * B, (J + 1) * B)
        X[s, s] = np.linalg.inv(Lp[s, s])

    def ltile(r, k):                                        # rows >= n of L read as zero off the diagonal
        t = np.zeros((B, B))
        rows = min(B, max(0, n - r * B))
        t[:rows] = L[r * B:r * B + rows, k * B:(k + 1) * B]
        return t
    launches = 0
    ts = 1
    while ts < nblk:
        outs = [(r, c) for r in range(nblk) for c in range(nblk) if (c // ts) % 2 == 0 and r // ts == c // ts + 1]
        for r, c in outs:                                   # inv_gemm_kernel<1>
            half = (c // ts + 1) * ts
            acc = np.zeros((B, B))
            for kt in range(c, half):
                assert (kt + 1) * B <= n                    # the kernel reads these columns of L unguarded
                acc += ltile(r, kt) @ X[kt * B:(kt + 1) * B, c * B:(c + 1) * B]
            X[c * B:(c + 1) * B, r * B:(r + 1) * B] = acc   # mirrored position (c, r)
        res = {}
        for r, c in outs:                                   # inv_gemm_kernel<2>
            half = (c // ts + 1) * ts
            acc = np.zeros((B, B))
            for kt in range(half, r + 1):
                acc += X[r * B:(r + 1) * B, kt * B:(kt + 1) * B] @ X[c * B:(c + 1) * B, kt * B:(kt + 1) * B]
            res[(r, c)] = -acc
        for (r, c), v in res.items():
            X[r * B:(r + 1) * B, c * B:(c + 1) * B] = v
        launches += 2
        ts *= 2
    return X, Lp, launches


@pytest.mark.parametrize('n', [5, 8, 9, 17, 24, 33, 40, 41, 63, 64, 65, 72, 100, 129, 136])
def test_recursive_doubling_inverse(n):
    B = 8
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n))
    L = np.linalg.cholesky(A @ A.T + n * np.eye(n))
    X, Lp, launches = tile_recursion(L, n, B)
    nblk = (n + B - 1) // B
    assert launches == 2 * int(np.ceil(np.log2(nblk))) if nblk > 1 else launches == 0
    ref = np.linalg.inv(Lp)
    for r in range(nblk):                                   # the packing kernels read the lower tiles only
        for c in range(r + 1):
            blk = (slice(r * B, (r + 1) * B), slice(c * B, (c + 1) * B))
            assert np.max(np.abs(X[blk] - ref[blk])) < 1e-13 * max(1.0, np.abs(ref).max())
    for r in range(nblk):                                   # diagonal tiles stay lower triangular
        blk = X[r * B:(r + 1) * B, r * B:(r + 1) * B]
        assert np.all(np.triu(blk, 1) == 0)
