"""GPU parity: Hamming kNN-2 through the C ABI vs the CPU oracle — bit-exact indices and distances,
including ties (lowest index first), ragged sizes and fewer than k train rows."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(ctx, d1, d2, n1=None, n2=None):
    import torch
    t1 = torch.from_numpy(d1).cuda()
    t2 = torch.from_numpy(d2).cuda()
    a1 = None if n1 is None else torch.tensor(n1, dtype=torch.int32).cuda()
    a2 = None if n2 is None else torch.tensor(n2, dtype=torch.int32).cuda()
    ctx.option("knn_impl", 0)          # this file covers the POPC kernel; test_gpu_knn_tc.py covers the tensor-core ones
    try:
        out = ctx.knn2_hamming(t1, t2, a1, a2)
        torch.cuda.synchronize()
    finally:
        ctx.option("knn_impl", 6)      # back to the default (auto)
    return [o.cpu().numpy() for o in out]


def _check_pair(oracle, d1, d2, got):
    i12, s12 = oracle.knn2_hamming(d1, d2)
    i21, s21 = oracle.knn2_hamming(d2, d1)
    np.testing.assert_array_equal(got[0], i12)
    np.testing.assert_array_equal(got[1], s12)
    np.testing.assert_array_equal(got[2], i21)
    np.testing.assert_array_equal(got[3], s21)


@pytest.mark.parametrize("n1,n2", [(1000, 1000), (64, 128), (65, 129), (1, 1), (2, 1), (1, 2), (3, 500), (500, 3),
                                   (257, 1023), (2000, 777)])
def test_knn_hamming_random(ctx, oracle, n1, n2):
    rng = np.random.default_rng(n1 * 7919 + n2)
    d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
    d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    _check_pair(oracle, d1, d2, _run(ctx, d1, d2))


def test_knn_hamming_ties_lowest_index(ctx, oracle):
    rng = np.random.default_rng(5)
    base = rng.integers(0, 256, (40, 32), dtype=np.uint8)
    d2 = np.concatenate([base, base, base[::-1], base])          # every row appears 4 times
    d1 = np.concatenate([base, rng.integers(0, 256, (30, 32), dtype=np.uint8)])
    d1[50:] &= 0xF0                                               # low-entropy rows: many equal distances
    d2[100:] &= 0xF0
    got = _run(ctx, d1, d2)
    _check_pair(oracle, d1, d2, got)
    assert (got[1][:40, 0] == 0).all() and (got[0][:40, 0] == np.arange(40)).all()
    assert (got[0][:40, 1] == np.arange(40) + 40).all()           # second copy wins the tie


def test_knn_hamming_extremes(ctx, oracle):
    d1 = np.zeros((70, 32), np.uint8)
    d2 = np.full((130, 32), 255, np.uint8)
    d2[7] = 0
    got = _run(ctx, d1, d2)
    _check_pair(oracle, d1, d2, got)
    assert got[1].max() == 256.0


def test_knn_hamming_batched_ragged(ctx, oracle):
    rng = np.random.default_rng(11)
    B, N1, N2 = 5, 300, 260
    d1 = rng.integers(0, 256, (B, N1, 32), dtype=np.uint8)
    d2 = rng.integers(0, 256, (B, N2, 32), dtype=np.uint8)
    n1 = [300, 17, 1, 0, 129]
    n2 = [260, 260, 5, 9, 0]
    got = _run(ctx, d1, d2, n1, n2)
    for b in range(B):
        i12, s12 = oracle.knn2_hamming(d1[b, :n1[b]], d2[b, :n2[b]])
        i21, s21 = oracle.knn2_hamming(d2[b, :n2[b]], d1[b, :n1[b]])
        np.testing.assert_array_equal(got[0][b, :n1[b]], i12)
        np.testing.assert_array_equal(got[1][b, :n1[b]], s12)
        np.testing.assert_array_equal(got[2][b, :n2[b]], i21)
        np.testing.assert_array_equal(got[3][b, :n2[b]], s21)
        assert (got[0][b, n1[b]:] == -1).all() and (got[2][b, n2[b]:] == -1).all()


def test_knn_hamming_full_size_properties(ctx):
    """BASELINE config sizes (5000 x 5000): size-independent properties instead of the O(N*M) oracle:
    a permuted copy is found at distance 0 by both directions, results are sorted, and the two
    directions agree on mutual nearest neighbours."""
    import torch
    rng = np.random.default_rng(99)
    n = 5000
    d1 = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    perm = rng.permutation(n)
    d2 = d1[perm]
    i12, s12, i21, s21 = _run(ctx, d1, d2)
    inv = np.empty(n, np.int64)
    inv[perm] = np.arange(n)
    assert (s12[:, 0] == 0).all() and (i12[:, 0] == inv).all()
    assert (s21[:, 0] == 0).all() and (i21[:, 0] == perm).all()
    assert (s12[:, 0] <= s12[:, 1]).all() and (s21[:, 0] <= s21[:, 1]).all()
    # second neighbours: spot-check 64 rows exhaustively with numpy
    rows = rng.choice(n, 64, replace=False)
    lut = np.array([bin(i).count("1") for i in range(256)], np.int32)
    for r in rows:
        d = lut[d1[r][None, :] ^ d2].sum(1)
        order = np.lexsort((np.arange(n), d))
        assert order[0] == i12[r, 0] and order[1] == i12[r, 1] and d[order[1]] == s12[r, 1]
