"""GPU parity of the tensor-core (tcgen05 / TMEM) Hamming kNN: raw int32 dot products against numpy, and the
kNN-2 results of both epilogues (32-bit keys, packed 16x2 keys) bit-exact against the CPU oracle — indices and
distances, ties to the lowest index, ragged and degenerate sizes, the extremes 0 and 256."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
# 1, 2: int8 tensor-core kernel (32-bit / packed epilogue, 2 = the default); 3, 4, 5: 4-bit operands (kind::mxf4): expanded in
# the kernel / pre-expanded rows copied by the producers / persistent CTAs fetching pre-swizzled tiles with cp.async.bulk;
# 6 = the default: 5 from 768 descriptors per set up, else 2
IMPLS = [1, 2, 3, 4, 5, 6] + [int(v) for v in os.environ.get("VSB_TEST_KNN_IMPLS", "").split(",") if v]


def _hamming_matrix(d1, d2):
    a = np.unpackbits(d1, axis=1).astype(np.int32)
    b = np.unpackbits(d2, axis=1).astype(np.int32)
    return a.sum(1)[:, None] + b.sum(1)[None, :] - 2 * (a @ b.T)


@pytest.mark.parametrize("n1,n2", [(128, 128), (130, 200), (5, 3), (1000, 1000)])
def test_tc_dot_products(ctx, n1, n2):
    import torch
    import vislam_b200 as vb
    rng = np.random.default_rng(n1 + 3 * n2)
    d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
    d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    d2[0] = d1[0]
    d2[-1] = ~d1[-1]
    t1, t2 = torch.from_numpy(d1).cuda(), torch.from_numpy(d2).cuda()
    dots = torch.full((n1, n2), -12345, dtype=torch.int32, device="cuda")
    fn = vb.lib().vsb_debug_knn_tc_dots
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    assert fn(ctx.handle, t1.data_ptr(), n1, t2.data_ptr(), n2, dots.data_ptr(), n2, None) == 0
    torch.cuda.synchronize()
    got = dots.cpu().numpy()
    want = 64 * (256 - 2 * _hamming_matrix(d1, d2))
    bad = np.argwhere(got != want)
    assert len(bad) == 0, (len(bad), got.size, bad[:8].tolist(), got[tuple(bad[0])], want[tuple(bad[0])])


def _run(ctx, impl, d1, d2, n1=None, n2=None):
    import torch
    ctx.option("knn_impl", impl)
    try:
        t1, t2 = torch.from_numpy(d1).cuda(), torch.from_numpy(d2).cuda()
        a1 = None if n1 is None else torch.tensor(n1, dtype=torch.int32).cuda()
        a2 = None if n2 is None else torch.tensor(n2, dtype=torch.int32).cuda()
        out = ctx.knn2_hamming(t1, t2, a1, a2)
        torch.cuda.synchronize()
    finally:
        ctx.option("knn_impl", 6)      # back to the default (auto)
    return [o.cpu().numpy() for o in out]


def _check_pair(oracle, d1, d2, got):
    i12, s12 = oracle.knn2_hamming(d1, d2)
    i21, s21 = oracle.knn2_hamming(d2, d1)
    for name, g, w in (("idx12", got[0], i12), ("dist12", got[1], s12), ("idx21", got[2], i21), ("dist21", got[3], s21)):
        bad = np.argwhere(g != w)
        assert len(bad) == 0, (name, len(bad), g.size, bad[:6].tolist(), g[tuple(bad[0])], w[tuple(bad[0])])


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("n1,n2", [(1000, 1000), (128, 128), (64, 128), (65, 129), (1, 1), (2, 1), (1, 2), (3, 500),
                                   (500, 3), (257, 1023), (2000, 777), (5000, 5000)])
def test_knn_tc_random(ctx, oracle, impl, n1, n2):
    rng = np.random.default_rng(n1 * 7919 + n2)
    d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
    d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    _check_pair(oracle, d1, d2, _run(ctx, impl, d1, d2))


@pytest.mark.parametrize("impl", IMPLS)
def test_knn_tc_ties_and_extremes(ctx, oracle, impl):
    rng = np.random.default_rng(5)
    base = rng.integers(0, 256, (40, 32), dtype=np.uint8)
    d2 = np.concatenate([base, base, base[::-1], base])          # every row appears 4 times
    d1 = np.concatenate([base, rng.integers(0, 256, (30, 32), dtype=np.uint8)])
    d1[50:] &= 0xF0                                               # low-entropy rows: many equal distances
    d2[100:] &= 0xF0
    got = _run(ctx, impl, d1, d2)
    _check_pair(oracle, d1, d2, got)
    assert (got[0][:40, 1] == np.arange(40) + 40).all()           # second copy wins the tie
    d1 = np.zeros((70, 32), np.uint8)
    d2 = np.full((300, 32), 255, np.uint8)                        # distance 256 everywhere ...
    d2[7] = 0                                                     # ... except one exact match
    d2[200] = 0
    got = _run(ctx, impl, d1, d2)
    _check_pair(oracle, d1, d2, got)
    assert got[3].max() == 256.0 and got[1].min() == 0.0 and got[1].max() == 0.0


@pytest.mark.parametrize("impl", IMPLS)
def test_knn_tc_batched_ragged(ctx, oracle, impl):
    rng = np.random.default_rng(11)
    B, N1, N2 = 6, 300, 270
    d1 = rng.integers(0, 256, (B, N1, 32), dtype=np.uint8)
    d2 = rng.integers(0, 256, (B, N2, 32), dtype=np.uint8)
    n1 = [300, 1, 129, 0, 299, 128]
    n2 = [270, 270, 2, 33, 1, 0]
    got = _run(ctx, impl, d1, d2, n1, n2)
    for b in range(B):
        a, c = d1[b, :n1[b]], d2[b, :n2[b]]
        if n1[b] and n2[b]:
            i12, s12 = oracle.knn2_hamming(a, c)
            i21, s21 = oracle.knn2_hamming(c, a)
        else:
            i12, s12 = np.full((n1[b], 2), -1, np.int32), np.zeros((n1[b], 2), np.float32)
            i21, s21 = np.full((n2[b], 2), -1, np.int32), np.zeros((n2[b], 2), np.float32)
        np.testing.assert_array_equal(got[0][b, :n1[b]], i12)
        np.testing.assert_array_equal(got[1][b, :n1[b]], s12)
        np.testing.assert_array_equal(got[2][b, :n2[b]], i21)
        np.testing.assert_array_equal(got[3][b, :n2[b]], s21)
        assert (got[0][b, n1[b]:] == -1).all() and (got[2][b, n2[b]:] == -1).all()


@pytest.mark.parametrize("impl", [5, 6])
@pytest.mark.parametrize("ragged", [False, True])
def test_knn_tc_chained_sequence(ctx, oracle, impl, ragged):
    """Set 2 of pair k IS set 1 of pair k + 1 (one [B+1,N,32] block, d2 = d1 + one set; counts likewise): the persistent 4-bit
    kernel expands every frame once. Same answers as independent pairs."""
    import torch
    rng = np.random.default_rng(23)
    B, N = 4, 800
    frames = rng.integers(0, 256, (B + 1, N, 32), dtype=np.uint8)
    counts = [800, 770, 1, 0, 799] if ragged else [N] * (B + 1)
    tf = torch.from_numpy(frames).cuda()
    tn = torch.tensor(counts, dtype=torch.int32).cuda() if ragged else None
    ctx.option("knn_impl", impl)
    try:
        d1, d2 = tf[:-1], tf[1:]
        assert d2.data_ptr() == d1.data_ptr() + N * 32 and d1.is_contiguous() and d2.is_contiguous()
        out = ctx.knn2_hamming(d1, d2, None if tn is None else tn[:-1], None if tn is None else tn[1:])
        torch.cuda.synchronize()
    finally:
        ctx.option("knn_impl", 6)
    got = [o.cpu().numpy() for o in out]
    for b in range(B):
        a, c = frames[b, :counts[b]], frames[b + 1, :counts[b + 1]]
        if len(a) and len(c):
            i12, s12 = oracle.knn2_hamming(a, c)
            i21, s21 = oracle.knn2_hamming(c, a)
        else:
            i12, s12 = np.full((len(a), 2), -1, np.int32), np.zeros((len(a), 2), np.float32)
            i21, s21 = np.full((len(c), 2), -1, np.int32), np.zeros((len(c), 2), np.float32)
        np.testing.assert_array_equal(got[0][b, :len(a)], i12)
        np.testing.assert_array_equal(got[1][b, :len(a)], s12)
        np.testing.assert_array_equal(got[2][b, :len(c)], i21)
        np.testing.assert_array_equal(got[3][b, :len(c)], s21)
        assert (got[0][b, len(a):] == -1).all() and (got[2][b, len(c):] == -1).all()


@pytest.mark.parametrize("impl", IMPLS)
def test_tracker_with_tensor_core_matcher(ctx, oracle, impl):
    import torch
    from vislam_b200 import synth
    from test_gpu_gn import TOL, rot_angle
    pairs = [synth.make_pair(n_feat=700, seed=s) for s in (1001, 1777, 1888)]
    ctx.option("knn_impl", impl)
    try:
        tr = ctx.tracker(752, 480, 700, pairs[0]["K"], n_cells=49, max_pairs=4)
        st = lambda k: torch.from_numpy(np.stack([p[k] for p in pairs])).cuda()
        pose, n_good = tr.track_pairs(st("prev"), st("cur"), st("d1"), st("d2"), st("kp1"), st("pose_prior"))
        torch.cuda.synchronize()
    finally:
        ctx.option("knn_impl", 6)      # back to the default (auto)
    pose, n_good = pose.cpu().numpy(), n_good.cpu().numpy()
    for b, p in enumerate(pairs):
        ref = oracle.track_pair(p["prev"], p["cur"], p["d1"], p["d2"], p["kp1"], p["K"], p["pose_prior"], n_cells=49)
        assert n_good[b] == len(ref["good_q"])
        assert rot_angle(pose[b][:4], ref["pose"][:4]) <= TOL
        assert np.abs(pose[b][4:] - ref["pose"][4:]).max() <= TOL
    tr.close()
