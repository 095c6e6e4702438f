"""GPU parity: float-descriptor (NORM_L2) kNN-2 through the C ABI vs the CPU oracle — bit-exact indices AND
distances (sqrtf of the float-difference, double-accumulated sum, reference src/Matcher.cpp:55 -> cv::BFMatcher),
including exact ties (lowest index first), ragged sizes, fewer than k train rows, odd descriptor lengths, and the
cv2 golden fixture (indices equal; distances to 1e-6 relative, cv2 accumulates in float SIMD lanes)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(autouse=True, params=[1, 0], ids=["tensor-core", "exact-fp64"])
def l2_impl(request, ctx):
    """Every case runs through both float-kNN implementations: knn_l2_impl 1 = tcgen05 TF32x3 distance GEMM + exact
    re-check (dim <= 64, dim % 8 == 0; other dimensions dispatch to the exact kernel) and 0 = the exact FP64 kernel."""
    ctx.option("knn_l2_impl", request.param)
    yield request.param
    ctx.option("knn_l2_impl", 1)


def _fallback_rows(ctx):
    import ctypes
    import vislam_b200 as vb
    v = ctypes.c_longlong()
    vb.check(vb.lib().vsb_debug_l2_fallback_rows(ctx.handle, ctypes.byref(v)), ctx.handle)
    return v.value


def _run(ctx, d1, d2, n1=None, n2=None):
    import torch
    t1 = torch.from_numpy(d1).cuda()
    t2 = torch.from_numpy(d2).cuda()
    a1 = None if n1 is None else torch.tensor(n1, dtype=torch.int32).cuda()
    a2 = None if n2 is None else torch.tensor(n2, dtype=torch.int32).cuda()
    out = ctx.knn2_l2(t1, t2, a1, a2)
    torch.cuda.synchronize()
    return [o.cpu().numpy() for o in out]


def _check_pair(oracle, d1, d2, got):
    i12, s12 = oracle.knn2_l2(d1, d2)
    i21, s21 = oracle.knn2_l2(d2, d1)
    np.testing.assert_array_equal(got[0], i12)
    np.testing.assert_array_equal(got[1], s12)      # bit-exact float distances
    np.testing.assert_array_equal(got[2], i21)
    np.testing.assert_array_equal(got[3], s21)


@pytest.mark.parametrize("n1,n2,dim", [(1000, 1000, 64), (64, 64, 64), (65, 129, 64), (1, 1, 64), (2, 1, 64),
                                       (1, 2, 64), (3, 500, 64), (500, 3, 64), (257, 1023, 128), (300, 200, 61),
                                       (130, 70, 7), (128, 128, 64), (129, 127, 32), (400, 300, 8), (333, 777, 40),
                                       (5000, 5000, 64), (2, 2, 16), (3, 3, 64), (4, 2, 24)])
def test_knn_l2_random(ctx, oracle, n1, n2, dim):
    rng = np.random.default_rng(n1 * 7919 + n2 + dim)
    d1 = rng.standard_normal((n1, dim)).astype(np.float32)
    d2 = rng.standard_normal((n2, dim)).astype(np.float32)
    d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
    d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    _check_pair(oracle, d1, d2, _run(ctx, d1, d2))


def test_knn_l2_synthetic_tum_shaped(ctx, oracle):
    from vislam_b200 import synth
    d1 = synth.float_descriptors(1500, 3001)
    d2, _ = synth.perturb_float(d1, 3002)
    _check_pair(oracle, d1, d2, _run(ctx, d1, d2))


def test_knn_l2_ties_lowest_index(ctx, oracle):
    rng = np.random.default_rng(5)
    base = rng.standard_normal((40, 64)).astype(np.float32)
    d2 = np.concatenate([base, base, base[::-1], base])           # every row appears 4 times
    d1 = np.concatenate([base, rng.standard_normal((30, 64)).astype(np.float32)])
    d1[50:] = np.round(d1[50:])                                    # coarse values: many equal distances
    d2[100:] = np.round(d2[100:])
    got = _run(ctx, d1, d2)
    _check_pair(oracle, d1, d2, got)
    assert (got[1][:40, 0] == 0).all() and (got[0][:40, 0] == np.arange(40)).all()
    assert (got[0][:40, 1] == np.arange(40) + 40).all()            # second copy wins the tie


def test_knn_l2_large_magnitudes(ctx, oracle):
    rng = np.random.default_rng(9)                                 # SIFT-like 0..255 integer-valued floats
    d1 = rng.integers(0, 256, (200, 128)).astype(np.float32)
    d2 = rng.integers(0, 256, (333, 128)).astype(np.float32)
    _check_pair(oracle, d1, d2, _run(ctx, d1, d2))


def test_knn_l2_batched_ragged(ctx, oracle):
    rng = np.random.default_rng(11)
    B, N1, N2 = 5, 150, 170
    d1 = rng.standard_normal((B, N1, 64)).astype(np.float32)
    d2 = rng.standard_normal((B, N2, 64)).astype(np.float32)
    n1 = [150, 1, 64, 0, 149]
    n2 = [170, 170, 2, 33, 1]
    got = _run(ctx, d1, d2, n1, n2)
    for b in range(B):
        a, c = d1[b, :n1[b]], d2[b, :n2[b]]
        i12, s12 = oracle.knn2_l2(a, c) if n1[b] else (np.zeros((0, 2), np.int32), np.zeros((0, 2), np.float32))
        i21, s21 = oracle.knn2_l2(c, a) if n2[b] else (np.zeros((0, 2), np.int32), np.zeros((0, 2), np.float32))
        if n2[b] == 0:
            i12, s12 = np.full((n1[b], 2), -1, np.int32), np.zeros((n1[b], 2), np.float32)
        if n1[b] == 0:
            i21, s21 = np.full((n2[b], 2), -1, np.int32), np.zeros((n2[b], 2), np.float32)
        np.testing.assert_array_equal(got[0][b, :n1[b]], i12)
        np.testing.assert_array_equal(got[1][b, :n1[b]], s12)
        np.testing.assert_array_equal(got[2][b, :n2[b]], i21)
        np.testing.assert_array_equal(got[3][b, :n2[b]], s21)
        assert (got[0][b, n1[b]:] == -1).all() and (got[2][b, n2[b]:] == -1).all()


def test_knn_l2_cv2_golden(ctx):
    g = np.load(os.path.join(HERE, "golden", "knn_cv2.npz"))
    got = _run(ctx, g["f1"], g["f2"])
    np.testing.assert_array_equal(got[0], g["fidx12"])
    np.testing.assert_array_equal(got[2], g["fidx21"])
    np.testing.assert_allclose(got[1], g["fdist12"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(got[3], g["fdist21"], rtol=1e-6, atol=1e-7)


def test_knn_l2_tensor_core_fallback_is_rare_and_exact(ctx, oracle, l2_impl):
    """The tensor-core path may only SELECT: rows whose candidate set cannot be proven complete are recomputed
    exhaustively.  On well-separated data that is (almost) never needed; on near-duplicate clusters it is, and the
    result is still bit-exact."""
    if l2_impl != 1:
        pytest.skip("tensor-core path only")
    from vislam_b200 import synth
    d1 = synth.float_descriptors(2000, 3001)
    d2, _ = synth.perturb_float(d1, 3002)
    got = _run(ctx, d1, d2)
    rows = _fallback_rows(ctx)
    _check_pair(oracle, d1, d2, got)
    assert rows <= 40, rows                                        # 4000 rows in total
    # clusters of near-duplicates (differences ~1e-6): the approximate scores cannot separate them
    rng = np.random.default_rng(21)
    centres = rng.standard_normal((25, 64)).astype(np.float32)
    c1 = (centres[rng.integers(0, 25, 600)] + 1e-6 * rng.standard_normal((600, 64))).astype(np.float32)
    c2 = (centres[rng.integers(0, 25, 700)] + 1e-6 * rng.standard_normal((700, 64))).astype(np.float32)
    got = _run(ctx, c1, c2)
    rows = _fallback_rows(ctx)
    _check_pair(oracle, c1, c2, got)
    assert rows > 100, rows


def test_knn_l2_scaled_and_offset_data(ctx, oracle):
    """Norms far from 1 and a large common offset (cancellation in |a|^2 + |b|^2 - 2ab): the bound scales with the data."""
    rng = np.random.default_rng(33)
    for scale, offset in ((1e3, 0.0), (1e-3, 0.0), (1.0, 50.0), (255.0, 128.0)):
        d1 = (rng.standard_normal((300, 64)) * scale + offset).astype(np.float32)
        d2 = (rng.standard_normal((260, 64)) * scale + offset).astype(np.float32)
        d2[:100] = d1[:100] + (0.01 * scale * rng.standard_normal((100, 64))).astype(np.float32)
        _check_pair(oracle, d1, d2, _run(ctx, d1, d2))
