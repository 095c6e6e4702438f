"""CPU: hand-checkable cases and property tests (hypothesis) for the oracle's Matcher chain, SE3 arithmetic
and GN fixed point."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st


def test_ratio_filter_hand_case(oracle):
    idx = np.array([[3, 1], [2, 0], [5, -1], [-1, -1]], np.int32)
    dist = np.array([[10, 20], [17, 20], [1, 0], [0, 0]], np.float32)
    keep = np.zeros(4, np.uint8)
    oracle.lib().vso_nn_filter(idx.reshape(-1), dist.reshape(-1), 4, oracle.RATIO, keep)
    assert list(keep) == [1, 0, 0, 0]          # 10 <= 0.8*20 keeps; 17 > 16 clears; < 2 neighbours clears
    # boundary: d0 == 0.8f*d1 is kept only if d0 <= double(0.8f)*d1  (0.8f widens to 0.800000011920929)
    dist = np.array([[8, 10]], np.float32)
    oracle.lib().vso_nn_filter(idx[:1].reshape(-1).copy(), dist.reshape(-1), 1, oracle.RATIO, keep)
    assert keep[0] == 1
    oracle.lib().vso_nn_filter(idx[:1].reshape(-1).copy(), dist.reshape(-1), 1, 0.8, keep)
    assert keep[0] == 1                         # 8 > 8.000000000000002 is false as well


def test_grid_filter_hand_case(oracle):
    """2x2 grid on a 100x100 image; strict '<' keeps the first of equal distances in y-sorted order."""
    kp = np.array([[10, 10], [20, 5], [60, 10], [10, 60], [90, 90], [95, 99]], np.float32)
    mq = np.arange(6, dtype=np.int32)
    mt = mq + 100
    md = np.array([5, 5, 7, 3, 9, 2], np.float32)
    order = oracle.sort_matches(mq, kp)
    assert list(order) == [1, 0, 2, 3, 4, 5]                  # y ascending, stable
    gq, gt, gd = oracle.grid_filter(mq, mt, md, order, kp, 100, 100, 4)
    assert list(gq) == [1, 2, 3, 5] and list(gt) == [101, 102, 103, 105] and list(gd) == [5, 7, 3, 2]
    assert oracle.grid_filter(mq[:0], mt[:0], md[:0], order[:0], kp, 100, 100, 4)[0].size == 0   # App. B-3


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2 ** 31), st.integers(2, 80), st.integers(2, 80))
def test_match_set_is_permutation_invariant(seed, n1, n2):
    """Permuting the train set permutes trainIdx and nothing else (the symmetric match SET is unchanged)."""
    from oracle import vso
    rng = np.random.default_rng(seed)
    d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
    d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    perm = rng.permutation(n2)
    a = vso.sym_matches(*vso.knn2_hamming(d1, d2), *vso.knn2_hamming(d2, d1))
    b = vso.sym_matches(*vso.knn2_hamming(d1, d2[perm]), *vso.knn2_hamming(d2[perm], d1))
    sa = {(int(q), int(t)) for q, t in zip(a[0], a[1])}
    sb = {(int(q), int(perm[t])) for q, t in zip(b[0], b[1])}
    # exact distance ties may resolve to a different (equally near) neighbour after permutation: compare where unique
    i12, s12 = vso.knn2_hamming(d1, d2)
    unique = {q for q in range(n1) if s12[q, 0] != s12[q, 1]}
    assert {m for m in sa if m[0] in unique} == {m for m in sb if m[0] in unique}


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2 ** 31))
def test_ratio_monotone(seed):
    """A stricter ratio never adds matches."""
    from oracle import vso
    rng = np.random.default_rng(seed)
    d1 = rng.integers(0, 256, (60, 32), dtype=np.uint8)
    d2 = d1.copy()
    d2 ^= np.packbits(rng.random((60, 256)) < 0.1, axis=1)
    k12, k21 = vso.knn2_hamming(d1, d2), vso.knn2_hamming(d2, d1)
    sizes = [len(vso.sym_matches(*k12, *k21, ratio=r)[0]) for r in (0.5, 0.7, 0.8, 0.95, 1.0)]
    assert sizes == sorted(sizes)


def _rot(q):
    x, y, z, w = [float(v) for v in q]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


@settings(max_examples=50, deadline=None)
@given(st.lists(st.floats(-0.3, 0.3), min_size=6, max_size=6))
def test_se3_exp_matches_closed_form(delta):
    """SE3::exp (se3.hpp:723-742) against the matrix exponential computed in float64."""
    from oracle import vso
    from scipy.linalg import expm
    d = np.array(delta, np.float32)
    pose = vso.se3_exp(d)
    w = d[3:].astype(np.float64)
    H = np.zeros((4, 4))
    H[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
    H[:3, 3] = d[:3]
    E = expm(H)
    np.testing.assert_allclose(_rot(pose[:4]), E[:3, :3], atol=3e-6)
    # Sophus evaluates (1 - cos(theta)) / theta^2 in float32 (se3.hpp:738): for small theta the subtraction cancels
    # (cos rounds to within 1 ulp of 1), so V carries an error of up to min(theta/2, ~2 ulp / theta) per unit of
    # upsilon.  That is reference behaviour the oracle must keep, so it is part of the tolerance.
    theta = float(np.linalg.norm(w))
    cancel = float(np.linalg.norm(d[:3])) * min(0.5 * theta, 2.4e-7 / max(theta, 1e-30))
    np.testing.assert_allclose(pose[4:], E[:3, 3], atol=3e-6 + cancel)
    assert abs(np.linalg.norm(pose[:4]) - 1) < 1e-5          # SOPHUS_ENSURE in so3.hpp:562-566


def test_se3_mul_and_matrix(oracle):
    rng = np.random.default_rng(1)
    for _ in range(20):
        a = oracle.se3_exp(rng.uniform(-0.2, 0.2, 6).astype(np.float32))
        b = oracle.se3_exp(rng.uniform(-0.2, 0.2, 6).astype(np.float32))
        c = oracle.se3_mul(a, b)
        np.testing.assert_allclose(oracle.se3_matrix(c), oracle.se3_matrix(a).astype(np.float64) @ oracle.se3_matrix(b),
                                   atol=2e-6)
    ident = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    np.testing.assert_array_equal(oracle.se3_mul(ident, a), a)
    np.testing.assert_array_equal(oracle.se3_exp(np.zeros(6, np.float32)), ident)      # Taylor branch, theta < 1e-5


def test_initial_pose_roundtrip(oracle):
    """VISystem.cpp:1135-1168: RPY2rotationMatrix(-rotationMatrix2RPY(R)) — the inverse rotation to FIRST order
    only (negating ZYX Euler angles is not an exact inverse; reproduced as the reference has it), t = -t_res."""
    from vislam_b200 import synth
    R = synth.so3_exp([0.01, -0.02, 0.015])
    pose = oracle.initial_pose(np.eye(3), R, [0.1, -0.2, 0.3])
    np.testing.assert_allclose(_rot(pose[:4]), R.T, atol=5e-4)
    Rz = synth.so3_exp([0, 0, 0.3])                       # single-axis rotation: exact inverse
    np.testing.assert_allclose(_rot(oracle.initial_pose(np.eye(3), Rz, [0, 0, 0])[:4]), Rz.T, atol=1e-6)
    np.testing.assert_allclose(pose[4:], [-0.1, 0.2, -0.3], atol=0)


def test_gn_fixed_point_on_identical_frames(oracle):
    """Noise-free: cur == prev and identity prior -> residuals are all zero, error 0, pose stays the identity."""
    from vislam_b200 import synth
    p = synth.make_pair(w=188, h=120, n_feat=120, K=(114.6635, 114.324, 91.42875, 61.71875), seed=77)
    ident = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    r = oracle.track_pair(p["prev"], p["prev"], p["d1"], p["d2"], p["kp1"], p["K"], ident, n_cells=49,
                          opts=oracle.default_opts(first_lvl=2))
    assert all(t["error"] == 0.0 for t in r["trace"])
    np.testing.assert_allclose(r["pose"], ident, atol=1e-7)


def test_warp_identity_and_translation(oracle):
    K = oracle.init_pyramid(752, 480, 458.654, 457.296, 367.215, 248.375)
    pts = np.array([[100, 50, 1, 1], [367.215, 248.375, 1, 1], [700, 400, 1, 1]], np.float32)
    ident = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    out = oracle.warp(pts, ident, K[0])
    np.testing.assert_allclose(out[:, :2], pts[:, :2], atol=1e-4)
    shifted = oracle.warp(pts, np.array([0, 0, 0, 1, 0.01, 0, 0], np.float32), K[0])
    np.testing.assert_allclose(shifted[:, 0] - pts[:, 0], 0.01 * 458.654, atol=1e-3)   # unit depth


def test_candidates_hand_case(oracle):
    c = oracle.candidates(np.array([[100.0, 50.0]], np.float32), 0, 752, 480)
    assert c.shape == (121, 4) and (c[:, 2:] == 1).all()
    assert c[0, 0] == 95 and c[0, 1] == 45 and c[1, 1] == 46 and c[-1, 0] == 105     # x outer, y inner
    c1 = oracle.candidates(np.array([[100.0, 50.0]], np.float32), 1, 376, 240)
    assert c1.shape == (49, 4)                                                        # half-size 3 at level 1
    edge = oracle.candidates(np.array([[0.0, 0.0]], np.float32), 0, 752, 480)
    assert edge.shape == (25, 4) and edge[:, :2].min() == 1                           # 0 < i, 0 < j (Camera.cpp:393)
    many = oracle.candidates(np.tile(np.array([[300.0, 200.0]], np.float32), (250, 1)), 0, 752, 480)
    assert many.shape[0] == 200 * 121                                                 # 200-feature cap (Camera.cpp:382)
