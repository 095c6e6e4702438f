"""CPU: the oracle against the committed golden fixtures (tests/golden/*.npz, made by make_golden.py):
cv2 4.13 for the third-party primitives, the reference's own Matcher.cpp for the filter chain, and the
oracle's pinned GN trace."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name))


def test_knn_hamming_matches_cv2_golden(oracle):
    g = load("knn_cv2.npz")
    i12, s12 = oracle.knn2_hamming(g["d1"], g["d2"])
    i21, s21 = oracle.knn2_hamming(g["d2"], g["d1"])
    np.testing.assert_array_equal(i12, g["idx12"])
    np.testing.assert_array_equal(s12, g["dist12"])
    np.testing.assert_array_equal(i21, g["idx21"])
    np.testing.assert_array_equal(s21, g["dist21"])
    # the tie fixture: d1[5] == d2[10] == d2[40] == d2[41] -> lowest train index first
    assert list(i12[5]) == [10, 40] and list(s12[5]) == [0.0, 0.0]


def test_knn_l2_matches_cv2_golden(oracle):
    g = load("knn_cv2.npz")
    for a, b, tag in ((g["f1"], g["f2"], "12"), (g["f2"], g["f1"], "21")):
        idx, dist = oracle.knn2_l2(a, b)
        np.testing.assert_array_equal(idx, g["fidx" + tag])
        np.testing.assert_allclose(dist, g["fdist" + tag], rtol=1e-6, atol=0)   # cv2 accumulates in float SIMD lanes


@pytest.mark.parametrize("tag", ["even", "odd", "kitti"])
def test_pyramid_scharr_match_cv2_golden(oracle, tag):
    g = load("camera_cv2.npz")
    pyr = oracle.pyramid(g[tag + "_l0"])
    for l in range(5):
        key = f"{tag}_l{l}"
        if key in g:
            np.testing.assert_array_equal(pyr[l], g[key])
    gx, gy = oracle.scharr3(g[tag + "_l0"])
    np.testing.assert_array_equal(gx, g[tag + "_gx"])
    np.testing.assert_array_equal(gy, g[tag + "_gy"])
    np.testing.assert_array_equal(oracle.grad_mag(gx, gy), g[tag + "_gm"])


def test_inv6_matches_cv2_golden(oracle):
    g = load("inv6_cv2.npz")
    for a, ref in zip(g["A"], g["Ainv"]):
        ok, inv = oracle.inv6(a)
        np.testing.assert_array_equal(inv, ref)          # bit-exact with cv2.invert(DECOMP_LU)
    assert not oracle.inv6(g["A"][-1])[0] and not g["Ainv"][-1].any()   # singular -> zeros


def test_filter_chain_matches_reference_matcher_cpp(oracle):
    """Golden outputs of the reference's own src/Matcher.cpp (compiled unmodified against oracle/cvshim)."""
    g = load("matcher_ref.npz")
    for i in range(int(g["n_cases"])):
        d1, d2, kp1 = g[f"c{i}_d1"], g[f"c{i}_d2"], g[f"c{i}_kp1"]
        nc = int(g[f"c{i}_ncells"])
        i12, s12 = oracle.knn2_hamming(d1, d2)
        i21, s21 = oracle.knn2_hamming(d2, d1)
        mq, mt, md = oracle.sym_matches(i12, s12, i21, s21, mode=0)
        np.testing.assert_array_equal(mq, g[f"c{i}_sym_q"])
        np.testing.assert_array_equal(mt, g[f"c{i}_sym_t"])
        order = oracle.sort_matches(mq, kp1)
        np.testing.assert_array_equal(mq[order], g[f"c{i}_sorted_q"])
        gq, gt, gd = oracle.grid_filter(mq, mt, md, order, kp1, 752, 480, nc)
        np.testing.assert_array_equal(gq, g[f"c{i}_good_q"])
        np.testing.assert_array_equal(gt, g[f"c{i}_good_t"])
        np.testing.assert_array_equal(gd, g[f"c{i}_good_d"])
        np.testing.assert_array_equal(kp1[gq], g[f"c{i}_prev_xy"])           # getGoodMatches
        np.testing.assert_array_equal(g[f"c{i}_kp2"][gt], g[f"c{i}_cur_xy"])


def test_defacto_vs_intended_symmetry(oracle):
    """Case 4 of the fixture: the 2->1 ratio test fails for d2 row 0 but the reference (reading the cleared
    vector's stale storage, Matcher.cpp:122-125) still reports the match; 'intended' mode drops it."""
    g = load("matcher_ref.npz")
    d1, d2 = g["c4_d1"], g["c4_d2"]
    i12, s12 = oracle.knn2_hamming(d1, d2)
    i21, s21 = oracle.knn2_hamming(d2, d1)
    assert i12[0, 0] == 0 and i21[0, 0] == 0 and s21[0, 0] > 0.8 * s21[0, 1]
    q0, _, _ = oracle.sym_matches(i12, s12, i21, s21, mode=0)
    q1, _, _ = oracle.sym_matches(i12, s12, i21, s21, mode=1)
    assert 0 in q0 and 0 in g["c4_sym_q"] and 0 not in q1


@pytest.mark.parametrize("tag", ["ref", "huber", "bilinear"])
def test_gn_oracle_pinned(oracle, tag):
    g = load("gn_oracle.npz")
    K = tuple(float(x) for x in g["K"])
    kw = {"ref": {}, "huber": dict(weight_mode=2, huber_k=12.0), "bilinear": dict(sample_mode=1)}[tag]
    r = oracle.track_pair(g["prev"], g["cur"], g["d1"], g["d2"], g["kp1"], K, g["prior"], n_cells=49,
                          opts=oracle.default_opts(first_lvl=2, **kw))
    np.testing.assert_array_equal(r["good_q"], g[tag + "_good_q"])
    np.testing.assert_array_equal(r["pose"], g[tag + "_pose"])
    np.testing.assert_array_equal(np.stack([t["pose"] for t in r["trace"]]), g[tag + "_trace_pose"])
    meta = np.array([[t["lvl"], t["iter"], t["n_valid"], t["updated"]] for t in r["trace"]], np.int32)
    np.testing.assert_array_equal(meta, g[tag + "_trace_meta"])


def test_intrinsics_known_answer(oracle):
    """EuRoC K through InitializePyramid (VISystem.cpp:1451-1493)."""
    K = oracle.init_pyramid(752, 480, 458.654, 457.296, 367.215, 248.375)
    fx = [458.654, 229.327, 114.6635, 57.33175, 28.665875]
    fy = [457.296, 228.648, 114.324, 57.162, 28.581]
    cx = [367.215, 183.3575, 91.42875, 45.464375, 22.4821875]
    cy = [248.375, 123.9375, 61.71875, 30.609375, 15.0546875]
    for l in range(5):
        assert K[l].fx == np.float32(np.float32(458.654) / 2 ** l)
        assert K[l].w == 752 >> l and K[l].h == 480 >> l
        np.testing.assert_allclose([K[l].fx, K[l].fy, K[l].cx, K[l].cy], [fx[l], fy[l], cx[l], cy[l]], rtol=2e-7)
        assert K[l].invfx == np.float32(1) / np.float32(K[l].fx)


def test_fast9_against_cv2_golden(oracle):
    """oracle/fast.c vs cv2.FastFeatureDetector (TYPE_9_16): corners, row-major order and scores, thresholds 0 / 20 / 50,
    with and without non-maximum suppression (tests/golden/fast_cv2.npz, make_fast_golden.py)."""
    g = load("fast_cv2.npz")
    names = [k[4:] for k in g.files if k.startswith("img_")]
    assert len(names) == 5
    total = 0
    for name in names:
        img = g["img_" + name]
        for thr in (0, 20, 50):
            for nm in (0, 1):
                xy, sc = oracle.fast9(img, thr, bool(nm))
                np.testing.assert_array_equal(xy, g[f"xy_{name}_{thr}_{nm}"], err_msg=f"{name} {thr} {nm}")
                np.testing.assert_array_equal(sc, g[f"sc_{name}_{thr}_{nm}"], err_msg=f"{name} {thr} {nm}")
                total += len(xy)
    assert total > 10000
