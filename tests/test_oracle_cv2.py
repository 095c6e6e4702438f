"""CPU: the oracle against a LIVE cv2 (when importable) on randomized inputs — the third-party primitives
the reference calls: BFMatcher.knnMatch, resize(0.5), Scharr(scale=3), addWeighted, invert."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("seed", range(4))
def test_knn_hamming_vs_cv2(oracle, seed):
    rng = np.random.default_rng(seed)
    n1, n2 = int(rng.integers(2, 400)), int(rng.integers(2, 400))
    d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
    d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    d2[rng.integers(0, n2, 10)] = d2[0]
    d1[rng.integers(0, n1, 5)] = d2[0]
    idx, dist = oracle.knn2_hamming(d1, d2)
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(d1, d2, 2)
    np.testing.assert_array_equal(idx, np.array([[a.trainIdx, b.trainIdx] for a, b in m]))
    np.testing.assert_array_equal(dist, np.array([[a.distance, b.distance] for a, b in m], np.float32))


def test_knn_fewer_than_k(oracle):
    d1 = np.random.default_rng(0).integers(0, 256, (5, 32), dtype=np.uint8)
    idx, dist = oracle.knn2_hamming(d1, d1[:1])
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(d1, d1[:1], 2)
    assert all(len(r) == 1 for r in m)                       # cv returns shorter lists
    assert (idx[:, 0] == 0).all() and (idx[:, 1] == -1).all()


@pytest.mark.parametrize("shape", [(480, 752), (376, 1241), (480, 640), (33, 67), (9, 5)])
def test_pyramid_vs_cv2(oracle, shape):
    img = np.random.default_rng(shape[0]).integers(0, 256, shape, dtype=np.uint8)
    cur = img
    for lvl, got in enumerate(oracle.pyramid(img)[1:], 1):
        if min(cur.shape) < 2:
            break
        cur = cv2.resize(cur, None, fx=0.5, fy=0.5)
        np.testing.assert_array_equal(got, cur, err_msg=f"level {lvl}")


@pytest.mark.parametrize("shape", [(480, 752), (30, 47), (3, 3), (1, 7), (7, 1)])
def test_scharr_vs_cv2(oracle, shape):
    img = np.random.default_rng(shape[1]).integers(0, 256, shape, dtype=np.uint8)
    gx, gy = oracle.scharr3(img)
    cx = cv2.Scharr(img, cv2.CV_16S, 1, 0, scale=3)
    cy = cv2.Scharr(img, cv2.CV_16S, 0, 1, scale=3)
    np.testing.assert_array_equal(gx, cx)
    np.testing.assert_array_equal(gy, cy)
    np.testing.assert_array_equal(oracle.grad_mag(gx, gy),
                                  cv2.addWeighted(cv2.convertScaleAbs(cx), 0.5, cv2.convertScaleAbs(cy), 0.5, 0))


def test_inv6_vs_cv2(oracle):
    rng = np.random.default_rng(3)
    for _ in range(100):
        J = rng.standard_normal((40, 6)).astype(np.float32) * np.array([1, 1, 0.5, 1e3, 1e3, 10], np.float32)
        A = (J.T.astype(np.float64) @ J.astype(np.float64)).astype(np.float32)
        ok, inv = oracle.inv6(A)
        _, ref = cv2.invert(A, flags=cv2.DECOMP_LU)
        np.testing.assert_array_equal(inv, ref)


@pytest.mark.parametrize("shape", [(480, 752), (33, 67), (7, 7), (6, 40)])
def test_fast9_vs_cv2(oracle, shape):
    rng = np.random.default_rng(shape[0] * 31 + shape[1])
    for img in (rng.integers(0, 256, shape, dtype=np.uint8),
                cv2.GaussianBlur(rng.integers(0, 256, shape, dtype=np.uint8), (5, 5), 1.0)):
        for thr, nm in ((20, True), (20, False), (5, True)):
            det = cv2.FastFeatureDetector_create(threshold=thr, nonmaxSuppression=nm, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
            kps = det.detect(img)
            xy, sc = oracle.fast9(img, thr, nm)
            np.testing.assert_array_equal(xy, np.array([[int(k.pt[0]), int(k.pt[1])] for k in kps], np.int32).reshape(-1, 2))
            np.testing.assert_array_equal(sc, np.array([int(k.response) for k in kps], np.int32))
