"""The ORB oracle (oracle/orb.c: cv::ORB::detectAndCompute, one pyramid level) against cv2 — golden fixture and live."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_orb_pattern_table_is_the_probed_one(oracle):
    """The .inc tables (oracle and product) are the table recovered from cv2 by tools/probe_orb_pattern.py."""
    probe = np.load(os.path.join(GOLD, "orb_pattern_probe.npy")).reshape(-1)
    root = os.path.dirname(os.path.dirname(__file__))
    for path in ("oracle/orb_pattern.inc", "vi-slam_b200/csrc/orb_pattern.inc"):
        txt = open(os.path.join(root, path)).read()
        body = txt[txt.index("*/") + 2:]
        vals = np.array([int(v) for v in body.replace("\n", " ").split(",") if v.strip()], np.int64)
        assert np.array_equal(vals, probe), path
    assert probe.min() >= -13 and probe.max() <= 13 and probe.size == 1024


@pytest.mark.parametrize("name", ["noise", "odd", "rects"])
@pytest.mark.parametrize("n", [60, 400, 5000])
def test_orb_matches_cv2_golden(oracle, name, n):
    g = np.load(os.path.join(GOLD, "orb_cv2.npz"))
    assert np.array_equal(oracle.orb_gauss_kernel(), g["gauss_kernel"])
    xy, resp, ang, desc = oracle.orb_detect_compute(g[f"{name}_img"], n)
    assert np.array_equal(xy, g[f"{name}_{n}_xy"])                 # same key-point set, row-major
    assert np.array_equal(resp, g[f"{name}_{n}_resp"])             # Harris responses, bit for bit
    assert np.array_equal(ang, g[f"{name}_{n}_angle"])             # orientations, bit for bit
    assert np.array_equal(desc, g[f"{name}_{n}_desc"])             # rBRIEF descriptors
    if n == 60:
        assert len(xy) >= 60                                       # ties at the threshold are kept, as in OpenCV


@pytest.mark.parametrize("name", ["noise", "odd", "rects"])
@pytest.mark.parametrize("tag", ["d", "e"])
def test_orb_pyramid_matches_cv2_golden(oracle, name, tag):
    """The full detector (scale pyramid; 'd' = the reference's setting: 8 levels, factor 1.2)."""
    g = np.load(os.path.join(GOLD, "orb_cv2.npz"))
    n, sf, nl = g[f"{name}_pyr{tag}_cfg"]
    xy, octv, resp, ang, desc = oracle.orb_detect_compute_pyr(g[f"{name}_img"], int(n), float(sf), int(nl))
    assert np.array_equal(xy, g[f"{name}_pyr{tag}_xy"])
    assert np.array_equal(octv, g[f"{name}_pyr{tag}_octave"])
    assert np.array_equal(resp, g[f"{name}_pyr{tag}_resp"])
    assert np.array_equal(ang, g[f"{name}_pyr{tag}_angle"])
    assert np.array_equal(desc, g[f"{name}_pyr{tag}_desc"])
    scale = np.array([oracle.lib().vso_orb_level_scale(float(sf), int(o)) for o in octv], np.float32)
    assert np.array_equal(np.float32(31.0) * scale, g[f"{name}_pyr{tag}_size"])        # cv::KeyPoint::size = patchSize * scale


def test_resize_linear_exact_matches_cv2_golden(oracle):
    g = np.load(os.path.join(GOLD, "orb_cv2.npz"))
    for i, (dw, dh) in enumerate(((220, 167), (132, 100), (263, 199), (97, 61))):
        assert np.array_equal(oracle.resize_linear_exact(g["noise_img"], dw, dh), g[f"resize{i}"])


def test_orb_live_vs_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(99)
    for w, h, n in ((320, 240, 300), (333, 201, 1200), (752, 480, 1000)):
        img = cv2.GaussianBlur((rng.random((h, w)) * 255).astype(np.uint8), (5, 5), 1.0 + rng.random())
        orb = cv2.ORB_create(nfeatures=n, nlevels=1, edgeThreshold=31, patchSize=31, fastThreshold=20)
        kps, des = orb.detectAndCompute(img, None)
        rows = sorted(range(len(kps)), key=lambda i: (kps[i].pt[1], kps[i].pt[0]))
        xy, resp, ang, desc = oracle.orb_detect_compute(img, n)
        assert np.array_equal(xy, np.array([[kps[i].pt[0], kps[i].pt[1]] for i in rows], np.int32))
        assert np.array_equal(resp, np.array([kps[i].response for i in rows], np.float32))
        assert np.array_equal(ang, np.array([kps[i].angle for i in rows], np.float32))
        assert np.array_equal(desc, des[rows])
        # the blur ORB applies is the generic float separable filter, not GaussianBlur's 8-bit fixed-point path
        k = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
        assert np.array_equal(oracle.orb_blur(img), cv2.sepFilter2D(img, cv2.CV_8U, k, k, borderType=cv2.BORDER_REFLECT_101))
        # with the scale pyramid (default cv::ORB)
        kps, des = cv2.ORB_create(nfeatures=n).detectAndCompute(img, None)
        rows = sorted(range(len(kps)), key=lambda i: (kps[i].octave, kps[i].pt[1], kps[i].pt[0]))
        xy, octv, resp, ang, desc = oracle.orb_detect_compute_pyr(img, n)
        assert np.array_equal(xy, np.array([kps[i].pt for i in rows], np.float32).reshape(-1, 2))
        assert np.array_equal(octv, [kps[i].octave for i in rows])
        assert np.array_equal(resp, np.array([kps[i].response for i in rows], np.float32))
        assert np.array_equal(ang, np.array([kps[i].angle for i in rows], np.float32))
        assert np.array_equal(desc, des[rows])
