#!/usr/bin/env python
"""Generates tests/golden/fast_cv2.npz: cv2.FastFeatureDetector (TYPE_9_16) on deterministic images — the pin of
oracle/fast.c (the algorithm lives in OpenCV, which the reference does not vendor).  Needs cv2; run from the repo root."""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def images():
    rng = np.random.default_rng(4242)
    out = {"noise_64x80": rng.integers(0, 256, (64, 80), dtype=np.uint8),
           "blur_120x160": cv2.GaussianBlur(rng.integers(0, 256, (120, 160), dtype=np.uint8), (5, 5), 1.2),
           "blocks_96x128": np.kron(rng.integers(0, 2, (12, 16), dtype=np.uint8) * 180 + 30, np.ones((8, 8), np.uint8)),
           "tiny_7x9": rng.integers(0, 256, (7, 9), dtype=np.uint8)}
    yy, xx = np.mgrid[0:90, 0:110]
    out["waves_90x110"] = (127 + 120 * np.sin(xx * 0.9) * np.cos(yy * 0.7)).astype(np.uint8)
    return out


if __name__ == "__main__":
    res = {}
    for name, img in images().items():
        res["img_" + name] = img
        for thr in (0, 20, 50):
            for nm in (0, 1):
                det = cv2.FastFeatureDetector_create(threshold=thr, nonmaxSuppression=bool(nm),
                                                     type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
                kps = det.detect(img)
                res[f"xy_{name}_{thr}_{nm}"] = np.array([[int(k.pt[0]), int(k.pt[1])] for k in kps], np.int32).reshape(-1, 2)
                res[f"sc_{name}_{thr}_{nm}"] = np.array([int(k.response) for k in kps], np.int32)
    np.savez_compressed(os.path.join(HERE, "fast_cv2.npz"), **res)
    print("wrote fast_cv2.npz:", sum(v.shape[0] for k, v in res.items() if k.startswith("xy_")), "corners")
