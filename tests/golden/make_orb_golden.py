#!/usr/bin/env python
"""Generates tests/golden/orb_cv2.npz (run in the BUILD container; needs cv2): cv2.ORB_create(nfeatures, nlevels=1,
edgeThreshold=31, patchSize=31, fastThreshold=20).detectAndCompute on small images — key points (x, y), Harris response,
angle and 32-byte descriptors, sorted row-major (cv2's own order is an artefact of std::nth_element) — plus the blur kernel."""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def images():
    rng = np.random.default_rng(20261018)
    yield "noise", cv2.GaussianBlur((rng.random((200, 264)) * 255).astype(np.uint8), (5, 5), 1.2)
    yield "odd", cv2.GaussianBlur((rng.random((157, 211)) * 255).astype(np.uint8), (3, 3), 0.8)
    im = np.full((180, 240), 30, np.uint8)
    for _ in range(40):
        x, y = int(rng.integers(10, 220)), int(rng.integers(10, 160))
        cv2.rectangle(im, (x, y), (x + int(rng.integers(6, 40)), y + int(rng.integers(6, 40))), int(rng.integers(60, 255)), -1)
    yield "rects", cv2.GaussianBlur(im, (3, 3), 0.6)


def main():
    out = {"gauss_kernel": cv2.getGaussianKernel(7, 2, cv2.CV_32F).ravel()}
    for name, img in images():
        out[f"{name}_img"] = img
        for n in (60, 400, 5000):
            orb = cv2.ORB_create(nfeatures=n, nlevels=1, edgeThreshold=31, patchSize=31, fastThreshold=20)
            kps, des = orb.detectAndCompute(img, None)
            des = des if des is not None else np.zeros((0, 32), np.uint8)
            rows = sorted(range(len(kps)), key=lambda i: (kps[i].pt[1], kps[i].pt[0]))
            out[f"{name}_{n}_xy"] = np.array([[kps[i].pt[0], kps[i].pt[1]] for i in rows], np.int32).reshape(-1, 2)
            out[f"{name}_{n}_resp"] = np.array([kps[i].response for i in rows], np.float32)
            out[f"{name}_{n}_angle"] = np.array([kps[i].angle for i in rows], np.float32)
            out[f"{name}_{n}_desc"] = des[rows] if len(rows) else des
            print(name, n, len(kps))
    # the full detector with its scale pyramid (the reference's ORB::create(n): 8 levels, factor 1.2), and two other settings
    for name, img in images():
        for tag, (n, sf, nl) in {"d": (300, 1.2, 8), "e": (150, 1.5, 4)}.items():
            kps, des = cv2.ORB_create(nfeatures=n, scaleFactor=sf, nlevels=nl).detectAndCompute(img, None)
            des = des if des is not None else np.zeros((0, 32), np.uint8)
            rows = sorted(range(len(kps)), key=lambda i: (kps[i].octave, kps[i].pt[1], kps[i].pt[0]))
            out[f"{name}_pyr{tag}_cfg"] = np.array([n, sf, nl], np.float64)
            out[f"{name}_pyr{tag}_xy"] = np.array([kps[i].pt for i in rows], np.float32).reshape(-1, 2)
            out[f"{name}_pyr{tag}_octave"] = np.array([kps[i].octave for i in rows], np.int32)
            out[f"{name}_pyr{tag}_resp"] = np.array([kps[i].response for i in rows], np.float32)
            out[f"{name}_pyr{tag}_angle"] = np.array([kps[i].angle for i in rows], np.float32)
            out[f"{name}_pyr{tag}_size"] = np.array([kps[i].size for i in rows], np.float32)
            out[f"{name}_pyr{tag}_desc"] = des[rows] if len(rows) else des
            print(name, "pyramid", tag, len(kps))
    rimg = next(images())[1]
    for i, (dw, dh) in enumerate(((220, 167), (132, 100), (263, 199), (97, 61))):
        out[f"resize{i}"] = cv2.resize(rimg, (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT)
    np.savez_compressed(os.path.join(HERE, "orb_cv2.npz"), **out)
    print(os.path.getsize(os.path.join(HERE, "orb_cv2.npz")))


if __name__ == "__main__":
    main()
