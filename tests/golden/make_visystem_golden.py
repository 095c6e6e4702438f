#!/usr/bin/env python
"""Generates tests/golden/visystem_ref.npz and solve6_cv2.npz (run in the BUILD container only; needs /root/reference).

visystem_ref.npz — outputs of the reference's OWN src/VISystem.cpp + src/Camera.cpp (+ Matcher/Plus/Imu), compiled
unmodified against oracle/refshim (`make -C oracle ref` -> oracle/_ref/libref_visystem.so): for a few small synthetic
frame pairs, the pyramid / Scharr gradients / candidate points built by Camera::Update, computeGradient and
ObtainPatchesPointsPreviousFrame, and the pose + per-iteration error printed by VISystem::EstimatePoseFeatures; plus
WarpFunctionSE3 and TukeyFunctionWeights on their own.  tests/test_ref_visystem.py holds the oracle to these bit for bit.

solve6_cv2.npz — cv2.solve(A, b, DECOMP_LU) on 6x6 float systems (what `A.inv() * b` evaluates to in OpenCV).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vi-slam_b200"))

import cv2  # noqa: E402
from oracle import ref_visystem as rv, vso  # noqa: E402
from vislam_b200 import synth  # noqa: E402

CASES = {
    # tag: (w, h, K, first seed to try, n_feat, n_cells, use the reference's own matcher)
    "a": (188, 120, (114.6635, 114.324, 91.42875, 61.71875), 4242, 200, 49, False),
    "b": (188, 120, (114.6635, 114.324, 91.42875, 61.71875), 60, 300, 49, True),
    "c": (155, 94, (89.857, 89.857, 75.9, 46.3), 911, 250, 225, True),      # odd sizes (KITTI-like level): clipped 2x2 blocks
}
# The reference reads current-frame pixels at (round(y2), round(x2)) after testing only y2 < rows (VISystem.cpp:1299, 1321):
# y2 in [rows - 0.5, rows) reads past the image — undefined behaviour upstream, which the stand-in counts
# (`oob_reads`).  About 40 % of random pairs of this size hit it; parity is only defined for runs that do not, so each
# case takes the first seed at or after the listed one with oob_reads == 0.


def prior_inputs(seed, identity_extrinsics=False):
    rng = np.random.default_rng(seed)
    rpy = rng.uniform(-0.01, 0.01, 3)
    rres = np.zeros(9, np.float32)
    vso.lib().vso_rpy_to_rot(np.ascontiguousarray(rpy, np.float64), rres)
    tres = rng.uniform(-0.01, 0.01, 3).astype(np.float32)
    i2c = np.zeros(9, np.float32)
    vso.lib().vso_rpy_to_rot(np.ascontiguousarray(rng.uniform(-0.5, 0.5, 3), np.float64), i2c)
    if identity_extrinsics:
        i2c = np.eye(3, dtype=np.float32).reshape(-1)
    return i2c.reshape(3, 3), rres.reshape(3, 3), tres


def visystem_ref():
    assert rv.available(), "oracle/_ref/libref_visystem.so cannot be built here"
    vso.build()
    out = {}
    for tag, (w, h, K, seed0, nf, n_cells, own_matcher) in CASES.items():
        for seed in range(seed0, seed0 + 50):
            p = synth.make_pair(w=w, h=h, n_feat=nf, K=K, seed=seed)
            i2c, rres, tres = prior_inputs(seed, identity_extrinsics=(tag == "a"))
            if own_matcher:
                kw = dict(kp_prev=p["kp1"], desc_prev=p["d1"], kp_cur=p["kp2"], desc_cur=p["d2"])
            else:
                gq, gt, _, _ = vso.match_pipeline(p["d1"], p["d2"], p["kp1"], w, h, n_cells, 1)
                kw = dict(good_prev=p["kp1"].reshape(-1, 2)[gq], good_cur=p["kp2"].reshape(-1, 2)[gt])
            r = rv.track_pair(p["prev"], p["cur"], K, i2c, rres, tres, n_cells=n_cells, **kw)
            if r["oob_reads"] == 0:
                break
        assert r["oob_reads"] == 0
        print("case", tag, "seed", seed, "iterations", len(r["trace"]), "good matches", len(r["good_prev"]))
        out.update({f"{tag}_prev": p["prev"], f"{tag}_cur": p["cur"], f"{tag}_K": np.array(K, np.float64),
                    f"{tag}_seed": np.int32(seed), f"{tag}_n_cells": np.int32(n_cells), f"{tag}_own_matcher": np.int32(own_matcher),
                    f"{tag}_kp1": p["kp1"], f"{tag}_kp2": p["kp2"], f"{tag}_d1": p["d1"], f"{tag}_d2": p["d2"],
                    f"{tag}_imu2cam": i2c, f"{tag}_r_imu_res": rres, f"{tag}_t_res": tres,
                    f"{tag}_good_prev": r["good_prev"], f"{tag}_good_cur": r["good_cur"],
                    f"{tag}_pose": r["pose"], f"{tag}_trace": r["trace"], f"{tag}_n_cand": r["n_cand"]})
        for l in range(5):
            out[f"{tag}_pyr{l}"] = r["pyr_prev"][l]
            if l >= 1:                                                        # level 0 gradients are 4x the rest; keep the file small
                out[f"{tag}_gx{l}"] = r["gx"][l]
                out[f"{tag}_gy{l}"] = r["gy"][l]
            assert np.array_equal(r["cands"][l][:, 2:], np.ones((len(r["cands"][l]), 2), np.float32))
            out[f"{tag}_cand{l}"] = r["cands"][l][:, :2].astype(np.int16)     # rows are (i, j, 1, 1) with integer i, j
    # WarpFunctionSE3 / TukeyFunctionWeights on their own
    rng = np.random.default_rng(5)
    K4 = (458.654, 457.296, 367.215, 248.375)
    pts = np.ones((400, 4), np.float32)
    pts[:, 0] = rng.integers(1, 94, 400)
    pts[:, 1] = rng.integers(1, 60, 400)
    pts[200:, 2] = rng.uniform(0.5, 3.0, 200).astype(np.float32)
    pose = vso.se3_exp(rng.uniform(-0.03, 0.03, 6).astype(np.float32))
    out.update({"warp_pts": pts, "warp_pose": pose, "warp_K": np.array(K4, np.float64), "warp_lvl": np.int32(3),
                "warp_out": rv.warp(pts, pose, 752, 480, K4, 3)})
    res = np.concatenate([rng.normal(0, 30, 600), np.round(rng.normal(0, 300, 100)), rng.integers(-3, 4, 100) + 0.5]).astype(np.float32)
    out.update({"tukey_r": res, "tukey_w": rv.tukey(res)})
    np.savez_compressed(os.path.join(HERE, "visystem_ref.npz"), **out)


def solve6_cv2():
    rng = np.random.default_rng(606)
    A, B, X = [], [], []
    for i in range(48):
        J = rng.normal(0, 10.0 ** rng.uniform(-1, 3), (200, 6)).astype(np.float32)
        J[:, 2] *= 0.002
        a = (J.astype(np.float64).T @ J.astype(np.float64)).astype(np.float32)
        if i % 6 == 5:
            a = rng.normal(0, 1, (6, 6)).astype(np.float32)           # non-symmetric: exercises the row exchanges
        b = rng.normal(0, 100, (6, 1)).astype(np.float32)
        ok, x = cv2.solve(a, b, flags=cv2.DECOMP_LU)
        A.append(a); B.append(b[:, 0]); X.append(x[:, 0] if ok else np.zeros(6, np.float32))
    a = np.ones((6, 6), np.float32)                                   # singular -> zeros
    b = np.arange(6, dtype=np.float32).reshape(6, 1)
    ok, x = cv2.solve(a, b, flags=cv2.DECOMP_LU)
    assert not ok
    A.append(a); B.append(b[:, 0]); X.append(np.zeros(6, np.float32))
    np.savez_compressed(os.path.join(HERE, "solve6_cv2.npz"), A=np.stack(A), b=np.stack(B), x=np.stack(X))


if __name__ == "__main__":
    visystem_ref(); solve6_cv2()
    for f in ("visystem_ref.npz", "solve6_cv2.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))
