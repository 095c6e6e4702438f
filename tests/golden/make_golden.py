#!/usr/bin/env python
"""Generates the committed golden fixtures in tests/golden/ (run in the BUILD container only).

Sources of truth, in decreasing authority:
  * matcher_ref.npz — outputs of the reference's OWN src/Matcher.cpp, compiled unmodified against
    oracle/cvshim (`make -C oracle ref`, needs /root/reference): nSymMatches, matches, sortedMatches,
    goodMatches for a handful of inputs, including one where the de-facto (stale-read)
    and the intended symmetry test differ.
  * knn_cv2.npz / camera_cv2.npz / inv6_cv2.npz — Python cv2 (4.13 here; the reference pins 3.2 by
    prose): BFMatcher.knnMatch k=2 (Hamming + L2), resize(0.5) chains, Scharr(scale 3), addWeighted,
    invert(DECOMP_LU).  These pin the third-party primitives the oracle restates.
  * gn_oracle.npz — the oracle's own GN trace on a small synthetic pair (regression pin: the GN solver has no
    executable reference, see oracle/vso.h).
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vi-slam_b200"))

import cv2  # noqa: E402
from oracle import vso  # noqa: E402
from vislam_b200 import synth  # noqa: E402


def knn_cv2():
    rng = np.random.default_rng(7)
    d1 = rng.integers(0, 256, (120, 32), dtype=np.uint8)
    d2 = rng.integers(0, 256, (90, 32), dtype=np.uint8)
    d2[10] = d2[40]; d2[41] = d2[40]; d1[5] = d2[40]          # exact ties -> lowest train index first
    d1[60:] &= 0xC0; d2[50:] &= 0xC0                          # many equal distances
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    out = {"d1": d1, "d2": d2}
    for name, (a, b) in {"12": (d1, d2), "21": (d2, d1)}.items():
        m = bf.knnMatch(a, b, 2)
        out["idx" + name] = np.array([[x.trainIdx for x in r] for r in m], np.int32)
        out["dist" + name] = np.array([[x.distance for x in r] for r in m], np.float32)
    f1 = synth.float_descriptors(80, 11)
    f2, _ = synth.perturb_float(f1, 12)
    f2 = f2[:70]
    bf = cv2.BFMatcher(cv2.NORM_L2)
    out["f1"], out["f2"] = f1, f2
    for name, (a, b) in {"12": (f1, f2), "21": (f2, f1)}.items():
        m = bf.knnMatch(a, b, 2)
        out["fidx" + name] = np.array([[x.trainIdx for x in r] for r in m], np.int32)
        out["fdist" + name] = np.array([[x.distance for x in r] for r in m], np.float32)
    np.savez_compressed(os.path.join(HERE, "knn_cv2.npz"), **out)


def camera_cv2():
    rng = np.random.default_rng(8)
    out = {}
    for tag, (h, w) in {"even": (48, 64), "odd": (47, 155), "kitti": (94, 311)}.items():
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        out[tag + "_l0"] = img
        cur = img
        for l in range(1, 5):
            if min(cur.shape) < 2:
                break
            cur = cv2.resize(cur, None, fx=0.5, fy=0.5)
            out[f"{tag}_l{l}"] = cur
        gx = cv2.Scharr(img, cv2.CV_16S, 1, 0, scale=3)
        gy = cv2.Scharr(img, cv2.CV_16S, 0, 1, scale=3)
        out[tag + "_gx"], out[tag + "_gy"] = gx, gy
        out[tag + "_gm"] = cv2.addWeighted(cv2.convertScaleAbs(gx), 0.5, cv2.convertScaleAbs(gy), 0.5, 0)
    np.savez_compressed(os.path.join(HERE, "camera_cv2.npz"), **out)


def inv6_cv2():
    rng = np.random.default_rng(9)
    A, I = [], []
    for i in range(24):
        J = rng.standard_normal((60, 6)).astype(np.float32) * np.array([4e6, 4e6, 3e6, 5e11, 6e11, 1e9], np.float32)
        a = (J.T.astype(np.float64) @ J.astype(np.float64)).astype(np.float32)
        if i == 23:
            a[:, 3] = 0; a[3, :] = 0                                  # singular -> zeros
        ok, inv = cv2.invert(a, flags=cv2.DECOMP_LU)
        A.append(a); I.append(inv)
    np.savez_compressed(os.path.join(HERE, "inv6_cv2.npz"), A=np.stack(A), Ainv=np.stack(I))


def matcher_ref():
    lib_path = os.path.join(ROOT, "oracle", "_ref", "libref_matcher.so")
    if not os.path.exists(lib_path):
        vso.build("ref")
    L = C.CDLL(lib_path)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rng = np.random.default_rng(10)
    cases = []
    specs = [(300, 280, 49, 0), (257, 301, 225, 1), (64, 64, 1, 0), (150, 40, 16, 0), (12, 12, 49, 2)]
    for n1, n2, nc, flavour in specs:
        d1 = synth.orb_descriptors(max(n1, n2), 100 + n1)
        d2, _ = synth.perturb_orb(d1, 200 + n2)
        d1, d2 = d1[:n1].copy(), d2[:n2].copy()
        kp1 = np.stack([rng.uniform(0, 751, n1), rng.uniform(0, 479, n1)], 1).astype(np.float32)
        kp2 = np.stack([rng.uniform(0, 751, n2), rng.uniform(0, 479, n2)], 1).astype(np.float32)
        if flavour == 1:
            kp1[:, 1] = np.floor(kp1[:, 1] / 16) * 16                  # equal-y ties
        if flavour == 2:
            # de-facto vs intended: row 0 of d1 matches row 0 of d2 cleanly (1->2 passes the ratio test), but
            # d1 row 1 is almost as close to d2 row 0 (2->1 FAILS the ratio test).  The reference still emits it.
            d2[0] = d1[0]
            d2[0, 0] ^= 0x1F                                           # distance 5 to d1[0]
            d1[1] = d2[0]
            d1[1, 1] ^= 0x3F                                           # distance 6 to d2[0]; far from d1[0]'s other candidates
        cases.append((d1, d2, kp1, kp2, nc))
    out = {"n_cases": np.array(len(cases))}
    for i, (d1, d2, kp1, kp2, nc) in enumerate(cases):
        n1, n2 = len(d1), len(d2)
        cap = max(n1, 1)
        gq = np.zeros(cap, np.int32); gt = np.zeros(cap, np.int32); gd = np.zeros(cap, np.float32)
        sq = np.zeros(cap, np.int32); st = np.zeros(cap, np.int32); so = np.zeros(cap, np.int32)
        pxy = np.zeros((cap, 2), np.float32); cxy = np.zeros((cap, 2), np.float32)
        ns = C.c_int(0)
        n = L.ref_matcher_run(p(d1), n1, p(d2), n2, 32, 1, p(kp1), p(kp2), 752, 480, nc, p(gq), p(gt), p(gd),
                              p(sq), p(st), C.byref(ns), p(so), p(pxy), p(cxy))
        out.update({f"c{i}_d1": d1, f"c{i}_d2": d2, f"c{i}_kp1": kp1, f"c{i}_kp2": kp2, f"c{i}_ncells": np.array(nc),
                    f"c{i}_good_q": gq[:n], f"c{i}_good_t": gt[:n], f"c{i}_good_d": gd[:n],
                    f"c{i}_sym_q": sq[:ns.value], f"c{i}_sym_t": st[:ns.value], f"c{i}_sorted_q": so[:ns.value],
                    f"c{i}_prev_xy": pxy[:n], f"c{i}_cur_xy": cxy[:n]})
    np.savez_compressed(os.path.join(HERE, "matcher_ref.npz"), **out)


def gn_oracle():
    """Small pair (188x120 frames, intrinsics of EuRoC level 2) so the fixture stays tiny."""
    w, h = 188, 120
    K = (114.6635, 114.324, 91.42875, 61.71875)
    p = synth.make_pair(w=w, h=h, n_feat=200, K=K, seed=4242)
    out = {"prev": p["prev"], "cur": p["cur"], "d1": p["d1"], "d2": p["d2"], "kp1": p["kp1"],
           "K": np.array(K, np.float64), "prior": p["pose_prior"]}
    for tag, kw in {"ref": {}, "huber": dict(weight_mode=2, huber_k=12.0), "bilinear": dict(sample_mode=1)}.items():
        r = vso.track_pair(p["prev"], p["cur"], p["d1"], p["d2"], p["kp1"], K, p["pose_prior"], n_cells=49,
                           opts=vso.default_opts(first_lvl=2, **kw))
        out[tag + "_pose"] = r["pose"]
        out[tag + "_good_q"] = r["good_q"]
        out[tag + "_trace_pose"] = np.stack([t["pose"] for t in r["trace"]])
        out[tag + "_trace_err"] = np.array([t["error"] for t in r["trace"]], np.float32)
        out[tag + "_trace_meta"] = np.array([[t["lvl"], t["iter"], t["n_valid"], t["updated"]] for t in r["trace"]], np.int32)
    np.savez_compressed(os.path.join(HERE, "gn_oracle.npz"), **out)


if __name__ == "__main__":
    knn_cv2(); camera_cv2(); inv6_cv2(); matcher_ref(); gn_oracle()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
