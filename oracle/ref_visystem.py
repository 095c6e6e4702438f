"""ctypes loader for oracle/_ref/libref_visystem.so (TEST INFRASTRUCTURE ONLY): the reference's own VISystem.cpp /
Camera.cpp / Matcher.cpp / Plus.cpp / Imu.cpp compiled unmodified against oracle/refshim (see its opencv2/core.hpp header
for what the stand-in supplies).  Built by `make -C oracle ref`, which needs /root/reference; the .so is git-ignored and
travels to the GPU box with the snapshot.  Used by tests/ and tests/golden/make_visystem_golden.py only."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "libref_visystem.so")
_lib = None


def available(build=True):
    if os.path.exists(PATH):
        return True
    if build and os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])
    return os.path.exists(PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(PATH)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def track_pair(prev, cur, K4, imu2cam, r_imu_res, t_res, good_prev=None, good_cur=None, kp_prev=None, desc_prev=None,
               kp_cur=None, desc_cur=None, n_cells=49):
    """One frame pair through the reference's Camera front end and VISystem::EstimatePoseFeatures.  Either the good
    matches (key-point coordinates, prev/cur) are given, or key points + ORB descriptors for Camera::computeGoodMatches."""
    prev = np.ascontiguousarray(prev, np.uint8)
    cur = np.ascontiguousarray(cur, np.uint8)
    h, w = prev.shape
    K4 = np.ascontiguousarray(K4, np.float32)
    lw = np.zeros(5, np.int32)
    lh = np.zeros(5, np.int32)
    tot = 2 * w * h + 64
    pyr_p = np.zeros(tot, np.uint8)
    pyr_c = np.zeros(tot, np.uint8)
    gx = np.zeros(tot, np.int16)
    gy = np.zeros(tot, np.int16)
    cap = 121 * 200 * 5
    cand = np.zeros((cap, 4), np.float32)
    nc = np.zeros(5, np.int32)
    pose = np.zeros(7, np.float32)
    trace = np.zeros((64, 3), np.float32)
    nt = C.c_int(0)
    ngo = C.c_int(0)
    gop = np.zeros((4096, 2), np.float32)
    goc = np.zeros((4096, 2), np.float32)
    err = C.create_string_buffer(256)
    f32 = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
    u8 = lambda a: None if a is None else np.ascontiguousarray(a, np.uint8)
    gp, gc, kp1, kp2, d1, d2 = f32(good_prev), f32(good_cur), f32(kp_prev), f32(kp_cur), u8(desc_prev), u8(desc_cur)
    n_good = -1 if gp is None else gp.reshape(-1, 2).shape[0]
    i2c, rr, tr = f32(imu2cam), f32(r_imu_res), f32(t_res)
    rc = lib().ref_track_pair(_p(prev), _p(cur), w, h, _p(K4), n_cells, _p(gp), _p(gc), n_good,
                              _p(kp1), _p(d1), 0 if kp1 is None else kp1.reshape(-1, 2).shape[0],
                              _p(kp2), _p(d2), 0 if kp2 is None else kp2.reshape(-1, 2).shape[0],
                              _p(i2c), _p(rr), _p(tr), _p(pyr_p), _p(pyr_c), _p(gx), _p(gy), _p(lw), _p(lh),
                              _p(cand), cap, _p(nc), _p(gop), _p(goc), C.byref(ngo),
                              _p(pose), _p(trace), 64, C.byref(nt), err, 256)
    if rc < 0:
        raise RuntimeError(err.value.decode())
    out = dict(oob_reads=rc, pose=pose, trace=trace[:nt.value].copy(), n_cand=nc, good_prev=gop[:ngo.value].copy(),
               good_cur=goc[:ngo.value].copy(), pyr_prev=[], pyr_cur=[], gx=[], gy=[], cands=[])
    off = c0 = 0
    for l in range(5):
        n = int(lw[l]) * int(lh[l])
        shp = (int(lh[l]), int(lw[l]))
        out["pyr_prev"].append(pyr_p[off:off + n].reshape(shp).copy())
        out["pyr_cur"].append(pyr_c[off:off + n].reshape(shp).copy())
        out["gx"].append(gx[off:off + n].reshape(shp).copy())
        out["gy"].append(gy[off:off + n].reshape(shp).copy())
        out["cands"].append(cand[c0:c0 + nc[l]].copy())
        off += n
        c0 += int(nc[l])
    return out


def warp(pts, pose, w, h, K4, lvl):
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 4)
    out = np.zeros_like(pts)
    rc = lib().ref_warp(_p(pts), pts.shape[0], _p(np.ascontiguousarray(pose, np.float32)), w, h,
                        _p(np.ascontiguousarray(K4, np.float32)), lvl, _p(out))
    if rc:
        raise RuntimeError("ref_warp failed")
    return out


def tukey(r):
    r = np.ascontiguousarray(r, np.float32).reshape(-1)
    out = np.zeros_like(r)
    if lib().ref_tukey(_p(r), r.size, _p(out)):
        raise RuntimeError("ref_tukey failed")
    return out
