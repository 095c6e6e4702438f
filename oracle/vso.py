"""ctypes loader for the CPU oracle (oracle/libvso.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package never does (see oracle/vso.h).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MAX_LEVELS = 5
RATIO = float(np.float32(0.8))  # Matcher.cpp:103  (0.8f widened to double)

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


class Intr(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("invfx", C.c_float), ("invfy", C.c_float), ("w", C.c_int), ("h", C.c_int)]


class GnOpts(C.Structure):
    _fields_ = [("first_lvl", C.c_int), ("last_lvl", C.c_int), ("max_iterations", C.c_int),
                ("epsilon", C.c_float), ("z_factor", C.c_float), ("weight_mode", C.c_int),
                ("sample_mode", C.c_int), ("huber_k", C.c_float)]


class GnTrace(C.Structure):
    _fields_ = [("lvl", C.c_int), ("iter", C.c_int), ("n_valid", C.c_int), ("updated", C.c_int),
                ("error", C.c_float), ("pose", C.c_float * 7), ("delta", C.c_float * 6)]


class GnFrames(C.Structure):
    _fields_ = [("prev_img", C.c_void_p * MAX_LEVELS), ("cur_img", C.c_void_p * MAX_LEVELS),
                ("prev_gx", C.c_void_p * MAX_LEVELS), ("prev_gy", C.c_void_p * MAX_LEVELS),
                ("cand", C.c_void_p * MAX_LEVELS), ("n_cand", C.c_int * MAX_LEVELS),
                ("img_w", C.c_int * MAX_LEVELS), ("img_h", C.c_int * MAX_LEVELS)]


def default_opts(**kw):
    """VISystem.cpp:1115-1121 literals."""
    o = GnOpts(3, 0, 10, 0.001, 0.002, 0, 0, 10.0)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def build(target="all"):
    subprocess.check_call(["make", "-s", "-C", _HERE, target])


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(_HERE, "libvso.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    L.vso_knn2_hamming.argtypes = [_u8p, C.c_int, _u8p, C.c_int, C.c_int, _i32p, _f32p]
    L.vso_knn2_l2.argtypes = [_f32p, C.c_int, _f32p, C.c_int, C.c_int, _i32p, _f32p]
    L.vso_nn_filter.argtypes = [_i32p, _f32p, C.c_int, C.c_double, _u8p]
    L.vso_sym_matches.argtypes = [_i32p, _f32p, C.c_int, _i32p, _f32p, C.c_int, C.c_double, C.c_int,
                                  _i32p, _i32p, _f32p]
    L.vso_sym_matches.restype = C.c_int
    L.vso_sort_matches.argtypes = [_i32p, C.c_int, _f32p, _i32p]
    L.vso_grid_filter.argtypes = [_i32p, _i32p, _f32p, _i32p, C.c_int, _f32p, C.c_int, C.c_int, C.c_int,
                                  _i32p, _i32p, _f32p]
    L.vso_grid_filter.restype = C.c_int
    L.vso_match_pipeline.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, _f32p,
                                     C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, _i32p, _i32p, _f32p,
                                     C.POINTER(C.c_int)]
    L.vso_match_pipeline.restype = C.c_int
    L.vso_pyr_size.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.vso_pyr_down.argtypes = [_u8p, C.c_int, C.c_int, _u8p]
    L.vso_scharr3.argtypes = [_u8p, C.c_int, C.c_int, _i16p, _i16p]
    L.vso_grad_mag.argtypes = [_i16p, _i16p, C.c_int, _u8p]
    L.vso_candidates.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p]
    L.vso_candidates.restype = C.c_int
    L.vso_init_pyramid.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                   C.POINTER(Intr)]
    L.vso_se3_exp.argtypes = [_f32p, _f32p]
    L.vso_se3_mul.argtypes = [_f32p, _f32p, _f32p]
    L.vso_se3_matrix.argtypes = [_f32p, _f32p]
    L.vso_rpy_to_rot.argtypes = [_f64p, _f32p]
    L.vso_rot_to_rpy.argtypes = [_f32p, _f64p]
    L.vso_rot_to_quat.argtypes = [_f32p, _f32p]
    L.vso_initial_pose.argtypes = [_f32p, _f32p, _f32p, _f32p]
    L.vso_warp.argtypes = [_f32p, C.c_int, _f32p, C.POINTER(Intr), _f32p]
    L.vso_inv6.argtypes = [_f32p, _f32p]
    L.vso_inv6.restype = C.c_int
    L.vso_solve6.argtypes = [_f32p, _f32p, _f32p]
    L.vso_solve6.restype = C.c_int
    L.vso_tukey_weights.argtypes = [_f32p, C.c_int, _f32p]
    L.vso_gn_solve.argtypes = [C.POINTER(GnFrames), C.POINTER(Intr), _f32p, C.POINTER(GnOpts), _f32p,
                               C.POINTER(GnTrace), C.c_int]
    L.vso_gn_solve.restype = C.c_int
    L.vso_fast9.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _i32p, C.c_int]
    L.vso_fast9.restype = C.c_int
    L.vso_orb_harris.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _i32p, C.c_int, _f32p]
    L.vso_orb_ic_angle.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _i32p, C.c_int, _f32p]
    L.vso_orb_gauss_kernel.argtypes = [_f32p]
    L.vso_orb_blur.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p]
    L.vso_orb_describe.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _i32p, _f32p, C.c_int, _u8p]
    L.vso_orb_detect_compute.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _f32p, _f32p, _u8p, C.c_int]
    L.vso_orb_detect_compute.restype = C.c_int
    L.vso_resize_linear_exact.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p, C.c_int, C.c_int]
    L.vso_orb_level_scale.argtypes = [C.c_float, C.c_int]
    L.vso_orb_level_scale.restype = C.c_float
    L.vso_orb_level_budget.argtypes = [C.c_int, C.c_float, C.c_int, _i32p]
    L.vso_orb_detect_compute_pyr.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, _f32p, _i32p,
                                             _f32p, _f32p, _u8p, C.c_int]
    L.vso_orb_detect_compute_pyr.restype = C.c_int
    _lib = L
    return L


# ----------------------------------------------------------------------------- numpy-level helpers
def knn2_hamming(q, t):
    q = np.ascontiguousarray(q, np.uint8)
    t = np.ascontiguousarray(t, np.uint8)
    nbytes = q.shape[1] if q.ndim == 2 and q.shape[0] else (t.shape[1] if t.ndim == 2 and t.shape[0] else 32)
    idx = np.full((q.shape[0], 2), -1, np.int32)
    dist = np.zeros((q.shape[0], 2), np.float32)
    lib().vso_knn2_hamming(q.reshape(-1), q.shape[0], t.reshape(-1), t.shape[0], nbytes, idx.reshape(-1),
                           dist.reshape(-1))
    return idx, dist


def knn2_l2(q, t):
    q = np.ascontiguousarray(q, np.float32)
    t = np.ascontiguousarray(t, np.float32)
    idx = np.full((q.shape[0], 2), -1, np.int32)
    dist = np.zeros((q.shape[0], 2), np.float32)
    lib().vso_knn2_l2(q.reshape(-1), q.shape[0], t.reshape(-1), t.shape[0], q.shape[1], idx.reshape(-1),
                      dist.reshape(-1))
    return idx, dist


def sym_matches(idx1, dist1, idx2, dist2, ratio=RATIO, mode=0):
    n1, n2 = idx1.shape[0], idx2.shape[0]
    mq = np.zeros(max(n1, 1), np.int32)
    mt = np.zeros(max(n1, 1), np.int32)
    md = np.zeros(max(n1, 1), np.float32)
    n = lib().vso_sym_matches(np.ascontiguousarray(idx1, np.int32).reshape(-1),
                              np.ascontiguousarray(dist1, np.float32).reshape(-1), n1,
                              np.ascontiguousarray(idx2, np.int32).reshape(-1),
                              np.ascontiguousarray(dist2, np.float32).reshape(-1), n2,
                              ratio, mode, mq, mt, md)
    return mq[:n].copy(), mt[:n].copy(), md[:n].copy()


def sort_matches(mq, kp1_xy):
    order = np.zeros(max(len(mq), 1), np.int32)
    lib().vso_sort_matches(np.ascontiguousarray(mq, np.int32), len(mq),
                           np.ascontiguousarray(kp1_xy, np.float32).reshape(-1), order)
    return order[:len(mq)].copy()


def grid_filter(mq, mt, md, order, kp1_xy, w, h, n_cells):
    n = len(mq)
    cap = max(n, 1)
    gq = np.zeros(cap, np.int32)
    gt = np.zeros(cap, np.int32)
    gd = np.zeros(cap, np.float32)
    pad = lambda a, dt: np.ascontiguousarray(a if n else np.zeros(1), dt)
    k = lib().vso_grid_filter(pad(mq, np.int32), pad(mt, np.int32), pad(md, np.float32), pad(order, np.int32), n,
                              np.ascontiguousarray(kp1_xy, np.float32).reshape(-1), w, h, n_cells, gq, gt, gd)
    return gq[:k].copy(), gt[:k].copy(), gd[:k].copy()


def match_pipeline(d1, d2, kp1_xy, w, h, n_cells, norm, ratio=RATIO, mode=0):
    """Camera::computeGoodMatches. norm: 1 Hamming (uint8 rows), 0 L2 (float32 rows)."""
    dt = np.uint8 if norm == 1 else np.float32
    d1 = np.ascontiguousarray(d1, dt)
    d2 = np.ascontiguousarray(d2, dt)
    n1, n2 = d1.shape[0], d2.shape[0]
    dim = d1.shape[1]
    cap = max(n1, 1)
    gq = np.zeros(cap, np.int32)
    gt = np.zeros(cap, np.int32)
    gd = np.zeros(cap, np.float32)
    nsym = C.c_int(0)
    k = lib().vso_match_pipeline(d1.ctypes.data, n1, d2.ctypes.data, n2, dim, norm,
                                 np.ascontiguousarray(kp1_xy, np.float32).reshape(-1), w, h, n_cells,
                                 ratio, mode, gq, gt, gd, C.byref(nsym))
    return gq[:k].copy(), gt[:k].copy(), gd[:k].copy(), nsym.value


def pyr_down(img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    dw, dh = C.c_int(), C.c_int()
    lib().vso_pyr_size(w, h, C.byref(dw), C.byref(dh))
    out = np.zeros((dh.value, dw.value), np.uint8)
    lib().vso_pyr_down(img.reshape(-1), w, h, out.reshape(-1))
    return out


def pyramid(img, levels=MAX_LEVELS):
    """Camera::Update (Camera.cpp:63-72): level 0 = copy, then 4x resize(0.5)."""
    out = [np.ascontiguousarray(img, np.uint8)]
    for _ in range(1, levels):
        out.append(pyr_down(out[-1]))
    return out


def scharr3(img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    gx = np.zeros((h, w), np.int16)
    gy = np.zeros((h, w), np.int16)
    lib().vso_scharr3(img.reshape(-1), w, h, gx.reshape(-1), gy.reshape(-1))
    return gx, gy


def grad_mag(gx, gy):
    g = np.zeros(gx.shape, np.uint8)
    lib().vso_grad_mag(np.ascontiguousarray(gx).reshape(-1), np.ascontiguousarray(gy).reshape(-1), gx.size,
                       g.reshape(-1))
    return g


def fast9(img, threshold=20, nonmax=True, cap=None):
    """cv::FAST(TYPE_9_16): returns (xy [n,2] int32, score [n] int32) in row-major order."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    cap = w * h if cap is None else cap
    xy = np.zeros((max(cap, 1), 2), np.int32)
    sc = np.zeros((max(cap, 1),), np.int32)
    n = lib().vso_fast9(img, w, h, w, int(threshold), int(bool(nonmax)), xy, sc, cap)
    n = min(n, cap)
    return xy[:n].copy(), sc[:n].copy()


def orb_harris(img, xy):
    img = np.ascontiguousarray(img, np.uint8)
    xy = np.ascontiguousarray(xy, np.int32).reshape(-1, 2)
    out = np.zeros(xy.shape[0], np.float32)
    lib().vso_orb_harris(img, img.shape[1], img.shape[0], img.shape[1], xy, xy.shape[0], out)
    return out


def orb_ic_angle(img, xy):
    img = np.ascontiguousarray(img, np.uint8)
    xy = np.ascontiguousarray(xy, np.int32).reshape(-1, 2)
    out = np.zeros(xy.shape[0], np.float32)
    lib().vso_orb_ic_angle(img, img.shape[1], img.shape[0], img.shape[1], xy, xy.shape[0], out)
    return out


def orb_gauss_kernel():
    k = np.zeros(7, np.float32)
    lib().vso_orb_gauss_kernel(k)
    return k


def orb_blur(img):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros_like(img)
    lib().vso_orb_blur(img, img.shape[1], img.shape[0], img.shape[1], out)
    return out


def orb_describe(blurred, xy, angle_deg):
    blurred = np.ascontiguousarray(blurred, np.uint8)
    xy = np.ascontiguousarray(xy, np.int32).reshape(-1, 2)
    ang = np.ascontiguousarray(angle_deg, np.float32)
    out = np.zeros((xy.shape[0], 32), np.uint8)
    lib().vso_orb_describe(blurred, blurred.shape[1], blurred.shape[0], blurred.shape[1], xy, ang, xy.shape[0], out)
    return out


def orb_detect_compute(img, nfeatures=500, fast_threshold=20, cap=None):
    """cv::ORB (nlevels = 1) detectAndCompute: (xy [n,2] int32, response, angle_deg, desc [n,32]) in row-major order."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    cap = cap or (w * h // 4 + 16)
    xy = np.zeros((cap, 2), np.int32)
    resp = np.zeros(cap, np.float32)
    ang = np.zeros(cap, np.float32)
    desc = np.zeros((cap, 32), np.uint8)
    n = lib().vso_orb_detect_compute(img, w, h, w, int(nfeatures), int(fast_threshold), xy, resp, ang, desc, cap)
    n = min(n, cap)
    return xy[:n].copy(), resp[:n].copy(), ang[:n].copy(), desc[:n].copy()


def resize_linear_exact(img, dw, dh):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros((dh, dw), np.uint8)
    lib().vso_resize_linear_exact(img, img.shape[1], img.shape[0], img.shape[1], out, dw, dh)
    return out


def orb_level_budget(nfeatures, scale_factor=1.2, nlevels=8):
    out = np.zeros(nlevels, np.int32)
    lib().vso_orb_level_budget(int(nfeatures), float(scale_factor), int(nlevels), out)
    return out


def orb_detect_compute_pyr(img, nfeatures=500, scale_factor=1.2, nlevels=8, fast_threshold=20, cap=None):
    """cv::ORB detectAndCompute with its scale pyramid: (xy [n,2] f32, octave [n] i32, response, angle_deg, desc [n,32]),
    level by level, row-major inside a level."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    cap = cap or (2 * nfeatures + 4096)
    xy = np.zeros((cap, 2), np.float32)
    octv = np.zeros(cap, np.int32)
    resp = np.zeros(cap, np.float32)
    ang = np.zeros(cap, np.float32)
    desc = np.zeros((cap, 32), np.uint8)
    n = lib().vso_orb_detect_compute_pyr(img, w, h, w, int(nfeatures), float(scale_factor), int(nlevels), int(fast_threshold),
                                         xy, octv, resp, ang, desc, cap)
    n = min(n, cap)
    return xy[:n].copy(), octv[:n].copy(), resp[:n].copy(), ang[:n].copy(), desc[:n].copy()


def candidates(good_xy, lvl, lw, lh):
    good_xy = np.ascontiguousarray(good_xy, np.float32).reshape(-1, 2)
    nf = good_xy.shape[0]
    out = np.zeros((121 * max(min(nf, 200), 1), 4), np.float32)
    n = lib().vso_candidates(good_xy.reshape(-1) if nf else np.zeros(2, np.float32), nf, lvl, lw, lh,
                             out.reshape(-1))
    return out[:n].copy()


def init_pyramid(w, h, fx, fy, cx, cy):
    K = (Intr * MAX_LEVELS)()
    lib().vso_init_pyramid(w, h, fx, fy, cx, cy, K)
    return K


def se3_exp(delta):
    out = np.zeros(7, np.float32)
    lib().vso_se3_exp(np.ascontiguousarray(delta, np.float32), out)
    return out


def se3_mul(a, b):
    out = np.zeros(7, np.float32)
    lib().vso_se3_mul(np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32), out)
    return out


def se3_matrix(p):
    out = np.zeros(16, np.float32)
    lib().vso_se3_matrix(np.ascontiguousarray(p, np.float32), out)
    return out.reshape(4, 4)


def initial_pose(imu2cam, r_imu_res, t_res):
    out = np.zeros(7, np.float32)
    lib().vso_initial_pose(np.ascontiguousarray(imu2cam, np.float32).reshape(-1),
                           np.ascontiguousarray(r_imu_res, np.float32).reshape(-1),
                           np.ascontiguousarray(t_res, np.float32), out)
    return out


def warp(pts, pose, K_lvl):
    pts = np.ascontiguousarray(pts, np.float32)
    out = np.zeros_like(pts)
    lib().vso_warp(pts.reshape(-1), pts.shape[0], np.ascontiguousarray(pose, np.float32), C.byref(K_lvl),
                   out.reshape(-1))
    return out


def inv6(a):
    out = np.zeros(36, np.float32)
    ok = lib().vso_inv6(np.ascontiguousarray(a, np.float32).reshape(-1), out)
    return ok, out.reshape(6, 6)


def solve6(a, b):
    """cv::solve(A, b, DECOMP_LU) — what `A.inv() * b` evaluates to (VISystem.cpp:1412)."""
    out = np.zeros(6, np.float32)
    ok = lib().vso_solve6(np.ascontiguousarray(a, np.float32).reshape(-1), np.ascontiguousarray(b, np.float32).reshape(-1), out)
    return ok, out


def tukey_weights(r):
    r = np.ascontiguousarray(r, np.float32).reshape(-1)
    out = np.zeros_like(r)
    lib().vso_tukey_weights(r, r.size, out)
    return out


def gn_solve(prev_pyr, cur_pyr, prev_gx, prev_gy, cands, K, pose_in, opts=None):
    """VISystem::EstimatePoseFeatures.  prev_pyr/cur_pyr/prev_gx/prev_gy/cands: lists over levels
    (None allowed for unused levels).  Returns (pose_out[7], list of trace dicts)."""
    opts = opts or default_opts()
    fr = GnFrames()
    keep = []
    for l in range(MAX_LEVELS):
        if l >= len(prev_pyr) or prev_pyr[l] is None:
            continue
        a = np.ascontiguousarray(prev_pyr[l], np.uint8)
        b = np.ascontiguousarray(cur_pyr[l], np.uint8)
        gx = np.ascontiguousarray(prev_gx[l], np.int16)
        gy = np.ascontiguousarray(prev_gy[l], np.int16)
        c = np.ascontiguousarray(cands[l], np.float32).reshape(-1, 4)
        keep += [a, b, gx, gy, c]
        fr.prev_img[l] = a.ctypes.data
        fr.cur_img[l] = b.ctypes.data
        fr.prev_gx[l] = gx.ctypes.data
        fr.prev_gy[l] = gy.ctypes.data
        fr.cand[l] = c.ctypes.data
        fr.n_cand[l] = c.shape[0]
        fr.img_w[l] = a.shape[1]
        fr.img_h[l] = a.shape[0]
    cap = (opts.first_lvl - opts.last_lvl + 1) * opts.max_iterations
    tr = (GnTrace * cap)()
    pose_out = np.zeros(7, np.float32)
    n = lib().vso_gn_solve(C.byref(fr), K, np.ascontiguousarray(pose_in, np.float32), C.byref(opts), pose_out,
                           tr, cap)
    trace = [dict(lvl=t.lvl, iter=t.iter, n_valid=t.n_valid, updated=t.updated, error=t.error,
                  pose=np.array(t.pose[:], np.float32), delta=np.array(t.delta[:], np.float32))
             for t in tr[:n]]
    return pose_out, trace


def track_pair(prev_img, cur_img, d1, d2, kp1_xy, K0, pose_in, n_cells=49, norm=1, mode=0, opts=None,
               prev_pyr=None, cur_pyr=None, prev_grad=None):
    """One frame pair through the intended loop (VISystemGPU.cpp:144-169): pyramid, match, gradient of the
    PREVIOUS frame, candidates, GN.  K0 = (fx, fy, cx, cy).  Returns dict."""
    h, w = prev_img.shape
    prev_pyr = prev_pyr or pyramid(prev_img)
    cur_pyr = cur_pyr or pyramid(cur_img)
    if prev_grad is None:
        prev_grad = [scharr3(p) for p in prev_pyr]
    gq, gt, gd, nsym = match_pipeline(d1, d2, kp1_xy, w, h, n_cells, norm, mode=mode)
    good_xy = np.ascontiguousarray(kp1_xy, np.float32).reshape(-1, 2)[gq]
    cands = [candidates(good_xy, l, w >> l, h >> l) for l in range(MAX_LEVELS)]
    K = init_pyramid(w, h, *K0)
    pose, trace = gn_solve(prev_pyr, cur_pyr, [g[0] for g in prev_grad], [g[1] for g in prev_grad], cands, K,
                           pose_in, opts)
    return dict(good_q=gq, good_t=gt, good_d=gd, n_sym=nsym, cands=cands, pose=pose, trace=trace)
