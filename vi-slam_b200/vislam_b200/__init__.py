"""ctypes binding of libvislam_b200.so (the C ABI in include/vislam_b200.h).

PyTorch is used only as plumbing here: device buffers (`tensor.data_ptr()`), streams and pinned host
memory.  There is NO CPU fallback: if the shared library is missing, or no CUDA device is visible, the
calls raise.  The reference-facing C++ class mirrors live in vi-slam_b200/host/.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvislam_b200.so")
MAX_LEVELS = 5
MAX_TRACE = 64
MAX_GN_FEATURES = 200


class VsbError(RuntimeError):
    pass


class PyrLayout(C.Structure):
    _fields_ = [("levels", C.c_int), ("w", C.c_int * MAX_LEVELS), ("h", C.c_int * MAX_LEVELS),
                ("offset", C.c_int64 * MAX_LEVELS), ("frame_stride", C.c_int64)]


class Intr(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("invfx", C.c_float), ("invfy", C.c_float), ("w", C.c_int), ("h", C.c_int)]


class GnOpts(C.Structure):
    _fields_ = [("first_lvl", C.c_int), ("last_lvl", C.c_int), ("max_iterations", C.c_int),
                ("epsilon", C.c_float), ("z_factor", C.c_float), ("weight_mode", C.c_int),
                ("sample_mode", C.c_int), ("huber_k", C.c_float), ("grad_mode", C.c_int),
                ("accum_mode", C.c_int)]


class GnTrace(C.Structure):
    _fields_ = [("lvl", C.c_int), ("iter", C.c_int), ("n_valid", C.c_int), ("updated", C.c_int),
                ("error", C.c_float), ("pose", C.c_float * 7), ("delta", C.c_float * 6)]


class TrackerCfg(C.Structure):
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("n_feat_max", C.c_int), ("desc_bytes", C.c_int),
                ("norm", C.c_int), ("n_cells", C.c_int), ("ratio", C.c_float), ("sym_mode", C.c_int),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("gn", GnOpts), ("max_pairs", C.c_int)]


# every symbol declared in include/vislam_b200.h (tests/test_capi_symbols.py checks the header against this)
EXPORTS = [
    "vsb_version", "vsb_error_string", "vsb_ctx_create", "vsb_ctx_destroy", "vsb_last_cuda_error",
    "vsb_sm_count", "vsb_launch_count", "vsb_knn2_hamming", "vsb_knn2_l2", "vsb_match_filter",
    "vsb_pyr_layout", "vsb_pyramid_build", "vsb_gradient_build", "vsb_candidates_build",
    "vsb_gather_keypoints", "vsb_init_pyramid", "vsb_gn_default_opts", "vsb_gn_solve", "vsb_initial_pose",
    "vsb_se3_mul", "vsb_tracker_create", "vsb_tracker_destroy", "vsb_track_sequence",
    "vsb_track_sequence_host", "vsb_track_sequence_orb", "vsb_track_sequence_orb_host", "vsb_track_pairs", "vsb_track_pairs_host", "vsb_kernel_count", "vsb_kernel_name", "vsb_profile_enable",
    "vsb_profile_reset", "vsb_profile_read", "vsb_popc_peak", "vsb_tracker_stats", "vsb_tracker_set_trace", "vsb_tracker_host_traffic", "vsb_fast_detect", "vsb_orb_detect_compute", "vsb_orb_detect_compute_pyr",
    "vsb_malloc", "vsb_free", "vsb_host_alloc", "vsb_host_free", "vsb_upload", "vsb_upload_2d", "vsb_download",
    "vsb_copy", "vsb_memset", "vsb_stream_create", "vsb_stream_destroy", "vsb_stream_sync",
    "vsb_nn_filter", "vsb_sym_matches", "vsb_sort_keys", "vsb_grid_best", "vsb_warp_se3", "vsb_se3_exp",
    "vsb_se3_matrix", "vsb_se3_from_rt", "vsb_ctx_option", "vsb_se3_update_batch",
]

_lib = None


def build():
    """Compile the CUDA extension in-tree (nvcc, sm_100a)."""
    subprocess.check_call(["make", "-s", "-C", os.path.dirname(_HERE)])


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VsbError(f"{LIB_PATH} is missing: build it with `make -C vi-slam_b200` "
                       "(there is no CPU fallback for this path)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
    L.vsb_version.restype = i32
    L.vsb_error_string.restype = C.c_char_p
    L.vsb_error_string.argtypes = [i32]
    L.vsb_ctx_create.argtypes = [i32, C.POINTER(vp)]
    L.vsb_ctx_destroy.argtypes = [vp]
    L.vsb_ctx_option.argtypes = [vp, C.c_char_p, i32]
    L.vsb_last_cuda_error.restype = C.c_char_p
    L.vsb_last_cuda_error.argtypes = [vp]
    L.vsb_sm_count.argtypes = [vp]
    L.vsb_launch_count.restype = C.c_longlong
    L.vsb_launch_count.argtypes = [vp]
    L.vsb_knn2_hamming.argtypes = [vp, vp, i32, vp, vp, i32, vp, i32, vp, vp, vp, vp, vp]
    L.vsb_knn2_l2.argtypes = [vp, vp, i32, vp, vp, i32, vp, i32, i32, vp, vp, vp, vp, vp]
    L.vsb_match_filter.argtypes = [vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, i32, i32, i32, i32, f32, i32,
                                   vp, vp, vp, i32, vp, vp, vp]
    L.vsb_pyr_layout.argtypes = [i32, i32, i32, C.POINTER(PyrLayout)]
    L.vsb_pyramid_build.argtypes = [vp, vp, i64, i32, i32, C.POINTER(PyrLayout), vp, vp]
    L.vsb_gradient_build.argtypes = [vp, vp, i32, C.POINTER(PyrLayout), vp, vp, vp, vp]
    L.vsb_candidates_build.argtypes = [vp, vp, i32, vp, i32, i32, C.POINTER(i32), C.POINTER(i32), vp, i32, vp, vp]
    L.vsb_gather_keypoints.argtypes = [vp, vp, i32, vp, i32, vp, i32, vp, vp]
    L.vsb_init_pyramid.argtypes = [i32, i32, f32, f32, f32, f32, C.POINTER(Intr)]
    L.vsb_gn_default_opts.argtypes = [C.POINTER(GnOpts)]
    L.vsb_gn_default_opts.restype = None
    L.vsb_gn_solve.argtypes = [vp, vp, vp, vp, vp, i64, C.POINTER(PyrLayout), vp, i32, vp, C.POINTER(Intr), vp,
                               C.POINTER(GnOpts), i32, vp, vp, vp, vp]
    L.vsb_initial_pose.argtypes = [C.POINTER(f32), C.POINTER(f32), C.POINTER(f32), C.POINTER(f32)]
    L.vsb_se3_mul.argtypes = [C.POINTER(f32), C.POINTER(f32), C.POINTER(f32)]
    L.vsb_se3_update_batch.argtypes = [vp, vp, vp, i32, vp, vp]
    L.vsb_tracker_create.argtypes = [vp, C.POINTER(TrackerCfg), C.POINTER(vp)]
    L.vsb_tracker_destroy.argtypes = [vp]
    L.vsb_track_sequence.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp, vp, vp]
    L.vsb_track_sequence_orb.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp, vp]
    L.vsb_track_sequence_host.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp, vp]
    L.vsb_track_sequence_orb_host.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp]
    L.vsb_track_pairs.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp]
    L.vsb_track_pairs_host.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp]
    L.vsb_kernel_count.restype = i32
    L.vsb_kernel_name.restype = C.c_char_p
    L.vsb_kernel_name.argtypes = [i32]
    L.vsb_profile_enable.argtypes = [vp, i32]
    L.vsb_profile_reset.argtypes = [vp]
    L.vsb_profile_read.argtypes = [vp, i32, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    L.vsb_popc_peak.argtypes = [vp, C.POINTER(C.c_double), vp]
    L.vsb_tracker_stats.argtypes = [vp, C.POINTER(C.c_longlong)]
    L.vsb_tracker_set_trace.argtypes = [vp, vp, vp]
    L.vsb_tracker_host_traffic.argtypes = [vp, C.POINTER(C.c_longlong)]
    L.vsb_fast_detect.argtypes = [vp, vp, i64, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp]
    L.vsb_orb_detect_compute.argtypes = [vp, vp, i64, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    L.vsb_orb_detect_compute_pyr.argtypes = [vp, vp, i64, i32, i32, i32, i32, i32, C.c_float, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]
    _lib = L
    return L


def check(rc, ctx=None):
    if rc != 0:
        msg = lib().vsb_error_string(rc).decode()
        if ctx is not None and rc == -2:
            msg += ": " + lib().vsb_last_cuda_error(ctx).decode()
        raise VsbError(f"vislam_b200 error {rc}: {msg}")


def default_gn_opts(**kw):
    o = GnOpts()
    lib().vsb_gn_default_opts(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def pyr_layout(w, h, levels=MAX_LEVELS):
    lay = PyrLayout()
    check(lib().vsb_pyr_layout(w, h, levels, C.byref(lay)))
    return lay


def init_pyramid(w, h, fx, fy, cx, cy):
    K = (Intr * MAX_LEVELS)()
    check(lib().vsb_init_pyramid(w, h, fx, fy, cx, cy, K))
    return K


def unpack_traces(raw, nt, count):
    """[count][MAX_TRACE] vsb_gn_trace_t bytes -> per pair a list of dicts."""
    traces = []
    for b in range(count):
        arr = (GnTrace * MAX_TRACE).from_buffer_copy(raw[b].tobytes())
        traces.append([dict(lvl=e.lvl, iter=e.iter, n_valid=e.n_valid, updated=e.updated, error=e.error,
                            pose=list(e.pose), delta=list(e.delta)) for e in arr[:nt[b]]])
    return traces


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


class Context:
    """One context per GPU (device buffers of the calls come from torch on that device)."""

    def __init__(self, device=0):
        import torch
        if not torch.cuda.is_available():
            raise VsbError("no CUDA device visible: the vislam_b200 path has no CPU fallback")
        self.device = device
        self.handle = C.c_void_p()
        check(lib().vsb_ctx_create(device, C.byref(self.handle)))
        self.torch = torch
        self.dev = torch.device("cuda", device)

    def close(self):
        if self.handle:
            lib().vsb_ctx_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return int(lib().vsb_launch_count(self.handle))

    @property
    def sm_count(self):
        return int(lib().vsb_sm_count(self.handle))

    def profile(self, on=True):
        check(lib().vsb_profile_enable(self.handle, 1 if on else 0))
        check(lib().vsb_profile_reset(self.handle), self.handle)

    def profile_read(self):
        """{kernel name: (total device ms, launches)} accumulated since profile()."""
        out = {}
        for k in range(lib().vsb_kernel_count()):
            ms, n = C.c_double(), C.c_longlong()
            check(lib().vsb_profile_read(self.handle, k, C.byref(ms), C.byref(n)), self.handle)
            if n.value:
                out[lib().vsb_kernel_name(k).decode()] = (ms.value, n.value)
        return out

    def popc_peak(self, stream=None):
        v = C.c_double()
        check(lib().vsb_popc_peak(self.handle, C.byref(v), _stream_ptr(stream)), self.handle)
        return v.value

    # ---- Matcher -------------------------------------------------------------------------------
    def option(self, name, value):
        """Tuning knob (vsb_ctx_option): "knn_impl" 0..6 (6 = auto, the default), "knn_l2_impl" 0/1, "gn_threads" 64/128/256, "pyr_impl" 0/1."""
        check(lib().vsb_ctx_option(self.handle, name.encode(), int(value)), self.handle)

    def knn2_hamming(self, d1, d2, n1=None, n2=None, stream=None):
        """d1 [B,N1,32] u8, d2 [B,N2,32] u8 (or 2-D for a single pair). Returns idx12, dist12, idx21, dist21."""
        t = self.torch
        single = d1.dim() == 2
        if single:
            d1, d2 = d1[None], d2[None]
        d1, d2 = d1.contiguous(), d2.contiguous()
        B, N1, N2 = d1.shape[0], d1.shape[1], d2.shape[1]
        idx12 = t.empty((B, N1, 2), dtype=t.int32, device=self.dev)
        dist12 = t.empty((B, N1, 2), dtype=t.float32, device=self.dev)
        idx21 = t.empty((B, N2, 2), dtype=t.int32, device=self.dev)
        dist21 = t.empty((B, N2, 2), dtype=t.float32, device=self.dev)
        check(lib().vsb_knn2_hamming(self.handle, _ptr(d1), N1, _ptr(n1), _ptr(d2), N2, _ptr(n2), B, _ptr(idx12),
                                     _ptr(dist12), _ptr(idx21), _ptr(dist21), _stream_ptr(stream)), self.handle)
        if single:
            return idx12[0], dist12[0], idx21[0], dist21[0]
        return idx12, dist12, idx21, dist21

    def knn2_l2(self, d1, d2, n1=None, n2=None, stream=None):
        t = self.torch
        single = d1.dim() == 2
        if single:
            d1, d2 = d1[None], d2[None]
        d1, d2 = d1.contiguous(), d2.contiguous()
        B, N1, N2, D = d1.shape[0], d1.shape[1], d2.shape[1], d1.shape[2]
        idx12 = t.empty((B, N1, 2), dtype=t.int32, device=self.dev)
        dist12 = t.empty((B, N1, 2), dtype=t.float32, device=self.dev)
        idx21 = t.empty((B, N2, 2), dtype=t.int32, device=self.dev)
        dist21 = t.empty((B, N2, 2), dtype=t.float32, device=self.dev)
        check(lib().vsb_knn2_l2(self.handle, _ptr(d1), N1, _ptr(n1), _ptr(d2), N2, _ptr(n2), D, B, _ptr(idx12),
                                _ptr(dist12), _ptr(idx21), _ptr(dist21), _stream_ptr(stream)), self.handle)
        if single:
            return idx12[0], dist12[0], idx21[0], dist21[0]
        return idx12, dist12, idx21, dist21

    def match_filter(self, idx12, dist12, idx21, dist21, kp1_xy, w, h, n_cells, ratio=0.8, sym_mode=0,
                     n1=None, n2=None, stream=None):
        """Batched [B,...] tensors. Returns good_q, good_t, good_d [B,cap], n_good [B], n_sym [B]."""
        t = self.torch
        B, N1, N2 = idx12.shape[0], idx12.shape[1], idx21.shape[1]
        root = int(n_cells ** 0.5)
        while (root + 1) * (root + 1) <= n_cells:
            root += 1
        while root * root > n_cells:
            root -= 1
        cap = max(root * root, 1)
        gq = t.zeros((B, cap), dtype=t.int32, device=self.dev)
        gt = t.zeros((B, cap), dtype=t.int32, device=self.dev)
        gd = t.zeros((B, cap), dtype=t.float32, device=self.dev)
        ng = t.zeros((B,), dtype=t.int32, device=self.dev)
        ns = t.zeros((B,), dtype=t.int32, device=self.dev)
        check(lib().vsb_match_filter(self.handle, _ptr(idx12.contiguous()), _ptr(dist12.contiguous()), N1, _ptr(n1),
                                     _ptr(idx21.contiguous()), _ptr(dist21.contiguous()), N2, _ptr(n2),
                                     _ptr(kp1_xy.contiguous()), B, w, h, n_cells, ratio, sym_mode, _ptr(gq), _ptr(gt),
                                     _ptr(gd), cap, _ptr(ng), _ptr(ns), _stream_ptr(stream)), self.handle)
        return gq, gt, gd, ng, ns

    # ---- feature detection (first stage) ----------------------------------------------------------
    def fast_detect(self, img, threshold=20, nonmax=True, cap=4096, stream=None):
        """img [B,h,w] u8 (device) -> kp_xy [B,cap,2] i32, score [B,cap] i32, n_found [B] i32 (cv::FAST TYPE_9_16)."""
        t = self.torch
        img = img.contiguous()
        B, h, w = img.shape
        xy = t.zeros((B, cap, 2), dtype=t.int32, device=self.dev)
        sc = t.zeros((B, cap), dtype=t.int32, device=self.dev)
        n = t.zeros((B,), dtype=t.int32, device=self.dev)
        check(lib().vsb_fast_detect(self.handle, _ptr(img), w * h, w, w, h, B, int(threshold), int(bool(nonmax)), cap,
                                    _ptr(xy), _ptr(sc), _ptr(n), _stream_ptr(stream)), self.handle)
        return xy, sc, n

    def orb_detect_compute(self, img, nfeatures=1000, fast_threshold=20, cap=None, describe=True, stream=None):
        """img [B,h,w] u8 (device) -> kp_xy [B,cap,2] i32, response [B,cap] f32, angle_deg [B,cap] f32, desc [B,cap,32] u8 (or
        None), n_kp [B] i32: cv::ORB (one pyramid level) detectAndCompute, key points in row-major order."""
        t = self.torch
        img = img.contiguous()
        B, h, w = img.shape
        cap = cap or 2 * nfeatures
        xy = t.zeros((B, cap, 2), dtype=t.int32, device=self.dev)
        resp = t.zeros((B, cap), dtype=t.float32, device=self.dev)
        ang = t.zeros((B, cap), dtype=t.float32, device=self.dev)
        desc = t.zeros((B, cap, 32), dtype=t.uint8, device=self.dev) if describe else None
        n = t.zeros((B,), dtype=t.int32, device=self.dev)
        check(lib().vsb_orb_detect_compute(self.handle, _ptr(img), w * h, w, w, h, B, int(nfeatures), int(fast_threshold), cap,
                                           _ptr(xy), _ptr(resp), _ptr(ang), _ptr(desc) if describe else None, _ptr(n),
                                           _stream_ptr(stream)), self.handle)
        return xy, resp, ang, desc, n

    def orb_detect_compute_pyr(self, img, nfeatures=1000, scale_factor=1.2, nlevels=8, fast_threshold=20, cap=None,
                               describe=True, stream=None):
        """cv::ORB with its scale pyramid: kp_xy [B,cap,2] f32, octave [B,cap] i32, response, angle_deg, desc, n_kp."""
        t = self.torch
        img = img.contiguous()
        B, h, w = img.shape
        cap = cap or 2 * nfeatures
        xy = t.zeros((B, cap, 2), dtype=t.float32, device=self.dev)
        octv = t.zeros((B, cap), dtype=t.int32, device=self.dev)
        resp = t.zeros((B, cap), dtype=t.float32, device=self.dev)
        ang = t.zeros((B, cap), dtype=t.float32, device=self.dev)
        desc = t.zeros((B, cap, 32), dtype=t.uint8, device=self.dev) if describe else None
        n = t.zeros((B,), dtype=t.int32, device=self.dev)
        check(lib().vsb_orb_detect_compute_pyr(self.handle, _ptr(img), w * h, w, w, h, B, int(nfeatures), float(scale_factor),
                                               int(nlevels), int(fast_threshold), cap, _ptr(xy), _ptr(octv), _ptr(resp), _ptr(ang),
                                               _ptr(desc) if describe else None, _ptr(n), _stream_ptr(stream)), self.handle)
        return xy, octv, resp, ang, desc, n

    # ---- Camera --------------------------------------------------------------------------------
    def pyramid_build(self, img, layout, stream=None):
        """img [B,h,w] u8 -> packed pyramid [B, frame_stride] u8."""
        t = self.torch
        img = img.contiguous()
        B, h, w = img.shape
        pyr = t.zeros((B, layout.frame_stride), dtype=t.uint8, device=self.dev)
        check(lib().vsb_pyramid_build(self.handle, _ptr(img), w * h, w, B, C.byref(layout), _ptr(pyr),
                                      _stream_ptr(stream)), self.handle)
        return pyr

    def gradient_build(self, pyr, layout, want_mag=False, stream=None):
        t = self.torch
        B = pyr.shape[0]
        gx = t.zeros((B, layout.frame_stride), dtype=t.int16, device=self.dev)
        gy = t.zeros((B, layout.frame_stride), dtype=t.int16, device=self.dev)
        gm = t.zeros((B, layout.frame_stride), dtype=t.uint8, device=self.dev) if want_mag else None
        check(lib().vsb_gradient_build(self.handle, _ptr(pyr), B, C.byref(layout), _ptr(gx), _ptr(gy), _ptr(gm),
                                       _stream_ptr(stream)), self.handle)
        return (gx, gy, gm) if want_mag else (gx, gy)

    def gather_keypoints(self, kp_xy, good_idx, n_good, stream=None):
        t = self.torch
        B, N = kp_xy.shape[0], kp_xy.shape[1]
        cap = good_idx.shape[1]
        out = t.zeros((B, cap, 2), dtype=t.float32, device=self.dev)
        check(lib().vsb_gather_keypoints(self.handle, _ptr(kp_xy.contiguous()), N, _ptr(good_idx), cap, _ptr(n_good),
                                         B, _ptr(out), _stream_ptr(stream)), self.handle)
        return out

    def candidates_build(self, good_xy, n_good, w, h, levels=MAX_LEVELS, stream=None):
        """good_xy [B,cap,2] f32, n_good [B] i32 -> cand [B,levels,cand_cap,4] f32, n_cand [B,levels] i32."""
        t = self.torch
        B, cap = good_xy.shape[0], good_xy.shape[1]
        cand_cap = 121 * max(min(cap, MAX_GN_FEATURES), 1)
        cand = t.zeros((B, levels, cand_cap, 4), dtype=t.float32, device=self.dev)
        n_cand = t.zeros((B, levels), dtype=t.int32, device=self.dev)
        lw = (C.c_int * levels)(*[w >> l for l in range(levels)])
        lh = (C.c_int * levels)(*[h >> l for l in range(levels)])
        check(lib().vsb_candidates_build(self.handle, _ptr(good_xy.contiguous()), cap, _ptr(n_good), B, levels, lw, lh,
                                         _ptr(cand), cand_cap, _ptr(n_cand), _stream_ptr(stream)), self.handle)
        return cand, n_cand

    # ---- VISystem ------------------------------------------------------------------------------
    def gn_solve(self, prev_pyr, cur_pyr, prev_gx, prev_gy, layout, cand, n_cand, K, pose_in, opts=None,
                 want_trace=True, pair_stride=None, stream=None):
        """prev_pyr/cur_pyr [B, frame_stride] u8 (pair c at row c), cand [B,levels,cap,4], pose_in [B,7]."""
        t = self.torch
        opts = opts or default_gn_opts()
        B = pose_in.shape[0]
        pose_out = t.zeros((B, 7), dtype=t.float32, device=self.dev)
        trace = t.zeros((B, MAX_TRACE, C.sizeof(GnTrace)), dtype=t.uint8, device=self.dev) if want_trace else None
        n_trace = t.zeros((B,), dtype=t.int32, device=self.dev) if want_trace else None
        stride = layout.frame_stride if pair_stride is None else pair_stride
        check(lib().vsb_gn_solve(self.handle, _ptr(prev_pyr), _ptr(cur_pyr), _ptr(prev_gx), _ptr(prev_gy), stride,
                                 C.byref(layout), _ptr(cand), cand.shape[2], _ptr(n_cand), K, _ptr(pose_in.contiguous()),
                                 C.byref(opts), B, _ptr(pose_out), _ptr(trace), _ptr(n_trace), _stream_ptr(stream)),
              self.handle)
        if not want_trace:
            return pose_out, None
        self.torch.cuda.synchronize(self.dev)
        return pose_out, unpack_traces(trace.cpu().numpy(), n_trace.cpu().numpy(), B)

    def tracker(self, w, h, n_feat_max, K, n_cells=49, max_pairs=256, norm=1, desc_bytes=32, ratio=0.8, sym_mode=0,
                gn_opts=None):
        return Tracker(self, w, h, n_feat_max, K, n_cells, max_pairs, norm, desc_bytes, ratio, sym_mode, gn_opts)


class Tracker:
    """Device-resident scratch for the whole loop of VISystemGPU::AddFrameGPU, batched over frame pairs."""

    def __init__(self, ctx, w, h, n_feat_max, K, n_cells, max_pairs, norm, desc_bytes, ratio, sym_mode, gn_opts):
        self.ctx = ctx
        cfg = TrackerCfg()
        cfg.w, cfg.h, cfg.n_feat_max, cfg.desc_bytes, cfg.norm = w, h, n_feat_max, desc_bytes, norm
        cfg.n_cells, cfg.ratio, cfg.sym_mode = n_cells, ratio, sym_mode
        cfg.fx, cfg.fy, cfg.cx, cfg.cy = K
        # the tracker owns the frames, so by default it evaluates Scharr at the candidate points (bit-identical to
        # reading the materialised gradient images, which cost 1.9 MB of writes per frame)
        cfg.gn = gn_opts or default_gn_opts(grad_mode=1)
        cfg.max_pairs = max_pairs
        self.cfg = cfg
        self.handle = C.c_void_p()
        check(lib().vsb_tracker_create(ctx.handle, C.byref(cfg), C.byref(self.handle)), ctx.handle)

    def close(self):
        if self.handle:
            lib().vsb_tracker_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stats(self):
        out = (C.c_longlong * 4)()
        check(lib().vsb_tracker_stats(self.handle, out), self.ctx.handle)
        return dict(pairs=out[0], iterations=out[1], point_visits=out[2], updates=out[3])

    def trace_on(self):
        """Record the solver's per-iteration trace of the following device-entry calls (read it with traces())."""
        t = self.ctx.torch
        P = self.cfg.max_pairs
        self._trace = t.zeros((P, MAX_TRACE, C.sizeof(GnTrace)), dtype=t.uint8, device=self.ctx.dev)
        self._n_trace = t.zeros((P,), dtype=t.int32, device=self.ctx.dev)
        check(lib().vsb_tracker_set_trace(self.handle, _ptr(self._trace), _ptr(self._n_trace)), self.ctx.handle)

    def trace_off(self):
        check(lib().vsb_tracker_set_trace(self.handle, None, None), self.ctx.handle)
        self._trace = self._n_trace = None

    def traces(self, count):
        self.ctx.torch.cuda.synchronize()
        return unpack_traces(self._trace.cpu().numpy(), self._n_trace.cpu().numpy(), count)

    def host_traffic(self):
        """Bytes moved by the last track_sequence_host call: dict(h2d, d2h, chunks)."""
        out = (C.c_longlong * 3)()
        check(lib().vsb_tracker_host_traffic(self.handle, out), self.ctx.handle)
        return dict(h2d=out[0], d2h=out[1], chunks=out[2])

    def track_sequence(self, frames, desc, kp_xy, prior, n_feat=None, pose=None, n_good=None, stream=None):
        """Device tensors: frames [T,h,w] u8, desc [T,N,D] u8, kp_xy [T,N,2] f32, prior [T-1,7] f32."""
        t = self.ctx.torch
        T = frames.shape[0]
        if pose is None:
            pose = t.zeros((T - 1, 7), dtype=t.float32, device=self.ctx.dev)
        if n_good is None:
            n_good = t.zeros((T - 1,), dtype=t.int32, device=self.ctx.dev)
        check(lib().vsb_track_sequence(self.handle, _ptr(frames), _ptr(desc), _ptr(kp_xy), _ptr(n_feat), _ptr(prior), T,
                                       _ptr(pose), _ptr(n_good), _stream_ptr(stream)), self.ctx.handle)
        return pose, n_good

    def track_sequence_orb(self, frames, prior, nfeatures=1000, stream=None):
        """From images alone: frames [T,h,w] u8 (device), prior [T-1,7] -> pose [T-1,7], n_good [T-1], n_feat [T] (ORB key
        points used per frame); cv::ORB::create(nfeatures) on the device feeds the matcher."""
        t = self.ctx.torch
        T = frames.shape[0]
        pose = t.zeros((T - 1, 7), dtype=t.float32, device=self.ctx.dev)
        n_good = t.zeros((T - 1,), dtype=t.int32, device=self.ctx.dev)
        n_feat = t.zeros((T,), dtype=t.int32, device=self.ctx.dev)
        check(lib().vsb_track_sequence_orb(self.handle, _ptr(frames), _ptr(prior), T, int(nfeatures), _ptr(pose), _ptr(n_good),
                                           _ptr(n_feat), _stream_ptr(stream)), self.ctx.handle)
        return pose, n_good, n_feat

    def track_sequence_host(self, frames, desc, kp_xy, prior, pose, n_good=None, n_feat=None):
        """HOST (pinned) tensors in, host tensors out; synchronous."""
        T = frames.shape[0]
        check(lib().vsb_track_sequence_host(self.handle, _ptr(frames), _ptr(desc), _ptr(kp_xy), _ptr(n_feat),
                                            _ptr(prior), T, _ptr(pose), _ptr(n_good)), self.ctx.handle)
        return pose, n_good

    def track_sequence_orb_host(self, frames, prior, pose, nfeatures=1000, n_good=None, n_feat=None):
        """From images alone with HOST (pinned) tensors: frames [T,h,w] u8, prior [T-1,7] -> pose [T-1,7] (host); synchronous."""
        T = frames.shape[0]
        check(lib().vsb_track_sequence_orb_host(self.handle, _ptr(frames), _ptr(prior), T, int(nfeatures), _ptr(pose),
                                                _ptr(n_good), _ptr(n_feat)), self.ctx.handle)
        return pose, n_good, n_feat

    def track_pairs_host(self, prev, cur, d1, d2, kp1_xy, prior, pose, n_good=None, n1=None, n2=None):
        """HOST (pinned) tensors in, host tensors out; synchronous.  Independent pairs, chunked over two streams."""
        check(lib().vsb_track_pairs_host(self.handle, _ptr(prev), _ptr(cur), _ptr(d1), _ptr(d2), _ptr(kp1_xy), _ptr(n1),
                                         _ptr(n2), _ptr(prior), prev.shape[0], _ptr(pose), _ptr(n_good)), self.ctx.handle)
        return pose, n_good

    def track_pairs(self, prev, cur, d1, d2, kp1_xy, prior, n1=None, n2=None, pose=None, n_good=None, stream=None):
        t = self.ctx.torch
        B = prev.shape[0]
        if pose is None:
            pose = t.zeros((B, 7), dtype=t.float32, device=self.ctx.dev)
        if n_good is None:
            n_good = t.zeros((B,), dtype=t.int32, device=self.ctx.dev)
        check(lib().vsb_track_pairs(self.handle, _ptr(prev), _ptr(cur), _ptr(d1), _ptr(d2), _ptr(kp1_xy), _ptr(n1),
                                    _ptr(n2), _ptr(prior), B, _ptr(pose), _ptr(n_good), _stream_ptr(stream)),
              self.ctx.handle)
        return pose, n_good
