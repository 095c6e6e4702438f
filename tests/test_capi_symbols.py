"""CPU: the C-ABI library loads, exports every symbol include/vislam_b200.h declares, its host-side helpers agree
with the oracle, and — with no GPU — compute entries fail loudly instead of falling back to the CPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def vb():
    import vislam_b200
    if not os.path.exists(vislam_b200.LIB_PATH):
        vislam_b200.build()
    vislam_b200.lib()
    return vislam_b200


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "vislam_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(vsb_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_are_exported(vb):
    out = subprocess.check_output(["nm", "-D", "--defined-only", vb.LIB_PATH], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    decl = declared_symbols()
    assert len(decl) >= 30
    missing = [s for s in decl if s not in exported]
    assert not missing, missing
    assert sorted(vb.EXPORTS) == decl          # the python binding list tracks the header
    for s in decl:
        assert hasattr(vb.lib(), s)


def test_library_does_not_link_the_oracle(vb):
    """The product must not route through oracle/: no vso_* symbol, no libvso dependency."""
    out = subprocess.check_output(["nm", "-D", vb.LIB_PATH], text=True)
    assert "vso_" not in out
    assert "libvso" not in subprocess.check_output(["ldd", vb.LIB_PATH], text=True)
    for root, _, files in os.walk(os.path.join(ROOT, "vi-slam_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                assert "libvso" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_no_gpu_means_error_not_fallback(vb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert vb.lib().vsb_ctx_create(0, C.byref(h)) == -2          # VSB_ERR_CUDA
    with pytest.raises(vb.VsbError):
        vb.Context(0)


def test_version_and_errors(vb):
    assert vb.lib().vsb_version() >= 100
    assert vb.lib().vsb_error_string(0) == b"ok"
    assert b"CUDA" in vb.lib().vsb_error_string(-2)
    names = [vb.lib().vsb_kernel_name(i).decode() for i in range(vb.lib().vsb_kernel_count())]
    assert "knn2_hamming" in names and "gn_solve" in names


@pytest.mark.parametrize("w,h", [(752, 480), (640, 480), (1241, 376), (1240, 376)])
def test_pyr_layout_matches_cv_resize_sizes(vb, oracle, w, h):
    lay = vb.pyr_layout(w, h)
    img = np.zeros((h, w), np.uint8)
    for l, lv in enumerate(oracle.pyramid(img)):
        assert (lay.h[l], lay.w[l]) == lv.shape
    assert lay.offset[0] == 0 and all(lay.offset[l] % 256 == 0 for l in range(5)) and lay.frame_stride % 256 == 0
    assert vb.lib().vsb_pyr_layout(17, 9, 5, C.byref(vb.PyrLayout())) == -1       # level 4 would be empty


def test_init_pyramid_matches_oracle(vb, oracle):
    for K0 in [(458.654, 457.296, 367.215, 248.375), (525.0, 525.0, 319.5, 239.5), (718.856, 718.856, 607.1928, 185.2157)]:
        a = vb.init_pyramid(752, 480, *K0)
        b = oracle.init_pyramid(752, 480, *K0)
        for l in range(5):
            for f in ("fx", "fy", "cx", "cy", "invfx", "invfy", "w", "h"):
                assert getattr(a[l], f) == getattr(b[l], f), (l, f)


def test_host_se3_helpers_match_oracle(vb, oracle):
    rng = np.random.default_rng(0)
    f7 = C.c_float * 7
    for _ in range(50):
        a = oracle.se3_exp(rng.uniform(-0.3, 0.3, 6).astype(np.float32))
        b = oracle.se3_exp(rng.uniform(-0.3, 0.3, 6).astype(np.float32))
        out = f7()
        assert vb.lib().vsb_se3_mul(f7(*a), f7(*b), out) == 0
        np.testing.assert_array_equal(np.array(out[:], np.float32), oracle.se3_mul(a, b))
        from vislam_b200 import synth
        R = synth.so3_exp(rng.uniform(-0.05, 0.05, 3)).astype(np.float32)
        M = synth.so3_exp(rng.uniform(-3, 3, 3)).astype(np.float32)       # imu2cam extrinsic rotation
        t = rng.uniform(-0.1, 0.1, 3).astype(np.float32)
        assert vb.lib().vsb_initial_pose((C.c_float * 9)(*M.reshape(-1)), (C.c_float * 9)(*R.reshape(-1)),
                                         (C.c_float * 3)(*t), out) == 0
        np.testing.assert_array_equal(np.array(out[:], np.float32), oracle.initial_pose(M, R, t))


def test_default_opts_are_the_reference_literals(vb):
    o = vb.default_gn_opts()
    assert (o.first_lvl, o.last_lvl, o.max_iterations) == (3, 0, 10)              # VISystem.cpp:1117-1120
    assert o.epsilon == np.float32(0.001) and o.z_factor == np.float32(0.002)     # :1115, :1121
    assert o.weight_mode == 0 and o.sample_mode == 0                              # identity weights, round()


def test_every_context_option_is_documented_in_the_header():
    """vsb_ctx_option's names (csrc/capi.cu) all appear in the header's description of the tuning knobs."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "vi-slam_b200", "csrc", "capi.cu")).read()
    hdr = open(os.path.join(root, "include", "vislam_b200.h")).read()
    names = sorted(set(re.findall(r'!strcmp\(name, "([a-z0-9_]+)"\)', src)))
    assert len(names) >= 10
    missing = [n for n in names if '"' + n + '"' not in hdr]
    assert not missing, missing
