"""The pose update of a Gauss-Newton iteration ON THE DEVICE (csrc/se3.cuh as the kernels compile it; entry
vsb_se3_update_batch): pose * exp(delta) against the reference's own Sophus (golden vectors generated from
thirdparty/sophus, tests/golden/se3_ref.npz), against the host build of the same header, and — because the device evaluates
sin / cos of small angles with its own polynomial — a sweep over the floats of (0, 0.5] against the float rounding of the
double functions."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "se3_ref.npz")


def _update(ctx, pose, delta):
    import torch
    import vislam_b200 as vb
    p = torch.from_numpy(np.ascontiguousarray(pose, np.float32)).cuda()
    d = torch.from_numpy(np.ascontiguousarray(delta, np.float32)).cuda()
    out = torch.empty_like(p)
    rc = vb.lib().vsb_se3_update_batch(ctx.handle, p.data_ptr(), d.data_ptr(), p.shape[0], out.data_ptr(), None)
    assert rc == 0
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_device_update_matches_the_references_sophus(ctx):
    """exp(a) * exp(b) of the golden file (computed by the reference's vendored Sophus, both Taylor / renormalisation branches
    covered — tests/test_se3_sophus.py::test_golden_covers_the_branches): the device gives the same bits."""
    g = np.load(GOLDEN)
    out = _update(ctx, g["out_exp_a"], g["delta_b"])
    bad = np.nonzero((_bits(out) != _bits(g["out_prod"])).any(axis=1))[0]
    assert bad.size == 0, f"{bad.size} of {out.shape[0]} products differ from Sophus, first {bad[:5]}"


def test_device_update_matches_host_build(ctx):
    """The same header compiled for the host (vsb_se3_exp / vsb_se3_mul, libm's double sin / cos) on random updates with
    rotation angles from 1e-7 to 3 rad."""
    import vislam_b200 as vb
    L = C.CDLL(vb.LIB_PATH)
    rng = np.random.default_rng(77)
    n = 60000
    delta = (rng.standard_normal((n, 6)) * 10.0 ** rng.uniform(-7, 0.5, (n, 1))).astype(np.float32)
    base = (rng.standard_normal((n, 6)) * 0.3).astype(np.float32)
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    pose = np.zeros((n, 7), np.float32)
    want = np.zeros((n, 7), np.float32)
    e = np.zeros(7, np.float32)
    for i in range(n):
        L.vsb_se3_exp(fp(base[i]), fp(pose[i]))
        L.vsb_se3_exp(fp(delta[i]), fp(e))
        L.vsb_se3_mul(fp(pose[i]), fp(e), fp(want[i]))
    out = _update(ctx, pose, delta)
    bad = np.nonzero((_bits(out) != _bits(want)).any(axis=1))[0]
    assert bad.size == 0, f"{bad.size} of {n} updates differ from the host build, first {bad[:5]}: {delta[bad[:2]]}"


@pytest.mark.parametrize("phase", [0, 1])
def test_device_small_angle_trig_sweep(ctx, phase):
    """Every 64th float (two phases) of the half angle in [2^-16, 0.5] and of the full angle in [2^-15, 0.5] (above the 1e-5 Taylor switch), 2 x 2 M
    arguments: a rotation about x applied to the identity.  q.w is cos(theta / 2) and q.x is sin(theta / 2) / theta * theta
    rounded as Sophus does; the translation of a unit step along y carries sin(theta) and cos(theta).  Expected values: the
    float rounding of numpy's (glibc's) double functions, the definition the oracle uses."""
    f32 = np.float32
    lo, hi = np.array([2.0 ** -16, 0.5], f32).view(np.uint32)
    half = np.arange(int(lo) + 7 + 31 * phase, int(hi) + 1, 64, dtype=np.uint32).view(f32)
    for theta in (half * f32(2), half):
        n = theta.size
        delta = np.zeros((n, 6), f32)
        delta[:, 1] = 1.0
        delta[:, 3] = theta
        ident = np.zeros((n, 7), f32)
        ident[:, 3] = 1.0
        out = _update(ctx, ident, delta)
        h = f32(0.5) * theta
        s_h = np.sin(h.astype(np.float64)).astype(f32)
        c_h = np.cos(h.astype(np.float64)).astype(f32)
        qx = (s_h / theta).astype(f32) * theta
        qw = c_h
        sn = ((qx * qx + f32(0)) + f32(0)) + qw * qw                     # SO3::operator*= renormalises when != 1
        sc = np.where(sn != f32(1), f32(2) / (f32(1) + sn), f32(1)).astype(f32)
        np.testing.assert_array_equal(_bits(out[:, 0]), _bits(qx * sc))
        np.testing.assert_array_equal(_bits(out[:, 3]), _bits(qw * sc))
        # V = I + c1 Om + c2 Om^2 applied to (0, 1, 0); Om = [[0,0,0],[0,0,-t],[0,t,0]] => Om^2[1][1] = -(t t), Om[2][1] = t
        th2 = theta * theta
        c1 = ((f32(1) - np.cos(theta.astype(np.float64)).astype(f32)) / th2).astype(f32)
        c2 = ((theta - np.sin(theta.astype(np.float64)).astype(f32)) / (th2 * theta).astype(f32)).astype(f32)
        om2 = -(theta * theta)
        ty = (f32(1) + c1 * f32(0)) + c2 * om2
        tz = (f32(0) + c1 * theta) + c2 * f32(0)
        # the identity pose rotates by the unit quaternion and adds the zero translation: exact
        np.testing.assert_array_equal(_bits(out[:, 5]), _bits(ty.astype(f32)))
        np.testing.assert_array_equal(_bits(out[:, 6]), _bits(tz.astype(f32)))
