"""GPU parity: pyramid (Camera::Update), Scharr gradients (Camera::computeGradient) and candidate patch
points (Camera::ObtainPatchesPointsPreviousFrame) through the C ABI vs the oracle — bit-exact."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _levels(arr, lay):
    return [arr[lay.offset[l]: lay.offset[l] + lay.w[l] * lay.h[l]].reshape(lay.h[l], lay.w[l])
            for l in range(lay.levels)]


@pytest.mark.parametrize("w,h", [(752, 480), (640, 480), (1241, 376), (100, 36), (67, 33), (40, 24)])
def test_pyramid_and_gradient(ctx, oracle, w, h):
    import torch
    import vislam_b200 as vb
    rng = np.random.default_rng(w * 1000 + h)
    B = 3
    img = rng.integers(0, 256, (B, h, w), dtype=np.uint8)
    img[1] = (np.add.outer(np.arange(h), np.arange(w)) * 3 % 256).astype(np.uint8)
    lay = vb.pyr_layout(w, h)
    pyr = ctx.pyramid_build(torch.from_numpy(img).cuda(), lay)
    gx, gy, gm = ctx.gradient_build(pyr, lay, want_mag=True)
    torch.cuda.synchronize()
    pyr, gx, gy, gm = pyr.cpu().numpy(), gx.cpu().numpy(), gy.cpu().numpy(), gm.cpu().numpy()
    for b in range(B):
        ref = oracle.pyramid(img[b])
        got = _levels(pyr[b], lay)
        ggx, ggy, ggm = _levels(gx[b], lay), _levels(gy[b], lay), _levels(gm[b], lay)
        for l in range(lay.levels):
            assert got[l].shape == ref[l].shape
            np.testing.assert_array_equal(got[l], ref[l])
            rx, ry = oracle.scharr3(ref[l])
            np.testing.assert_array_equal(ggx[l], rx)
            np.testing.assert_array_equal(ggy[l], ry)
            np.testing.assert_array_equal(ggm[l], oracle.grad_mag(rx, ry))


@pytest.mark.parametrize("w,h", [(752, 480), (640, 480), (64, 48), (16, 16), (2048, 1024), (1241, 376), (1243, 379), (155, 47), (33, 31), (37, 21)])
@pytest.mark.parametrize("levels", [5, 3, 1])
def test_pyramid_register_blocked_kernel(ctx, oracle, w, h, levels):
    """The register-blocked kernels (pyr_impl 1: one thread per 16x16 block; aligned multiples of 16, and the any-size form with
    its edge fix-up for level sizes that cvRound rounds up: 155 -> 78, 47 -> 24) against the shared-memory tile kernel
    (pyr_impl 0) and the oracle, for every level count, with the frame copied in and with level 0 already in place."""
    import torch
    import vislam_b200 as vb
    rng = np.random.default_rng(w + h + levels)
    B = 2
    img = rng.integers(0, 256, (B, h, w), dtype=np.uint8)
    img[1, ::2] = 255                                       # rounding of (a+b+c+d+2)>>2 at the extremes
    img[1, 1::2, ::3] = 254
    lay = vb.pyr_layout(w, h, levels)
    dimg = torch.from_numpy(img).cuda()
    out = {}
    for impl in (0, 1):
        ctx.option("pyr_impl", impl)
        out[impl] = ctx.pyramid_build(dimg, lay)
        # level 0 already in place (the host-buffer tracker entry uploads frames straight into the pyramid)
        inplace = torch.zeros_like(out[impl])
        inplace[:, :w * h] = dimg.reshape(B, -1)
        vb.check(vb.lib().vsb_pyramid_build(ctx.handle, None, w * h, w, B, ctypes.byref(lay), inplace.data_ptr(),
                                            vb._stream_ptr()), ctx.handle)
        torch.cuda.synchronize()
        assert torch.equal(inplace, out[impl])
    ctx.option("pyr_impl", 1)
    assert torch.equal(out[0], out[1])
    got = out[1].cpu().numpy()
    for b in range(B):
        ref = oracle.pyramid(img[b])
        for l, g in enumerate(_levels(got[b], lay)):
            np.testing.assert_array_equal(g, ref[l])


def test_candidates(ctx, oracle):
    import torch
    rng = np.random.default_rng(8)
    w, h = 752, 480
    B, cap = 4, 225
    n_good = np.array([225, 49, 0, 3], np.int32)
    xy = np.zeros((B, cap, 2), np.float32)
    for b in range(B):
        xy[b, :, 0] = rng.uniform(0, w - 1, cap)
        xy[b, :, 1] = rng.uniform(0, h - 1, cap)
    xy[0, :8] = [[0, 0], [751, 479], [0.4, 479], [751, 0.6], [3.5, 3.5], [4.49, 10.51], [15.5, 15.5], [747.9, 475.2]]
    xy[3, :3] = [[1, 1], [750.5, 2.25], [376.0, 240.0]]
    cand, n_cand = ctx.candidates_build(torch.from_numpy(xy).cuda(), torch.from_numpy(n_good).cuda(), w, h)
    torch.cuda.synchronize()
    cand, n_cand = cand.cpu().numpy(), n_cand.cpu().numpy()
    for b in range(B):
        for l in range(5):
            ref = oracle.candidates(xy[b, :n_good[b]], l, w >> l, h >> l)
            assert n_cand[b, l] == ref.shape[0]
            np.testing.assert_array_equal(cand[b, l, :ref.shape[0]], ref)
