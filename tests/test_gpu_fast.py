"""GPU parity: FAST-9/16 corner detection (vsb_fast_detect — the key-point stage of cv::ORB / cv::cuda::ORB,
Camera.cpp:124-129, CameraGPU.cpp:99-104) through the C ABI vs the oracle (itself pinned to cv2): same corners in the
same row-major order with the same scores — bit-exact — for batches, odd sizes, every threshold regime, with and
without suppression, and with a capacity smaller than the number of corners."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _detect(ctx, imgs, thr, nm, cap):
    import torch
    xy, sc, n = ctx.fast_detect(torch.from_numpy(np.ascontiguousarray(imgs)).cuda(), thr, nm, cap)
    torch.cuda.synchronize()
    return xy.cpu().numpy(), sc.cpu().numpy(), n.cpu().numpy()


@pytest.mark.parametrize("shape", [(480, 752), (376, 1241), (33, 67), (7, 7), (6, 40), (64, 64), (17, 200)])
@pytest.mark.parametrize("thr,nm", [(20, True), (20, False), (0, True), (60, True), (255, True)])
def test_fast_vs_oracle(ctx, oracle, shape, thr, nm):
    rng = np.random.default_rng(shape[0] * 7 + shape[1] + thr)
    B = 3
    imgs = rng.integers(0, 256, (B,) + shape, dtype=np.uint8)
    yy, xx = np.mgrid[0:shape[0], 0:shape[1]]
    imgs[1] = (127 + 120 * np.sin(xx * 0.9) * np.cos(yy * 0.7)).astype(np.uint8)                  # smooth: equal scores, plateaus
    imgs[2] = np.kron(rng.integers(0, 2, ((shape[0] + 7) // 8, (shape[1] + 7) // 8), dtype=np.uint8) * 180 + 30,
                      np.ones((8, 8), np.uint8))[:shape[0], :shape[1]]
    cap = shape[0] * shape[1]
    xy, sc, n = _detect(ctx, imgs, thr, nm, cap)
    for b in range(B):
        rxy, rsc = oracle.fast9(imgs[b], thr, nm)
        assert n[b] == len(rxy), (b, n[b], len(rxy))
        np.testing.assert_array_equal(xy[b, :n[b]], rxy)
        np.testing.assert_array_equal(sc[b, :n[b]], rsc)


def test_fast_capacity_truncates_in_order(ctx, oracle):
    rng = np.random.default_rng(9)
    imgs = rng.integers(0, 256, (2, 120, 160), dtype=np.uint8)
    xy, sc, n = _detect(ctx, imgs, 20, True, 50)
    for b in range(2):
        rxy, rsc = oracle.fast9(imgs[b], 20, True)
        assert n[b] == len(rxy) and n[b] > 50                     # n reports what was FOUND
        np.testing.assert_array_equal(xy[b], rxy[:50])
        np.testing.assert_array_equal(sc[b], rsc[:50])


def test_fast_cv2_golden(ctx):
    g = np.load(os.path.join(HERE, "golden", "fast_cv2.npz"))
    for name in [k[4:] for k in g.files if k.startswith("img_")]:
        img = g["img_" + name]
        for thr in (0, 20, 50):
            for nm in (0, 1):
                xy, sc, n = _detect(ctx, img[None], thr, bool(nm), img.size)
                ref = g[f"xy_{name}_{thr}_{nm}"]
                assert n[0] == len(ref)
                np.testing.assert_array_equal(xy[0, :n[0]], ref)
                np.testing.assert_array_equal(sc[0, :n[0]], g[f"sc_{name}_{thr}_{nm}"])


def test_fast_euroc_shaped_frames(ctx, oracle):
    """Rendered textured frames of the synthetic EuRoC-shaped sequence (what the tracker sees)."""
    from vislam_b200 import synth
    seq = synth.make_sequence(3, n_feat=10, seed=2001)
    xy, sc, n = _detect(ctx, seq["frames"], 20, True, 20000)
    for b in range(3):
        rxy, rsc = oracle.fast9(seq["frames"][b], 20, True)
        assert n[b] == len(rxy) and n[b] > 100
        np.testing.assert_array_equal(xy[b, :n[b]], rxy)
        np.testing.assert_array_equal(sc[b, :n[b]], rsc)
