"""GPU parity: vsb_orb_detect_compute (cv::ORB, one pyramid level) against the oracle and the cv2 golden fixture.
Everything goes through the C ABI (vislam_b200.Context); bit-exact: key-point set and order, responses, angles, descriptors."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def run(ctx, imgs, n, cap=None, describe=True):
    import torch
    d = torch.from_numpy(np.ascontiguousarray(imgs)).cuda()
    xy, resp, ang, desc, nk = ctx.orb_detect_compute(d, nfeatures=n, cap=cap, describe=describe)
    torch.cuda.synchronize()
    out = []
    for b in range(imgs.shape[0]):
        k = int(nk[b])
        m = min(k, xy.shape[1])
        out.append((k, xy[b, :m].cpu().numpy(), resp[b, :m].cpu().numpy(), ang[b, :m].cpu().numpy(),
                    desc[b, :m].cpu().numpy() if describe else None))
    return out


@pytest.mark.parametrize("name", ["noise", "odd", "rects"])
@pytest.mark.parametrize("n", [60, 400, 5000])
def test_orb_matches_cv2_golden(ctx, name, n):
    g = np.load(os.path.join(GOLD, "orb_cv2.npz"))
    img = g[f"{name}_img"]
    (k, xy, resp, ang, desc), = run(ctx, img[None], n, cap=6000)
    assert k == len(g[f"{name}_{n}_xy"])
    assert np.array_equal(xy, g[f"{name}_{n}_xy"])
    assert np.array_equal(resp, g[f"{name}_{n}_resp"])
    assert np.array_equal(ang, g[f"{name}_{n}_angle"])
    assert np.array_equal(desc, g[f"{name}_{n}_desc"])


def test_orb_batch_vs_oracle_euroc_shaped(ctx, oracle):
    """A batch of 752x480 frames (textured synthetic scene + noise frames), 1000 features each."""
    from vislam_b200 import synth
    rng = np.random.default_rng(5)
    frames = [synth.make_pair(w=752, h=480, n_feat=50, seed=s)["prev"] for s in (11, 12)]
    for s in (1.0, 2.0):
        f = (rng.random((480, 752)) * 255).astype(np.float32)
        k = int(4 * s) | 1
        # separable box smoothing keeps plenty of corners without cv2 on the GPU box
        f = np.apply_along_axis(lambda r: np.convolve(r, np.ones(k) / k, mode="same"), 1, f)
        f = np.apply_along_axis(lambda r: np.convolve(r, np.ones(k) / k, mode="same"), 0, f)
        frames.append(np.clip(f, 0, 255).astype(np.uint8))
    imgs = np.stack(frames)
    res = run(ctx, imgs, 1000, cap=4000)
    for b, (k, xy, resp, ang, desc) in enumerate(res):
        oxy, oresp, oang, odesc = oracle.orb_detect_compute(imgs[b], 1000)
        assert k == len(oxy), (b, k, len(oxy))
        assert np.array_equal(xy, oxy[:len(xy)])
        assert np.array_equal(resp, oresp[:len(xy)])
        assert np.array_equal(ang, oang[:len(xy)])
        assert np.array_equal(desc, odesc[:len(xy)])
    assert max(r[0] for r in res) >= 1000


@pytest.mark.parametrize("w,h", [(640, 480), (1241, 376), (333, 201), (100, 80)])
def test_orb_sizes_and_caps(ctx, oracle, w, h):
    rng = np.random.default_rng(w * 7 + h)
    f = (rng.random((h, w)) * 255).astype(np.float32)
    f = (f + np.roll(f, 1, 0) + np.roll(f, 1, 1) + np.roll(f, (1, 1), (0, 1))) / 4
    img = f.astype(np.uint8)
    for n, cap in ((500, 2000), (50, 16), (0, 8)):
        (k, xy, resp, ang, desc), = run(ctx, img[None], n, cap=cap)
        oxy, oresp, oang, odesc = oracle.orb_detect_compute(img, n)
        assert k == len(oxy)
        m = min(k, cap)
        assert np.array_equal(xy, oxy[:m]) and np.array_equal(resp, oresp[:m])
        assert np.array_equal(ang, oang[:m]) and np.array_equal(desc, odesc[:m])


def test_orb_detect_only_and_invalid(ctx):
    import torch
    import vislam_b200 as vb
    img = torch.zeros((1, 100, 100), dtype=torch.uint8, device="cuda")
    xy, resp, ang, desc, n = ctx.orb_detect_compute(img, nfeatures=10, describe=False)
    assert desc is None and int(n[0]) == 0
    small = torch.zeros((1, 60, 200), dtype=torch.uint8, device="cuda")
    with pytest.raises(vb.VsbError):
        ctx.orb_detect_compute(small, nfeatures=10)


def test_orb_descriptors_feed_the_matcher(ctx, oracle):
    """Frame -> ORB -> kNN: the descriptors of a frame and of a shifted copy match each other through the Hamming matcher."""
    import torch
    rng = np.random.default_rng(3)
    f = (rng.random((300, 400)) * 255).astype(np.float32)
    f = (f + np.roll(f, 1, 0) + np.roll(f, 1, 1) + np.roll(f, (1, 1), (0, 1))) / 4
    a = f.astype(np.uint8)
    b = np.roll(a, (3, 5), (0, 1))
    (ka, xya, _, _, da), (kb, xyb, _, _, db) = run(ctx, np.stack([a, b]), 500, cap=1500)
    idx, dist = oracle.knn2_hamming(da, db)
    good = dist[:, 0] == 0
    assert good.sum() > 200                                         # interior points reappear with identical descriptors
    shift = xyb[idx[good, 0]] - xya[good]
    assert np.all(shift == np.array([5, 3]))


def run_pyr(ctx, imgs, n, sf=1.2, nl=8, cap=None):
    import torch
    d = torch.from_numpy(np.ascontiguousarray(imgs)).cuda()
    xy, octv, resp, ang, desc, nk = ctx.orb_detect_compute_pyr(d, nfeatures=n, scale_factor=sf, nlevels=nl, cap=cap)
    torch.cuda.synchronize()
    out = []
    for b in range(imgs.shape[0]):
        k = int(nk[b])
        m = min(k, xy.shape[1])
        out.append((k, xy[b, :m].cpu().numpy(), octv[b, :m].cpu().numpy(), resp[b, :m].cpu().numpy(), ang[b, :m].cpu().numpy(),
                    desc[b, :m].cpu().numpy()))
    return out


@pytest.mark.parametrize("name", ["noise", "odd", "rects"])
@pytest.mark.parametrize("tag", ["d", "e"])
def test_orb_pyramid_matches_cv2_golden(ctx, name, tag):
    """The full detector with its scale pyramid ('d' = the reference's ORB::create(n): 8 levels, factor 1.2) vs cv2."""
    g = np.load(os.path.join(GOLD, "orb_cv2.npz"))
    n, sf, nl = g[f"{name}_pyr{tag}_cfg"]
    (k, xy, octv, resp, ang, desc), = run_pyr(ctx, g[f"{name}_img"][None], int(n), float(sf), int(nl), cap=2000)
    assert k == len(g[f"{name}_pyr{tag}_xy"])
    assert np.array_equal(xy, g[f"{name}_pyr{tag}_xy"])
    assert np.array_equal(octv, g[f"{name}_pyr{tag}_octave"])
    assert np.array_equal(resp, g[f"{name}_pyr{tag}_resp"])
    assert np.array_equal(ang, g[f"{name}_pyr{tag}_angle"])
    assert np.array_equal(desc, g[f"{name}_pyr{tag}_desc"])


def test_orb_pyramid_batch_vs_oracle(ctx, oracle):
    """A batch of EuRoC- and KITTI-shaped frames through the default detector, plus a capacity smaller than the result."""
    rng = np.random.default_rng(8)
    for (w, h, n) in ((752, 480, 1000), (1241, 376, 2000)):
        frames = []
        for s in range(3):
            f = (rng.random((h, w)) * 255).astype(np.float32)
            f = (f + np.roll(f, 1, 0) + np.roll(f, 1, 1) + np.roll(f, (1, 1), (0, 1))) / 4
            frames.append(f.astype(np.uint8))
        imgs = np.stack(frames)
        for cap in (4 * n, n // 2):
            res = run_pyr(ctx, imgs, n, cap=cap)
            for b, (k, xy, octv, resp, ang, desc) in enumerate(res):
                oxy, ooct, oresp, oang, odesc = oracle.orb_detect_compute_pyr(imgs[b], n)
                assert k == len(oxy), (b, k, len(oxy))
                m = min(k, cap)
                assert np.array_equal(xy, oxy[:m]) and np.array_equal(octv, ooct[:m])
                assert np.array_equal(resp, oresp[:m]) and np.array_equal(ang, oang[:m])
                assert np.array_equal(desc, odesc[:m])


@pytest.mark.parametrize("w,h,sf,nl", [(752, 480, 1.2, 8), (1241, 376, 1.2, 8), (640, 480, 2.0, 3), (333, 201, 1.05, 6),
                                       (501, 303, 1.7, 4), (752, 480, 2.5, 3), (750, 470, 1.002, 3)])
def test_orb_pyramid_forms_agree(ctx, w, h, sf, nl):
    """The word-based resize (four columns x eight rows per thread, folded border taps) and the per-warp trigonometry pre-pass
    against the per-pixel / per-warp forms they replace ("orb_impl" bit mask), incl. unaligned level-0 rows (1241), scale
    factors up to 2 and beyond (2.5: the generic kernel serves both settings) — every output array identical."""
    rng = np.random.default_rng(w + h)
    frames = []
    for s in range(2):
        f = (rng.random((h, w)) * 255).astype(np.float32)
        f = (f + np.roll(f, 1, 0) + np.roll(f, 1, 1) + np.roll(f, (1, 1), (0, 1))) / 4
        frames.append(f.astype(np.uint8))
    imgs = np.stack(frames)
    try:
        ctx.option("orb_impl", 15)
        old = run_pyr(ctx, imgs, 800, sf, nl, cap=1600)
    finally:
        ctx.option("orb_impl", 0)
    new = run_pyr(ctx, imgs, 800, sf, nl, cap=1600)
    for a, b in zip(old, new):
        assert a[0] == b[0] and a[0] > 100
        for x, y in zip(a[1:], b[1:]):
            assert np.array_equal(x, y)


def test_orb_levels_on_separate_streams(ctx):
    """Small batches run the pyramid levels on separate streams ("orb_lp" 1, the default: one block of scratch per level, key
    points appended in level order at the end); the one-stream form (0) must give the same arrays, call after call."""
    rng = np.random.default_rng(77)
    frames = []
    for s in range(3):
        f = (rng.random((480, 752)) * 255).astype(np.float32)
        f = (f + np.roll(f, 1, 0) + np.roll(f, 1, 1) + np.roll(f, (1, 1), (0, 1))) / 4
        frames.append(f.astype(np.uint8))
    imgs = np.stack(frames)
    try:
        ctx.option("orb_lp", 0)
        one = run_pyr(ctx, imgs, 1000, cap=2000)
    finally:
        ctx.option("orb_lp", 1)
    for _ in range(3):
        lp = run_pyr(ctx, imgs, 1000, cap=2000)
        for a, b in zip(one, lp):
            assert a[0] == b[0] and a[0] > 500
            for x, y in zip(a[1:], b[1:]):
                assert np.array_equal(x, y)
