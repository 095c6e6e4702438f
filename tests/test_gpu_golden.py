"""GPU vs the committed golden fixtures: cv2 4.13 outputs, the reference's own Matcher.cpp outputs and the
pinned oracle GN trace — through the C ABI, without calling the oracle at all."""
import os

import numpy as np
import pytest

from test_gpu_gn import TOL, rot_angle

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_knn_vs_cv2_golden(ctx):
    import torch
    g = np.load(os.path.join(G, "knn_cv2.npz"))
    out = ctx.knn2_hamming(torch.from_numpy(g["d1"]).cuda(), torch.from_numpy(g["d2"]).cuda())
    torch.cuda.synchronize()
    i12, s12, i21, s21 = [o.cpu().numpy() for o in out]
    np.testing.assert_array_equal(i12, g["idx12"])
    np.testing.assert_array_equal(s12, g["dist12"])
    np.testing.assert_array_equal(i21, g["idx21"])
    np.testing.assert_array_equal(s21, g["dist21"])


def test_filter_vs_reference_matcher_cpp_golden(ctx):
    import torch
    g = np.load(os.path.join(G, "matcher_ref.npz"))
    for i in range(int(g["n_cases"])):
        d1, d2, kp1 = g[f"c{i}_d1"], g[f"c{i}_d2"], g[f"c{i}_kp1"]
        k = ctx.knn2_hamming(torch.from_numpy(d1).cuda()[None], torch.from_numpy(d2).cuda()[None])
        gq, gt, gd, ng, ns = ctx.match_filter(*k, torch.from_numpy(kp1).cuda()[None], 752, 480, int(g[f"c{i}_ncells"]))
        torch.cuda.synchronize()
        n = int(ng[0])
        assert int(ns[0]) == len(g[f"c{i}_sym_q"])
        np.testing.assert_array_equal(gq[0, :n].cpu().numpy(), g[f"c{i}_good_q"])
        np.testing.assert_array_equal(gt[0, :n].cpu().numpy(), g[f"c{i}_good_t"])
        np.testing.assert_array_equal(gd[0, :n].cpu().numpy(), g[f"c{i}_good_d"])
        xy = ctx.gather_keypoints(torch.from_numpy(kp1).cuda()[None], gq, ng)
        np.testing.assert_array_equal(xy[0, :n].cpu().numpy(), g[f"c{i}_prev_xy"])


@pytest.mark.parametrize("tag", ["even", "odd", "kitti"])
def test_camera_vs_cv2_golden(ctx, tag):
    import torch
    import vislam_b200 as vb
    g = np.load(os.path.join(G, "camera_cv2.npz"))
    img = g[tag + "_l0"]
    h, w = img.shape
    levels = sum(1 for l in range(5) if f"{tag}_l{l}" in g)
    lay = vb.pyr_layout(w, h, levels)
    pyr = ctx.pyramid_build(torch.from_numpy(img).cuda()[None], lay)
    gx, gy, gm = ctx.gradient_build(pyr, lay, want_mag=True)
    torch.cuda.synchronize()
    pyr, gx, gy, gm = [t[0].cpu().numpy() for t in (pyr, gx, gy, gm)]
    for l in range(levels):
        lv = pyr[lay.offset[l]: lay.offset[l] + lay.w[l] * lay.h[l]].reshape(lay.h[l], lay.w[l])
        np.testing.assert_array_equal(lv, g[f"{tag}_l{l}"])
    n0 = w * h
    np.testing.assert_array_equal(gx[:n0].reshape(h, w), g[tag + "_gx"])
    np.testing.assert_array_equal(gy[:n0].reshape(h, w), g[tag + "_gy"])
    np.testing.assert_array_equal(gm[:n0].reshape(h, w), g[tag + "_gm"])


@pytest.mark.parametrize("tag", ["ref", "huber", "bilinear"])
def test_tracker_vs_pinned_oracle_trace(ctx, tag):
    import torch
    import vislam_b200 as vb
    g = np.load(os.path.join(G, "gn_oracle.npz"))
    K = tuple(float(x) for x in g["K"])
    kw = {"ref": {}, "huber": dict(weight_mode=2, huber_k=12.0), "bilinear": dict(sample_mode=1)}[tag]
    h, w = g["prev"].shape
    tr = ctx.tracker(w, h, g["d1"].shape[0], K, n_cells=49, max_pairs=1, gn_opts=vb.default_gn_opts(first_lvl=2, **kw))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()[None]
    pose, ng = tr.track_pairs(dev(g["prev"]), dev(g["cur"]), dev(g["d1"]), dev(g["d2"]), dev(g["kp1"]), dev(g["prior"]))
    torch.cuda.synchronize()
    pose = pose[0].cpu().numpy()
    assert int(ng[0]) == len(g[tag + "_good_q"])
    assert rot_angle(pose[:4], g[tag + "_pose"][:4]) <= TOL
    assert np.abs(pose[4:] - g[tag + "_pose"][4:]).max() <= TOL
    tr.close()
