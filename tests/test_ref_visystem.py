"""The oracle's Gauss-Newton solve and Camera front end against the REFERENCE'S OWN source text.

`make -C oracle ref` compiles /root/reference/src/{VISystem,Camera,CameraModel,Matcher,Plus,Imu}.cpp unmodified, in
place, against oracle/refshim (a functional OpenCV stand-in that follows OpenCV 3.2's evaluation rules — Mat views,
lazy cv::MatExpr folding, double-accumulating gemm, LU solve, area-fast resize, Scharr) into
oracle/_ref/libref_visystem.so.  tests/golden/visystem_ref.npz holds what that library computed
(tests/golden/make_visystem_golden.py); the tests below hold the oracle to it bit for bit, and — where the library is
present (this container; it also travels to the GPU box) — compare live on larger inputs.

Doing this surfaced two places where the restatement had followed the C++ spelling instead of what OpenCV evaluates:
`(col - cx) * invfx` is ONE scaled conversion x*invfx + (-cx*invfx), and `A.inv() * b` is cv::solve(A, b).
"""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["a", "b", "c"]


def load(name):
    return np.load(os.path.join(GOLD, name))


def test_solve6_matches_cv2_golden(oracle):
    g = load("solve6_cv2.npz")
    for a, b, x in zip(g["A"], g["b"], g["x"]):
        ok, got = oracle.solve6(a, b)
        assert np.array_equal(got, x)
    assert not oracle.solve6(g["A"][-1], g["b"][-1])[0]          # singular -> zeros


def test_solve6_is_not_inverse_times_b(oracle):
    """The two forms differ in the last bits — which is why the distinction matters for bit parity."""
    g = load("solve6_cv2.npz")
    differ = 0
    for a, b in zip(g["A"][:-1], g["b"][:-1]):
        _, x = oracle.solve6(a, b)
        _, inv = oracle.inv6(a)
        y = (inv.astype(np.float64) @ b.astype(np.float64)).astype(np.float32)
        differ += not np.array_equal(x, y)
        assert np.allclose(x, y, rtol=2e-2, atol=1e-6)
    assert differ > 0


def _oracle_run(oracle, g, tag, good_prev):
    prev, cur = g[f"{tag}_prev"], g[f"{tag}_cur"]
    h, w = prev.shape
    K = tuple(float(v) for v in g[f"{tag}_K"])
    prev_pyr, cur_pyr = oracle.pyramid(prev), oracle.pyramid(cur)
    grads = [oracle.scharr3(p) for p in prev_pyr]
    cands = [oracle.candidates(good_prev, l, w >> l, h >> l) for l in range(5)]
    pose0 = oracle.initial_pose(g[f"{tag}_imu2cam"], g[f"{tag}_r_imu_res"], g[f"{tag}_t_res"])
    pose, trace = oracle.gn_solve(prev_pyr, cur_pyr, [x[0] for x in grads], [x[1] for x in grads], cands,
                                  oracle.init_pyramid(w, h, *K), pose0, oracle.default_opts(first_lvl=3))
    return prev_pyr, grads, cands, pose, trace


def _check_trace(ref_trace, trace):
    """The reference prints `lvl = <l>Error it <k> =<error>` with 6 significant digits (VISystem.cpp:1351)."""
    assert len(ref_trace) == len(trace)
    for r, t in zip(ref_trace, trace):
        assert (int(r[0]), int(r[1])) == (t["lvl"], t["iter"])
        assert float("%g" % t["error"]) == pytest.approx(float(r[2]), rel=2e-6)


@pytest.mark.parametrize("tag", CASES)
def test_oracle_matches_reference_visystem_golden(oracle, tag):
    g = load("visystem_ref.npz")
    prev = g[f"{tag}_prev"]
    h, w = prev.shape
    if int(g[f"{tag}_own_matcher"]):
        # the good matches came out of the reference's Camera::computeGoodMatches (its own Matcher on these descriptors)
        gq, gt, _, _ = oracle.match_pipeline(g[f"{tag}_d1"], g[f"{tag}_d2"], g[f"{tag}_kp1"], w, h, int(g[f"{tag}_n_cells"]), 1)
        assert np.array_equal(g[f"{tag}_kp1"].reshape(-1, 2)[gq], g[f"{tag}_good_prev"])
        assert np.array_equal(g[f"{tag}_kp2"].reshape(-1, 2)[gt], g[f"{tag}_good_cur"])
    prev_pyr, grads, cands, pose, trace = _oracle_run(oracle, g, tag, g[f"{tag}_good_prev"])
    for l in range(5):
        assert np.array_equal(prev_pyr[l], g[f"{tag}_pyr{l}"]), f"pyramid level {l}"       # Camera::Update
        if l >= 1:
            assert np.array_equal(grads[l][0], g[f"{tag}_gx{l}"]), f"gx level {l}"          # Camera::computeGradient
            assert np.array_equal(grads[l][1], g[f"{tag}_gy{l}"]), f"gy level {l}"
        c = np.asarray(cands[l], np.float32).reshape(-1, 4)                                 # ObtainPatchesPointsPreviousFrame
        assert c.shape[0] == int(g[f"{tag}_n_cand"][l])
        assert np.array_equal(c[:, :2], g[f"{tag}_cand{l}"].astype(np.float32))
        assert np.all(c[:, 2:] == 1.0)
    assert np.array_equal(pose, g[f"{tag}_pose"]), (pose, g[f"{tag}_pose"])                  # EstimatePoseFeatures, bit for bit
    _check_trace(g[f"{tag}_trace"], trace)


def test_warp_and_tukey_match_reference_golden(oracle):
    g = load("visystem_ref.npz")
    K = oracle.init_pyramid(752, 480, *[float(v) for v in g["warp_K"]])
    out = oracle.warp(g["warp_pts"], g["warp_pose"], K[int(g["warp_lvl"])])
    assert np.array_equal(out, g["warp_out"])                                               # WarpFunctionSE3
    assert np.array_equal(oracle.tukey_weights(g["tukey_r"]), g["tukey_w"])                 # TukeyFunctionWeights


# ------------------------------------------------------------------------------------------------- live comparisons
@pytest.fixture(scope="module")
def ref():
    from oracle import ref_visystem as rv
    if not rv.available():
        pytest.skip("oracle/_ref/libref_visystem.so is not built (needs /root/reference)")
    rv.lib()
    return rv


def test_reference_library_reproduces_its_golden(ref):
    """Guards the fixture against a stale or differently-built library."""
    g = load("visystem_ref.npz")
    for tag in CASES:
        K = tuple(float(v) for v in g[f"{tag}_K"])
        r = ref.track_pair(g[f"{tag}_prev"], g[f"{tag}_cur"], K, g[f"{tag}_imu2cam"], g[f"{tag}_r_imu_res"], g[f"{tag}_t_res"],
                           good_prev=g[f"{tag}_good_prev"], good_cur=g[f"{tag}_good_cur"], n_cells=int(g[f"{tag}_n_cells"]))
        assert r["oob_reads"] == 0
        assert np.array_equal(r["pose"], g[f"{tag}_pose"])
        assert np.array_equal(r["trace"], g[f"{tag}_trace"])


def test_reference_reads_past_the_image_on_some_inputs(ref):
    """SURVEY App. B-4, shown by the reference's own code: VISystem.cpp:1299 tests y2 < rows, :1321 reads row round(y2), so
    y2 in [rows - 0.5, rows) reads past the current image.  Upstream that is undefined behaviour (whatever follows the
    image on the heap); the oracle and the kernels reject such a point.  Bit parity is therefore only defined — and only
    asserted — for runs without such reads; about 40 % of random small pairs have some."""
    from vislam_b200 import synth
    from golden.make_visystem_golden import prior_inputs
    K = (114.6635, 114.324, 91.42875, 61.71875)
    hits = 0
    for seed in (66, 69, 77, 79, 80):
        p = synth.make_pair(w=188, h=120, n_feat=300, K=K, seed=seed)
        i2c, rres, tres = prior_inputs(seed)
        r = ref.track_pair(p["prev"], p["cur"], K, i2c, rres, tres, n_cells=49,
                           kp_prev=p["kp1"], desc_prev=p["d1"], kp_cur=p["kp2"], desc_cur=p["d2"])
        hits += r["oob_reads"] > 0
    assert hits > 0


@pytest.mark.parametrize("w,h,K,n_cells,seed", [
    (752, 480, (458.654, 457.296, 367.215, 248.375), 49, 1001),          # BASELINE configs[0]: EuRoC-shaped pair
    (752, 480, (458.654, 457.296, 367.215, 248.375), 225, 1002),         # 200-feature cap
    (640, 480, (525.0, 525.0, 319.5, 239.5), 100, 3001),                 # TUM-shaped
    (621, 188, (359.428, 359.428, 303.5964, 92.60785), 64, 4001),        # KITTI level-1-shaped, odd width
])
def test_full_size_pair_live(oracle, ref, w, h, K, n_cells, seed):
    from vislam_b200 import synth
    from golden.make_visystem_golden import prior_inputs
    for seed in range(seed, seed + 20):                 # first seed whose run is well defined upstream (no App. B-4 reads)
        p = synth.make_pair(w=w, h=h, n_feat=1000, K=K, seed=seed)
        i2c, rres, tres = prior_inputs(seed)
        r = ref.track_pair(p["prev"], p["cur"], K, i2c, rres, tres, n_cells=n_cells,
                           kp_prev=p["kp1"], desc_prev=p["d1"], kp_cur=p["kp2"], desc_cur=p["d2"])
        if r["oob_reads"] == 0:
            break
    assert r["oob_reads"] == 0
    g = {"x_prev": p["prev"], "x_cur": p["cur"], "x_K": np.array(K), "x_imu2cam": i2c, "x_r_imu_res": rres, "x_t_res": tres}
    gq, gt, _, _ = oracle.match_pipeline(p["d1"], p["d2"], p["kp1"], w, h, n_cells, 1)
    assert np.array_equal(p["kp1"].reshape(-1, 2)[gq], r["good_prev"])
    prev_pyr, grads, cands, pose, trace = _oracle_run(oracle, g, "x", r["good_prev"])
    for l in range(5):
        assert np.array_equal(prev_pyr[l], r["pyr_prev"][l])
        assert np.array_equal(grads[l][0], r["gx"][l]) and np.array_equal(grads[l][1], r["gy"][l])
        assert np.array_equal(np.asarray(cands[l], np.float32).reshape(-1, 4), r["cands"][l])
    assert np.array_equal(pose, r["pose"]), (pose, r["pose"])
    _check_trace(r["trace"], trace)


def test_warp_and_tukey_live(oracle, ref):
    rng = np.random.default_rng(12)
    K4 = (458.654, 457.296, 367.215, 248.375)
    K = oracle.init_pyramid(752, 480, *K4)
    for lvl in range(5):
        w, h = 752 >> lvl, 480 >> lvl
        pts = np.ones((2000, 4), np.float32)
        pts[:, 0] = rng.integers(1, w, 2000)
        pts[:, 1] = rng.integers(1, h, 2000)
        if lvl % 2:
            pts[:, 2] = rng.uniform(0.3, 4.0, 2000).astype(np.float32)
        pose = oracle.se3_exp(rng.uniform(-0.05, 0.05, 6).astype(np.float32))
        assert np.array_equal(oracle.warp(pts, pose, K[lvl]), ref.warp(pts, pose, 752, 480, K4, lvl))
    for n in (1, 2, 9, 1000, 20000):
        r = rng.normal(0, 10.0 ** rng.uniform(0, 2.5), n).astype(np.float32)
        assert np.array_equal(oracle.tukey_weights(r), ref.tukey(r))
