"""GPU parity of the C++ class mirrors (vi-slam_b200/host: Matcher / CameraGPU / VISystemGPU with the
reference's member names) against the CPU oracle.  The classes are driven by host_runner (plain g++ code that
calls only the C ABI); this file writes its inputs, runs it, and checks every public result it dumps."""
import os
import subprocess

import numpy as np
import pytest

from test_gpu_gn import TOL, rot_angle

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNNER = os.path.join(ROOT, "vi-slam_b200", "vislam_b200", "host_runner")


def _runner():
    if not os.path.exists(RUNNER):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "vi-slam_b200", "host")])
    return RUNNER


def _write_meta(d, **kw):
    with open(os.path.join(d, "meta.txt"), "w") as f:
        for k, v in kw.items():
            f.write(f"{k} {float(v):.9g}\n")


def _rd(d, name, dt=np.float32):
    return np.fromfile(os.path.join(d, name), dt)


def _knn_dump(raw):
    """rows of (size, 2 x (queryIdx, trainIdx, imgIdx, distance))"""
    return raw.reshape(-1, 9)


@pytest.mark.parametrize("norm,n_cells,sym_mode", [(1, 49, 0), (1, 225, 1), (0, 49, 0)])
def test_matcher_class_vs_oracle(tmp_path, oracle, norm, n_cells, sym_mode):
    from vislam_b200 import synth
    rng = np.random.default_rng(7 + n_cells)
    n1, n2, w, h = 700, 640, 752, 480
    if norm == 1:
        d1 = synth.orb_descriptors(n1, 11)
        d2 = synth.perturb_orb(d1, 12)[0][:n2]
        dim = 32
    else:
        d1 = synth.float_descriptors(n1, 11)
        d2 = synth.perturb_float(d1, 12)[0][:n2]
        dim = 64
    kp1 = np.stack([rng.uniform(0, w - 1, n1), rng.uniform(0, h - 1, n1)], 1).astype(np.float32)
    kp2 = np.stack([rng.uniform(0, w - 1, n2), rng.uniform(0, h - 1, n2)], 1).astype(np.float32)
    kp1[::9, 1] = np.floor(kp1[::9, 1])   # equal-y ties: sort stability
    d = str(tmp_path)
    d1.tofile(os.path.join(d, "d1.bin")); d2.tofile(os.path.join(d, "d2.bin"))
    kp1.tofile(os.path.join(d, "kp1.bin")); kp2.tofile(os.path.join(d, "kp2.bin"))
    _write_meta(d, n1=n1, n2=n2, dim=dim, norm=norm, w=w, h=h, n_cells=n_cells, sym_mode=sym_mode)
    out = subprocess.run([_runner(), "matcher", d], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr

    knn = oracle.knn2_hamming if norm == 1 else oracle.knn2_l2
    i12, s12 = knn(d1, d2)
    i21, s21 = knn(d2, d1)
    # computeMatches: vector<vector<DMatch>> exactly as BFMatcher::knnMatch fills it (imgIdx 0)
    a1, a2 = _knn_dump(_rd(d, "aux1_raw.bin")), _knn_dump(_rd(d, "aux2_raw.bin"))
    assert a1.shape[0] == n1 and a2.shape[0] == n2
    for a, idx, dist in ((a1, i12, s12), (a2, i21, s21)):
        assert (a[:, 0] == 2).all()
        np.testing.assert_array_equal(a[:, 1], np.arange(a.shape[0]))        # queryIdx
        np.testing.assert_array_equal(a[:, 2].astype(np.int32), idx[:, 0])
        np.testing.assert_array_equal(a[:, 6].astype(np.int32), idx[:, 1])
        np.testing.assert_array_equal(a[:, 3], 0)                            # imgIdx
        if norm == 1:
            np.testing.assert_array_equal(a[:, 4], dist[:, 0])
            np.testing.assert_array_equal(a[:, 8], dist[:, 1])
        else:
            np.testing.assert_allclose(a[:, 4], dist[:, 0], rtol=1e-6)
            np.testing.assert_allclose(a[:, 8], dist[:, 1], rtol=1e-6)
    # nnFilter cleared the same rows the oracle's filter drops
    keep1 = np.zeros(n1, np.uint8); keep2 = np.zeros(n2, np.uint8)
    oracle.lib().vso_nn_filter(i12.reshape(-1), s12.reshape(-1), n1, oracle.RATIO, keep1)
    oracle.lib().vso_nn_filter(i21.reshape(-1), s21.reshape(-1), n2, oracle.RATIO, keep2)
    f1, f2 = _knn_dump(_rd(d, "aux1_filtered.bin")), _knn_dump(_rd(d, "aux2_filtered.bin"))
    np.testing.assert_array_equal(f1[:, 0] == 2, keep1.astype(bool))
    np.testing.assert_array_equal(f2[:, 0] == 2, keep2.astype(bool))
    # computeSymMatches / sortMatches / bestMatchesFilter
    mq, mt, md = oracle.sym_matches(i12, s12, i21, s21, mode=sym_mode)
    m = _rd(d, "matches.bin").reshape(-1, 4)
    np.testing.assert_array_equal(m[:, 0].astype(np.int32), mq)
    np.testing.assert_array_equal(m[:, 1].astype(np.int32), mt)
    np.testing.assert_array_equal(m[:, 2], -1)                               # DMatch(q, t, d): imgIdx -1
    np.testing.assert_array_equal(m[:, 3], md)
    order = oracle.sort_matches(mq, kp1)
    s = _rd(d, "sorted.bin").reshape(-1, 4)
    np.testing.assert_array_equal(s[:, 0].astype(np.int32), mq[order])
    gq, gt, gd = oracle.grid_filter(mq, mt, md, order, kp1, w, h, n_cells)
    g = _rd(d, "good.bin").reshape(-1, 4)
    np.testing.assert_array_equal(g[:, 0].astype(np.int32), gq)
    np.testing.assert_array_equal(g[:, 1].astype(np.int32), gt)
    np.testing.assert_array_equal(g[:, 3], gd)
    gk = _rd(d, "good_kp.bin").reshape(-1, 4)
    np.testing.assert_array_equal(gk[:, :2], kp1[gq])
    np.testing.assert_array_equal(gk[:, 2:], kp2[gt])
    counts = _rd(d, "counts.bin")
    assert counts[0] == len(mq) and counts[1] == len(gq)


@pytest.mark.parametrize("mirror_host,grad_images", [(1, 1), (0, 0)])
def test_visystem_gpu_sequence_vs_oracle(tmp_path, oracle, mirror_host, grad_images):
    """VISystemGPU::AddFrameGPU over a short synthetic sequence: per-pair GN pose, good-match and candidate
    counts, composed trajectory, host mirrors of pyramid / gradient, and WarpFunctionSE3, against the oracle."""
    from vislam_b200 import synth
    T, N, n_cells = 5, 300, 49
    seq = synth.make_sequence(T, n_feat=N, seed=2001)
    w, h, K = seq["w"], seq["h"], seq["K"]
    d = str(tmp_path)
    np.ascontiguousarray(seq["frames"], np.uint8).tofile(os.path.join(d, "frames.bin"))
    np.ascontiguousarray(seq["desc"], np.uint8).tofile(os.path.join(d, "desc.bin"))
    np.ascontiguousarray(seq["kp"], np.float32).tofile(os.path.join(d, "kp.bin"))
    np.ascontiguousarray(seq["R_imu_res"], np.float32).tofile(os.path.join(d, "rimu.bin"))
    np.ascontiguousarray(seq["t_res"], np.float32).tofile(os.path.join(d, "tres.bin"))
    _write_meta(d, frames=T, n_feat=N, w=w, h=h, n_cells=n_cells, mirror_host=mirror_host, grad_images=grad_images,
                fx=K[0], fy=K[1], cx=K[2], cy=K[3])
    out = subprocess.run([_runner(), "sequence", d], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    poses = _rd(d, "poses.bin").reshape(T - 1, 7)
    finals = _rd(d, "final.bin").reshape(T - 1, 7)
    ngood = _rd(d, "ngood.bin")
    ncand = _rd(d, "ncand.bin").reshape(T - 1, 5)
    niter = _rd(d, "niter.bin")
    acc = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    ref = None
    for k in range(T - 1):
        prior = oracle.initial_pose(np.eye(3), seq["R_imu_res"][k], seq["t_res"][k])
        ref = oracle.track_pair(seq["frames"][k], seq["frames"][k + 1], seq["desc"][k], seq["desc"][k + 1], seq["kp"][k],
                                K, prior, n_cells=n_cells)
        assert ngood[k] == len(ref["good_q"])
        np.testing.assert_array_equal(ncand[k], [len(c) for c in ref["cands"]])
        assert niter[k] == len(ref["trace"])
        assert rot_angle(poses[k][:4], ref["pose"][:4]) <= TOL          # 1e-5 rad
        assert np.abs(poses[k][4:] - ref["pose"][4:]).max() <= TOL      # 1e-5 m
        acc = oracle.se3_mul(acc, poses[k])
        np.testing.assert_array_equal(finals[k], acc)                   # Track(): final_poseCam *= estimate
    if mirror_host:
        last = seq["frames"][T - 1]
        pyr = oracle.pyramid(last)
        np.testing.assert_array_equal(_rd(d, "last_gray4.bin", np.uint8), pyr[4].reshape(-1))
        np.testing.assert_array_equal(_rd(d, "last_gx3.bin", np.int16), oracle.scharr3(pyr[3])[0].reshape(-1))
        pts = _rd(d, "warp2_in.bin").reshape(-1, 4)
        np.testing.assert_array_equal(pts, ref["cands"][2])
        Kl = oracle.init_pyramid(w, h, *K)
        np.testing.assert_array_equal(_rd(d, "warp2.bin").reshape(-1, 4), oracle.warp(pts, poses[T - 2], Kl[2]))


@pytest.mark.parametrize("gpu", [0, 1])
def test_camera_orb_detector_vs_oracle(tmp_path, oracle, gpu):
    """Camera / CameraGPU::detectAndComputeFeatures with the ORB detector = cv::ORB::create(200) / cuda::ORB::create(1000)
    (src/Camera.cpp:127, src/CameraGPU.cpp:99) on the device: cv::KeyPoint fields and descriptors against the oracle."""
    rng = np.random.default_rng(21)
    w, h, T = 376, 240, 3
    frames = []
    for _ in range(T):
        f = (rng.random((h, w)) * 255).astype(np.float32)
        f = (f + np.roll(f, 1, 0) + np.roll(f, 1, 1) + np.roll(f, (1, 1), (0, 1))) / 4
        frames.append(f.astype(np.uint8))
    frames = np.stack(frames)
    d = str(tmp_path)
    frames.tofile(os.path.join(d, "frames.bin"))
    _write_meta(d, frames=T, w=w, h=h, gpu=gpu)
    out = subprocess.run([_runner(), "orb", d], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    n = 1000 if gpu else 200
    for i in range(T):
        kp = _rd(d, f"orb_kp{i}.bin").reshape(-1, 6)
        desc = _rd(d, f"orb_desc{i}.bin", np.uint8).reshape(-1, 32)
        xy, octv, resp, ang, odesc = oracle.orb_detect_compute_pyr(frames[i], n)
        assert len(kp) == len(xy) >= n // 2
        np.testing.assert_array_equal(kp[:, :2], xy)
        np.testing.assert_array_equal(kp[:, 5].astype(np.int32), octv)
        np.testing.assert_array_equal(kp[:, 4], resp)
        np.testing.assert_array_equal(kp[:, 3], ang)
        size = np.array([np.float32(31.0) * np.float32(oracle.lib().vso_orb_level_scale(1.2, int(o))) for o in octv], np.float32)
        np.testing.assert_array_equal(kp[:, 2], size)
        np.testing.assert_array_equal(desc, odesc)


def test_visystem_gpu_from_raw_frames_vs_oracle(tmp_path, oracle):
    """The whole loop from images alone: VISystemGPU::AddFrameGPU with the ORB detector — device ORB (cuda::ORB::create(1000)
    as CameraGPU.cpp:99), Hamming kNN + filters, pyramid, candidates, GN — against the oracle's chain on the same frames."""
    from vislam_b200 import synth
    T, n_cells = 4, 49
    seq = synth.make_sequence(T, n_feat=10, seed=2001)
    w, h, K = seq["w"], seq["h"], seq["K"]
    d = str(tmp_path)
    np.ascontiguousarray(seq["frames"], np.uint8).tofile(os.path.join(d, "frames.bin"))
    np.ascontiguousarray(seq["R_imu_res"], np.float32).tofile(os.path.join(d, "rimu.bin"))
    np.ascontiguousarray(seq["t_res"], np.float32).tofile(os.path.join(d, "tres.bin"))
    _write_meta(d, frames=T, n_feat=0, w=w, h=h, n_cells=n_cells, mirror_host=0, grad_images=1, orb=1,
                fx=K[0], fy=K[1], cx=K[2], cy=K[3])
    out = subprocess.run([_runner(), "sequence", d], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    poses = _rd(d, "poses.bin").reshape(T - 1, 7)
    ngood = _rd(d, "ngood.bin")
    feats = [oracle.orb_detect_compute_pyr(seq["frames"][t], 1000) for t in range(T)]
    assert min(len(f[0]) for f in feats) > 100
    for k in range(T - 1):
        prior = oracle.initial_pose(np.eye(3), seq["R_imu_res"][k], seq["t_res"][k])
        r = oracle.track_pair(seq["frames"][k], seq["frames"][k + 1], feats[k][4], feats[k + 1][4], feats[k][0], K, prior,
                              n_cells=n_cells)
        assert int(ngood[k]) == len(r["good_q"]) > 10
        np.testing.assert_array_equal(poses[k], r["pose"])


def test_reference_gpu_main_runs_from_files(tmp_path):
    """The reproduced GPU executable of the reference (host/tests/main_vi_slamGPU_caller.cpp, built against the forwarding
    headers) on an EuRoC-layout dataset on disk: DataReader -> VISystemGPU::InitializeSystemGPU (calibration XML) ->
    AddFrameGPU per frame (IMU samples -> imuCore.estimate -> GN prior; ORB on the device -> match -> GN) -> CSV."""
    import subprocess
    from test_host_build import CAL_XML, PKG, _build
    from vislam_b200 import synth
    _build()
    T = 18
    seq = synth.make_sequence(T, n_feat=10, seed=2001)
    img_dir = tmp_path / "cam0" / "data"
    img_dir.mkdir(parents=True)
    t0, cam_dt, imu_dt = 1403636579763555584, 50_000_000, 5_000_000
    for k in range(T):
        with open(img_dir / f"{t0 + k * cam_dt}.pgm", "wb") as f:
            f.write(b"P5\n752 480\n255\n" + seq["frames"][k].tobytes())
    rng = np.random.default_rng(3)
    with open(tmp_path / "imu0.csv", "w") as f:
        f.write("#timestamp [ns],w_x,w_y,w_z,a_x,a_y,a_z\n")
        for k in range(T - 1):
            for i in range(10):
                w = seq["gyro"][k, i]
                a = np.array([0, 0, 9.68]) + rng.normal(0, 2e-3, 3)
                f.write(str(t0 + k * cam_dt + i * imu_dt) + "," + ",".join(repr(float(x)) for x in np.concatenate([w, a])) + "\n")
    with open(tmp_path / "gt.csv", "w") as f:
        f.write("#timestamp,p_x,p_y,p_z,q_w,q_x,q_y,q_z,v_x,v_y,v_z,bw_x,bw_y,bw_z,ba_x,ba_y,ba_z\n")
        for k in range((T - 1) * 10):
            f.write(str(t0 + k * imu_dt) + "," + ",".join(repr(float(x)) for x in [0.001 * k, 0, 1, 1, 0, 0, 0, 0.2, 0, 0] + [0] * 6) + "\n")
    (tmp_path / "cal.xml").write_text(CAL_XML % "0 0 0 0")
    out_csv = tmp_path / "out.csv"
    r = subprocess.run([os.path.join(PKG, "ref_main_gpu"), str(img_dir) + "/", str(tmp_path / "imu0.csv"), str(tmp_path / "gt.csv"),
                        str(tmp_path / "cal.xml"), str(out_csv), "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rows = np.loadtxt(out_csv, delimiter=",", ndmin=2)
    assert rows.shape[0] >= T - 4 and rows.shape[1] == 14
    assert np.isfinite(rows).all()
    q = rows[:, 3:7]
    np.testing.assert_allclose(np.linalg.norm(q, axis=1), 1.0, atol=1e-4)      # toQuaternion of the tracked orientation
