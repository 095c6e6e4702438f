"""CPU: dataset ingestion and the Plus angle helpers of the host-side class mirrors
against the reference's OWN sources — src/Plus.cpp, GroundTruth.cpp, ImageReader.cpp, DataReader.cpp compiled
unmodified against the type shim into oracle/_ref/libref_io.so — bit for bit: live where that library exists (the build
container), and through tests/golden/io_ref.npz (its recorded outputs, tests/golden/make_io_golden.py) everywhere."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import dataset_fixture as fx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_SO = os.path.join(ROOT, "vi-slam_b200", "vislam_b200", "libvislam_host.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_io.so")
GOLDEN = os.path.join(ROOT, "tests", "golden", "io_ref.npz")

dbl, flt = C.c_double, C.c_float


def _load(path):
    L = C.CDLL(path)
    return L


@pytest.fixture(scope="module")
def host():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "vi-slam_b200")])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "vi-slam_b200", "host")])
    return _load(HOST_SO)


@pytest.fixture(scope="module")
def ref():
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    if not os.path.exists(REF_SO):
        return None
    return _load(REF_SO)


def plus_inputs():
    rng = np.random.default_rng(2024)
    rpy = rng.uniform(-np.pi, np.pi, (200, 3))
    rpy[:8] = [[0, 0, 0], [np.pi, 0, 0], [0, np.pi / 2, 0], [0, -np.pi / 2, 0], [1e-9, -1e-9, 3.0], [-3.1, 1.5, -3.1],
               [0.1, 1.5707963, 0.2], [2.5, -1.2, -0.7]]
    pos = rng.normal(0, 2, (200, 3))
    return rpy, pos


def run_plus(L, prefix):
    """Every Plus helper on the shared inputs -> dict of arrays."""
    rpy, pos = plus_inputs()
    n = len(rpy)
    out = {k: [] for k in ("quat", "rpy_from_quat", "rpy360", "diff", "rot", "rpy_from_rot", "tm", "tm_rpy", "tm_pos")}
    f = lambda name: getattr(L, prefix + name)
    f("computeDiff").restype = dbl
    f("computeDiff").argtypes = [dbl, dbl]
    f("toQuaternion").argtypes = [dbl, dbl, dbl, C.POINTER(dbl)]
    for i in range(n):
        q = (dbl * 4)()
        f("toQuaternion")(rpy[i, 0], rpy[i, 1], rpy[i, 2], q)
        out["quat"].append(list(q))
        a = (dbl * 3)()
        f("toRPY")(q, a)
        out["rpy_from_quat"].append(list(a))
        b = (dbl * 3)()
        f("toRPY360")((dbl * 3)(*rpy[i]), b)
        out["rpy360"].append(list(b))
        out["diff"].append(f("computeDiff")(rpy[i, 0], rpy[i, 2]))
        m = (flt * 9)()
        f("RPY2rotationMatrix")((dbl * 3)(*rpy[i]), m)
        out["rot"].append(list(m))
        c = (dbl * 3)()
        f("rotationMatrix2RPY")(m, c)
        out["rpy_from_rot"].append(list(c))
        t = (flt * 16)()
        f("RPYAndPosition2transformationMatrix")((dbl * 3)(*rpy[i]), (dbl * 3)(*pos[i]), t)
        out["tm"].append(list(t))
        r2, p2 = (dbl * 3)(), (dbl * 3)()
        f("transformationMatrix2RPY_position")(t, r2, p2)
        out["tm_rpy"].append(list(r2))
        out["tm_pos"].append(list(p2))
    return {k: np.array(v) for k, v in out.items()}


def run_dataset(L, prefix, root):
    img_dir, imu_csv, gt_csv, times = fx.write(root)
    f = lambda name: getattr(L, prefix + name)
    res = {}
    # GroundTruth
    for tag, path in (("imu", imu_csv), ("gt", gt_csv)):
        cols, ts = C.c_int(), dbl()
        buf = (dbl * 8192)()
        f("groundtruth_read").restype = C.c_int
        rows = f("groundtruth_read")(path.encode(), C.c_char(b","), C.byref(cols), C.byref(ts), buf, 8192)
        assert rows > 0
        data = np.array(buf[: rows * cols.value]).reshape(rows, cols.value)
        res[f"{tag}_rows"], res[f"{tag}_cols"], res[f"{tag}_timestep"] = rows, cols.value, ts.value
        res[f"{tag}_data"] = data[: rows - 1]        # the last row belongs to the header line (zeros / upstream UB)
    # ImageReader
    tbuf = (C.c_long * 64)()
    ts = dbl()
    f("imagereader_list").restype = C.c_int
    n = f("imagereader_list")(img_dir.encode(), tbuf, 64, C.byref(ts))
    res["image_times"] = np.array(tbuf[:n], np.int64)
    res["image_timestep"] = ts.value
    # DataReader
    idx, tt = (C.c_int * 4)(), (dbl * 6)()
    f("datareader_open").restype = C.c_void_p
    h = f("datareader_open")(img_dir.encode(), imu_csv.encode(), gt_csv.encode(), C.c_char(b","), idx, tt)
    assert h
    res["sync_idx"] = np.array(idx[:])
    res["sync_t"] = np.array(tt[:5])
    steps = []
    for k in range(0, 8):
        counts, misc, cks = (C.c_int * 6)(), (dbl * 4)(), (C.c_long * 2)()
        imu, gt = (dbl * (6 * 64))(), (dbl * (16 * 64))()
        f("datareader_update")(C.c_void_p(h), k, k + 1, counts, imu, gt, misc, cks, 64)
        steps.append(dict(counts=np.array(counts[:]), misc=np.array(misc[:]), cks=np.array(cks[:]),
                          imu=np.array(imu[: 6 * counts[0]]), gt=np.array(gt[: 16 * counts[1]])))
    f("datareader_close")(C.c_void_p(h))
    for key in ("counts", "misc", "cks"):
        res["step_" + key] = np.stack([s[key] for s in steps])
    res["step_imu"] = np.concatenate([s["imu"] for s in steps])
    res["step_gt"] = np.concatenate([s["gt"] for s in steps])
    return res


def _assert_same(a, b):
    assert set(a) == set(b)
    for k in a:
        np.testing.assert_array_equal(np.asarray(a[k]), np.asarray(b[k]), err_msg=k)


def test_plus_helpers_match_reference_sources(host, ref):
    got = run_plus(host, "vih_")
    gold = dict(np.load(GOLDEN))
    _assert_same(got, {k[5:]: v for k, v in gold.items() if k.startswith("plus_")})
    if ref is not None:
        _assert_same(got, run_plus(ref, "ref_"))


def test_oracle_rpy_helpers_match_reference_sources(oracle):
    """The C oracle's rotation <-> RPY restatement (used for the GN initial pose, VISystem.cpp:1135-1168) against the
    recorded outputs of the reference's own Plus.cpp: pins SURVEY.md 8a row G-7 by the reference itself."""
    gold = dict(np.load(GOLDEN))
    rpy, _ = plus_inputs()
    L = oracle.lib()
    for i in range(len(rpy)):                    # the oracle's ctypes signatures take numpy arrays
        m = np.zeros(9, np.float32)
        L.vso_rpy_to_rot(np.ascontiguousarray(rpy[i], np.float64), m)
        np.testing.assert_array_equal(m, gold["plus_rot"][i].astype(np.float32))
        a = np.zeros(3, np.float64)
        L.vso_rot_to_rpy(m, a)
        np.testing.assert_array_equal(a, gold["plus_rpy_from_rot"][i])


def test_dataset_ingestion_matches_reference_sources(host, ref, tmp_path):
    got = run_dataset(host, "vih_", str(tmp_path / "host"))
    gold = dict(np.load(GOLDEN))
    _assert_same(got, {k[3:]: v for k, v in gold.items() if k.startswith("ds_")})
    if ref is not None:
        _assert_same(got, run_dataset(ref, "ref_", str(tmp_path / "ref")))
    # the synchronisation did real work: images before the first IMU / GT sample are skipped
    assert got["sync_idx"][0] == 4 and got["step_counts"][:, 0].min() >= 9


def test_image_decode_png_and_pgm(host, tmp_path):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    host.vih_imread_gray.restype = C.c_int

    def read(path):
        r, c = C.c_int(), C.c_int()
        assert host.vih_imread_gray(path.encode(), C.byref(r), C.byref(c), None, 0) == 0
        buf = (C.c_ubyte * (r.value * c.value))()
        assert host.vih_imread_gray(path.encode(), C.byref(r), C.byref(c), buf, r.value * c.value) == 0
        return np.frombuffer(buf, np.uint8).reshape(r.value, c.value).copy()

    for shape in ((480, 752), (37, 53), (1, 1), (5, 300)):
        g = rng.integers(0, 256, shape, dtype=np.uint8)
        g[: shape[0] // 2] = np.arange(shape[1]) % 256            # smooth rows exercise the Sub/Up/Paeth filters
        for ext in ("png", "pgm"):
            p = str(tmp_path / f"g_{shape[0]}_{shape[1]}.{ext}")
            assert cv2.imwrite(p, g)
            np.testing.assert_array_equal(read(p), g)
            np.testing.assert_array_equal(read(p), cv2.imread(p, cv2.IMREAD_GRAYSCALE))
    # colour PNG (TUM rgb/): grayscale conversion within one level of cv2's (OpenCV 3.2 lets libpng convert with fixed-point
    # 0.299 / 0.587 weights; cv2 4.x converts after decoding)
    bgr = rng.integers(0, 256, (48, 64, 3), dtype=np.uint8)
    p = str(tmp_path / "c.png")
    assert cv2.imwrite(p, bgr)
    diff = read(p).astype(int) - cv2.imread(p, cv2.IMREAD_GRAYSCALE).astype(int)
    assert np.abs(diff).max() <= 1
    bgra = np.dstack([bgr, rng.integers(0, 256, (48, 64), dtype=np.uint8)])
    assert cv2.imwrite(p, bgra)
    assert np.abs(read(p).astype(int) - cv2.imread(p, cv2.IMREAD_GRAYSCALE).astype(int)).max() <= 1
    # unsupported / missing files fail loudly
    r, c = C.c_int(), C.c_int()
    assert host.vih_imread_gray(str(tmp_path / "missing.png").encode(), C.byref(r), C.byref(c), None, 0) != 0


def test_dataset_ingestion_reads_png_frames(host, tmp_path):
    """The same dataset with PNG frames (the real EuRoC format) gives the same pixel checksums."""
    pytest.importorskip("cv2")
    img_dir, imu_csv, gt_csv, _ = fx.write(str(tmp_path / "png"), ext="png")
    idx, tt = (C.c_int * 4)(), (dbl * 6)()
    host.vih_datareader_open.restype = C.c_void_p
    h = host.vih_datareader_open(img_dir.encode(), imu_csv.encode(), gt_csv.encode(), C.c_char(b","), idx, tt)
    assert h
    counts, misc, cks = (C.c_int * 6)(), (dbl * 4)(), (C.c_long * 2)()
    imu, gt = (dbl * (6 * 64))(), (dbl * (16 * 64))()
    assert host.vih_datareader_update(C.c_void_p(h), 0, 1, counts, imu, gt, misc, cks, 64) == 0
    host.vih_datareader_close(C.c_void_p(h))
    assert list(counts[2:]) == [fx.H, fx.W, fx.H, fx.W]
    assert cks[0] == int(fx.image(idx[0]).sum()) and cks[1] == int(fx.image(idx[0] + 1).sum())


def test_too_few_images_and_missing_files_fail_loudly(host, tmp_path):
    d = tmp_path / "few"
    d.mkdir()
    for i in range(5):
        (d / f"{i}.pgm").write_bytes(b"P5\n1 1\n255\n\x00")
    ts = dbl()
    host.vih_imagereader_list.restype = C.c_int
    host.vih_last_error.restype = C.c_char_p
    assert host.vih_imagereader_list((str(d) + "/").encode(), (C.c_long * 8)(), 8, C.byref(ts)) < 0
    assert b"insufficient" in host.vih_last_error()
    cols = C.c_int()
    host.vih_groundtruth_read.restype = C.c_int
    assert host.vih_groundtruth_read(str(tmp_path / "nope.csv").encode(), C.c_char(b","), C.byref(cols), C.byref(ts),
                                     (dbl * 4)(), 4) < 0


def test_trajectory_csv_row_format(host, tmp_path):
    """27 comma-separated columns per frame in the order of src/main_vi_slam.cpp:183-210."""
    rows = np.arange(54, dtype=np.float64).reshape(2, 27) * 0.5 + 0.25
    p = str(tmp_path / "traj.csv")
    host.vih_trajectory_write.restype = C.c_int
    assert host.vih_trajectory_write(p.encode(), rows.ctypes.data_as(C.POINTER(dbl)), 2) == 0
    back = np.loadtxt(p, delimiter=",")
    assert back.shape == (2, 27)
    np.testing.assert_allclose(back, rows, rtol=1e-5)          # ostream default precision: 6 significant digits
