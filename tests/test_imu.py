"""CPU: the IMU prior.
  * Imu::initializate / estimate chain of the host-side mirror against the reference's OWN src/Imu.cpp, compiled
    unmodified against the cv + ROS shims (oracle/_ref/libref_imu.so), on the same samples and the same filter answers —
    every public field bit for bit; live where the reference exists, through tests/golden/imu_ref.npz everywhere.
  * MadgwickFilter (restatement of the external imu_filter_madgwick node's algorithm; the package is not vendored, so
    it is checked through properties: unit norm, convergence to the gravity direction, gyro integration, gradient)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_SO = os.path.join(ROOT, "vi-slam_b200", "vislam_b200", "libvislam_host.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_imu.so")
GOLDEN = os.path.join(ROOT, "tests", "golden", "imu_ref.npz")
dbl = C.c_double
DT, N0, N_PER, STEPS = 0.005, 40, 10, 12


def ptr(a):
    return a.ctypes.data_as(C.POINTER(dbl))


@pytest.fixture(scope="module")
def host():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "vi-slam_b200")])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "vi-slam_b200", "host")])
    return C.CDLL(HOST_SO)


@pytest.fixture(scope="module")
def ref():
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    return C.CDLL(REF_SO) if os.path.exists(REF_SO) else None


def samples(seed=7):
    """200 Hz gyro / accelerometer of a slowly rotating, gently accelerating body (z-up world, g = 9.68 as upstream)."""
    rng = np.random.default_rng(seed)
    n = N0 + N_PER * STEPS
    t = np.arange(n) * DT
    w = np.stack([0.3 * np.sin(2.0 * t), 0.2 * np.cos(1.3 * t), 0.4 * np.ones(n)], 1) + rng.normal(0, 2e-3, (n, 3))
    a = np.stack([0.4 * np.sin(3 * t), 0.3 * np.cos(2 * t), 9.68 + 0.2 * np.sin(5 * t)], 1) + rng.normal(0, 2e-2, (n, 3))
    return np.ascontiguousarray(w), np.ascontiguousarray(a)


def run_host(host, w, a, q_in=None, want_q=False):
    n = len(w)
    out = np.zeros((STEPS + 1, 60))
    q_out = np.zeros((n, 4)) if want_q else None
    vel = np.array([0.2, -0.1, 0.05])
    rc = host.vih_imu_run(dbl(DT), dbl(0.7), ptr(vel), ptr(w), ptr(a), ptr(q_in) if q_in is not None else None,
                          ptr(q_out) if want_q else None, N0, N_PER, STEPS, ptr(out))
    assert rc == 0
    return out, q_out


def run_ref(ref, w, a, q):
    out = np.zeros((STEPS + 1, 60))
    vel = np.array([0.2, -0.1, 0.05])
    assert ref.ref_imu_run(dbl(DT), dbl(0.7), ptr(vel), ptr(w), ptr(a), ptr(q), N0, N_PER, STEPS, ptr(out)) == 0
    return out


def test_imu_estimate_matches_reference_sources(host, ref):
    w, a = samples()
    out_builtin, q = run_host(host, w, a, want_q=True)
    out_table, _ = run_host(host, w, a, q_in=q)
    np.testing.assert_array_equal(out_builtin, out_table)        # the filter hook is transparent
    gold = np.load(GOLDEN)
    np.testing.assert_array_equal(q, gold["q"])
    np.testing.assert_array_equal(out_builtin, gold["out"])
    if ref is not None:
        np.testing.assert_array_equal(out_builtin, run_ref(ref, w, a, q))
    # the chain did something: residual rotation is a proper rotation close to the integrated gyro, biases are learnt
    R = out_builtin[-1, 42:51].reshape(3, 3)
    np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-5)
    assert np.linalg.norm(out_builtin[3, 15:18]) > 0.1          # angBias re-estimated while currentTimeMs < 2500
    # upstream re-estimates the gyro bias as the mean rate of every block while currentTimeMs < 2500 (Imu.cpp:425-430), so
    # the constant 0.4 rad/s yaw rate of these samples is "calibrated" away and the residual yaw per block stays ~0
    np.testing.assert_allclose(out_builtin[-1, 15:18], [0, 0, 0.4], atol=0.35)
    assert abs(out_builtin[-1, 2]) < 2e-3


def test_madgwick_filter_properties(host):
    rng = np.random.default_rng(3)
    n = 4000
    # (1) static, tilted sensor: converges to the tilt that maps the measured specific force onto world +z, unit norm
    roll, pitch = 0.3, -0.2
    cr, sr, cp, sp = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch)
    Rwb = np.array([[cp, sp * sr, sp * cr], [0, cr, -sr], [-sp, cp * sr, cp * cr]])     # Ry(pitch) Rx(roll), yaw 0
    acc = np.tile(Rwb.T @ np.array([0, 0, 9.68]), (n, 1)) + rng.normal(0, 1e-3, (n, 3))
    gyr = rng.normal(0, 1e-4, (n, 3))
    q = np.zeros((n, 4))
    host.vih_madgwick_run(dbl(0.1), dbl(DT), ptr(gyr), ptr(acc), n, ptr(q))
    np.testing.assert_allclose(np.linalg.norm(q, axis=1), 1.0, atol=1e-12)
    w_, x, y, z = q[-1]
    up_in_body = np.array([2 * (x * z - w_ * y), 2 * (w_ * x + y * z), 1 - 2 * (x * x + y * y)])      # R(q)^T e_z
    np.testing.assert_allclose(up_in_body, acc[-1] / np.linalg.norm(acc[-1]), atol=2e-3)
    # the very first answer is already the accelerometer tilt (stateless initialisation), zero yaw
    w0, x0, y0, z0 = q[0]
    assert abs(np.arctan2(2 * (w0 * x0 + y0 * z0), 1 - 2 * (x0 * x0 + y0 * y0)) - roll) < 2e-3
    assert abs(np.arctan2(2 * (w0 * z0 + x0 * y0), 1 - 2 * (y0 * y0 + z0 * z0))) < 2e-3
    # (2) pure yaw rotation about gravity: yaw integrates the gyro (the accelerometer cannot see yaw)
    gyr = np.tile([0.0, 0.0, 0.5], (n, 1))
    acc = np.tile([0.0, 0.0, 9.68], (n, 1))
    host.vih_madgwick_run(dbl(0.1), dbl(DT), ptr(gyr), ptr(acc), n, ptr(q))
    yaw = np.unwrap(np.arctan2(2 * (q[:, 0] * q[:, 3] + q[:, 1] * q[:, 2]), 1 - 2 * (q[:, 2] ** 2 + q[:, 3] ** 2)))
    np.testing.assert_allclose(yaw[-1], 0.5 * DT * n, rtol=1e-6)
    # (3) gain 0 == plain first-order quaternion integration of the gyro
    gyr = rng.normal(0, 0.5, (200, 3))
    acc = np.tile([0.0, 0.0, 9.68], (200, 1))
    q2 = np.zeros((200, 4))
    host.vih_madgwick_run(dbl(0.0), dbl(DT), ptr(gyr), ptr(acc), 200, ptr(q2))
    p = np.array([1.0, 0, 0, 0])
    for i in range(200):
        gx, gy, gz = gyr[i]
        d = 0.5 * np.array([-p[1] * gx - p[2] * gy - p[3] * gz, p[0] * gx + p[2] * gz - p[3] * gy,
                            p[0] * gy - p[1] * gz + p[3] * gx, p[0] * gz + p[1] * gy - p[2] * gx])
        p = p + d * DT
        p /= np.linalg.norm(p)
        np.testing.assert_allclose(q2[i], p, atol=1e-13)
