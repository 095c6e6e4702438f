"""Deterministic synthetic EuRoC-shaped dataset on disk: a directory of <timestamp_ns>.pgm
frames at 20 Hz, imu0-style CSV at 200 Hz (t, w xyz, a xyz) and state_groundtruth_estimate0-style CSV at 200 Hz
(t, p xyz, q wxyz, v xyz, bw xyz, ba xyz), each with one '#' header line.  The three streams start at different times
so that DataReader's synchronisation has work to do."""
import os

import numpy as np

N_IMAGES, W, H = 24, 16, 12
T0 = 1403636579763555584          # EuRoC-like nanosecond epoch
CAM_DT, IMU_DT = 50_000_000, 5_000_000


def image(i):
    yy, xx = np.mgrid[0:H, 0:W]
    return ((xx * 7 + yy * 13 + i * 29) % 256).astype(np.uint8)


def write(root, ext="pgm"):
    img_dir = os.path.join(root, "cam0", "data") + "/"
    os.makedirs(img_dir, exist_ok=True)
    times = []
    for i in range(N_IMAGES):
        t = T0 + i * CAM_DT
        times.append(t)
        if ext == "pgm":
            with open(os.path.join(img_dir, f"{t}.pgm"), "wb") as f:
                f.write(b"P5\n%d %d\n255\n" % (W, H))
                f.write(image(i).tobytes())
        else:
            import cv2
            cv2.imwrite(os.path.join(img_dir, f"{t}.png"), image(i))
    rng = np.random.default_rng(77)
    # IMU starts 3 camera periods + 2 samples after the first image, ground truth 1 period + 1 sample after it
    imu_t = T0 + 3 * CAM_DT + 2 * IMU_DT + np.arange(0, 230) * IMU_DT
    gt_t = T0 + 1 * CAM_DT + 1 * IMU_DT + np.arange(0, 200) * IMU_DT
    imu_csv = os.path.join(root, "imu0.csv")
    with open(imu_csv, "w") as f:
        f.write("#timestamp [ns],w_RS_S_x [rad s^-1],w_RS_S_y,w_RS_S_z,a_RS_S_x [m s^-2],a_RS_S_y,a_RS_S_z\n")
        for t in imu_t:
            v = rng.normal(0, 1, 6) * [0.1, 0.1, 0.1, 0.5, 0.5, 0.5] + [0, 0, 0, 0, 0, 9.68]
            f.write(str(int(t)) + "," + ",".join(repr(float(x)) for x in v) + "\n")
    gt_csv = os.path.join(root, "gt.csv")
    with open(gt_csv, "w") as f:
        f.write("#timestamp,p_x,p_y,p_z,q_w,q_x,q_y,q_z,v_x,v_y,v_z,bw_x,bw_y,bw_z,ba_x,ba_y,ba_z\n")
        for k, t in enumerate(gt_t):
            p = np.array([0.01 * k, 0.5 * np.sin(0.03 * k), 1.0 + 0.002 * k])
            ang = 0.01 * k
            q = np.array([np.cos(ang / 2), 0.0, np.sin(ang / 2) * 0.6, np.sin(ang / 2) * 0.8])
            v = np.array([0.2, 0.3 * np.cos(0.03 * k), 0.04])
            b = rng.normal(0, 1e-3, 6)
            f.write(str(int(t)) + "," + ",".join(repr(float(x)) for x in np.concatenate([p, q, v, b])) + "\n")
    return img_dir, imu_csv, gt_csv, times
