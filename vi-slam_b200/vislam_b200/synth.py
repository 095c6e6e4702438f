"""Synthetic EuRoC / TUM / KITTI-shaped inputs for the frame-tracking path.

Everything here is input generation only: rendered textured plane at unit depth seen from known SE3
poses, random ORB-like 256-bit descriptors (or unit-norm float descriptors), keypoints that follow the
scene, and a 200 Hz gyro stream whose integral is the rotation prior that replaces the reference's
ROS Madgwick node (Imu.cpp:401-433).  Seeds are part of the contract.
"""
import math

import numpy as np

EUROC_K = (458.654, 457.296, 367.215, 248.375)   # calibration/calibrationEUROC.xml:20
TUM_K = (525.0, 525.0, 319.5, 239.5)             # calibration/calibrationTUM.xml:19
KITTI_K = (718.856, 718.856, 607.1928, 185.2157)  # src/main_vi_slam.1.cpp:97-101


# ------------------------------------------------------------------------------------------ SE3 helpers (float64)
def so3_exp(w):
    w = np.asarray(w, np.float64)
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + math.sin(th) / th * K + (1 - math.cos(th)) / th ** 2 * (K @ K)


def rot_to_quat_xyzw(R):
    t = np.trace(R)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        return np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    i = int(np.argmax(np.diag(R)))
    j, k = (i + 1) % 3, (i + 2) % 3
    s = math.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
    q = np.zeros(4)
    q[i] = 0.25 * s
    q[3] = (R[k, j] - R[j, k]) / s
    q[j] = (R[j, i] + R[i, j]) / s
    q[k] = (R[k, i] + R[i, k]) / s
    return q


def pose7(R, t):
    """{qx,qy,qz,qw,tx,ty,tz} float32, the layout used across the C ABI."""
    return np.concatenate([rot_to_quat_xyzw(R), t]).astype(np.float32)


# ------------------------------------------------------------------------------------------ texture + rendering
def make_texture(h, w, seed):
    """Band-limited noise: sum of bilinearly up-sampled white-noise octaves, u8."""
    rng = np.random.default_rng(seed)
    acc = np.zeros((h, w), np.float64)
    amp_sum = 0.0
    for cell, amp in ((64, 1.0), (32, 0.8), (16, 0.7), (8, 0.5), (4, 0.35)):
        gh, gw = h // cell + 2, w // cell + 2
        g = rng.random((gh, gw))
        ys = np.arange(h) / cell
        xs = np.arange(w) / cell
        y0 = ys.astype(int)
        x0 = xs.astype(int)
        fy = (ys - y0)[:, None]
        fx = (xs - x0)[None, :]
        a = g[y0][:, x0]
        b = g[y0][:, x0 + 1]
        c = g[y0 + 1][:, x0]
        d = g[y0 + 1][:, x0 + 1]
        acc += amp * ((1 - fy) * ((1 - fx) * a + fx * b) + fy * ((1 - fx) * c + fx * d))
        amp_sum += amp
    acc /= amp_sum
    acc = (acc - acc.min()) / (acc.max() - acc.min())
    return np.clip(np.rint(acc * 255), 0, 255).astype(np.uint8)


class Scene:
    """Fronto-parallel textured plane at z = 1 in the frame of camera 0."""

    def __init__(self, w, h, K=EUROC_K, seed=1001, margin=128):
        self.w, self.h, self.K, self.margin = w, h, K, margin
        self.tex = make_texture(h + 2 * margin, w + 2 * margin, seed)

    def render(self, G_list, device=None):
        """Render frames for world->camera transforms G (4x4 float64 each). Returns uint8 [T,h,w] (numpy).
        Uses torch on `device` when given (plumbing only; never timed)."""
        import torch
        dev = torch.device(device or "cpu")
        fx, fy, cx, cy = self.K
        tex = torch.from_numpy(self.tex.astype(np.float32)).to(dev)
        th, tw = tex.shape
        v, u = torch.meshgrid(torch.arange(self.h, device=dev, dtype=torch.float64),
                              torch.arange(self.w, device=dev, dtype=torch.float64), indexing="ij")
        d = torch.stack([(u - cx) / fx, (v - cy) / fy, torch.ones_like(u)], -1)  # rays in camera coords
        out = np.empty((len(G_list), self.h, self.w), np.uint8)
        for i, G in enumerate(G_list):
            R = torch.from_numpy(np.asarray(G[:3, :3], np.float64)).to(dev)
            t = torch.from_numpy(np.asarray(G[:3, 3], np.float64)).to(dev)
            Rt_d = d @ R          # (R^T d) for every pixel
            Rt_t = R.T @ t
            lam = (1.0 + Rt_t[2]) / Rt_d[..., 2]
            P = lam[..., None] * Rt_d - Rt_t   # point on the plane, world coords (z == 1)
            tu = (P[..., 0] * fx + cx + self.margin).clamp(0, tw - 1.001)
            tv = (P[..., 1] * fy + cy + self.margin).clamp(0, th - 1.001)
            u0 = tu.floor().long()
            v0 = tv.floor().long()
            au = (tu - u0).float()
            av = (tv - v0).float()
            i00 = tex[v0, u0]
            i01 = tex[v0, u0 + 1]
            i10 = tex[v0 + 1, u0]
            i11 = tex[v0 + 1, u0 + 1]
            val = (1 - av) * ((1 - au) * i00 + au * i01) + av * ((1 - au) * i10 + au * i11)
            out[i] = val.round().clamp(0, 255).to(torch.uint8).cpu().numpy()
        return out

    def project(self, G, XY):
        """Project plane points (X,Y,1) into the camera with world->camera transform G. Returns pixels [n,2]."""
        fx, fy, cx, cy = self.K
        P = np.concatenate([XY, np.ones((XY.shape[0], 1))], 1) @ G[:3, :3].T + G[:3, 3]
        return np.stack([P[:, 0] / P[:, 2] * fx + cx, P[:, 1] / P[:, 2] * fy + cy], 1)


# ------------------------------------------------------------------------------------------ descriptors
def orb_descriptors(n, seed, nbytes=32):
    return np.random.default_rng(seed).integers(0, 256, (n, nbytes), dtype=np.uint8)


def perturb_orb(base, seed, p_flip=0.05, frac_fresh=0.3, permute=True):
    """Set 2 of SURVEY §8d config 1: permuted copy, i.i.d. bit flips on 70 % of rows, fresh rows for 30 %.
    Returns (descriptors, perm) with descriptors[i] derived from base[perm[i]] (perm[i] = -1 for fresh rows)."""
    rng = np.random.default_rng(seed)
    n, nb = base.shape
    perm = rng.permutation(n) if permute else np.arange(n)
    out = base[perm].copy()
    flips = rng.random((n, nb * 8)) < p_flip
    out ^= np.packbits(flips, axis=1)
    fresh = rng.random(n) < frac_fresh
    out[fresh] = rng.integers(0, 256, (int(fresh.sum()), nb), dtype=np.uint8)
    perm = perm.copy()
    perm[fresh] = -1
    return out, perm


def float_descriptors(n, seed, dim=64):
    d = np.random.default_rng(seed).standard_normal((n, dim))
    return (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)


def perturb_float(base, seed, sigma=0.05, frac_fresh=0.3, permute=True):
    rng = np.random.default_rng(seed)
    n, dim = base.shape
    perm = rng.permutation(n) if permute else np.arange(n)
    out = base[perm].astype(np.float64) + rng.standard_normal((n, dim)) * sigma
    fresh = rng.random(n) < frac_fresh
    out[fresh] = rng.standard_normal((int(fresh.sum()), dim))
    out /= np.linalg.norm(out, axis=1, keepdims=True)
    perm = perm.copy()
    perm[fresh] = -1
    return out.astype(np.float32), perm


# ------------------------------------------------------------------------------------------ trajectories + IMU
def smooth_trajectory(n_frames, seed, frame_hz=20.0, max_rot=0.02, max_pos=0.03):
    """Camera-to-world poses C_k = (R_k, p_k) on a smooth closed curve (sum of a few sinusoids), k = 0..n-1.
    Also returns the analytic body angular velocity function for the IMU."""
    rng = np.random.default_rng(seed)
    fr = rng.uniform(0.05, 0.4, (2, 3, 3))      # Hz
    ph = rng.uniform(0, 2 * math.pi, (2, 3, 3))
    am = rng.uniform(0.3, 1.0, (2, 3, 3))
    am /= am.sum(-1, keepdims=True)

    def angles(t):  # rotation vector of the orientation (small) and position
        s = np.sin(2 * math.pi * fr * t + ph) - np.sin(ph)
        v = (am * s).sum(-1)
        return v[0] * max_rot * 4, v[1] * max_pos * 4

    ts = np.arange(n_frames) / frame_hz
    Rs, ps = [], []
    for t in ts:
        a, p = angles(t)
        Rs.append(so3_exp(a))
        ps.append(p)
    return ts, np.array(Rs), np.array(ps), angles


def world_to_cam(R, p):
    G = np.eye(4)
    G[:3, :3] = R.T
    G[:3, 3] = -R.T @ p
    return G


def gyro_samples(angles_fn, t0, t1, n, seed, sigma_g=1.7e-4):
    """n body-rate samples on [t0,t1): finite-difference log of the analytic orientation + white noise
    (sigma_g rad/s, EuRoC ADIS16448 figure).  Returns (n,3) float64."""
    rng = np.random.default_rng(seed)
    dt = (t1 - t0) / n
    out = np.zeros((n, 3))
    for i in range(n):
        Ra = so3_exp(angles_fn(t0 + i * dt)[0])
        Rb = so3_exp(angles_fn(t0 + (i + 1) * dt)[0])
        dR = Ra.T @ Rb
        w = np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]]) / 2
        out[i] = w / dt + rng.standard_normal(3) * sigma_g
    return out


def integrate_gyro(w, dt):
    """R0^T Rn = prod exp(w_i dt): the quantity Imu::estimate exposes as residual_rotationMatrix (Imu.cpp:412-415)."""
    R = np.eye(3)
    for wi in w:
        R = R @ so3_exp(wi * dt)
    return R


# ------------------------------------------------------------------------------------------ configs
def make_pair(w=752, h=480, n_feat=1000, K=EUROC_K, seed=1001, device=None, desc="orb", n_feat2=None):
    """SURVEY §8d config 1 (single frame pair).  Returns a dict of numpy arrays."""
    rng = np.random.default_rng(seed + 1)
    scene = Scene(w, h, K, seed)
    om = rng.uniform(-1, 1, 3)
    om *= 0.02 * rng.uniform(0.3, 1.0) / np.linalg.norm(om)
    tt = rng.uniform(-1, 1, 3)
    tt *= 0.02 * rng.uniform(0.3, 1.0) / np.linalg.norm(tt)
    T = np.eye(4)              # X_cur = T X_prev
    T[:3, :3] = so3_exp(om)
    T[:3, 3] = tt
    frames = scene.render([np.eye(4), T], device)
    n2 = n_feat2 if n_feat2 is not None else n_feat
    if desc == "orb":
        d1 = orb_descriptors(n_feat, seed + 2)
        d2, perm = perturb_orb(d1, seed + 3)
    else:
        d1 = float_descriptors(n_feat, seed + 2)
        d2, perm = perturb_float(d1, seed + 3)
    krng = np.random.default_rng(seed + 4)
    kp1 = np.stack([krng.uniform(1, w - 2, n_feat), krng.uniform(1, h - 2, n_feat)], 1)
    fx, fy, cx, cy = K
    XY = np.stack([(kp1[:, 0] - cx) / fx, (kp1[:, 1] - cy) / fy], 1)
    kp2_all = scene.project(T, XY)
    kp2 = np.where(perm[:, None] >= 0, kp2_all[np.maximum(perm, 0)],
                   np.stack([krng.uniform(1, w - 2, n_feat), krng.uniform(1, h - 2, n_feat)], 1))
    kp2 = np.clip(kp2, 0, [w - 1, h - 1])
    if n2 != n_feat:
        d2, kp2 = d2[:n2], kp2[:n2]
    # rotation prior = true rotation (noise-free gyro), translation prior = truth + 2 mm noise
    prior = pose7(T[:3, :3], tt + rng.standard_normal(3) * 0.002)
    return dict(w=w, h=h, K=K, prev=frames[0], cur=frames[1], d1=d1, d2=d2,
                kp1=kp1.astype(np.float32), kp2=kp2.astype(np.float32), T_true=T,
                pose_true=pose7(T[:3, :3], tt), pose_prior=prior)


def make_sequence(n_frames, w=752, h=480, n_feat=1000, K=EUROC_K, seed=2001, device=None, desc="orb",
                  frame_hz=20.0, imu_hz=200.0):
    """frames along a smooth trajectory, 10 gyro samples per frame interval, per-frame
    descriptors/keypoints of a fixed landmark set.  Returns dict with frames [T,h,w] u8, desc [T,N,D],
    kp [T,N,2] f32, gyro [T-1,10,3], R_imu_res [T-1,3,3] f32, t_res [T-1,3] f32, T_true [T-1,4,4]."""
    scene = Scene(w, h, K, seed)
    ts, Rs, ps, ang = smooth_trajectory(n_frames, seed + 1, frame_hz)
    Gs = [world_to_cam(Rs[k], ps[k]) for k in range(n_frames)]
    frames = scene.render(Gs, device)
    base = orb_descriptors(n_feat, seed + 2) if desc == "orb" else float_descriptors(n_feat, seed + 2)
    lrng = np.random.default_rng(seed + 3)
    fx, fy, cx, cy = K
    lm = np.stack([(lrng.uniform(0.08 * w, 0.92 * w, n_feat) - cx) / fx,
                   (lrng.uniform(0.08 * h, 0.92 * h, n_feat) - cy) / fy], 1)
    descs, kps = [], []
    for k in range(n_frames):
        if desc == "orb":
            dk, perm = perturb_orb(base, seed + 100 + k)
        else:
            dk, perm = perturb_float(base, seed + 100 + k)
        pk = scene.project(Gs[k], lm)
        krng = np.random.default_rng(seed + 50000 + k)
        rnd = np.stack([krng.uniform(1, w - 2, n_feat), krng.uniform(1, h - 2, n_feat)], 1)
        kk = np.where(perm[:, None] >= 0, pk[np.maximum(perm, 0)], rnd)
        descs.append(dk)
        kps.append(np.clip(kk, 0, [w - 1.001, h - 1.001]))
    n_imu = int(round(imu_hz / frame_hz))
    gyro = np.zeros((n_frames - 1, n_imu, 3))
    R_res = np.zeros((n_frames - 1, 3, 3), np.float32)
    t_res = np.zeros((n_frames - 1, 3), np.float32)
    T_true = np.zeros((n_frames - 1, 4, 4))
    trng = np.random.default_rng(seed + 4)
    for k in range(n_frames - 1):
        gyro[k] = gyro_samples(ang, ts[k], ts[k + 1], n_imu, seed + 2 + 7919 * k)
        R_res[k] = integrate_gyro(gyro[k], 1.0 / imu_hz).astype(np.float32)
        T = Gs[k + 1] @ np.linalg.inv(Gs[k])
        T_true[k] = T
        # reference: pose0.t = -TranslationResidual (GT-injected, VISystem.cpp:1150-1162)
        t_res[k] = (-(T[:3, 3]) + trng.standard_normal(3) * 0.002).astype(np.float32)
    return dict(w=w, h=h, K=K, frames=frames, desc=np.stack(descs), kp=np.stack(kps).astype(np.float32),
                gyro=gyro, R_imu_res=R_res, t_res=t_res, T_true=T_true)
