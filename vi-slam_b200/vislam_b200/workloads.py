"""The five BASELINE.json configs as concrete synthetic inputs.  Input generation only — nothing here is
timed, and nothing here loads the CUDA library (the reference arm of bench.py imports this module too).

    configs[0]  single EuRoC-shaped 752x480 pair, 1000 ORB features            synth.make_pair(seed 1001)
    configs[1]  EuRoC MH-like sequence, 2000 frames, 1000 ORB features/frame   sequence(CFG1)
    configs[2]  TUM-shaped 640x480 sequence, 64-d float descriptors (L2)       sequence(CFG2)
    configs[3]  KITTI-shaped 1241x376 sequence, 5000 ORB features, 5 levels    sequence(CFG3)
    configs[4]  8192 independent 752x480 pairs, 5000 ORB features each         pairs_on_device(CFG4) — generated on the GPU
"""
import math
import time

import numpy as np

from . import synth

CFG0 = dict(name="configs[0]", w=752, h=480, n_feat=1000, K=synth.EUROC_K, n_cells=49, first_lvl=3, seed=1001,
            what="single synthetic EuRoC-shaped 752x480 frame pair, 1000 ORB features, kNN k=2 ratio 0.8 + 4-level GN")
CFG1 = dict(name="configs[1]", w=752, h=480, n_feat=1000, K=synth.EUROC_K, n_cells=49, first_lvl=3, seed=2001, frames=2000,
            desc="orb", norm=1, desc_bytes=32,
            what="synthetic EuRoC MH-like sequence 752x480, 2000 frames, 1000 ORB features/frame, 200 Hz IMU prior, "
                 "num_cells=49, GN levels 3->0")
CFG2 = dict(name="configs[2]", w=640, h=480, n_feat=1000, K=synth.TUM_K, n_cells=49, first_lvl=3, seed=3001, frames=500,
            desc="float", norm=0, desc_bytes=256,
            what="synthetic TUM-shaped 640x480 sequence, 500 frames, 1000 float (SURF-like 64-d) descriptors/frame, L2 kNN "
                 "on the tensor cores (tf32 split GEMM + exact re-check), num_cells=49, GN levels 3->0")
CFG3 = dict(name="configs[3]", w=1241, h=376, n_feat=5000, K=synth.KITTI_K, n_cells=225, first_lvl=4, seed=4001, frames=200,
            desc="orb", norm=1, desc_bytes=32,
            what="synthetic KITTI-shaped 1241x376 sequence, 200 frames, 5000 ORB features/frame, 5-level pyramid (GN levels "
                 "4->0), num_cells=225 (200-feature cap)")
CFG4 = dict(name="configs[4]", w=752, h=480, n_feat=5000, K=synth.EUROC_K, n_cells=225, first_lvl=3, seed=5000, pairs=8192,
            desc="orb", norm=1, desc_bytes=32,
            what="8192 independent synthetic 752x480 frame pairs, 5000 ORB features each (seeds 5000+i), num_cells=225, "
                 "sharded across the ranks in contiguous blocks")


def sequence(cfg, initial_pose, n_frames=None, seed=None, device=None, log=None):
    """A sequence config as numpy arrays (frames rendered with torch on `device` when given): synth.make_sequence plus the GN
    prior of every pair formed exactly as VISystem.cpp:1135-1168 does, by `initial_pose(imu2cam 3x3, R_imu_res 3x3, t_res 3)
    -> pose[7]` (the product's host helper in the GPU arm, the oracle's in the reference arm — bit-identical)."""
    t0 = time.time()
    n = n_frames or cfg["frames"]
    seq = synth.make_sequence(n, w=cfg["w"], h=cfg["h"], n_feat=cfg["n_feat"], K=cfg["K"],
                              seed=cfg["seed"] if seed is None else seed, device=device, desc=cfg["desc"])
    eye = np.eye(3, dtype=np.float32)
    seq["prior"] = np.stack([np.asarray(initial_pose(eye, seq["R_imu_res"][k], seq["t_res"][k]), np.float32)
                             for k in range(n - 1)])
    if log:
        log(f"[bench] {cfg['name']}: {n} frames {cfg['w']}x{cfg['h']}, {cfg['n_feat']} {cfg['desc']} features in {time.time() - t0:.1f}s")
    return seq


def _render_batch(scene, G, device, chunk=32):
    """Scene.render for a batch of world->camera transforms G [B,4,4] (float64 numpy), in float32 on the device, `chunk` frames
    per pass.  Returns a uint8 torch tensor [B,h,w] on the device."""
    import torch
    dev = torch.device(device)
    fx, fy, cx, cy = scene.K
    tex = torch.from_numpy(scene.tex.astype(np.float32)).to(dev)
    th, tw = tex.shape
    v, u = torch.meshgrid(torch.arange(scene.h, device=dev, dtype=torch.float32),
                          torch.arange(scene.w, device=dev, dtype=torch.float32), indexing="ij")
    d = torch.stack([(u - cx) / fx, (v - cy) / fy, torch.ones_like(u)], -1)              # rays, [h,w,3]
    out = torch.empty((G.shape[0], scene.h, scene.w), dtype=torch.uint8, device=dev)
    Gt = torch.from_numpy(np.asarray(G, np.float32)).to(dev)
    for b0 in range(0, G.shape[0], chunk):
        R = Gt[b0:b0 + chunk, :3, :3]                                                     # [b,3,3]
        t = Gt[b0:b0 + chunk, :3, 3]                                                      # [b,3]
        Rt_d = torch.einsum("hwk,bkj->bhwj", d, R)                                       # R^T d per pixel
        Rt_t = torch.einsum("bkj,bk->bj", R, t)                                          # R^T t
        lam = (1.0 + Rt_t[:, 2])[:, None, None] / Rt_d[..., 2]
        P = lam[..., None] * Rt_d - Rt_t[:, None, None, :]
        tu = (P[..., 0] * fx + cx + scene.margin).clamp(0, tw - 1.001)
        tv = (P[..., 1] * fy + cy + scene.margin).clamp(0, th - 1.001)
        u0, v0 = tu.floor().long(), tv.floor().long()
        au, av = tu - u0, tv - v0
        val = (1 - av) * ((1 - au) * tex[v0, u0] + au * tex[v0, u0 + 1]) + av * ((1 - au) * tex[v0 + 1, u0] + au * tex[v0 + 1, u0 + 1])
        out[b0:b0 + chunk] = val.round().clamp(0, 255).to(torch.uint8)
    return out


def pairs_on_device(cfg, lo, hi, device, log=None):
    """Pairs [lo, hi) of a batched config, generated ON the device (torch; 8192 pairs x 5000 features would take minutes in
    numpy).  Pair i: its own small base pose and relative motion drawn from seed cfg.seed + i (the motion statistics of
    synth.make_pair), both frames rendered from one shared textured scene; descriptors set 1 uniform random bytes, set 2 =
    permuted copy with i.i.d. bit flips p = 0.05 on 70 % of the rows and fresh rows for 30 %; key points
    uniform in the image; prior = true rotation, translation + 2 mm noise.
    Returns torch tensors on the device: prev, cur [n,h,w] u8; d1, d2 [n,N,32] u8; kp1 [n,N,2] f32; prior [n,7] f32."""
    import torch
    t0 = time.time()
    n, N, w, h = hi - lo, cfg["n_feat"], cfg["w"], cfg["h"]
    dev = torch.device(device)
    scene = synth.Scene(w, h, cfg["K"], 1001)
    G0 = np.tile(np.eye(4), (n, 1, 1))
    G1 = np.tile(np.eye(4), (n, 1, 1))
    prior = np.zeros((n, 7), np.float32)
    for i in range(n):
        rng = np.random.default_rng(cfg["seed"] + lo + i)
        def small(mag):
            v = rng.uniform(-1, 1, 3)
            return v * (mag * rng.uniform(0.3, 1.0) / np.linalg.norm(v))
        G0[i, :3, :3] = synth.so3_exp(small(0.01)); G0[i, :3, 3] = small(0.02)
        T = np.eye(4)
        T[:3, :3] = synth.so3_exp(small(0.02)); T[:3, 3] = small(0.02)
        G1[i] = T @ G0[i]
        prior[i] = synth.pose7(T[:3, :3], T[:3, 3] + rng.standard_normal(3) * 0.002)
    prev = _render_batch(scene, G0, dev)
    cur = _render_batch(scene, G1, dev)
    g = torch.Generator(device=dev)
    g.manual_seed(int(cfg["seed"]) * 7919 + lo)
    d1 = torch.randint(0, 256, (n, N, 32), dtype=torch.uint8, device=dev, generator=g)
    d2 = torch.empty_like(d1)
    weights = torch.tensor([128, 64, 32, 16, 8, 4, 2, 1], dtype=torch.int32, device=dev)
    step = max(1, (1 << 26) // (N * 256))
    for b0 in range(0, n, step):
        b1 = min(n, b0 + step)
        perm = torch.rand((b1 - b0, N), device=dev, generator=g).argsort(1)
        src = torch.gather(d1[b0:b1], 1, perm[..., None].expand(-1, -1, 32))
        flips = (torch.rand((b1 - b0, N, 32, 8), device=dev, generator=g) < 0.05).to(torch.int32)
        src = src ^ (flips * weights).sum(-1).to(torch.uint8)
        fresh = torch.rand((b1 - b0, N), device=dev, generator=g) < 0.3
        rnd = torch.randint(0, 256, (b1 - b0, N, 32), dtype=torch.uint8, device=dev, generator=g)
        d2[b0:b1] = torch.where(fresh[..., None], rnd, src)
    kp1 = torch.rand((n, N, 2), device=dev, generator=g)
    kp1[..., 0] = 1 + kp1[..., 0] * (w - 3)
    kp1[..., 1] = 1 + kp1[..., 1] * (h - 3)
    if log:
        log(f"[bench] {cfg['name']}: pairs [{lo}, {hi}) x {N} features generated on the device in {time.time() - t0:.1f}s")
    return dict(prev=prev, cur=cur, d1=d1, d2=d2, kp1=kp1.contiguous(), prior=torch.from_numpy(prior).to(dev))


def sum_levels(w, h, levels):
    """Pixels of pyramid levels 0..levels-1 with cv::resize's cvRound(size * 0.5) level sizes."""
    tot = 0
    for _ in range(levels):
        tot += w * h
        w, h = _half(w), _half(h)
    return tot


def _half(v):
    k = v >> 1
    return k + (k & 1) if v & 1 else k
