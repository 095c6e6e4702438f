#!/usr/bin/env python
"""Micro-benchmark of ORB detect + describe on the device (CUDA events) on 752x480 frames, one level and with the scale
pyramid (the reference's ORB::create(n): 8 levels, factor 1.2), with the CPU references beside it: cv2.ORB (if importable)
and the oracle.  Frames are smoothed noise (thousands of FAST corners per frame, so every selection stage has work).
Usage: python tools/kbench_orb.py [frames] [nfeatures]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vi-slam_b200"))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import vislam_b200 as vb

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
W, H = 752, 480
rng = np.random.default_rng(1)
base = []
for s in range(8):
    f = (rng.random((H, W)) * 255).astype(np.float32)
    f = (f + np.roll(f, 1, 0) + np.roll(f, 1, 1) + np.roll(f, (1, 1), (0, 1))) / 4
    base.append(f.astype(np.uint8))
base = np.stack(base)
img = torch.from_numpy(base).cuda()[torch.arange(B, device="cuda") % 8].contiguous()
ctx = vb.Context(0)
for name, fn in (("one level", lambda: ctx.orb_detect_compute(img, nfeatures=N, cap=2 * N)),
                 ("8 levels, factor 1.2", lambda: ctx.orb_detect_compute_pyr(img, nfeatures=N, cap=2 * N))):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    ctx.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    prof = {k: round(v[0] / reps, 3) for k, v in ctx.profile_read().items()}
    ctx.profile(False)
    print(f"ORB detect+describe ({name}), {B} frames {W}x{H}, nfeatures {N}: {ms:.3f} ms  {B / ms * 1e3:.0f} frames/s  "
          f"{ms / B * 1e3:.1f} us/frame  key points/frame={float(out[-1].float().mean()):.0f}  kernels_ms={prof}")
f = base[0]
try:
    import cv2
    cv2.setNumThreads(1)
    for name, orb in (("one level", cv2.ORB_create(nfeatures=N, nlevels=1)), ("8 levels", cv2.ORB_create(nfeatures=N))):
        t0 = time.perf_counter()
        for _ in range(10):
            orb.detectAndCompute(f, None)
        print(f"cv2 ORB ({name}), 1 thread: {(time.perf_counter() - t0) / 10 * 1e3:.3f} ms/frame")
except ImportError:
    pass
from oracle import vso
t0 = time.perf_counter()
for _ in range(3):
    vso.orb_detect_compute_pyr(f, N)
print(f"oracle (plain C restatement, 8 levels), 1 thread: {(time.perf_counter() - t0) / 3 * 1e3:.3f} ms/frame")
