#!/usr/bin/env python
"""Micro-benchmark of FAST-9/16 detection (CUDA events) on rendered 752x480 frames, with the CPU references beside it:
cv2.FastFeatureDetector (if importable) and the oracle.  Usage: python tools/kbench_fast.py [frames]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vi-slam_b200"))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import vislam_b200 as vb
from vislam_b200 import synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
ctx = vb.Context(0)
seq = synth.make_sequence(8, n_feat=10, seed=2001)
frames = torch.from_numpy(seq["frames"]).cuda()
img = frames[torch.arange(B, device="cuda") % 8].contiguous()
for _ in range(3):
    xy, sc, n = ctx.fast_detect(img, 20, True, 4096)
torch.cuda.synchronize()
ctx.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    xy, sc, n = ctx.fast_detect(img, 20, True, 4096)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
prof = {k: round(v[0] / reps, 3) for k, v in ctx.profile_read().items()}
px = B * 752 * 480
print(f"FAST-9/16 + NMS, {B} frames 752x480: {ms:.3f} ms  {B / ms * 1e3:.0f} frames/s  {px / ms / 1e6:.1f} Gpixel/s  "
      f"{2 * px / ms / 1e6:.0f} GB/s algorithmic (1 B read + 1 B score written per pixel)  corners/frame={float(n.float().mean()):.0f}  kernels_ms={prof}")
f = seq["frames"][0]
try:
    import cv2
    det = cv2.FastFeatureDetector_create(threshold=20, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    cv2.setNumThreads(1)
    t0 = time.perf_counter()
    for _ in range(50):
        det.detect(f)
    print(f"cv2 FAST, 1 thread: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms/frame")
except ImportError:
    pass
from oracle import vso
t0 = time.perf_counter()
for _ in range(5):
    vso.fast9(f, 20, True)
print(f"oracle (plain C restatement), 1 thread: {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms/frame")
