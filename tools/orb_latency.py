#!/usr/bin/env python
"""Latency of cv::ORB::create(1000) (8 levels) on the device for small batches, profiling OFF (CUDA events around 50 calls),
with the pyramid levels on separate streams (orb_lp 1, default) and on one stream.  Usage: python tools/orb_latency.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vi-slam_b200"))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import vislam_b200 as vb

rng = np.random.default_rng(1)
f = (rng.random((480, 752)) * 255).astype(np.float32)
f = ((f + np.roll(f, 1, 0) + np.roll(f, 1, 1) + np.roll(f, (1, 1), (0, 1))) / 4).astype(np.uint8)
ctx = vb.Context(0)
for B in (1, 2, 8, 32):
    img = torch.from_numpy(f).cuda()[None].repeat(B, 1, 1).contiguous()
    for lp in (1, 0):
        ctx.option("orb_lp", lp)
        for _ in range(5):
            ctx.orb_detect_compute_pyr(img, nfeatures=1000, cap=2000)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            ctx.orb_detect_compute_pyr(img, nfeatures=1000, cap=2000)
        e1.record()
        torch.cuda.synchronize()
        print(f"{B} frame(s), orb_lp={lp}: {e0.elapsed_time(e1) / 50 * 1e3:.0f} us per call", flush=True)
