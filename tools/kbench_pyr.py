#!/usr/bin/env python
"""Micro-benchmark of the pyramid kernels (CUDA events): GB/s of algorithmic traffic (read w*h, write every level)
for pyr_impl 0 (shared-memory tile kernel) and 1 (register-blocked).  Usage: python tools/kbench_pyr.py [frames] [w h]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vi-slam_b200"))
import torch
import vislam_b200 as vb

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
w, h = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (752, 480)
ctx = vb.Context(0)
lay = vb.pyr_layout(w, h)
img = torch.randint(0, 256, (B, h, w), dtype=torch.uint8, device="cuda")
pyr = torch.zeros((B, lay.frame_stride), dtype=torch.uint8, device="cuda")
px_all = sum(lay.w[l] * lay.h[l] for l in range(lay.levels))
for impl in (0, 1):
    ctx.option("pyr_impl", impl)
    for inplace in (False, True):
        src = None if inplace else img.data_ptr()
        nbytes = B * ((w * h + px_all - w * h) if inplace else (w * h + px_all))
        run = lambda: vb.check(vb.lib().vsb_pyramid_build(ctx.handle, src, w * h, w, B, ctypes.byref(lay), pyr.data_ptr(),
                                                          vb._stream_ptr()), ctx.handle)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"pyr_impl={impl} in_place={inplace}: {ms:.3f} ms for {B} frames  {nbytes / ms / 1e6:.0f} GB/s algorithmic")
