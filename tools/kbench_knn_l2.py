#!/usr/bin/env python
"""Micro-benchmark of the float-descriptor kNN kernels (CUDA events): knn_l2_impl 0 (exact FP64 kernel) vs 1
(tcgen05 TF32x3 distance GEMM + exact re-check).  Usage: python tools/kbench_knn_l2.py [N:batch ...]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vi-slam_b200"))
import torch
import vislam_b200 as vb

cfgs = [a.split(":") for a in sys.argv[1:]] or [("5000", "16"), ("1000", "256"), ("500", "512")]
ctx = vb.Context(0)
D = 64
for n, b in cfgs:
    n, b = int(n), int(b)
    g = torch.Generator(device="cuda").manual_seed(1)
    d1 = torch.randn((b, n, D), device="cuda", generator=g)
    d1 = d1 / d1.norm(dim=2, keepdim=True)
    d2 = d1[:, torch.randperm(n, device="cuda", generator=g)] + 0.05 * torch.randn((b, n, D), device="cuda", generator=g)
    d2 = (d2 / d2.norm(dim=2, keepdim=True)).contiguous()
    ref = None
    for impl in (0, 1):
        ctx.option("knn_l2_impl", impl)
        for _ in range(2):
            out = ctx.knn2_l2(d1, d2)
        torch.cuda.synchronize()
        reps = 5
        ctx.profile_read()
        ctx.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = ctx.knn2_l2(d1, d2)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        prof = {k: round(v[0] / reps, 3) for k, v in ctx.profile_read().items()}
        ctx.profile(False)
        if ref is None:
            ref = [o.clone() for o in out]
        same = all(torch.equal(a, c) for a, c in zip(out, ref))
        fb = ctypes.c_longlong()
        vb.lib().vsb_debug_l2_fallback_rows(ctx.handle, ctypes.byref(fb))
        flops = 2.0 * n * n * D * b
        print(f"N={n} D={D} batch={b} impl={impl}: {ms:.3f} ms  {ms * 1e3 / b:.1f} us/pair  {flops / ms / 1e9:.1f} TFLOP/s (2NMD)  "
              f"{n * n * b / ms / 1e6:.1f} Gdist/s  fallback_rows={fb.value if impl else 0}  same_as_exact={same}  kernels_ms={prof}")
ctx.option("knn_l2_impl", 1)
