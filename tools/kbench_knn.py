#!/usr/bin/env python
"""Micro-benchmark of the Hamming kNN kernels (CUDA events): distance evaluations/s for every implementation
(vsb_ctx_option "knn_impl") at several problem sizes.  Usage: python tools/kbench_knn.py [N:batch ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vi-slam_b200"))
import numpy as np
import torch
import vislam_b200 as vb

cfgs = [a.split(":") for a in sys.argv[1:]] or [("1000", "2000"), ("5000", "80"), ("256", "8000"), ("2000", "500")]
ctx = vb.Context(0)
peak = ctx.popc_peak()
print(f"popc peak {peak/1e12:.2f} TPOPC/s")
for n, b in cfgs:
    n, b = int(n), int(b)
    g = torch.Generator(device="cuda").manual_seed(1)
    d1 = torch.randint(0, 256, (b, n, 32), dtype=torch.uint8, device="cuda", generator=g)
    d2 = torch.randint(0, 256, (b, n, 32), dtype=torch.uint8, device="cuda", generator=g)
    ref = None
    for impl in (0, 1, 2, 3, 4, 5):
        ctx.option("knn_impl", impl)
        for _ in range(3):
            out = ctx.knn2_hamming(d1, d2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            out = ctx.knn2_hamming(d1, d2)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        if ref is None:
            ref = [o.clone() for o in out]
        same = all(torch.equal(a, c) for a, c in zip(out, ref))
        pairs = b * n * n
        print(f"N={n} batch={b} impl={impl}: {ms:.3f} ms  {pairs/ms/1e6:.1f} Gdist/s  "
              f"({8*pairs/ms/1e3/peak*1e0:.3f} of POPC roofline)  {ms*1e3/b:.2f} us/pair  same_as_impl0={same}")
ctx.option("knn_impl", 6)
