#!/usr/bin/env python
"""Latency of the whole tracking step for small batches of independent 752x480 pairs (BASELINE configs[0]: a single
pair) through vsb_track_pairs, device-resident inputs, CUDA events; per-kernel split from the context profiler.
Usage: python tools/latency_pairs.py [n_feat] [batch ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vi-slam_b200"))
import numpy as np
import torch
import vislam_b200 as vb
from vislam_b200 import synth

n_feat = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
batches = [int(x) for x in sys.argv[2:]] or [1, 8, 64, 512]
ctx = vb.Context(0)
p = synth.make_pair(n_feat=n_feat, seed=1001)
for B in batches:
    tr = ctx.tracker(752, 480, n_feat, p["K"], n_cells=49, max_pairs=B)
    st = lambda k: torch.from_numpy(np.stack([p[k]] * B)).cuda()
    args = (st("prev"), st("cur"), st("d1"), st("d2"), st("kp1"), st("pose_prior"))
    for _ in range(5):
        tr.track_pairs(*args)
    torch.cuda.synchronize()
    ctx.profile(True)
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        tr.track_pairs(*args)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    prof = {k: round(v[0] / reps * 1e3, 1) for k, v in sorted(ctx.profile_read().items(), key=lambda kv: -kv[1][0])}
    ctx.profile(False)
    print(f"batch {B:5d} x {n_feat} features: {ms * 1e3:9.1f} us/step  {ms * 1e3 / B:8.2f} us/pair   kernels_us={prof}")
    tr.close()
