#!/usr/bin/env python
"""vsb_track_sequence_orb on BASELINE configs[1]'s frames (ORB on the device, then match + GN): ms per 2000 frames and the
per-kernel split, for a list of ORB scratch budgets (MB).  Usage: python tools/kbench_raw.py [frames] [budget_mb ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vi-slam_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import vislam_b200 as vb
import bench
from vislam_b200 import workloads as wl

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
budgets = [int(x) for x in sys.argv[2:]] or [8192, 384]
seq = wl.sequence(wl.CFG1, bench.product_initial_pose(vb), n_frames=n, device="cuda")
ctx = vb.Context(0)
tr = ctx.tracker(752, 480, 1000, seq["K"], n_cells=49, max_pairs=n - 1)
frames, prior = torch.from_numpy(seq["frames"]).cuda(), torch.from_numpy(seq["prior"]).cuda()
ref = None
for mb in budgets:
    ctx.option("orb_scratch_mb", mb)
    tr.track_sequence_orb(frames, prior, nfeatures=1000)
    torch.cuda.synchronize()
    ctx.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        pose, ng, nf = tr.track_sequence_orb(frames, prior, nfeatures=1000)
    e1.record()
    torch.cuda.synchronize()
    prof = {k: round(v[0] / 3, 2) for k, v in ctx.profile_read().items()}
    ctx.profile(False)
    p = pose.cpu().numpy()
    same = "ref" if ref is None else ("same bits" if np.array_equal(p, ref) else "DIFFERENT")
    ref = p if ref is None else ref
    print(f"orb_scratch_mb={mb}: {e0.elapsed_time(e1) / 3:.2f} ms per {n} frames  key points/frame {float(nf.float().mean()):.0f}  {prof}  [{same}]", flush=True)
