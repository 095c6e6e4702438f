#!/usr/bin/env python
"""One BASELINE config through the tracker, a few passes, for ncu launch lists / captures.
Usage: python tools/leg_once.py <1|2|3|4> [passes=3] [frames or pairs]"""
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

which = sys.argv[1] if len(sys.argv) > 1 else "1"
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = {"1": wl.CFG1, "2": wl.CFG2, "3": wl.CFG3, "4": wl.CFG4}[which]
n = int(sys.argv[3]) if len(sys.argv) > 3 else None
ctx = vb.Context(0)
dev = torch.device("cuda", 0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if which == "4":
    n = n or 1024
    data = wl.pairs_on_device(cfg, 0, n, dev)
    tr = ctx.tracker(cfg["w"], cfg["h"], cfg["n_feat"], cfg["K"], n_cells=cfg["n_cells"], max_pairs=n)
    fn = lambda: tr.track_pairs(data["prev"], data["cur"], data["d1"], data["d2"], data["kp1"], data["prior"])
else:
    seq = wl.sequence(cfg, bench.product_initial_pose(vb), n_frames=n, device=dev)
    T = seq["frames"].shape[0]
    n = T - 1
    tr = ctx.tracker(cfg["w"], cfg["h"], cfg["n_feat"], cfg["K"], n_cells=cfg["n_cells"], max_pairs=n, norm=cfg["norm"],
                     desc_bytes=cfg["desc_bytes"], gn_opts=vb.default_gn_opts(grad_mode=1, first_lvl=cfg["first_lvl"]))
    desc = np.ascontiguousarray(seq["desc"]).view(np.uint8).reshape(T, cfg["n_feat"], cfg["desc_bytes"])
    d = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (seq["frames"], desc, seq["kp"], seq["prior"])]
    fn = lambda: tr.track_sequence(*d)
fn()
torch.cuda.synchronize()
ctx.profile(True)
e0.record()
for _ in range(passes):
    fn()
e1.record()
torch.cuda.synchronize()
prof = {k: round(v[0] / passes, 4) for k, v in ctx.profile_read().items()}
print(f"{cfg['name']}: {n} pairs, {e0.elapsed_time(e1) / passes:.3f} ms per pass, kernels_ms={prof}")
