#!/usr/bin/env python
"""BASELINE configs[4]: batched throughput over independent 752x480 frame pairs with 5000 ORB features each, on one GPU
(vsb_track_pairs, inputs resident in HBM, CUDA events).  The batch repeats 8 distinct synthetic pairs (seeds 5000+i); a pair's
result does not depend on its position in the batch (tests/test_gpu_tracker.py), so the timing is that of B independent pairs.
Usage: python tools/bench_config5.py [pairs=1024] [n_cells=225]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vi-slam_b200"))
import numpy as np
import torch
import vislam_b200 as vb
from vislam_b200 import synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n_cells = int(sys.argv[2]) if len(sys.argv) > 2 else 225
uniq = [synth.make_pair(n_feat=5000, seed=5000 + i) for i in range(8)]
ctx = vb.Context(0)
tr = ctx.tracker(752, 480, 5000, uniq[0]["K"], n_cells=n_cells, max_pairs=B)
idx = torch.arange(B, device="cuda") % 8
st = lambda k: torch.from_numpy(np.stack([u[k] for u in uniq])).cuda()[idx].contiguous()
args = [st(k) for k in ("prev", "cur", "d1", "d2", "kp1", "pose_prior")]
for _ in range(2):
    pose, n_good = tr.track_pairs(*args)
torch.cuda.synchronize()
tr.stats()
ctx.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
e0.record()
for _ in range(reps):
    pose, n_good = tr.track_pairs(*args)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
prof = {k: round(v[0] / reps, 3) for k, v in ctx.profile_read().items()}
s = tr.stats()
print(f"configs[4]: {B} independent pairs, 5000 features, n_cells {n_cells}: {ms:.2f} ms  {B / ms * 1e3:.0f} pairs/s  "
      f"{ms / B * 1e3:.1f} us/pair  good matches/pair {float(n_good.float().mean()):.0f}  GN points/pair "
      f"{s['point_visits'] / max(1, s['pairs']):.0f}  kernels_ms={prof}")
knn_ms = prof.get("knn2_hamming", 0.0)
if knn_ms:
    print(f"Hamming kNN at 5000 x 5000: {knn_ms / B * 1e3:.1f} us/pair, {2 * 256 * 5000 * 5000 * B / (knn_ms * 1e-3) / 1e12:.0f} TOP/s "
          f"algorithmic (one distance matrix per pair; the kernel computes it once per direction)")
