import os, sys
sys.path.insert(0, "vi-slam_b200"); sys.path.insert(0, ".")
import numpy as np, torch
import vislam_b200 as vb
import bench
from vislam_b200 import workloads as wl
seq = wl.sequence(wl.CFG1, lambda a, b, c: np.zeros(7, np.float32), n_frames=500, device="cuda")
ctx = vb.Context(0)
img = torch.from_numpy(seq["frames"]).cuda()
for _ in range(2):
    out = ctx.orb_detect_compute_pyr(img, nfeatures=1000, cap=2000)
torch.cuda.synchronize()
print("kp/frame", float(out[-1].float().mean()))
