import sys, ctypes as C
sys.path.insert(0, "vi-slam_b200"); sys.path.insert(0, ".")
import numpy as np, torch
import vislam_b200 as vb, bench
from vislam_b200 import workloads as wl
cfg = wl.CFG2
seq = wl.sequence(cfg, bench.product_initial_pose(vb), device="cuda")
T = seq["frames"].shape[0]
ctx = vb.Context(0)
d = torch.from_numpy(np.ascontiguousarray(seq["desc"])).cuda()
q, t = d[:-1].contiguous(), d[1:].contiguous()
out = ctx.knn2_l2(q, t)
torch.cuda.synchronize()
n = C.c_longlong()
vb.lib().vsb_debug_l2_fallback_rows.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
vb.lib().vsb_debug_l2_fallback_rows(ctx.handle, C.byref(n))
print("fallback rows", n.value, "of", 2 * (T - 1) * cfg["n_feat"], f"= {100.0 * n.value / (2 * (T - 1) * cfg['n_feat']):.2f} %")
