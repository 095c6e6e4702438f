#!/usr/bin/env python
"""Times the tracker's Gauss-Newton kernel on BASELINE configs[1] (1999 pairs of a 2000-frame 752x480 sequence) for a
matrix of knobs (vsb_ctx_option): threads per pair, tail launch, staged-level budget, kernel.  One process, sequence
rendered once.  Usage: python tools/kbench_gn.py [n_frames] [spec ...]   spec = impl:threads:tail:stage_bytes[:variant]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vi-slam_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import vislam_b200 as vb
import bench


def main():
    n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    specs = sys.argv[2:] or ["1:0:1:8192", "1:0:0:8192", "1:0:1:0", "1:0:1:32768", "1:128:0:8192", "1:256:0:8192",
                             "1:512:0:8192", "0:0:0:0"]
    n_cells = int(os.environ.get("KB_CELLS", "49"))
    from vislam_b200 import workloads as wl
    seq = wl.sequence(wl.CFG1, bench.product_initial_pose(vb), n_frames=n_frames, device="cuda")
    ctx = vb.Context(0)
    tr = ctx.tracker(wl.CFG1["w"], wl.CFG1["h"], wl.CFG1["n_feat"], seq["K"], n_cells=n_cells, max_pairs=n_frames - 1)
    dev = lambda a: a.cuda() if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a)).cuda()
    frames, desc, kp, prior = dev(seq["frames"]), dev(seq["desc"]), dev(seq["kp"]), dev(seq["prior"])
    ref = None
    for spec in specs:
        f = [int(x) for x in spec.split(":")] + [0]
        impl, threads, tail, stage, variant = f[:5]
        ctx.option("gn_variant", variant)
        ctx.option("gn_impl", impl); ctx.option("gn_threads", threads); ctx.option("gn_tail", tail)
        ctx.option("gn_stage_bytes", stage)
        for _ in range(3):
            pose, _ng = tr.track_sequence(frames, desc, kp, prior)
        torch.cuda.synchronize()
        ctx.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            pose, _ng = tr.track_sequence(frames, desc, kp, prior)
        e1.record()
        torch.cuda.synchronize()
        prof = ctx.profile_read()
        ctx.profile(False)
        p = pose.cpu().numpy()
        same = "ref" if ref is None else ("same bits" if np.array_equal(p, ref) else f"DIFF max {np.abs(p - ref).max():.2e}")
        if ref is None:
            ref = p
        gn = prof.get("gn_solve", (0, 1))
        cd = prof.get("candidates", (0, 1))
        print(f"variant={variant} impl={impl} threads={threads:4d} tail={tail} stage={stage:6d}: gn {gn[0] / reps:.3f} ms  candidates "
              f"{cd[0] / reps:.3f} ms  step {e0.elapsed_time(e1) / reps:.3f} ms  [{same}]", flush=True)
    tr.close()
    ctx.close()


if __name__ == "__main__":
    main()
