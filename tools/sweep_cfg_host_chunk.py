#!/usr/bin/env python
"""End-to-end time of the short-sequence configs (BASELINE configs[2], configs[3]) through vsb_track_sequence_host for several
chunk sizes of the host entry (tracker max_pairs): how deep the upload / compute pipeline should be for a 200- or 500-frame
sequence.  Usage (GPU box): python tools/sweep_cfg_host_chunk.py [2|3] [chunk ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vi-slam_b200")):
    sys.path.insert(0, p)
import bench
from vislam_b200 import workloads as wl

which = sys.argv[1] if len(sys.argv) > 1 else "3"
chunks = [int(a) for a in sys.argv[2:]] or [25, 34, 50, 67, 100, 128, 250]
cfg = {"2": wl.CFG2, "3": wl.CFG3}[which]
g = bench.Gpu(0, 1, 0)
seq = wl.sequence(cfg, bench.product_initial_pose(g.vb), device=g.dev)
for ch in chunks:
    lg = bench.leg_sequence(g, cfg, seq, 3, 2, ch)
    e2e = bench.timed_wall(g, lg["step_host"], 5, 2)
    n = lg["n_pairs"]
    print(f"configs[{which}] host chunk {ch:4d}: e2e {e2e:.3f} ms per {n} pairs = {(n + 1) / e2e:.1f} k frames/s; device-resident "
          f"{lg['ms']:.3f} ms", flush=True)
    lg["tracker"].close(); lg["tr_host"].close()
