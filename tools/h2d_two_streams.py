#!/usr/bin/env python
"""Does a second copy stream raise the host-to-device rate of one GPU?  Pinned 722 MB copied (a) in one piece, (b) as two halves
on two streams at once, (c) as 8 pieces alternating over two streams.  Usage: python tools/h2d_two_streams.py"""
import torch
n = 722 * 1000 * 1000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
s = [torch.cuda.Stream(), torch.cuda.Stream()]


def run(pieces, streams):
    step = n // pieces
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for st in s[:streams]:
        st.wait_event(e0)
    for i in range(pieces):
        with torch.cuda.stream(s[i % streams]):
            d[i * step:(i + 1) * step].copy_(h[i * step:(i + 1) * step], non_blocking=True)
    for st in s[:streams]:
        torch.cuda.current_stream().wait_stream(st)
    e1.record()
    torch.cuda.synchronize()
    return n / (e0.elapsed_time(e1) * 1e-3) / 1e9


for name, p, k in (("one piece, one stream", 1, 1), ("two halves, two streams", 2, 2), ("8 pieces, two streams", 8, 2), ("8 pieces, one stream", 8, 1)):
    run(p, k)
    print(f"{name}: {max(run(p, k) for _ in range(5)):.2f} GB/s")
