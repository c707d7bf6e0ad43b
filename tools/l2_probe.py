#!/usr/bin/env python
"""What an L2-RESIDENT copy reaches on this GPU (torch's copy kernel, the kernel behind MEASURED_PEAKS.json's HBM number, on
buffers that fit the 126 MB L2): the ceiling to hold a small box's iteration against — 64^3 D3Q19 fp32 moves 2 x 19.9 MB per
iteration.  50 copies per CUDA-graph replay so that launch gaps do not dominate; also one copy per launch for comparison."""
import json
import sys

import torch

assert torch.cuda.is_available()
dev = torch.device("cuda:0")
# bring the clocks up first: an idle GPU runs its first milliseconds far below its boost clocks (the first version of this probe
# reported 3.5 TB/s for what is an 8.8 TB/s copy)
w = torch.ones(256 * 1000 * 1000 // 4, device=dev)
for _ in range(2000):
    w.mul_(1.0000001)
torch.cuda.synchronize()
out = {}
for mb in (5, 10, 20, 40):
    n = mb * 1000 * 1000 // 4
    a, b = torch.empty(n, device=dev), torch.ones(n, device=dev)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(5):
            a.copy_(b)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        g.capture_begin()
        for _ in range(25):
            a.copy_(b)
            b.copy_(a)
        g.capture_end()
        g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(20):
            g.replay()
        e1.record(s)
        s.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (20 * 50)
        e0.record(s)
        for _ in range(200):
            a.copy_(b)
        e1.record(s)
        s.synchronize()
        us1 = e0.elapsed_time(e1) * 1e3 / 200
    out[f"{mb}MB"] = {"us_per_copy_graph": round(us, 2), "GBps_graph": round(2 * n * 4 / us / 1e3, 1), "us_per_copy_launch": round(us1, 2),
                     "GBps_launch": round(2 * n * 4 / us1 / 1e3, 1)}
print(json.dumps(out, indent=1))
