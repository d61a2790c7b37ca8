#!/usr/bin/env python
"""What a bidirectional pinned-copy stream achieves as a function of copy size and of the number of
distinct host buffers (the end-to-end path moves 25 MB copies between ~30 different pinned buffers; the
p265_pcie_probe ceiling uses two 256 MB buffers over and over)."""
import sys, time
import torch

def run(size_mb, n_buf, total_mb=2048, both=True):
    dev = torch.device("cuda", 0)
    n = size_mb << 20
    h_in = [torch.empty(n, dtype=torch.uint8, pin_memory=True).fill_(1) for _ in range(n_buf)]
    h_out = [torch.empty(n, dtype=torch.uint8, pin_memory=True).fill_(2) for _ in range(n_buf)]
    d_in = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(2)]
    d_out = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(2)]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    reps = max(4, total_mb // size_mb)
    for warm in (True, False):
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record(s1); e[2].record(s2)
        for r in range(2 if warm else reps):
            with torch.cuda.stream(s1):
                d_in[r & 1].copy_(h_in[r % n_buf], non_blocking=True)
            if both:
                with torch.cuda.stream(s2):
                    h_out[r % n_buf].copy_(d_out[r & 1], non_blocking=True)
        e[1].record(s1); e[3].record(s2)
        torch.cuda.synchronize()
    h2d = n * reps / (e[0].elapsed_time(e[1]) * 1e-3) / 1e9
    d2h = n * reps / (e[2].elapsed_time(e[3]) * 1e-3) / 1e9 if both else 0.0
    print("copy %4d MB x %3d, %2d host buffers per direction: H2D %.1f GB/s  D2H %.1f GB/s  sum %.1f" %
          (size_mb, reps, n_buf, h2d, d2h, h2d + d2h), flush=True)

for size, nb in ((256, 1), (25, 1), (25, 8), (25, 32), (4, 1), (4, 32), (2, 64)):
    run(size, nb)
