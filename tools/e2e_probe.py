#!/usr/bin/env python
"""Where an end-to-end step spends its time: host issue loop (validation + enqueue) vs waiting for the
device (copies + kernels).  python tools/e2e_probe.py [--pics 8] [--ctx 6]"""
import argparse, os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from p265_b200 import synth
from p265_b200.engine import Engine
from p265_b200.picture import ResidualBatch

ap = argparse.ArgumentParser()
ap.add_argument("--pics", type=int, default=8)
ap.add_argument("--ctx", type=int, default=6)
ap.add_argument("--steps", type=int, default=6)
args = ap.parse_args()
engs = [Engine(0) for _ in range(args.ctx)]
for e in engs:
    e.set_async(True)


def pin(a):
    a = np.ascontiguousarray(a)
    t = torch.empty(a.nbytes, dtype=torch.uint8).pin_memory()
    v = t.numpy().view(a.dtype).reshape(a.shape)
    v[...] = a
    return t, v


keep, items = [], []
r1 = synth.residual_batch("4k10", n_pics=1, n_unique=1)
g1, rec1, par1 = synth.sao_batch(3840, 2160, 10, n_pics=1, n_unique=1)
k1, h_tus = pin(r1.tus); k2, h_co = pin(r1.coeffs); k3, h_rec = pin(rec1); k4, h_par = pin(par1)
keep += [k1, k2, k3, k4]
hb = ResidualBatch(r1.geom, h_tus, h_co, r1.scaling_factor, r1.covers_all)
for p in range(args.pics):
    a = torch.empty(hb.geom.total_elems() * 2, dtype=torch.uint8).pin_memory()
    b = torch.empty(h_rec.nbytes, dtype=torch.uint8).pin_memory()
    keep += [a, b]
    items.append((a.numpy().view(np.int16), b.numpy().view(h_rec.dtype)))
for step in range(args.steps):
    t0 = time.perf_counter()
    tr = ts = 0.0
    for p, (h_ro, h_so) in enumerate(items):
        e = engs[p % args.ctx]
        a = time.perf_counter(); e.residual(hb, h_ro); b = time.perf_counter(); e.sao(h_rec, g1, 6, h_par, out=h_so); c = time.perf_counter()
        tr += b - a; ts += c - b
    t1 = time.perf_counter()
    for e in engs:
        e.sync()
    t2 = time.perf_counter()
    print("step %d: issue %.2f ms (residual calls %.2f, sao calls %.2f), wait %.2f ms, total %.2f ms for %d pictures" %
          (step, (t1 - t0) * 1e3, tr * 1e3, ts * 1e3, (t2 - t1) * 1e3, (t2 - t0) * 1e3, args.pics), flush=True)
