#!/usr/bin/env python
"""Where does the end-to-end step spend its time?  Per-call pipeline times of the host entry points on
ONE asynchronous context (calls serialise on its stream), host issue time, and the pipelined rate over
several contexts with a single synchronisation at the end.  Run on the GPU box."""
import os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from p265_b200 import synth
from p265_b200.engine import Engine
from p265_b200.picture import PackedResidualBatch, ResidualBatch


def pin(a):
    a = np.ascontiguousarray(a)
    t = torch.empty(a.nbytes, dtype=torch.uint8, pin_memory=True)
    v = t.numpy().view(a.dtype).reshape(a.shape)
    v[...] = a
    return t, v


def main():
    n_ctx = int(os.environ.get("N_CTX", "6"))
    n_pic = int(os.environ.get("N_PIC", "48"))
    r = synth.residual_batch("4k10", n_pics=1, seed=26610)
    g, rec, par = synth.sao_batch(3840, 2160, 10, n_pics=1, seed=27610)
    p = r.packed()
    keep = []
    k, h_tus = pin(p.tus); keep.append(k)
    k, h_st = pin(p.stream); keep.append(k)
    k, d_tus = pin(r.tus); keep.append(k)
    k, d_co = pin(r.coeffs); keep.append(k)
    k, h_par = pin(par); keep.append(k)
    pb = PackedResidualBatch(r.geom, h_tus, h_st, r.scaling_factor, r.covers_all, bins=p.bins)
    db = ResidualBatch(r.geom, d_tus, d_co, r.scaling_factor, r.covers_all)
    bufs = []
    for i in range(max(n_ctx, 8)):
        k1, ro = pin(np.zeros(r.geom.total_elems(), np.int16))
        k2, rc = pin(rec)
        k3, so = pin(np.zeros_like(rec))
        keep += [k1, k2, k3]
        bufs.append((ro, rc, so))
    eng = Engine(0); eng.set_async(True)

    def run(label, fn, n=12):
        fn(0); eng.sync()
        t0 = time.perf_counter()
        for i in range(n):
            fn(i % len(bufs))
        t1 = time.perf_counter()
        eng.sync()
        t2 = time.perf_counter()
        print("%-46s issue %.3f ms/call   pipeline %.3f ms/call" % (label, (t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3), flush=True)

    run("residual packed  (H2D 5.7 MB, D2H 25 MB)", lambda i: eng.residual(pb, bufs[i][0]))
    run("residual dense   (H2D 27 MB, D2H 25 MB)", lambda i: eng.residual(db, bufs[i][0]))
    run("sao in place     (H2D 25 MB, write-back)", lambda i: eng.sao(bufs[i][1], g, 6, h_par, inplace=True))
    run("sao out of place (H2D 25 MB, D2H 25 MB)", lambda i: eng.sao(bufs[i][1], g, 6, h_par, out=bufs[i][2]))
    off = par.copy(); off["type"][:] = 0
    k, h_off = pin(off); keep.append(k)
    run("sao in place, all CTBs off (no write-back)", lambda i: eng.sao(bufs[i][1], g, 6, h_off, inplace=True))
    on = par.copy(); on["type"][:] = 1
    k, h_on = pin(on); keep.append(k)
    run("sao in place, all CTBs band (full write-back)", lambda i: eng.sao(bufs[i][1], g, 6, h_on, inplace=True))
    engs = [Engine(0) for _ in range(n_ctx)]
    for e in engs:
        e.set_async(True)
    for inplace in (True, False):
        for packed in (True, False):
            def step():
                for q in range(n_pic):
                    e = engs[q % n_ctx]
                    ro, rc, so = bufs[q % n_ctx]
                    e.residual(pb if packed else db, ro)
                    if inplace:
                        e.sao(rc, g, 6, h_par, inplace=True)
                    else:
                        e.sao(rc, g, 6, h_par, out=so)
                t1 = time.perf_counter()
                for e in engs:
                    e.sync()
                return t1
            step()
            t0 = time.perf_counter()
            t1 = step()
            t2 = time.perf_counter()
            print("pipelined %d ctx, %d pictures, one sync: packed=%s inplace=%s  issue %.3f ms/pic  total %.3f ms/pic = %.0f Mpixel/s"
                  % (n_ctx, n_pic, packed, inplace, (t1 - t0) / n_pic * 1e3, (t2 - t0) / n_pic * 1e3,
                     3840 * 2160 * n_pic / (t2 - t0) / 1e6), flush=True)


if __name__ == "__main__":
    main()
