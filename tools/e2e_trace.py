#!/usr/bin/env python
"""Timeline of the pipelined end-to-end path (Engine.set_trace): which phase of which call occupies the
H2D / D2H direction when, how busy each direction is, and where it idles.  Run on the GPU box.

    python tools/e2e_trace.py [--ctx 6] [--pics 24] [--split]"""
import argparse, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from p265_b200 import synth
from p265_b200.engine import Engine
from p265_b200.picture import PackedResidualBatch


def pin(a):
    a = np.ascontiguousarray(a)
    t = torch.empty(a.nbytes, dtype=torch.uint8, pin_memory=True)
    v = t.numpy().view(a.dtype).reshape(a.shape)
    v[...] = a
    return t, v


def union_len(iv):
    iv = sorted(iv)
    tot, cur_s, cur_e = 0.0, None, None
    for s, e in iv:
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                tot += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    if cur_e is not None:
        tot += cur_e - cur_s
    return tot


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ctx", type=int, default=6)
    ap.add_argument("--pics", type=int, default=24)
    ap.add_argument("--split", action="store_true")
    ap.add_argument("--show", type=int, default=0, help="print the first N marks")
    args = ap.parse_args()
    r = synth.residual_batch("4k10", n_pics=1, seed=26610)
    g, rec, par = synth.sao_batch(3840, 2160, 10, n_pics=1, seed=27610)
    p = r.packed()
    keep = []
    k, h_tus = pin(p.tus); keep.append(k)
    k, h_st = pin(p.stream); keep.append(k)
    k, h_par = pin(par); keep.append(k)
    pb = PackedResidualBatch(r.geom, h_tus, h_st, r.scaling_factor, r.covers_all, bins=p.bins)
    bufs = []
    for i in range(args.pics):
        k1, ro = pin(np.zeros(r.geom.total_elems(), np.int16))
        k2, rc = pin(rec)
        keep += [k1, k2]
        bufs.append((ro, rc))
    engs = [Engine(0) for _ in range(args.ctx)]
    for e in engs:
        e.set_async(True)

    def step():
        for q in range(args.pics):
            ro, rc = bufs[q]
            e = engs[(2 * q) % args.ctx] if args.split else engs[q % args.ctx]
            e.residual(pb, ro)
            e = engs[(2 * q + 1) % args.ctx] if args.split else e
            e.sao(rc, g, 6, h_par, inplace=True)
        for e in engs:
            e.sync()
    step()
    for e in engs:
        e.set_trace(True)
    step()
    marks = []
    for i, e in enumerate(engs):
        m = e.trace()
        for j in range(0, len(m), 4):
            kind = m[j][0]
            t = [x[2] for x in m[j:j + 4]]
            marks.append((i, kind, t))
    t0 = min(t[0] for _, _, t in marks)
    t1 = max(t[3] for _, _, t in marks)
    h2d = [(t[0], t[1]) for _, _, t in marks]
    d2h = [(t[2], t[3]) for _, _, t in marks]
    ker = [(t[1], t[2]) for _, _, t in marks]
    span = t1 - t0
    print("pictures %d, contexts %d, split %s: span %.3f ms = %.3f ms per picture = %.0f Mpixel/s" %
          (args.pics, args.ctx, args.split, span, span / args.pics, 3840 * 2160 * args.pics / span / 1e3))
    print("  NOTE: a phase's interval runs from the previous mark on ITS stream to its own mark: it includes the time "
          "the copy waited for its engine")
    for name, iv in (("H2D phase (mark0->1)", h2d), ("kernels (mark1->2)", ker), ("D2H / write-back phase (mark2->3)", d2h)):
        print("  %-36s sum %.3f ms  union %.3f ms (%.0f %% of the span)  mean %.3f ms" %
              (name, sum(e - s for s, e in iv), union_len(iv), 100 * union_len(iv) / span, np.mean([e - s for s, e in iv])))
    for kind, nm in ((1, "residual"), (2, "SAO")):
        sel = [t for _, k2, t in marks if k2 == kind]
        a = np.array(sel)
        print("  %-9s mean phase lengths: H2D %.3f  kernels %.3f  out %.3f ms" %
              (nm, (a[:, 1] - a[:, 0]).mean(), (a[:, 2] - a[:, 1]).mean(), (a[:, 3] - a[:, 2]).mean()))
    # idle gaps of the D2H direction
    iv = sorted(d2h)
    gaps, cur = [], iv[0][1]
    for s, e in iv[1:]:
        if s > cur:
            gaps.append((cur - t0, s - cur))
        cur = max(cur, e)
    print("  D2H direction idle gaps: %d, total %.3f ms, largest %s" %
          (len(gaps), sum(gl for _, gl in gaps), ["%.3f@%.2f" % (gl, at) for at, gl in sorted(gaps, key=lambda x: -x[1])[:5]]))
    if args.show:
        for i, kind, t in sorted(marks, key=lambda m: m[2][0])[:args.show]:
            print("   ctx %2d %-8s %s" % (i, "residual" if kind == 1 else "sao", " ".join("%8.3f" % (x - t0) for x in t)))


if __name__ == "__main__":
    main()
