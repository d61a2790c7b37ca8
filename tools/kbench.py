#!/usr/bin/env python
"""Kernel-only timing of the residual / SAO kernels on device-resident synthetic inputs.
Quick A/B tool for tuning runs (not the contract benchmark -- that is bench.py).

    python tools/kbench.py [--pics 8] [--reps 20]
"""
import argparse, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from p265_b200 import synth
from p265_b200.engine import Engine
from p265_b200.picture import ResidualBatch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pics", type=int, default=8)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--tag", default="")
    ap.add_argument("--only", default="", help="comma list of sections: residual,sao,recon,deblock")
    ap.add_argument("--quick", action="store_true", help="residual: the SF-replicated mix and the per-size runs with a table only")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    eng = Engine(0, stream.cuda_stream)

    def to_dev(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)

    def time_residual(batch, label, force_general=False):
        d_tus, d_co = to_dev(batch.tus), to_dev(batch.coeffs)
        d_sf = to_dev(batch.scaling_factor) if batch.scaling_factor is not None else None
        d_out = torch.empty(batch.geom.total_elems() * 2, dtype=torch.uint8, device=dev)
        bins = batch.bin_counts()
        rep = bool(batch.sf_replicated) and not force_general
        dense = batch.dense_small_bins() and not os.environ.get("P265_KB_NO_DENSE")
        zext = bool((batch.tus["rsvd"] >> 11).any())

        def run():
            eng.residual_dev(d_tus.data_ptr(), bins, d_co.data_ptr(), d_sf.data_ptr() if d_sf is not None else None,
                             batch.geom, d_out.data_ptr(), zero_fill=False, sf_replicated=rep, dense_arena=dense,
                             zero_extents=zext)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(args.reps):
                run()
            e1.record(stream)
        e1.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        samples = batch.samples()
        gbs = (4 * samples + 16 * len(batch.tus)) / (ms * 1e-3) / 1e9
        print("%-34s %8.4f ms  %7.1f GB/s alg (%.3f of 6555)  %6.2f Gsample/s" %
              (label, ms, gbs, gbs / 6554.9, samples / (ms * 1e-3) / 1e9), flush=True)
        return ms

    only = set(x for x in args.only.split(",") if x)

    def want(name):
        return (not only and name != "lowfreq") or name in only

    if want("residual"):
        bench_residual(args, time_residual)
    if "config2" in only:   # BASELINE config 2: 1080p 8-bit intra mix, flat lists (4x as many pictures: same working set)
        c2 = synth.residual_batch("1080p8", n_pics=4 * args.pics, n_unique=2)
        if os.environ.get("P265_KB_RASTER"):
            c2 = raster_small_bins(c2)
        c2 = c2.densified()
        time_residual(c2, "1080p8 mix, flat lists")
        if not os.environ.get("P265_KB_MIX_ONLY"):
            for l2 in (5, 4, 3, 2):
                sel = c2.tus[c2.tus["log2n"] == l2]
                time_residual(ResidualBatch(c2.geom, np.ascontiguousarray(sel), c2.coeffs, None, covers_all=True),
                              "1080p8 only %2dx%-2d (%7d TBs)" % (1 << l2, 1 << l2, len(sel)))
    if want("lowfreq"):
        bench_lowfreq(args, time_residual)
    if "zprof" in only:   # ncu target: the 32x32 / 16x16 bins without codes and with every TB promising the first quarter
        full = synth.residual_batch("4k10_lowfreq", n_pics=args.pics, n_unique=min(2, args.pics), extents=True).densified()
        for l2 in (5, 4):
            sel = full.tus[full.tus["log2n"] == l2].copy()
            for z in (0, 2):
                sel["rsvd"] = (z << 11) | (z << 13)
                time_residual(ResidualBatch(full.geom, np.ascontiguousarray(sel.copy()), full.coeffs, full.scaling_factor,
                                            covers_all=True), "zprof %2dx%-2d codes (%d,%d)" % (1 << l2, 1 << l2, z, z))
    if want("sao") or want("recon"):
        bench_sao_recon(args, eng, dev, stream, to_dev, want)
    if want("deblock"):
        bench_deblock(args, eng, dev, stream, to_dev)
    if want("overlap"):
        bench_overlap(args, eng, dev, stream, to_dev)
    if want("split"):
        bench_split(args, eng, dev, stream, to_dev)


def bench_split(args, eng, dev, stream, to_dev):
    """Big bins (32x32 + 16x16) on one stream, small bins (8x8 + 4x4) on another, concurrently
    (run with P265_GRID_PCT=50 so that both kernels' CTAs fit an SM together)."""
    full = synth.residual_batch("4k10", n_pics=args.pics, n_unique=min(2, args.pics))
    d_co, d_sf = to_dev(full.coeffs), to_dev(full.scaling_factor)
    d_out = torch.empty(full.geom.total_elems() * 2, dtype=torch.uint8, device=dev)
    stream2 = torch.cuda.Stream(device=dev)
    eng2 = Engine(0, stream2.cuda_stream)
    l2 = full.tus["log2n"]
    for label, sel_a, sel_b in (("{32,16} || {8,4}", l2 >= 4, l2 < 4), ("{32,4} || {16,8}", (l2 == 5) | (l2 == 2), (l2 == 4) | (l2 == 3))):
        ba = ResidualBatch(full.geom, np.ascontiguousarray(full.tus[sel_a]), full.coeffs, full.scaling_factor, covers_all=True)
        bb = ResidualBatch(full.geom, np.ascontiguousarray(full.tus[sel_b]), full.coeffs, full.scaling_factor, covers_all=True)
        ta, tb = to_dev(ba.tus), to_dev(bb.tus)
        bins_a, bins_b = ba.bin_counts(), bb.bin_counts()

        def run_a():
            eng.residual_dev(ta.data_ptr(), bins_a, d_co.data_ptr(), d_sf.data_ptr(), full.geom, d_out.data_ptr(),
                             zero_fill=False, sf_replicated=True)

        def run_b():
            eng2.residual_dev(tb.data_ptr(), bins_b, d_co.data_ptr(), d_sf.data_ptr(), full.geom, d_out.data_ptr(),
                              zero_fill=False, sf_replicated=True)
        for _ in range(3):
            run_a(); run_b()
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(stream)
        stream2.wait_event(e0)
        for _ in range(args.reps):
            run_a(); run_b()
        e2.record(stream2)
        stream.wait_event(e2)
        e1.record(stream)
        e1.synchronize()
        print("split residual %-20s %8.4f ms per batch (P265_GRID_PCT=%s)" % (label, e0.elapsed_time(e1) / args.reps,
                                                                              os.environ.get("P265_GRID_PCT", "100")), flush=True)


def bench_overlap(args, eng, dev, stream, to_dev):
    """Residual on one stream, SAO on another: issue-bound and bandwidth-bound work overlap."""
    batch = synth.residual_batch("4k10", n_pics=args.pics, n_unique=min(2, args.pics))
    d_tus, d_co, d_sf = to_dev(batch.tus), to_dev(batch.coeffs), to_dev(batch.scaling_factor)
    d_out = torch.empty(batch.geom.total_elems() * 2, dtype=torch.uint8, device=dev)
    bins = batch.bin_counts()
    geom, rec, params = synth.sao_batch(3840, 2160, 10, n_pics=args.pics, n_unique=min(2, args.pics))
    d_rec, d_par = to_dev(rec), to_dev(params)
    d_o = torch.empty_like(d_rec)
    for prio in (0, -1):
        stream2 = torch.cuda.Stream(device=dev, priority=prio)
        eng2 = Engine(0, stream2.cuda_stream)

        def res():
            eng.residual_dev(d_tus.data_ptr(), bins, d_co.data_ptr(), d_sf.data_ptr(), batch.geom, d_out.data_ptr(),
                             zero_fill=False, sf_replicated=True)

        def sao():
            eng2.sao_dev(d_rec.data_ptr(), d_o.data_ptr(), geom, 6, d_par.data_ptr())
        for order in ("res first", "sao first"):
            for _ in range(3):
                res(); sao()
            torch.cuda.synchronize()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(stream)
            stream2.wait_event(e0)
            for _ in range(args.reps):
                if order == "res first":
                    res(); sao()
                else:
                    sao(); res()
            e2.record(stream2)
            stream.wait_event(e2)
            e1.record(stream)
            e1.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            print("overlap residual || SAO (%s, sao prio %d)   %8.4f ms per pair" % (order, prio, ms), flush=True)


def raster_small_bins(batch):
    """Diagnostic (P265_KB_RASTER=1): the 8x8 / 4x4 TBs of a batch re-ordered by (picture, plane, y, x) -- raster order
    inside a plane instead of decoding order -- so that the 32 TBs of a work item lie side by side in a few plane rows."""
    t = batch.tus
    lim = 5 if os.environ.get("P265_KB_RASTER") == "2" else 4      # "2": the 16x16 bin as well
    key = np.where(t["log2n"] >= lim, 0, 1).astype(np.int64)
    sub = ((t["pic"].astype(np.int64) * 4 + t["c_idx"]) << 32) | (t["y"].astype(np.int64) << 16) | t["x"].astype(np.int64)
    order = np.lexsort((np.where(key == 1, sub, np.arange(len(t))), (t["flags"] & 7).astype(np.int64) * key, -t["log2n"].astype(np.int64)))
    return ResidualBatch(batch.geom, np.ascontiguousarray(t[order]), batch.coeffs, batch.scaling_factor, batch.covers_all,
                         batch.sf_replicated)


def bench_residual(args, time_residual):
    full = synth.residual_batch("4k10", n_pics=args.pics, n_unique=min(2, args.pics))
    if os.environ.get("P265_KB_RASTER"):
        full = raster_small_bins(full)
    if not os.environ.get("P265_KB_NO_DENSE"):
        full = full.densified()   # arena in descriptor order: P265_RES_DENSE_ARENA applies
    time_residual(full, "4k10 mix, SF replicated")
    if os.environ.get("P265_KB_MIX_ONLY"):
        return
    if not args.quick:
        time_residual(full, "4k10 mix, SF general", force_general=True)
        flat = ResidualBatch(full.geom, full.tus, full.coeffs, None, covers_all=True)
        time_residual(flat, "4k10 mix, flat lists")
    for l2 in (5, 4, 3, 2):
        sel = full.tus[full.tus["log2n"] == l2]
        for sf, name in ((full.scaling_factor, "SF repl"), (None, "flat")):
            if args.quick and sf is None:
                continue
            b = ResidualBatch(full.geom, np.ascontiguousarray(sel), full.coeffs, sf, covers_all=True)
            time_residual(b, "only %2dx%-2d (%7d TBs) %s" % (1 << l2, 1 << l2, len(sel), name))



def bench_lowfreq(args, time_residual):
    """Zero-aware passes: config 3's mix with confined coefficients (synth.SANITY_EXTENT_MIX), the big bins
    with and without zero-extent codes in the descriptors, and one uniform code pair at a time."""
    full = synth.residual_batch("4k10_lowfreq", n_pics=args.pics, n_unique=min(2, args.pics), extents=True).densified()
    plain = ResidualBatch(full.geom, full.tus.copy(), full.coeffs, full.scaling_factor, covers_all=True)
    plain.tus["rsvd"] = 0
    time_residual(plain, "lowfreq mix, no codes")
    time_residual(full, "lowfreq mix, codes")
    for l2 in (5, 4):
        for b, name in ((plain, "no codes"), (full, "codes")):
            sel = b.tus[b.tus["log2n"] == l2]
            time_residual(ResidualBatch(full.geom, np.ascontiguousarray(sel), full.coeffs, full.scaling_factor,
                                        covers_all=True), "only %2dx%-2d lowfreq, %s" % (1 << l2, 1 << l2, name))
        sel = full.tus[full.tus["log2n"] == l2].copy()
        for z in (1, 2):   # timing only: every TB PROMISED the same extent (results are wrong for fuller TBs)
            sel["rsvd"] = (z << 11) | (z << 13)
            time_residual(ResidualBatch(full.geom, np.ascontiguousarray(sel), full.coeffs, full.scaling_factor,
                                        covers_all=True), "only %2dx%-2d all codes (%d,%d) [timing]" % (1 << l2, 1 << l2, z, z))


def bench_sao_recon(args, eng, dev, stream, to_dev, want):
    geom, rec, params = synth.sao_batch(3840, 2160, 10, n_pics=args.pics, n_unique=min(2, args.pics))
    d_rec, d_par = to_dev(rec), to_dev(params)
    d_o = torch.empty_like(d_rec)
    for label, mod in (("SAO config 4 mix", None), ("SAO all off (copy)", 0), ("SAO all band", 1), ("SAO all edge", 2)):
        if not want("sao"):
            break
        p = params.copy()
        if mod is not None:
            p["type"][:] = mod
            if mod:
                p["offset_val"][:] = (3, 1, -1, -3)
        d_par = to_dev(p)
        for _ in range(3):
            eng.sao_dev(d_rec.data_ptr(), d_o.data_ptr(), geom, 6, d_par.data_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(args.reps):
                eng.sao_dev(d_rec.data_ptr(), d_o.data_ptr(), geom, 6, d_par.data_ptr())
            e1.record(stream)
        e1.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        b = 4 * geom.n_pics * (geom.width * geom.height * 3 // 2)
        print("%-34s %8.4f ms  %7.1f GB/s alg (%.3f of 6555)" % (label, ms, b / (ms * 1e-3) / 1e9,
                                                                 b / (ms * 1e-3) / 1e9 / 6554.9), flush=True)


    # reconstruction: 6 B/sample at 10 bits (pred 2 + residual 2 in, rec 2 out)
    d_pred = to_dev(rec)
    d_res = torch.zeros(geom.total_elems() * 2, dtype=torch.uint8, device=dev)
    d_out = torch.empty_like(d_pred)
    for _ in range(3):
        eng.reconstruct_dev(d_pred.data_ptr(), d_res.data_ptr(), d_out.data_ptr(), geom)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.reps):
            eng.reconstruct_dev(d_pred.data_ptr(), d_res.data_ptr(), d_out.data_ptr(), geom)
        e1.record(stream)
    e1.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    b = 6 * geom.n_pics * (geom.width * geom.height * 3 // 2)
    print("%-34s %8.4f ms  %7.1f GB/s alg (%.3f of 6555)" % ("reconstruct (pred + residual)", ms,
                                                             b / (ms * 1e-3) / 1e9, b / (ms * 1e-3) / 1e9 / 6554.9))



def bench_deblock(args, eng, dev, stream, to_dev):
    # deblocking: in place, 4 B/sample at 10 bits when every block is touched
    for label, dense in (("deblock, TU-grid edge map", False), ("deblock, every 8x8 edge", True)):
        g2, rec2, blk, ctb = synth.deblock_batch(3840, 2160, 10, n_pics=args.pics, n_unique=min(2, args.pics),
                                                 dense=dense)
        d_pix, d_blk, d_ctb = to_dev(rec2), to_dev(blk), to_dev(ctb)
        d_work = torch.empty_like(d_pix)
        times = []
        for _ in range(args.reps + 3):
            d_work.copy_(d_pix)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                stream.wait_stream(torch.cuda.current_stream())
                e0.record(stream)
                eng.deblock_dev(d_work.data_ptr(), g2, 6, d_blk.data_ptr(), d_ctb.data_ptr())
                e1.record(stream)
            e1.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = float(np.median(times[3:]))
        b = 4 * g2.n_pics * (g2.width * g2.height * 3 // 2)
        print("%-34s %8.4f ms  %7.1f GB/s alg (%.3f of 6555)" % (label, ms, b / (ms * 1e-3) / 1e9,
                                                                 b / (ms * 1e-3) / 1e9 / 6554.9), flush=True)


if __name__ == "__main__":
    main()
