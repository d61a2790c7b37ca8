#!/usr/bin/env python
"""Benchmark of the residual + SAO path at 4K 10-bit (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--pics P] [--impl reference]

A *step* is one pass of the hot path over one batch of P synthetic 4K 10-bit 4:2:0
pictures: one residual launch (config 3: TB mix skewed to 32x32, default scaling
lists, transform-skip, bypass) + one SAO launch (config 4: band + all four edge
classes, per-CTB parameters).  Metric: Mpixel/s, pixel = luma sample position of the
3840x2160 picture.

  value     device-resident inputs/outputs, CUDA events on the launching stream, max
            over ranks (N > 1: one process per GPU, pictures/streams partitioned by
            rank, no data-path collective -> weak scaling)
  e2e       same step through the public host API (Engine.residual on the packed
            coefficient stream / Engine.sao in place -> C-ABI p265_residual_batch_packed /
            p265_sao_batch) from pinned host buffers, H2D and D2H inside the timed region;
            e2e.pcie = plain pinned copies both ways at once on the same box (all ranks
            together), e2e.pcie_frac = how close the step is to that ceiling
  sustained the same device-resident step repeated for >= 2 s (clocks and power sampled)
  verify    after the timed region, pictures of the timed device buffers are compared
            with oracle/spec_oracle.c (checker only; --no-verify skips it)
  roofline  dominant kernel vs measured HBM bandwidth (MEASURED_PEAKS.json), dense
            algorithmic bytes: residual 4 B/sample + 16 B/TB, SAO
            4 B/sample at 10 bits
  cpu_baseline  the reference's own pure-Python scaling.py + transform.py (through the
            py3 shim, 1 core) on a stratified TB sample, + the oracle's SAO (the
            reference has none); N=1 only

`--impl reference` times that CPU path with all host cores.  Nothing here reads
/root/reference at run time.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

PIC_W, PIC_H = 3840, 2160
METRIC = "Mpixel/s residual+SAO at 4K 10-bit"


# ------------------------------------------------------------------------ helpers
def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML (same counters nvidia-smi
    prints, ~1 ms per query) while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()
        self.max_mhz, self.err = None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            uuid = None
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                tok = vis.split(",")[index].strip()
                if tok.startswith("GPU-"):
                    uuid = tok
                else:
                    index = int(tok)
            self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid) if uuid else pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, repr(e)

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                try:
                    mw = nv.nvmlDeviceGetPowerUsage(self.h)
                except Exception:
                    mw = 0
                self.samples.append((float(mhz), int(reasons), mw / 1000.0))
            except Exception as e:  # pragma: no cover
                self.err = repr(e)
                break
            self._stop_evt.wait(0.002)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                "sw_power_cap": 0x4}
        sm = [s[0] for s in self.samples]
        reasons = sorted(n for n, b in bits.items() if any(s[1] & b for s in self.samples))
        pw = [s[2] for s in self.samples if s[2] > 0]
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz,
               "reasons": reasons, "samples": len(sm), "power_w": round(float(np.median(pw)), 1) if pw else None,
               "how": "NVML, every 2 ms during the timed region"}
        if self.err:
            out["error"] = self.err
        return out


def host_placement(local: int):
    """Pin this rank to the CPUs next to its GPU (sysfs local_cpulist / numa_node of the PCI
    device) so that page-locked buffers are first-touched on that NUMA node.  Best effort: a
    VM that exposes one node (numa_node = -1) leaves nothing to choose."""
    info = {"numa_node": None, "cpus": None, "pinned": False}
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        devn = torch.cuda.get_device_properties(local).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0" % (dom, bus, devn)
        node = int(open(path + "/numa_node").read())
        cpus = open(path + "/local_cpulist").read().strip()
        info["numa_node"], info["cpus"] = node, cpus
        ids = set()
        for part in cpus.split(","):
            if part:
                lo, _, hi = part.partition("-")
                ids.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        if node >= 0 and ids and (ids & allowed) and (ids & allowed) != allowed:
            os.sched_setaffinity(0, ids & allowed)
            info["pinned"] = True
    except Exception as e:  # pragma: no cover
        info["error"] = repr(e)
    return info


def pinned(n_bytes):
    import torch
    return torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)


# ------------------------------------------------------------------ workload
def make_workload(n_pics: int, seed_base: int):
    from p265_b200 import synth
    # arena in descriptor order: the layout the product path produces on the device (unpack_kernel writes
    # the expanded arena TB after TB as sorted), P265_RES_DENSE_ARENA
    res = synth.residual_batch("4k10", n_pics=n_pics, seed=seed_base, n_unique=min(2, n_pics)).densified()
    geom, rec, params = synth.sao_batch(PIC_W, PIC_H, 10, n_pics=n_pics, seed=seed_base + 1000,
                                        n_unique=min(2, n_pics))
    return res, geom, rec, params


def alg_bytes(res, sao_geom):
    """Dense algorithmic bytes per launch."""
    samples = res.samples()
    r = 4 * samples + 16 * len(res.tus)
    s_samples = sao_geom.n_pics * (sao_geom.width * sao_geom.height * 3 // 2)
    return r, 4 * s_samples, samples, s_samples


def alg_int_ops(res):
    """Dense even/odd butterfly op count: per sample 20/23/28.5/39.25."""
    ops = {2: 20.0, 3: 23.0, 4: 28.5, 5: 39.25}
    l2 = res.tus["log2n"]
    return float(sum(ops[k] * int((l2 == k).sum()) * (1 << (2 * k)) for k in (2, 3, 4, 5)))


def verify_device_buffers(res, d_res, sgeom, rec, params, d_sao):
    """Compare pictures of the buffers the timed launches wrote with oracle/spec_oracle.c: the first
    two (the two distinct synthetic pictures) and the last one of the batch.  Raises on a mismatch."""
    from oracle import c_oracle
    from p265_b200.picture import PicGeom, ResidualBatch
    c_oracle.build()
    g = res.geom
    pics = sorted({0, min(1, g.n_pics - 1), g.n_pics - 1})
    one = PicGeom(g.width, g.height, 1, g.bit_depth_y, g.bit_depth_c)
    got_res = d_res.cpu().numpy().view(np.int16)
    got_sao = d_sao.cpu().numpy().view(rec.dtype)
    sg1 = PicGeom(sgeom.width, sgeom.height, 1, sgeom.bit_depth_y, sgeom.bit_depth_c)
    for p in pics:
        t = np.ascontiguousarray(res.tus[res.tus["pic"] == p])
        t["pic"] = 0
        sub = ResidualBatch(one, t, res.coeffs, res.scaling_factor, res.covers_all, res.sf_replicated)
        want = c_oracle.residual_batch(sub, zero_fill=False)
        have = got_res[p * g.pic_stride:(p + 1) * g.pic_stride]
        for c in range(3):
            if not np.array_equal(one.plane_view(have, 0, c), one.plane_view(want, 0, c)):
                raise SystemExit("bench.py --verify: residual planes of picture %d differ from the oracle" % p)
        r1 = rec[p * sgeom.pic_stride:(p + 1) * sgeom.pic_stride]
        want = c_oracle.sao_batch(r1, sg1, 6, params[p:p + 1])
        have = got_sao[p * sgeom.pic_stride:(p + 1) * sgeom.pic_stride]
        for c in range(3):
            if not np.array_equal(sg1.plane_view(have, 0, c), sg1.plane_view(want, 0, c)):
                raise SystemExit("bench.py --verify: SAO planes of picture %d differ from the oracle" % p)
    return {"pictures": pics, "residual_equal_oracle": True, "sao_equal_oracle": True,
            "oracle": "oracle/spec_oracle.c (checker; after the timed region)"}


# ------------------------------------------------------------------ GPU arm
def run_gpu(args):
    # rank 0's stdout must hold the JSON line only: everything written to fd 1 before the final print
    # (library banners included) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
    import torch
    import torch.distributed as dist
    from p265_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d does not match WORLD_SIZE %d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    # The path has no collective (north star: no NCCL): the ranks only meet for the benchmark's barrier
    # and the MAX over ranks of the timed region -- host scalars, exchanged over gloo (TCP, 127.0.0.1)
    if world > 1:
        dist.init_process_group("gloo")
    dev = torch.device("cuda", local)
    host = host_placement(local)   # CPU affinity / NUMA node of this rank's GPU, before any pinned allocation

    # stream s -> GPU s mod G: with G streams on G GPUs every rank owns
    # exactly one synthetic stream (seed 26510 + s); per-GPU work is fixed -> weak scaling
    from p265_b200 import partition
    stream_id = partition.streams_of_rank(world, rank, world)[0]
    res, sgeom, rec, params = make_workload(args.pics, 26510 + stream_id)
    stream = torch.cuda.Stream(device=dev)
    eng = Engine(local, stream.cuda_stream)

    def to_dev(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)

    d_tus, d_co = to_dev(res.tus), to_dev(res.coeffs)
    d_sf = to_dev(res.scaling_factor) if res.scaling_factor is not None else None
    d_res = torch.empty(res.geom.total_elems() * 2, dtype=torch.uint8, device=dev)
    d_rec, d_par = to_dev(rec), to_dev(params)
    d_sao = torch.empty_like(d_rec)
    bins = res.bin_counts()
    torch.cuda.synchronize()

    def residual():
        eng.residual_dev(d_tus.data_ptr(), bins, d_co.data_ptr(), d_sf.data_ptr() if d_sf is not None else None,
                         res.geom, d_res.data_ptr(), zero_fill=not res.covers_all,
                         sf_replicated=bool(res.sf_replicated), dense_arena=res.dense_small_bins())

    def sao():
        eng.sao_dev(d_rec.data_ptr(), d_sao.data_ptr(), sgeom, 6, d_par.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
        e1.synchronize()
        return e0.elapsed_time(e1)

    def step():
        residual()
        sao()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = eng.launch_count
    ms_total = timed(step, args.steps)
    launches = eng.launch_count - l0
    barrier()
    # per-kernel durations for the roofline (same stream, same buffers, > L2 working set)
    ms_res = timed(residual, args.steps) / args.steps
    ms_sao = timed(sao, args.steps) / args.steps
    clocks = sampler.stop()
    # ---- sustained leg: the same step back to back for >= args.sustain seconds (INT-bound kernels follow
    # the SM clock: this shows what the burst number is worth under a seconds-long load)
    sustained = None
    if args.sustain > 0:
        per_chunk = max(args.steps, int(0.25 / max(ms_total / args.steps * 1e-3, 1e-6)))
        sam2 = ClockSampler(local)
        sam2.start()
        t_ms, n_steps = 0.0, 0
        while t_ms < args.sustain * 1e3:
            t_ms += timed(step, per_chunk)
            n_steps += per_chunk
        c2 = sam2.stop()
        sustained = {"seconds": round(t_ms / 1e3, 3), "steps": n_steps, "ms_per_step": round(t_ms / n_steps, 4),
                     "value": round(PIC_W * PIC_H * args.pics * n_steps / (t_ms * 1e-3) / 1e6, 1),
                     "sm_mhz_median": c2["sm_mhz"], "power_w_median": c2["power_w"], "reasons": c2["reasons"],
                     "what": "this rank's device-resident step repeated back to back (events per ~0.25 s chunk, "
                             "no host synchronisation inside a chunk)"}
    # ---- verify: the buffers the timed kernels wrote, against the oracle (checker only, after the timing)
    verify = None
    if args.verify:
        verify = verify_device_buffers(res, d_res, sgeom, rec, params, d_sao)
    # the neighbouring kernels of the path (reconstruction, deblocking), rank 0 only,
    # outside the metric: same pictures, device resident, in place / out of place as they run
    other = {}
    if rank == 0 and not args.no_other:
        from p265_b200 import synth
        n_o = min(args.pics, 8)
        dg, drec, dblk, dctb = synth.deblock_batch(PIC_W, PIC_H, 10, n_pics=n_o, n_unique=min(2, n_o))
        t_pix, t_blk, t_ctb = to_dev(drec), to_dev(dblk), to_dev(dctb)
        t_work = torch.empty_like(t_pix)
        ts = []
        for _ in range(max(5, args.steps // 2) + 2):
            t_work.copy_(t_pix)                       # deblocking is in place: fresh input every time
            torch.cuda.synchronize()
            ts.append(timed(lambda: eng.deblock_dev(t_work.data_ptr(), dg, 6, t_blk.data_ptr(), t_ctb.data_ptr()), 1))
        ms = float(np.median(ts[2:]))
        b = 4 * n_o * (PIC_W * PIC_H * 3 // 2)
        other["deblock_kernel"] = {"ms": round(ms, 4), "pics": n_o, "alg_bytes": b,
                                   "frac_hbm": round(b / (ms * 1e-3) / 1e9 / measured_peaks()[0]["hbm_gbs"], 4)}
        t_res = torch.zeros(dg.total_elems() * 2, dtype=torch.uint8, device=dev)
        ms = timed(lambda: eng.reconstruct_dev(t_pix.data_ptr(), t_res.data_ptr(), t_work.data_ptr(), dg),
                   args.steps) / args.steps
        b = 6 * n_o * (PIC_W * PIC_H * 3 // 2)
        other["recon_kernel"] = {"ms": round(ms, 4), "pics": n_o, "alg_bytes": b,
                                 "frac_hbm": round(b / (ms * 1e-3) / 1e9 / measured_peaks()[0]["hbm_gbs"], 4)}
        del t_pix, t_work, t_res
        # BASELINE config 2: residual of the synthetic 1080p 8-bit intra mix, flat lists,
        # DST 4x4 -- outside the metric; 4x as many pictures as the 4K batch (same working set, > L2)
        # (arena in descriptor order, like the main workload: the layout unpack_kernel and a sorting packer produce)
        c2 = synth.residual_batch("1080p8", n_pics=4 * n_o, n_unique=2).densified()
        c_tus, c_co = to_dev(c2.tus), to_dev(c2.coeffs)
        c_out = torch.empty(c2.geom.total_elems() * 2, dtype=torch.uint8, device=dev)
        c_bins = c2.bin_counts()

        def residual_c2():
            eng.residual_dev(c_tus.data_ptr(), c_bins, c_co.data_ptr(), None, c2.geom, c_out.data_ptr(),
                             zero_fill=not c2.covers_all, sf_replicated=False, dense_arena=c2.dense_small_bins())
        for _ in range(3):
            residual_c2()
        ms = timed(residual_c2, args.steps) / args.steps
        b = 4 * c2.samples() + 16 * len(c2.tus)
        other["config2_residual_1080p8"] = {
            "ms": round(ms, 4), "pics": 4 * n_o, "tbs": int(len(c2.tus)), "alg_bytes": int(b),
            "mpixel_s": round(4 * n_o * c2.geom.width * c2.geom.height / (ms * 1e-3) / 1e6, 1),
            "frac_hbm": round(b / (ms * 1e-3) / 1e9 / measured_peaks()[0]["hbm_gbs"], 4),
            "alg_int_ops": alg_int_ops(c2)}
        del c_tus, c_co, c_out
        # Zero-aware passes (include/p265_b200.h: zero-extent codes): config 3's TB mix with the coefficients of
        # every 16x16 / 32x32 TB confined as in the reference's only real stream (synth.SANITY_EXTENT_MIX) --
        # the same launch once without codes in the descriptors and once with them.  Outside the metric: the
        # SURVEY 8(d) coefficient model of config 3 itself leaves nothing to skip.
        zb = synth.residual_batch("4k10_lowfreq", n_pics=n_o, n_unique=2, extents=True).densified()
        z_co, z_sf = to_dev(zb.coeffs), to_dev(zb.scaling_factor)
        z_out = torch.empty(zb.geom.total_elems() * 2, dtype=torch.uint8, device=dev)
        z_bins = zb.bin_counts()
        z_plain = zb.tus.copy()
        z_plain["rsvd"] = 0
        zo = {"pics": n_o, "tbs": int(len(zb.tus)), "alg_bytes": int(4 * zb.samples() + 16 * len(zb.tus))}
        for key, tus in (("ms_without_codes", z_plain), ("ms_with_codes", zb.tus)):
            z_tus = to_dev(tus)

            def residual_z():
                eng.residual_dev(z_tus.data_ptr(), z_bins, z_co.data_ptr(), z_sf.data_ptr(), zb.geom, z_out.data_ptr(),
                                 zero_fill=False, sf_replicated=bool(zb.sf_replicated), dense_arena=zb.dense_small_bins(),
                                 zero_extents=key == "ms_with_codes")
            for _ in range(3):
                residual_z()
            zo[key] = round(timed(residual_z, args.steps) / args.steps, 4)
            del z_tus
        zo["speedup"] = round(zo["ms_without_codes"] / zo["ms_with_codes"], 3)
        zo["frac_hbm_with_codes"] = round(zo["alg_bytes"] / (zo["ms_with_codes"] * 1e-3) / 1e9 / measured_peaks()[0]["hbm_gbs"], 4)
        zo["what"] = ("config-3 TB mix, coefficients of the 16x16 / 32x32 TBs confined to the (rows, columns) extents "
                      "measured on sanity.bin's 16x16 TBs: 41 % full, 29 % first quarter both ways, rest mixed")
        other["zero_aware_residual_4k10_lowfreq"] = zo
        del z_co, z_sf, z_out

    ms_total_max = partition.max_over_ranks(ms_total)
    pixels_step = PIC_W * PIC_H * args.pics * world
    value = pixels_step * args.steps / (ms_total_max * 1e-3) / 1e6

    # ---- end to end through the host API (pinned host buffers, copies timed) ----
    # A decoder hands pictures over one at a time, so every picture is its own pair of host calls:
    #   Engine.residual(packed batch)  H2D descriptors + packed coefficient stream -> unpack + residual
    #                                  kernels -> D2H residual planes
    #   Engine.sao(..., inplace=True)  H2D reconstructed planes + parameters -> SAO kernel -> the CTBs SAO
    #                                  modified are written back into the same page-locked buffer
    # The calls go round-robin to a few ASYNCHRONOUS contexts (p265_ctx_set_async: one stream each), so
    # the H2D copy of one picture overlaps the D2H copy of another; every context is synchronised before
    # the step ends.  --e2e-dense times the round-1 transport (dense arena, out-of-place SAO) instead.
    e2e_pics = max(1, args.e2e_pics)
    n_ctx = max(1, min(args.e2e_ctx, e2e_pics))
    engs = [Engine(local) for _ in range(n_ctx)]
    for e in engs:
        e.set_async(True)

    def pin(a):
        a = np.ascontiguousarray(a)
        t_ = pinned(a.nbytes)
        v = t_.numpy().view(a.dtype).reshape(a.shape)
        v[...] = a
        return t_, v

    from p265_b200.picture import PackedResidualBatch, ResidualBatch
    keep, uniq = [], []
    for u in range(min(2, e2e_pics)):           # two distinct pictures, inputs shared read-only
        r1, g1, rec1, par1 = make_workload(1, 26510 + stream_id + 100 * (u + 1))
        if args.e2e_dense:
            k_tus, h_tus = pin(r1.tus)
            k_co, h_co = pin(r1.coeffs)
            hb = ResidualBatch(r1.geom, h_tus, h_co, r1.scaling_factor, r1.covers_all)
        else:
            p1 = r1.packed()
            k_tus, h_tus = pin(p1.tus)
            k_co, h_co = pin(p1.stream)
            hb = PackedResidualBatch(r1.geom, h_tus, h_co, r1.scaling_factor, r1.covers_all, bins=p1.bins)
        k_par, h_par = pin(par1)
        keep += [k_tus, k_co, k_par]
        uniq.append((hb, g1, rec1, h_par, r1))
    items = []
    for p in range(e2e_pics):
        hb, g1, rec1, h_par, r1 = uniq[p % len(uniq)]
        k_ro = pinned(hb.geom.total_elems() * 2)
        k_rec, h_rec = pin(rec1)                # SAO works in place: every picture owns its planes
        keep += [k_ro, k_rec]
        h_so = None
        if args.e2e_dense:
            k_so = pinned(h_rec.nbytes)
            keep.append(k_so)
            h_so = k_so.numpy().view(h_rec.dtype)
        items.append((hb, g1, h_rec, h_par, k_ro.numpy().view(np.int16), h_so))

    pool = None
    if args.e2e_pool:
        # the in-process dispatcher (p265_b200/pool.py): one host thread per visible GPU, picture p -> GPU
        # p mod G; single-process runs only (under torchrun every rank already owns one GPU)
        from p265_b200.pool import EnginePool
        if world > 1:
            raise SystemExit("--e2e-pool is the single-process multi-GPU mode; do not combine it with torchrun")
        pool = EnginePool(contexts_per_device=n_ctx)

    futs = []

    def e2e_issue():
        """Queue one step: every picture's two host calls (they return once copies and kernels are queued)."""
        for p, (hb, g1, h_rec, h_par, h_ro, h_so) in enumerate(items):
            if pool is not None:
                futs.append(pool.residual(hb, h_ro, picture=p))
                futs.append(pool.sao(h_rec, g1, 6, h_par, out=h_so, inplace=h_so is None, picture=p))
                continue
            # the two calls of a picture are independent of each other (a decoder filters picture n while
            # it computes the residual of picture n+1): they go to DIFFERENT contexts, so the H2D copy of the
            # reconstructed planes is not ordered behind the D2H copy of the residual planes
            e = engs[(2 * p) % n_ctx] if args.e2e_split else engs[p % n_ctx]
            e.residual(hb, h_ro)
            e = engs[(2 * p + 1) % n_ctx] if args.e2e_split else e
            if h_so is None:
                e.sao(h_rec, g1, 6, h_par, inplace=True)
            else:
                e.sao(h_rec, g1, 6, h_par, out=h_so)

    def e2e_wait():
        """Every output of every queued step is in host memory when this returns."""
        for f in futs:
            f.result()
        futs.clear()
        for e in engs:
            e.sync()

    def e2e_step():
        e2e_issue()
        e2e_wait()

    # the pipelined outputs of a first step on pristine inputs are the synchronous call's outputs and the
    # oracle's (outside the timed region; later steps filter their own output again, which changes sample
    # values but not the work: SAO's cost does not depend on them)
    e2e_step()
    chk = Engine(local)
    hb, g1, h_rec, h_par, h_ro, h_so = items[-1]
    rec_first = uniq[(e2e_pics - 1) % len(uniq)][2]
    want_sao = chk.sao(rec_first, g1, 6, h_par)
    if not (np.array_equal(chk.residual(hb), h_ro) and np.array_equal(want_sao, h_so if h_so is not None else h_rec)):
        raise SystemExit("bench.py: asynchronous end-to-end outputs differ from the synchronous call")
    if args.verify:
        from oracle import c_oracle
        r1 = uniq[(e2e_pics - 1) % len(uniq)][4]
        if not (np.array_equal(c_oracle.residual_batch(r1, zero_fill=False), h_ro)
                and np.array_equal(c_oracle.sao_batch(rec_first, g1, 6, np.asarray(h_par)), want_sao)):
            raise SystemExit("bench.py: end-to-end outputs differ from the oracle")
        verify["e2e_outputs_equal_oracle"] = True
    del chk

    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    e2e_step()
    barrier()
    l_e2e = sum(e.launch_count for e in engs)
    # The timed region: exactly e2e_steps steps queued back to back (a picture's calls of step k+1 go to the
    # same context as in step k, so its stream orders them behind step k's use of the same buffers), then
    # every context synchronised -- all inputs copied in and all outputs read back inside the region.
    # A decoder does not drain its pipeline between pictures either.
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_issue()
    t_issue = time.perf_counter() - t0
    e2e_wait()
    dt = (time.perf_counter() - t0) / e2e_steps
    torch.cuda.synchronize()
    e2e_launches = (sum(e.launch_count for e in engs) - l_e2e) // e2e_steps
    if pool is not None:
        e2e_launches = sum(pool.launch_counts().values()) // (e2e_steps + 2)
    # secondary: every step synchronised on its own (pipeline drained between steps), median step time
    barrier()
    step_s = []
    for _ in range(max(3, e2e_steps // 3)):
        t0 = time.perf_counter()
        e2e_step()
        step_s.append(time.perf_counter() - t0)
    dt_max = partition.max_over_ranks(dt)
    dt_drained_max = partition.max_over_ranks(float(np.median(step_s)))
    e2e_value = PIC_W * PIC_H * e2e_pics * world / dt_max / 1e6

    def sao_writeback_bytes(g1, par):
        """Bytes p265_sao_batch writes back in place: the CTB components with sao type != 0."""
        total = 0
        par = np.asarray(par).reshape(-1, (g1.height + 63) // 64, (g1.width + 63) // 64)
        for c in range(3):
            h, w = g1.plane_shape(c)
            cs = 64 >> (1 if c else 0)
            ys = np.minimum(cs, h - np.arange(par.shape[1]) * cs)
            xs = np.minimum(cs, w - np.arange(par.shape[2]) * cs)
            area = ys[:, None] * xs[None, :]
            total += int((area[None] * (par["type"][..., c] != 0)).sum()) * 2
        return total

    h2d = d2h = 0
    for hb, g1, h_rec, h_par, h_ro, h_so in items:
        h2d += hb.tus.nbytes + (hb.coeffs.nbytes if args.e2e_dense else hb.stream.nbytes) + h_rec.nbytes + h_par.nbytes
        h2d += hb.scaling_factor.nbytes if hb.scaling_factor is not None else 0
        d2h += h_ro.nbytes + (h_so.nbytes if h_so is not None else sao_writeback_bytes(g1, h_par))
    # ---- the box's ceiling: plain pinned copies in both directions at once, every rank at the same time
    # twice: one 256 MB buffer per direction copied over and over (stays in the host's last-level cache: the
    # number copy benchmarks report) and 25 MB copies cycling through 16 buffers per direction (the size and
    # the cache-cold buffers of this path: every picture has its own).  pcie_frac is against the second.
    barrier()
    hot_h2d, hot_d2h = eng.pcie_probe(256 << 20, reps=8)
    barrier()
    pc_h2d, pc_d2h = eng.pcie_probe(25 << 20, reps=64, n_buffers=16)
    barrier()
    hot_h2d_sum, hot_d2h_sum = partition.sum_over_ranks(hot_h2d), partition.sum_over_ranks(hot_d2h)
    pc_h2d_min = -partition.max_over_ranks(-pc_h2d)
    pc_d2h_min = -partition.max_over_ranks(-pc_d2h)
    pc_h2d_sum, pc_d2h_sum = partition.sum_over_ranks(pc_h2d), partition.sum_over_ranks(pc_d2h)
    # time the step's own bytes would need at the probed rates (the slower direction decides)
    t_floor = max(h2d / pc_h2d, d2h / pc_d2h)
    pcie_frac = partition.max_over_ranks(t_floor) / dt_max
    pcie_frac_hot = partition.max_over_ranks(max(h2d / hot_h2d, d2h / hot_d2h)) / dt_max
    for e in engs:
        e.close()
    pool_info = None
    if pool is not None:
        pool_info = {"devices": pool.devices, "launches_per_device": pool.launch_counts(), "calls_per_device": pool.calls()}
        pool.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_kind = measured_peaks()
    b_res, b_sao, n_samples, s_samples = alg_bytes(res, sgeom)
    kernels = {"residual_kernel": (b_res, ms_res), "sao_kernel": (b_sao, ms_sao)}
    dom = max(kernels, key=lambda k: kernels[k][1])
    ach = kernels[dom][0] / (kernels[dom][1] * 1e-3) / 1e9
    # DRAM traffic of the dominant kernel from the committed ncu capture (profiles/), scaled
    # from the capture's batch size to this run's (bytes per picture x pictures per launch)
    traffic = None
    tp = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            per_pic = json.load(open(tp)).get(dom, {}).get("dram_bytes_per_picture")
            traffic = int(per_pic * args.pics) if per_pic else None
        except Exception:
            traffic = None
    int_peak = {}
    for kind, name in ((0, "imad"), (1, "iadd3"), (2, "imad+iadd3"), (3, "dp2a"), (4, "shf"), (5, "dp2a+iadd3")):
        try:
            int_peak[name] = round(eng.int_peak(kind)[0] / 1e12, 3)
        except Exception as e:  # pragma: no cover
            int_peak[name] = str(e)
    best_int = max(v for v in int_peak.values() if isinstance(v, float))
    ops = alg_int_ops(res)
    c2o = other.get("config2_residual_1080p8")
    if c2o:
        ops2 = c2o.pop("alg_int_ops")
        t2 = max(c2o["alg_bytes"] / (peaks["hbm_gbs"] * 1e9), ops2 / (best_int * 1e12))
        c2o["frac_slower_of_int32_and_hbm"] = round(t2 / (c2o["ms"] * 1e-3), 4)
    # roofline time of the residual launch = the slower of dense algorithmic
    # INT32 ops at the measured dual-issue peak and algorithmic bytes at the measured HBM copy
    # bandwidth; SAO is HBM-bound
    t_res_roof = max(b_res / (peaks["hbm_gbs"] * 1e9), ops / (best_int * 1e12))
    t_sao_roof = b_sao / (peaks["hbm_gbs"] * 1e9)
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": "Mpixel/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms_total_max / args.steps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32 (int16 data x int8 basis, dp2a)",
        "data": "synthetic",
        "config": {"workload": "4K 10-bit 4:2:0: residual (config 3: 5/10/25/60 % 4/8/16/32 luma area, default "
                               "scaling lists, TS 10 % of 4x4, bypass 1 %) + SAO (config 4) per picture",
                   "pics_per_step_per_gpu": args.pics, "tbs_per_step": int(len(res.tus)),
                   "partition": "stream s -> GPU s mod G, no collective",
                   "l2": "inputs larger than L2 (%.0f MB touched per step)" % ((b_res + b_sao) / 1e6)},
        "e2e": {"value": round(e2e_value, 1), "unit": "Mpixel/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "pics_per_step_per_gpu": e2e_pics, "steps": e2e_steps,
                "contexts": n_ctx, "gpu_launches_per_step": int(e2e_launches),
                "ms_per_step": round(dt_max * 1e3, 3),
                "transport": "dense int16 arena, out-of-place SAO (round 1)" if args.e2e_dense else
                             "packed coefficient stream (bitmap + int8/int16 levels), SAO in place with "
                             "write-back of the modified CTBs only",
                "pcie": {"h2d_gbs": round(pc_h2d / 1e9, 2), "d2h_gbs": round(pc_d2h / 1e9, 2),
                         "min_over_ranks_gbs": [round(pc_h2d_min / 1e9, 2), round(pc_d2h_min / 1e9, 2)],
                         "sum_over_ranks_gbs": [round(pc_h2d_sum / 1e9, 2), round(pc_d2h_sum / 1e9, 2)],
                         "one_hot_buffer_gbs": [round(hot_h2d / 1e9, 2), round(hot_d2h / 1e9, 2)],
                         "one_hot_buffer_sum_over_ranks_gbs": [round(hot_h2d_sum / 1e9, 2), round(hot_d2h_sum / 1e9, 2)],
                         "what": "p265_pcie_probe, page-locked H2D and D2H copies at the same time on two streams, "
                                 "all ranks at once (rank 0's rates; min / sum over ranks): 64 x 25 MB per direction "
                                 "cycling through 16 host buffers each (cache-cold, like the path's own buffers); "
                                 "one_hot_buffer = 8 x 256 MB from / to ONE buffer per direction (last-level-cache "
                                 "resident: the usual copy benchmark)"},
                "pcie_frac": round(pcie_frac, 4), "pcie_frac_one_hot_buffer": round(pcie_frac_hot, 4),
                "pcie_frac_def": "max(h2d_bytes / probed H2D rate, d2h_bytes / probed D2H rate) / step time, "
                                 "slowest rank",
                "achieved_gbs": {"h2d": round(h2d / dt / 1e9, 2), "d2h": round(d2h / dt / 1e9, 2)},
                "host": host, "pool": pool_info,
                "timing": "wall clock around %d steps queued back to back and one synchronisation of every context "
                          "(host issue %.2f ms per step); with a synchronisation after every step: median %.2f ms per "
                          "step (min %.2f, max %.2f)" % (e2e_steps, t_issue / e2e_steps * 1e3,
                                                         float(np.median(step_s)) * 1e3, min(step_s) * 1e3,
                                                         max(step_s) * 1e3),
                "drained_step_value": round(PIC_W * PIC_H * e2e_pics * world / dt_drained_max / 1e6, 1),
                "how": "one Engine.residual + one Engine.sao call per picture (C-ABI host entry points "
                       "p265_residual_batch_packed / p265_sao_batch), round-robin over asynchronous contexts "
                       "(p265_ctx_set_async), pinned host buffers, every H2D / D2H byte inside the timed "
                       "region, all contexts synchronised before the region ends; first step on pristine "
                       "inputs checked against the synchronous call" + (" and the oracle" if args.verify else "")},
        "gpu_launches": int(launches),
        "gpu_launches_per_step": {"expand_kernel": 1, "residual_kernel<bin 32/16/8/4>": 4, "sao_kernel": 1},
        "clocks": clocks,
        "sustained": sustained,
        "verify": verify,
        "ranks": {"backend": "gloo (barrier + MAX of host scalars only)" if world > 1 else None,
                  "collective_on_data_path": None},
        "other_kernels": other,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": round(ach, 1),
                     "peak": peaks["hbm_gbs"], "peak_source": peak_kind, "unit": "GB/s",
                     "frac": round(ach / peaks["hbm_gbs"], 4), "traffic": traffic,
                     "alg_bytes_per_launch": int(kernels[dom][0]),
                     "kernels": {k: {"ms": round(v[1], 4), "alg_gbs": round(v[0] / (v[1] * 1e-3) / 1e9, 1),
                                     "frac_hbm": round(v[0] / (v[1] * 1e-3) / 1e9 / peaks["hbm_gbs"], 4)}
                                 for k, v in kernels.items()},
                     "combined_frac_hbm": round((b_res + b_sao) / ((ms_res + ms_sao) * 1e-3) / 1e9
                                                / peaks["hbm_gbs"], 4),
                     "slower_of_int32_and_hbm": {
                         "definition": "SURVEY 8(d): max(alg bytes / measured HBM, dense alg int ops / measured "
                                       "INT32 dual-issue peak) / measured time",
                         "residual_bound": "int32" if ops / (best_int * 1e12) > b_res / (peaks["hbm_gbs"] * 1e9)
                         else "hbm",
                         "residual_frac": round(t_res_roof / (ms_res * 1e-3), 4),
                         "sao_frac": round(t_sao_roof / (ms_sao * 1e-3), 4),
                         "combined_frac": round((t_res_roof + t_sao_roof) / ((ms_res + ms_sao) * 1e-3), 4)},
                     "int32": {"peak_tops_measured": int_peak,
                               "residual_alg_gops_per_launch": round(ops / 1e9, 2),
                               "residual_alg_tops": round(ops / (ms_res * 1e-3) / 1e12, 2),
                               "residual_frac_of_best_int_peak": round(ops / (ms_res * 1e-3) / 1e12 / best_int, 4)}},
    }
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(single_core=True)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------ CPU arms
def _ref_worker(task):
    """Times the reference's own inverse_scaling + inverse_transform on a share of the
    stratified sample (runs in a worker process)."""
    seed, counts = task
    sys.path.insert(0, REPO)
    from oracle import refshim
    import tempfile
    ns = refshim.load(tempfile.mkdtemp(prefix="p265ref_"))
    sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
    from make_fixtures import fake_pu
    from p265_b200 import scaling_list
    sf = scaling_list.as_reference_object(scaling_list.default_scaling_factor())
    rng = np.random.default_rng(seed)
    out = {}
    for l2, cnt in counts.items():
        n = 1 << l2
        t = 0.0
        for _ in range(cnt):
            lv = (rng.laplace(0, 6, (n, n)) * (rng.random((n, n)) < 0.2)).astype(np.int64)
            pu = fake_pu(lv, 0, int(rng.integers(34, 50)), 10, sf=sf)
            t0 = time.perf_counter()
            ns.scaling.inverse_scaling(pu=pu, x0=0, y0=0, log2size=l2)
            ns.transform.inverse_transform(pu=pu, x0=0, y0=0, log2size=l2)
            t += time.perf_counter() - t0
        out[l2] = (cnt, t)
    return out


def cpu_residual_reference(cores: int):
    """Seconds per 4K picture of the reference's pure-Python residual path, extrapolated
    from a stratified TB sample by the config-3 TB counts.  Returns (sec, kind, sample)."""
    from oracle import refshim
    from p265_b200 import synth
    pic = synth.residual_batch("4k10", n_pics=1, seed=26510)
    l2 = pic.tus["log2n"]
    tb_counts = {k: int((l2 == k).sum()) for k in (2, 3, 4, 5)}
    if not refshim.shim_available():
        return None, "port", tb_counts
    per_core = {5: 24, 4: 48, 3: 96, 2: 192}
    tasks = [(100 + i, per_core) for i in range(cores)]
    t0 = time.perf_counter()
    if cores == 1:
        results = [_ref_worker(tasks[0])]
    else:
        import multiprocessing as mp
        with mp.get_context("spawn").Pool(cores) as pool:
            results = pool.map(_ref_worker, tasks)
    wall = time.perf_counter() - t0
    sec = 0.0
    for k in (2, 3, 4, 5):
        cnt = sum(r[k][0] for r in results)
        tt = sum(r[k][1] for r in results)
        sec += tb_counts[k] * (tt / cnt)
    sample = ("stratified %s TBs per size per core (32/16/8/4), %d core(s), %.1f s wall; per-TB time x "
              "config-3 TB counts %s" % (per_core, cores, wall, tb_counts))
    return sec / cores, "reference", sample


def cpu_sao_port(threads: int):
    """Seconds per 4K picture of the oracle's SAO (the reference has no SAO filter)."""
    from oracle import c_oracle
    from p265_b200 import synth
    c_oracle.build()
    c_oracle.lib().oracle_set_threads(threads)
    geom, rec, params = synth.sao_batch(PIC_W, PIC_H, 10, n_pics=1, seed=27510)
    t0 = time.perf_counter()
    c_oracle.sao_batch(rec, geom, 6, params)
    dt = time.perf_counter() - t0
    c_oracle.lib().oracle_set_threads(0)
    return dt


def cpu_residual_port(threads: int):
    from oracle import c_oracle
    from p265_b200 import synth
    c_oracle.build()
    c_oracle.lib().oracle_set_threads(threads)
    pic = synth.residual_batch("4k10", n_pics=1, seed=26510)
    t0 = time.perf_counter()
    c_oracle.residual_batch(pic, zero_fill=False)
    dt = time.perf_counter() - t0
    c_oracle.lib().oracle_set_threads(0)
    return dt


def cpu_baseline(single_core: bool):
    cores = 1 if single_core else (os.cpu_count() or 1)
    sec_res, kind, sample = cpu_residual_reference(cores)
    if sec_res is None:
        sec_res = cpu_residual_port(cores)
        sample = "oracle/spec_oracle.c on one full 4K picture (reference shim not present on this box)"
    sec_sao = cpu_sao_port(cores)
    port_all = cpu_residual_port(0) + cpu_sao_port(0)
    return {"value": round(PIC_W * PIC_H / (sec_res + sec_sao) / 1e6, 6), "unit": "Mpixel/s", "cores": cores,
            "kind": kind,
            "sample": "residual: %s; SAO: oracle/spec_oracle.c on one 4K picture (the reference has no SAO "
                      "filter)" % sample,
            "sec_per_picture": {"residual": round(sec_res, 3), "sao": round(sec_sao, 4)},
            "port_all_cores": {"value": round(PIC_W * PIC_H / port_all / 1e6, 3), "unit": "Mpixel/s",
                               "cores": os.cpu_count(), "what": "oracle/spec_oracle.c residual + SAO, one 4K "
                                                                "picture, pthreads"}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    values, last = [], None
    for i in range(max(1, min(args.steps, 3)) + min(args.warmup, 1)):
        last = cpu_baseline(single_core=False)
        if i >= min(args.warmup, 1):
            values.append(last["value"])
    v = float(np.mean(values))
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 6), "unit": "Mpixel/s",
            "n_gpus": args.gpus, "steps": len(values), "warmup": min(args.warmup, 1),
            "ms_per_step": round(PIC_W * PIC_H / (v * 1e6) * 1e3, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64 (numpy / Python int)", "data": "synthetic",
            "config": {"workload": "4K 10-bit 4:2:0 residual (config 3) + SAO (config 4), one picture per step, "
                                   "bounded stratified sample (see cpu_baseline.sample)"},
            "cpu_baseline": dict(last, value=round(v, 6), cores=cores),
            "e2e": {"value": round(v, 6), "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pics", type=int, default=16, help="4K pictures per step per GPU")
    ap.add_argument("--e2e-pics", type=int, default=8)
    ap.add_argument("--e2e-ctx", type=int, default=6)
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-other", action="store_true", help="skip the deblocking / reconstruction kernel timings")
    ap.add_argument("--sustain", type=float, default=2.0, help="seconds of the sustained leg (0 = off)")
    ap.add_argument("--no-verify", dest="verify", action="store_false",
                    help="skip the oracle comparison of the timed buffers")
    ap.add_argument("--e2e-dense", action="store_true", help="e2e leg with the round-1 transport (A/B)")
    ap.add_argument("--no-e2e-split", dest="e2e_split", action="store_false",
                    help="queue a picture's residual and SAO call on the same context (round-1 behaviour)")
    ap.add_argument("--e2e-pool", action="store_true",
                    help="e2e leg through EnginePool over every visible GPU in this one process")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
