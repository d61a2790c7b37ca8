"""Shared driver of the fuzz-stream tests (CPU oracle backend and GPU backend).

The streams under tests/golden/fuzz/ were written by tests/golden/make_fuzz_streams.py (random
bins through the reference's own parser, CABAC-encoded) and decoded by libavcodec; the answers are
in tests/golden/fuzz_ffmpeg.npz.  A test parses a stream with the reference parser (py3 shim),
packs it with the product's packer and runs

    residual (backend) -> host intra prediction + reconstruction  == libavcodec, loop filters skipped
    -> deblocking (backend) -> SAO (backend)                       == libavcodec's output pictures

Two places where libavcodec 62.11 deviates from the standard are reproduced here, so that the
comparison stays bit-exact everywhere and the deviations are characterised exactly:

  * slice_loop_filter_across_slices_enabled_flag in SAO: 8.7.3 uses the flag of the LATER of the
    two slices (current sample's slice for left / upper neighbours, the neighbour's slice for
    right / lower ones); libavcodec uses the current CTB's flag for all eight directions.
  * cu_transquant_bypass + SAO on chroma: libavcodec restores the unfiltered samples of bypass
    CUs only in the top-left (CtbSize/2)^2 luma area of a CTB for the chroma planes (the
    restore loop takes the chroma width / height as luma extents).

  * per-slice deblocking override, chroma only: the tc offset of an edge segment is the current or the left CTB's by
    loop position, not that of the slice holding the q0 sample (`lav_chroma_tc_mask`; one stream, differences
    confined to the mask, luma exact).

`out_spec` (the standard's rules, what the product implements) is returned next to `out_lav`
(libavcodec's rules) so the tests can also state where the two differ.
"""
import json
import os
import sys
import tempfile

import numpy as np

from conftest import GOLDEN

sys.path.insert(0, GOLDEN)
FUZZ_DIR = os.path.join(GOLDEN, "fuzz")
COMPS = ("y", "cb", "cr")


def manifest():
    with open(os.path.join(FUZZ_DIR, "manifest.json")) as fh:
        return json.load(fh)


def answers():
    return np.load(os.path.join(GOLDEN, "fuzz_ffmpeg.npz"))


_ns = None


def parse(name, cfg):
    """(images, sps, pps) of a fuzz stream through the reference's parser."""
    global _ns
    import make_fuzz_streams as gen
    from oracle import refshim
    if not refshim.shim_available():
        if not refshim.reference_available():
            import pytest
            pytest.skip("baseline/_ref shim not present")
        refshim.build()
    if _ns is None:
        _ns = refshim.load(tempfile.mkdtemp(prefix="p265ref_"))
    gen.prepare(_ns, cfg)
    return gen.run_parser(_ns, os.path.join(FUZZ_DIR, name + ".bin"))


def lav_sao_avail(img, sps, pps=None):
    """Per-CTB neighbour masks under libavcodec's rule (current CTB's flag for every direction).
    Tile boundaries with loop_filter_across_tiles_enabled_flag = 0 follow the standard there (a
    neighbouring CTB of another tile is unusable, diagonals included)."""
    wc, hc = int(sps.pic_width_in_ctbs_y), int(sps.pic_height_in_ctbs_y)
    sl = np.zeros((hc, wc), np.int64)
    for a, ctu in img.ctus.items():
        sl[a // wc, a % wc] = int(ctu.slice_addr)
    tile = None
    if pps is not None and getattr(pps, "tiles_enabled_flag", 0) and \
            not getattr(pps, "loop_filter_across_tiles_enabled_flag", 1):
        tile = np.asarray(pps.tile_id_rs).reshape(hc, wc)
    flag = {int(h.slice_segment_address): int(getattr(h, "slice_loop_filter_across_slices_enabled_flag", 1))
            for h in img.slice_hdrs}
    out = np.zeros((hc, wc), np.uint16)
    for ry in range(hc):
        for rx in range(wc):
            m = 0
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    ny, nx = ry + dy, rx + dx
                    if 0 <= ny < hc and 0 <= nx < wc and (sl[ny, nx] == sl[ry, rx] or flag[int(sl[ry, rx])]) \
                            and (tile is None or tile[ny, nx] == tile[ry, rx]):
                        m |= 1 << ((dy + 1) * 3 + (dx + 1))
            out[ry, rx] = m
    return out


def lav_chroma_restored(nf, width, height, ctb_log2):
    """Chroma-plane mask of bypass samples libavcodec restores after SAO (see module docstring)."""
    hc, wc = height // 2, width // 2
    yc, xc = np.mgrid[0:hc, 0:wc]
    yl, xl = 2 * yc, 2 * xc
    ctb = 1 << ctb_log2
    x0, y0 = (xl >> ctb_log2) << ctb_log2, (yl >> ctb_log2) << ctb_log2
    lim_x = np.minimum(ctb // 2, (width - x0) // 2)
    lim_y = np.minimum(ctb // 2, (height - y0) // 2)
    bypass = nf[yl >> 3, xl >> 3].astype(bool)
    return bypass, bypass & (xl - x0 < lim_x) & (yl - y0 < lim_y)


def decode_picture(img, sps, pps, backend):
    """Returns (rec, out_spec, out_lav): three (Y, Cb, Cr) tuples."""
    from p265_b200 import deblock_api, intra_host, packer, sao_api, scaling_list
    from p265_b200.picture import PicGeom
    w, h = int(sps.pic_width_in_luma_samples), int(sps.pic_height_in_luma_samples)
    bd = int(sps.bit_depth_y)
    ctb_log2 = int(sps.ctb_log2_size_y)
    batch = packer.pack_pictures([img], sps, scaling_list.active_table(sps, pps))
    res = backend.residual(batch)
    rec = intra_host.reconstruct_intra_picture(img, sps, pps, [batch.geom.plane_view(res, 0, c) for c in range(3)])
    geom = PicGeom(w, h, 1, bd, int(sps.bit_depth_c))
    buf = np.zeros(geom.total_elems(), np.uint8 if bd <= 8 else np.uint16)
    for c in range(3):
        geom.plane_view(buf, 0, c)[:] = rec[c]
    blk, ctb = deblock_api.edge_map_from_picture(img, sps, pps)
    dbk = backend.deblock(buf, geom, ctb_log2, blk, ctb)
    nf = sao_api.no_filter_from_picture(img, sps)

    def sao(avail, no_filter):
        params = packer.sao_params_from_picture(img, sps, avail, pps)
        o = backend.sao(dbk, geom, ctb_log2, params, no_filter)
        return [geom.plane_view(o, 0, c).copy() for c in range(3)]

    out_spec = sao(sao_api.availability_from_picture(img, sps, pps), nf)
    out_lav = sao(lav_sao_avail(img, sps, pps), nf)
    if nf is not None:
        unfiltered = sao(lav_sao_avail(img, sps, pps), None)
        bypass, restored = lav_chroma_restored(nf, w, h, ctb_log2)
        for c in (1, 2):
            out_lav[c] = np.where(bypass & ~restored, unfiltered[c], out_lav[c])
    return rec, out_spec, out_lav


def check_stream(name, cfg, backend, want):
    """Runs every picture of a stream; returns per-picture counts of samples where the standard's
    rules and libavcodec's differ (for the assertions the tests make about them)."""
    imgs, sps, pps = parse(name, cfg)
    assert len(imgs) == cfg["pictures"]
    diffs = []
    for p, img in enumerate(imgs):
        rec, out_spec, out_lav = decode_picture(img, sps, pps, backend)
        tc_dev = [0, 0, 0]
        for c, n in enumerate(COMPS):
            assert np.array_equal(rec[c], want["%s/rec%d_%s" % (name, p, n)]), (name, "rec", p, n)
            if c and cfg.get("lav_chroma_tc_dev"):
                d = out_lav[c] != want["%s/out%d_%s" % (name, p, n)]
                assert not (d & ~lav_chroma_tc_mask(img, sps, pps)).any(), (name, "out", p, n)
                tc_dev[c] = int(d.sum())
                continue
            assert np.array_equal(out_lav[c], want["%s/out%d_%s" % (name, p, n)]), (name, "out", p, n)
        diffs.append([int((out_spec[c] != out_lav[c]).sum()) + tc_dev[c] for c in range(3)])
    return diffs


def lav_chroma_tc_mask(img, sps, pps):
    """Third libavcodec deviation (per-slice deblocking override with different slice_tc_offset_div2 in neighbouring
    CTBs): for chroma it takes the tc offset of an edge segment from the current or the left CTB by loop position
    instead of from the slice of the q0 sample.  Where it can matter: the p0 / q0 samples of the 8x8 chroma edge grid
    inside CTBs whose 3x3 CTB neighbourhood does not share one tc offset."""
    from p265_b200 import deblock_api
    wc, hc = int(sps.pic_width_in_ctbs_y), int(sps.pic_height_in_ctbs_y)
    w, h = int(sps.pic_width_in_luma_samples) // 2, int(sps.pic_height_in_luma_samples) // 2
    cs = 1 << (int(sps.ctb_log2_size_y) - 1)
    slices = deblock_api._slice_params(img, pps)
    tc = np.zeros((hc + 2, wc + 2), np.int64)
    for a, ctu in img.ctus.items():
        tc[a // wc + 1, a % wc + 1] = slices[int(ctu.slice_addr)][2]
    tc[0], tc[-1], tc[:, 0], tc[:, -1] = tc[1], tc[-2], tc[:, 1], tc[:, -2]
    tc[0, 0], tc[0, -1], tc[-1, 0], tc[-1, -1] = tc[1, 1], tc[1, -2], tc[-2, 1], tc[-2, -2]
    mixed = np.zeros((hc, wc), bool)
    for dy in range(3):
        for dx in range(3):
            mixed |= tc[dy:dy + hc, dx:dx + wc] != tc[1:-1, 1:-1]
    m = np.kron(mixed, np.ones((cs, cs), bool))[:h, :w]
    grown = m.copy()                                         # the p0 sample lies in the neighbouring CTB
    grown[1:] |= m[:-1]; grown[:-1] |= m[1:]; grown[:, 1:] |= m[:, :-1]; grown[:, :-1] |= m[:, 1:]
    y, x = np.mgrid[0:h, 0:w]
    on_edge = (x % 8 == 0) | (x % 8 == 7) | (y % 8 == 0) | (y % 8 == 7)
    return grown & on_edge


class OracleBackend:
    def __init__(self, c_oracle):
        self.co = c_oracle

    def residual(self, batch):
        return self.co.residual_batch(batch)

    def deblock(self, buf, geom, ctb_log2, blk, ctb):
        return self.co.deblock_batch(buf, geom, ctb_log2, blk, ctb)

    def sao(self, buf, geom, ctb_log2, params, nf):
        return self.co.sao_batch(buf, geom, ctb_log2, params, nf)


class GpuBackend:
    def __init__(self, engine):
        self.eng = engine

    def residual(self, batch):
        return self.eng.residual(batch)

    def deblock(self, buf, geom, ctb_log2, blk, ctb):
        return self.eng.deblock(buf, geom, ctb_log2, blk, ctb)

    def sao(self, buf, geom, ctb_log2, params, nf):
        return self.eng.sao(buf, geom, ctb_log2, params, no_filter=nf)
