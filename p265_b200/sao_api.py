"""SAO with the reference's surface: `Sao(ctx)` per-CTU parameter object (decoder/sao.py)
plus the picture-level filter the reference does not have.

`Sao.parse()` restates the sao() syntax (H.265 7.3.8.3, binarisation 9.3.3) against the
parser objects the reference hands it (`ctx.cabac`, `ctx.img`, `ctx.pps`, `ctx.sps`) and
fills the same five arrays (sao.py:43-47); it logs the same syntax-element lines, so the
reference's golden logs stay byte-identical with this module swapped in
(tests/test_dropin_golden.py).  The CABAC engine itself stays the reference's (host,
sequential).  `filter_picture()` runs 8.7.3 on the GPU through the C-ABI.
"""
from __future__ import annotations

import numpy as np

from . import packer
from .engine import get_engine
from .picture import AVAIL_ALL, PicGeom

try:                                   # the reference's logger when it is importable
    import log as _log                 # decoder/log.py (bare-name import like sao.py:2)
    _syntax = _log.syntax
except Exception:                      # stand-alone use
    import logging
    _syntax = logging.getLogger("p265_b200.sao")


class Sao:
    """Per-CTU SAO syntax (sao.py:4-136): sao_type_idx[3] (0 off, 1 band, 2 edge; Cr
    copies Cb), sao_offset_abs[3][4], sao_offset_sign[3][4], sao_band_position[3],
    sao_eo_class[3] (Cr copies Cb)."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.cabac = ctx.cabac
        self.img = ctx.img
        self.pps = ctx.pps
        self.sps = ctx.sps
        self.sao_merge_left_flag = 0
        self.sao_merge_up_flag = 0

    # -- binarisations (9.3.3.x / sao.py:138-220) -------------------------------------
    def _ctx_idx(self):
        # Tables 9-5 / 9-6 hold one context per initType; the reference always passes 0
        # (sao.py:139,144,149), which is what an I slice (initType 0) needs.
        return int(getattr(self.img.slice_hdr, "init_type", 0) or 0)

    def _merge_flag(self, name):
        bit = self.cabac.decode_decision("sao_merge_leftup_flag", self._ctx_idx())
        _syntax.info("%s = %d" % (name, bit))
        return bit

    def _type_idx(self, name):
        if self.cabac.decode_decision("sao_type_idx_lumachroma_flag", self._ctx_idx()) == 0:
            value = 0                                   # TR cMax 2: first bin context coded,
        else:                                           # second bin bypass
            value = 1 if self.cabac.decode_bypass() == 0 else 2
        _syntax.info("%s = %d" % (name, value))
        return value

    def _offset_abs(self, c_idx, i):
        bit_depth = self.sps.bit_depth_y if c_idx == 0 else self.sps.bit_depth_c
        c_max = (1 << (min(bit_depth, 10) - 5)) - 1     # 7.4.9.3.2
        value = 0
        while value < c_max and self.cabac.decode_bypass():
            value += 1
        _syntax.info("sao_offset_abs[%s][%d] = %d" % ("luma" if c_idx == 0 else "chroma", i, value))
        return value

    def _offset_sign(self, c_idx, i):
        bit = self.cabac.decode_bypass()
        _syntax.info("sao_offset_sign[%s][%d] = %d" % ("luma" if c_idx == 0 else "chroma", i, bit))
        return bit

    def _band_position(self, c_idx):
        value = 0
        for _ in range(5):                              # FL, 5 bits
            value = (value << 1) | self.cabac.decode_bypass()
        _syntax.info("sao_band_position[%s] = %d" % ("luma" if c_idx == 0 else "chroma", value))
        return value

    def _eo_class(self, name):
        value = (self.cabac.decode_bypass() << 1) | self.cabac.decode_bypass()   # FL, 2 bits
        _syntax.info("%s = %d" % (name, value))
        return value

    # -- sao() syntax (7.3.8.3) ---------------------------------------------------------
    def parse(self):
        img, pps, sps = self.img, self.pps, self.sps
        self.slice_hdr = hdr = img.slice_hdr
        ctu = img.ctu
        rx, ry, width = ctu.x_ctb, ctu.y_ctb, sps.pic_width_in_ctbs_y
        self.sao_merge_left_flag = self.sao_merge_up_flag = 0
        if rx > 0:
            in_slice = ctu.addr_rs > ctu.slice_addr
            in_tile = pps.tile_id[ctu.addr_ts] == pps.tile_id[pps.ctb_addr_rs2ts[ctu.addr_rs - 1]]
            if in_slice and in_tile:
                self.sao_merge_left_flag = self._merge_flag("sao_merge_left_flag")
        if ry > 0 and not self.sao_merge_left_flag:
            in_slice = (ctu.addr_rs - width) >= ctu.slice_addr
            in_tile = pps.tile_id[ctu.addr_ts] == pps.tile_id[pps.ctb_addr_rs2ts[ctu.addr_rs - width]]
            if in_slice and in_tile:
                self.sao_merge_up_flag = self._merge_flag("sao_merge_up_flag")

        self.sao_type_idx = np.zeros(3, int)
        self.sao_offset_abs = np.zeros((3, 4), int)
        self.sao_offset_sign = np.zeros((3, 4), int)
        self.sao_band_position = np.zeros(3, int)
        self.sao_eo_class = np.zeros(3, int)

        if self.sao_merge_left_flag or self.sao_merge_up_flag:
            src_addr = ctu.addr_rs - 1 if self.sao_merge_left_flag else ctu.addr_rs - width
            src = img.ctus[src_addr].sao               # syntax elements are inherited (7.4.9.3.2)
            self.sao_type_idx[:] = src.sao_type_idx
            self.sao_offset_abs[:] = src.sao_offset_abs
            self.sao_offset_sign[:] = src.sao_offset_sign
            self.sao_band_position[:] = src.sao_band_position
            self.sao_eo_class[:] = src.sao_eo_class
            return

        for c_idx in range(3):
            if not ((hdr.slice_sao_luma_flag and c_idx == 0) or (hdr.slice_sao_chroma_flag and c_idx > 0)):
                continue
            if c_idx == 0:
                self.sao_type_idx[0] = self._type_idx("sao_type_idx_luma")
            elif c_idx == 1:
                self.sao_type_idx[1] = self.sao_type_idx[2] = self._type_idx("sao_type_idx_chroma")
            if self.sao_type_idx[c_idx] == 0:
                continue
            for i in range(4):
                self.sao_offset_abs[c_idx][i] = self._offset_abs(c_idx, i)
            if self.sao_type_idx[c_idx] == 1:
                for i in range(4):
                    if self.sao_offset_abs[c_idx][i] != 0:
                        self.sao_offset_sign[c_idx][i] = self._offset_sign(c_idx, i)
                self.sao_band_position[c_idx] = self._band_position(c_idx)
            elif c_idx == 0:
                self.sao_eo_class[0] = self._eo_class("sao_eo_class_luma")
            elif c_idx == 1:
                self.sao_eo_class[1] = self.sao_eo_class[2] = self._eo_class("sao_eo_class_chroma")


# ------------------------------------------------------------------ picture-level filter
def availability_from_picture(img, sps, pps=None) -> np.ndarray:
    """Per-CTB neighbour masks from the parsed picture: slice addresses (ctu.slice_addr,
    slice.py:252), slice_loop_filter_across_slices_enabled_flag (slice.py:177-179) and
    tiles (pps.tile_id_rs, pps.loop_filter_across_tiles_enabled_flag, pps.py:93)."""
    wc, hc = int(sps.pic_width_in_ctbs_y), int(sps.pic_height_in_ctbs_y)
    slice_addr = np.zeros((hc, wc), dtype=np.int64)
    for addr, ctu in img.ctus.items():
        slice_addr[addr // wc, addr % wc] = int(getattr(ctu, "slice_addr", 0))
    flags = {}
    for hdr in getattr(img, "slice_hdrs", []):
        if getattr(hdr, "dependent_slice_segment_flag", 0):
            continue
        flags[int(hdr.slice_segment_address)] = int(
            getattr(hdr, "slice_loop_filter_across_slices_enabled_flag", 1))
    for a in np.unique(slice_addr):
        flags.setdefault(int(a), 1)
    tile_id = np.zeros((hc, wc), dtype=np.int64)
    across_tiles, order = True, None
    if pps is not None and getattr(pps, "tiles_enabled_flag", 0):
        tile_id = np.asarray(pps.tile_id_rs).reshape(hc, wc)
        across_tiles = bool(getattr(pps, "loop_filter_across_tiles_enabled_flag", 1))
        order = np.asarray(pps.ctb_addr_rs2ts).reshape(hc, wc)
    if len(flags) == 1 and across_tiles:
        return np.full((hc, wc), AVAIL_ALL, dtype=np.uint16)
    return packer.ctb_availability(slice_addr, flags, tile_id, across_tiles, order)


def no_filter_from_picture(img, sps):
    """8x8-luma-granular map of blocks SAO must leave untouched (8.7.3: pcm_flag with
    pcm_loop_filter_disabled_flag, cu_transquant_bypass_flag) or None when there is none."""
    w8, h8 = (int(sps.pic_width_in_luma_samples) + 7) // 8, (int(sps.pic_height_in_luma_samples) + 7) // 8
    out, any_set = np.zeros((h8, w8), dtype=np.uint8), False
    pcm_off = bool(getattr(sps, "pcm_loop_filter_disabled_flag", 0))
    for ctu in img.ctus.values():
        for cu in ctu.get_leaves():
            if not hasattr(cu, "pred_mode"):
                continue
            if getattr(cu, "cu_transquant_bypass_flag", 0) or (pcm_off and getattr(cu, "pcm_flag", 0)):
                out[cu.y >> 3:(cu.y + cu.size + 7) >> 3, cu.x >> 3:(cu.x + cu.size + 7) >> 3] = 1
                any_set = True
    return out if any_set else None


def filter_picture(planes, img, sps, pps=None, device: int = 0):
    """SAO (8.7.3) of one reconstructed (deblocked) picture.

    planes: (Y, Cb, Cr) 2-D arrays [row][col] (uint8 or uint16); img: the parsed picture
    whose CTUs carry `sao` objects; returns three new arrays (out of place)."""
    y, cb, cr = [np.asarray(p) for p in planes]
    h, w = y.shape
    if cb.shape != (h // 2, w // 2) or cr.shape != cb.shape:
        raise ValueError("planes must be 4:2:0")
    geom = PicGeom(w, h, 1, int(sps.bit_depth_y), int(sps.bit_depth_c))
    dtype = np.uint8 if max(geom.bit_depth_y, geom.bit_depth_c) <= 8 else np.uint16
    buf = np.zeros(geom.total_elems(), dtype=dtype)
    for c, p in enumerate((y, cb, cr)):
        geom.plane_view(buf, 0, c)[:] = p
    params = packer.sao_params_from_picture(img, sps, availability_from_picture(img, sps, pps), pps)
    out = get_engine(device).sao(buf, geom, int(sps.ctb_log2_size_y), params,
                                 no_filter=no_filter_from_picture(img, sps))
    return tuple(geom.plane_view(out, 0, c).copy() for c in range(3))
