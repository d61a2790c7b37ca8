"""Deblocking filter (H.265 8.7.2) -- the step between reconstruction and SAO that the
reference does not have (only its control flags are parsed,
pps.py:121-131, slice.py:170-179).

Host side: `edge_map_from_picture` walks the parsed CU / TU trees once and derives, per 8x8
luma block, the boundary strengths of the two edges that start in it (8.7.2.3 transform /
prediction block edges, 8.7.2.4 Bs), the CU's QpY and the pcm / bypass "no filter" bit;
per CTB, the slice's beta / tc offsets and the PPS chroma QP offsets.  The filter itself
(decisions 8.7.2.5.3/.6, luma 8.7.2.5.7, chroma 8.7.2.5.8) runs on the GPU
(`csrc/deblock.cu`) through the C-ABI, in place, one launch per batch of pictures.
"""
from __future__ import annotations

import numpy as np

from .engine import get_engine
from .picture import (DBK_BS_H0, DBK_BS_H1, DBK_BS_V0, DBK_BS_V1, DBK_CTB, DBK_NO_FILTER,
                      DBK_QP_SHIFT, PicGeom)

MODE_INTRA = 1     # cu.py:29


def _leaf_cus(ctu):
    for cu in ctu.get_leaves():
        if hasattr(cu, "pred_mode"):
            yield cu


def _slice_params(img, pps):
    """{SliceAddrRs: (deblocking disabled, beta_offset_div2, tc_offset_div2, across_slices)}"""
    out = {}
    pps_off = int(getattr(pps, "pps_deblocking_filter_disabled_flag", 0)) if pps is not None else 0
    pps_beta = int(getattr(pps, "pps_beta_offset_div2", 0)) if pps is not None else 0
    pps_tc = int(getattr(pps, "pps_tc_offset_div2", 0)) if pps is not None else 0
    hdrs = list(getattr(img, "slice_hdrs", [])) or [img.slice_hdr]
    for hdr in hdrs:
        if getattr(hdr, "dependent_slice_segment_flag", 0):
            continue
        out[int(getattr(hdr, "slice_segment_address", 0))] = (
            int(getattr(hdr, "slice_deblocking_filter_disabled_flag", pps_off)),
            int(getattr(hdr, "slice_beta_offset_div2", pps_beta)),
            int(getattr(hdr, "slice_tc_offset_div2", pps_tc)),
            int(getattr(hdr, "slice_loop_filter_across_slices_enabled_flag", 1)))
    return out


def edge_map_from_picture(img, sps, pps=None):
    """(blk uint16 (H/8, W/8), ctb DBK_CTB (ctbs_h, ctbs_w)) of one parsed picture."""
    w, h = int(sps.pic_width_in_luma_samples), int(sps.pic_height_in_luma_samples)
    if w % 8 or h % 8:
        raise ValueError("picture size must be a multiple of the minimum CB size (8)")
    ctb_log2 = int(sps.ctb_log2_size_y)
    wc, hc = int(sps.pic_width_in_ctbs_y), int(sps.pic_height_in_ctbs_y)
    w4, h4, w8, h8 = w // 4, h // 4, w // 8, h // 8
    intra4 = np.zeros((h4, w4), bool)
    cbf4 = np.zeros((h4, w4), bool)
    slice4 = np.zeros((h4, w4), np.int64)
    tile4 = np.zeros((h4, w4), np.int64)
    # edge present (TU or PU edge, filterEdgeFlag not yet applied); tu_* = transform edge
    v_edge = np.zeros((h4, w8), bool)
    v_tu = np.zeros((h4, w8), bool)
    h_edge = np.zeros((h8, w4), bool)
    h_tu = np.zeros((h8, w4), bool)
    qp8 = np.zeros((h8, w8), np.int64)
    nof8 = np.zeros((h8, w8), bool)
    ctb = np.zeros((hc, wc), DBK_CTB)
    slices = _slice_params(img, pps)
    pcm_off = bool(getattr(sps, "pcm_loop_filter_disabled_flag", 0))
    tiles = pps is not None and bool(getattr(pps, "tiles_enabled_flag", 0))

    def mark(x, y, size, tu):
        if x % 8 == 0:
            v_edge[y >> 2:(y + size) >> 2, x >> 3] = True
            if tu:
                v_tu[y >> 2:(y + size) >> 2, x >> 3] = True
        if y % 8 == 0:
            h_edge[y >> 3, x >> 2:(x + size) >> 2] = True
            if tu:
                h_tu[y >> 3, x >> 2:(x + size) >> 2] = True

    def walk_tu(tu):
        if tu.children:
            for ch in tu.children:
                walk_tu(ch)
            return
        mark(tu.x, tu.y, tu.size, True)
        if getattr(tu, "cbf_luma", 0):
            cbf4[tu.y >> 2:(tu.y + tu.size) >> 2, tu.x >> 2:(tu.x + tu.size) >> 2] = True

    for addr, ctu in img.ctus.items():
        sa = int(getattr(ctu, "slice_addr", 0))
        off, beta, tc, _ = slices.get(sa, (0, 0, 0, 1))
        e = ctb[addr // wc, addr % wc]
        e["beta_offset_div2"], e["tc_offset_div2"] = beta, tc
        e["cb_qp_offset"] = int(getattr(pps, "pps_cb_qp_offset", 0)) if pps is not None else 0
        e["cr_qp_offset"] = int(getattr(pps, "pps_cr_qp_offset", 0)) if pps is not None else 0
        x0, y0 = (addr % wc) << ctb_log2, (addr // wc) << ctb_log2
        slice4[y0 >> 2:(y0 >> 2) + (1 << (ctb_log2 - 2)), x0 >> 2:(x0 >> 2) + (1 << (ctb_log2 - 2))] = sa
        if tiles:
            tile4[y0 >> 2:(y0 >> 2) + (1 << (ctb_log2 - 2)), x0 >> 2:(x0 >> 2) + (1 << (ctb_log2 - 2))] = \
                int(pps.tile_id[pps.ctb_addr_rs2ts[addr]])
        for cu in _leaf_cus(ctu):
            ys, xs = slice(cu.y >> 2, (cu.y + cu.size) >> 2), slice(cu.x >> 2, (cu.x + cu.size) >> 2)
            intra4[ys, xs] = cu.pred_mode == MODE_INTRA
            qp8[cu.y >> 3:(cu.y + cu.size) >> 3, cu.x >> 3:(cu.x + cu.size) >> 3] = int(cu.qp_y)
            if getattr(cu, "cu_transquant_bypass_flag", 0) or (pcm_off and getattr(cu, "pcm_flag", 0)):
                nof8[cu.y >> 3:(cu.y + cu.size) >> 3, cu.x >> 3:(cu.x + cu.size) >> 3] = True
            if off:                      # slice_deblocking_filter_disabled_flag: no edge of this CU
                continue
            mark(cu.x, cu.y, cu.size, True)            # coding block edge = transform block edge
            root = getattr(cu, "tu", None)
            if root is not None:
                walk_tu(root)
            if cu.pred_mode != MODE_INTRA:
                # prediction block edges of the inter partitions (8.7.2.3); Bs from motion data
                # is outside this decoder's scope (the reference parses no motion compensation)
                half, quarter = cu.size >> 1, cu.size >> 2
                pm = int(getattr(cu, "part_mode", 0))   # 0 2Nx2N 1 2NxN 2 Nx2N 3 NxN 4 2NxnU 5 2NxnD 6 nLx2N 7 nRx2N
                vx = {2: half, 3: half, 6: quarter, 7: cu.size - quarter}.get(pm)
                if vx is not None and (cu.x + vx) % 8 == 0:
                    v_edge[ys, (cu.x + vx) >> 3] = True
                hy = {1: half, 3: half, 4: quarter, 5: cu.size - quarter}.get(pm)
                if hy is not None and (cu.y + hy) % 8 == 0:
                    h_edge[(cu.y + hy) >> 3, xs] = True

    # ---- filterEdgeFlag (8.7.2.2): picture, slice and tile boundaries ------------------
    across_tiles = True if not tiles else bool(getattr(pps, "loop_filter_across_tiles_enabled_flag", 1))
    v_edge[:, 0] = False
    h_edge[0, :] = False
    xq = np.arange(1, w8) * 2                       # 4x4 column of q0 for edge x8
    ok = np.ones((h4, w8 - 1), bool)
    sl_q, sl_p = slice4[:, xq], slice4[:, xq - 1]
    across = np.vectorize(lambda a: slices.get(int(a), (0, 0, 0, 1))[3])(sl_q).astype(bool) if len(slices) > 1 \
        else np.ones_like(ok)
    ok &= (sl_q == sl_p) | across
    if not across_tiles:
        ok &= tile4[:, xq] == tile4[:, xq - 1]
    v_edge[:, 1:] &= ok
    yq = np.arange(1, h8) * 2
    ok = np.ones((h8 - 1, w4), bool)
    sl_q, sl_p = slice4[yq, :], slice4[yq - 1, :]
    across = np.vectorize(lambda a: slices.get(int(a), (0, 0, 0, 1))[3])(sl_q).astype(bool) if len(slices) > 1 \
        else np.ones_like(ok)
    ok &= (sl_q == sl_p) | across
    if not across_tiles:
        ok &= tile4[yq, :] == tile4[yq - 1, :]
    h_edge[1:, :] &= ok

    # ---- boundary strength (8.7.2.4) -------------------------------------------------
    bs_v = np.zeros((h4, w8), np.uint16)
    xq = np.arange(w8) * 2
    xp = np.maximum(xq - 1, 0)
    bs_v[v_tu & (cbf4[:, xq] | cbf4[:, xp])] = 1
    bs_v[intra4[:, xq] | intra4[:, xp]] = 2
    bs_v[~v_edge] = 0
    bs_h = np.zeros((h8, w4), np.uint16)
    yq = np.arange(h8) * 2
    yp = np.maximum(yq - 1, 0)
    bs_h[h_tu & (cbf4[yq, :] | cbf4[yp, :])] = 1
    bs_h[intra4[yq, :] | intra4[yp, :]] = 2
    bs_h[~h_edge] = 0

    blk = (bs_v[0::2, :] << DBK_BS_V0) | (bs_v[1::2, :] << DBK_BS_V1) | \
          (bs_h[:, 0::2] << DBK_BS_H0) | (bs_h[:, 1::2] << DBK_BS_H1)
    blk = blk.astype(np.uint16) | ((qp8 & 0x7F) << DBK_QP_SHIFT).astype(np.uint16)
    blk[nof8] |= DBK_NO_FILTER
    return np.ascontiguousarray(blk, dtype=np.uint16), ctb


def filter_picture(planes, img, sps, pps=None, device: int = 0):
    """Deblocked copy of one reconstructed picture: planes = (Y, Cb, Cr) [row][col] arrays."""
    y, cb, cr = [np.asarray(p) for p in planes]
    h, w = y.shape
    if cb.shape != (h // 2, w // 2) or cr.shape != cb.shape:
        raise ValueError("planes must be 4:2:0")
    geom = PicGeom(w, h, 1, int(sps.bit_depth_y), int(sps.bit_depth_c))
    dtype = np.uint8 if max(geom.bit_depth_y, geom.bit_depth_c) <= 8 else np.uint16
    buf = np.zeros(geom.total_elems(), dtype=dtype)
    for c, p in enumerate((y, cb, cr)):
        geom.plane_view(buf, 0, c)[:] = p
    blk, ctb = edge_map_from_picture(img, sps, pps)
    out = get_engine(device).deblock(buf, geom, int(sps.ctb_log2_size_y), blk, ctb)
    return tuple(geom.plane_view(out, 0, c).copy() for c in range(3))
