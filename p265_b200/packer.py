"""Walk the reference parser's picture objects once and emit the packed formats.

Input is duck-typed: the objects the reference's parser builds (`image.Image` with
`ctus`, `cu.Cu` leaves carrying `tu`, `tu.Tu` leaves carrying `trans_coeff_level` --
image.py:5-20, cu.py:99-173, tu.py:84-135) or anything shaped like them.  Nothing from
the reference is imported here.

The reference calls `tu.get_trans_coeff_level` once per coefficient, an O(#leaves)
tree walk each time (tu.py:667-684); this packer replaces that with one pass per
picture.
"""
from __future__ import annotations

import numpy as np

from .picture import (AVAIL_ALL, SAO_CTB, TU_BYPASS, TU_DESC, TU_DST, TU_INTRA, TU_SKIP,
                      PicGeom, ResidualBatch, extent_codes, set_extents, sort_by_size)

MODE_INTRA = 1   # cu.py:29


def _leaf_cus(ctu):
    """Leaf CUs that were actually parsed (children outside the picture are created
    but never parsed, cu.py:86-94)."""
    for cu in ctu.get_leaves():
        if hasattr(cu, "pred_mode"):
            yield cu


def iter_tbs(img, sps):
    """Yield (c_idx, x, y, log2n, qp, flags, coeffs_yx) for every coded TB of a picture,
    in decoding order.  Coordinates are in the component's plane; `coeffs_yx` is the
    (N, N) row-major [y][x] view of the reference's [x][y] array."""
    for addr in sorted(img.ctus):
        for cu in _leaf_cus(img.ctus[addr]):
            yield from iter_cu_tbs(cu, sps)


def iter_cu_tbs(cu, sps):
    """The coded TBs of ONE leaf CU (same tuples as iter_tbs): what the parser-side emitter
    (emit.PictureSink) consumes the moment the CU's QP is known (cu.py:487)."""
    root = getattr(cu, "tu", None)
    if root is None:
        return
    intra = cu.pred_mode == MODE_INTRA
    bypass = bool(getattr(cu, "cu_transquant_bypass_flag", 0))
    qps = (cu.qp_y + sps.qp_bd_offset_y, cu.qp_cb + sps.qp_bd_offset_c,
           cu.qp_cr + sps.qp_bd_offset_c)
    base = (TU_INTRA if intra else 0) | (TU_BYPASS if bypass else 0)
    for leaf in root.get_leaves():
        levels = getattr(leaf, "trans_coeff_level", None)
        if levels is None:
            continue
        ts = getattr(leaf, "transform_skip_flag", (0, 0, 0))
        if getattr(leaf, "cbf_luma", 1):
            fl = base | (TU_SKIP if ts[0] else 0)
            if leaf.log2size == 2 and intra:
                fl |= TU_DST
            yield 0, leaf.x, leaf.y, leaf.log2size, qps[0], fl, levels[0].T
        if leaf.log2size > 2:
            cx, cy, cl2 = leaf.x >> 1, leaf.y >> 1, leaf.log2size - 1
            cbf = (leaf.cbf_cb, leaf.cbf_cr)
        elif getattr(leaf, "idx", 0) == 3 and leaf.parent is not None:
            # four 4x4 luma TBs share one 4x4 Cb + Cr pair, parsed with the 4th
            # sibling at the first sibling's position (tu.py:128-135)
            first = leaf.parent.children[0]
            cx, cy, cl2 = first.x >> 1, first.y >> 1, 2
            cbf = (first.cbf_cb, first.cbf_cr)
        else:
            continue
        for c in (1, 2):
            if cbf[c - 1]:
                fl = base | (TU_SKIP if ts[c] else 0)
                yield c, cx, cy, cl2, qps[c], fl, levels[c].T


def geom_from_sps(sps, n_pics: int = 1) -> PicGeom:
    return PicGeom(width=int(sps.pic_width_in_luma_samples),
                   height=int(sps.pic_height_in_luma_samples), n_pics=n_pics,
                   bit_depth_y=int(sps.bit_depth_y), bit_depth_c=int(sps.bit_depth_c))


def pack_pictures(imgs, sps, scaling_factor=None) -> ResidualBatch:
    """Pack the coded TBs of `imgs` (a list of parsed pictures) into one batch."""
    geom = geom_from_sps(sps, len(imgs))
    recs, blocks, off = [], [], 0
    for p, img in enumerate(imgs):
        for c, x, y, l2, qp, fl, lv in iter_tbs(img, sps):
            n = 1 << l2
            if lv.shape != (n, n):
                raise ValueError("TB at (%d,%d) c_idx=%d: coefficient block is %r, expected "
                                 "%dx%d" % (x, y, c, lv.shape, n, n))
            recs.append((x, y, l2, c, qp, fl, off >> 4, p, 0))
            blocks.append(np.ascontiguousarray(lv, dtype=np.int16).reshape(-1))
            off += n * n
    tus = np.array(recs, dtype=TU_DESC) if recs else np.zeros(0, dtype=TU_DESC)
    coeffs = np.concatenate(blocks) if blocks else np.zeros(0, dtype=np.int16)
    set_extents(tus, *extent_codes(tus, coeffs))     # zero-extent codes of the 16x16 / 32x32 TBs (picture.py)
    return ResidualBatch(geom=geom, tus=sort_by_size(tus, geom), coeffs=coeffs,
                         scaling_factor=scaling_factor, covers_all=tbs_cover_planes(tus, geom))


def tbs_cover_planes(tus, geom) -> bool:
    """True when the (non-overlapping) TBs tile every plane of every picture completely, so the
    residual launch needs no zero fill (P265_RES_ZERO_FILL)."""
    area = int((1 << (2 * tus["log2n"].astype(np.int64))).sum())
    return area == geom.n_pics * (geom.width * geom.height + 2 * geom.width_c * geom.height_c)


# ------------------------------------------------------------------------- SAO
def sao_offset_val(type_idx: int, offset_abs, offset_sign, bit_depth: int, log2_offset_scale: int = 0):
    """SaoOffsetVal[1..4] (7.4.9.3.2) from the fields `sao.Sao.parse` fills
    (sao.py:43-77).  Edge offsets have fixed signs (+,+,-,-); the reference only writes
    them into `sao_offset_sign` on the merge path (sao.py:111-116).  The scale is
    log2_sao_offset_scale_{luma,chroma} (PPS range extension, 0 when absent): for bit depths up
    to 10 that equals the 04/2013 edition's bitDepth - Min(bitDepth, 10)."""
    shift = int(log2_offset_scale)
    out = []
    for i in range(4):
        if type_idx == 2:
            neg = i >= 2
        else:
            neg = bool(offset_sign[i])
        v = int(offset_abs[i]) << shift
        out.append(-v if neg else v)
    return out


def sao_params_from_picture(img, sps, avail=None, pps=None) -> np.ndarray:
    """(ctbs_h, ctbs_w) SAO_CTB table from `img.ctus[addr].sao` (ctu.py:22)."""
    wc, hc = int(sps.pic_width_in_ctbs_y), int(sps.pic_height_in_ctbs_y)
    scale = (int(getattr(pps, "log2_sao_offset_scale_luma", 0)),) + \
        (int(getattr(pps, "log2_sao_offset_scale_chroma", 0)),) * 2
    tab = np.zeros((hc, wc), dtype=SAO_CTB)
    tab["avail"] = AVAIL_ALL
    for addr, ctu in img.ctus.items():
        s = getattr(ctu, "sao", None)
        if s is None or not hasattr(s, "sao_type_idx"):
            continue                                    # SAO disabled for the slice
        e = tab[addr // wc, addr % wc]
        for c in range(3):
            t = int(s.sao_type_idx[c])
            e["type"][c] = t
            e["band_pos"][c] = int(s.sao_band_position[c])
            e["eo_class"][c] = int(s.sao_eo_class[c])
            bd = int(sps.bit_depth_y if c == 0 else sps.bit_depth_c)
            e["offset_val"][c] = sao_offset_val(t, s.sao_offset_abs[c], s.sao_offset_sign[c], bd,
                                                scale[c]) if t else 0
    if avail is not None:
        tab["avail"] = avail
    return tab


def ctb_availability(slice_addr, slice_lf_across, tile_id, lf_across_tiles: bool,
                     ctb_addr_rs2ts=None) -> np.ndarray:
    """Per-CTB neighbour masks for SAO edge offset (8.7.3).

    slice_addr: (ctbs_h, ctbs_w) SliceAddrRs of every CTB (ctu.slice_addr,
    slice.py:252); slice_lf_across: {SliceAddrRs: slice_loop_filter_across_slices_
    enabled_flag} (slice.py:177-179); tile_id: (ctbs_h, ctbs_w) TileId (pps.py:93,
    tile_id_rs); ctb_addr_rs2ts: optional (ctbs_h, ctbs_w) tile-scan addresses giving
    the decoding order (defaults to raster order).

    A neighbour sample in another slice is unusable when the *later* of the two
    slices (in decoding order) has its flag equal to 0; in another tile when
    loop_filter_across_tiles_enabled_flag is 0."""
    slice_addr = np.asarray(slice_addr)
    tile_id = np.asarray(tile_id)
    hc, wc = slice_addr.shape
    order = (np.arange(hc * wc).reshape(hc, wc) if ctb_addr_rs2ts is None
             else np.asarray(ctb_addr_rs2ts))
    out = np.zeros((hc, wc), dtype=np.uint16)
    for ry in range(hc):
        for rx in range(wc):
            m = 0
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    ny, nx = ry + dy, rx + dx
                    if not (0 <= ny < hc and 0 <= nx < wc):
                        continue
                    ok = True
                    if slice_addr[ny, nx] != slice_addr[ry, rx]:
                        later = (ny, nx) if order[ny, nx] > order[ry, rx] else (ry, rx)
                        ok = bool(slice_lf_across[int(slice_addr[later])])
                    if ok and tile_id[ny, nx] != tile_id[ry, rx] and not lf_across_tiles:
                        ok = False
                    if ok:
                        m |= 1 << ((dy + 1) * 3 + (dx + 1))
            out[ry, rx] = m
    return out
