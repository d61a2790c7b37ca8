"""Both in-loop filters of one reconstructed picture in ONE host round trip.

The reference parses the controls of the two filters (deblocking: pps.py:121-131,
slice.py:170-179; SAO: sao.py, slice.py:121-126) and implements neither.  `deblock_api` and
`sao_api` each run one filter with their own H2D + D2H of the planes; the decode flow needs
both back to back on the same samples, so this entry ships the planes once, deblocks in place
on the device (8.7.2), applies SAO to the deblocked samples (8.7.3) and brings the final planes
back once (C-ABI p265_loop_filter_batch).
"""
from __future__ import annotations

import numpy as np

from . import deblock_api, packer, sao_api
from .engine import get_engine
from .picture import PicGeom


def filter_picture(planes, img, sps, pps=None, device: int = 0, deblock: bool = True, sao: bool = True):
    """(Y, Cb, Cr) [row][col] arrays of a reconstructed picture -> deblocked + SAO-filtered copies.

    `sao` is ignored for pictures whose slices have SAO switched off for every CTB."""
    y, cb, cr = [np.asarray(p) for p in planes]
    h, w = y.shape
    if cb.shape != (h // 2, w // 2) or cr.shape != cb.shape:
        raise ValueError("planes must be 4:2:0")
    geom = PicGeom(w, h, 1, int(sps.bit_depth_y), int(sps.bit_depth_c))
    dtype = np.uint8 if max(geom.bit_depth_y, geom.bit_depth_c) <= 8 else np.uint16
    buf = np.zeros(geom.total_elems(), dtype=dtype)
    for c, p in enumerate((y, cb, cr)):
        geom.plane_view(buf, 0, c)[:] = p
    blk = ctb = params = nf = None
    if deblock:
        blk, ctb = deblock_api.edge_map_from_picture(img, sps, pps)
    if sao:
        params = packer.sao_params_from_picture(img, sps, sao_api.availability_from_picture(img, sps, pps), pps)
        nf = sao_api.no_filter_from_picture(img, sps)
    if blk is None and params is None:
        return tuple(p.copy() for p in (y, cb, cr))
    get_engine(device).loop_filter(buf, geom, int(sps.ctb_log2_size_y), blk, ctb, params, nf)
    return tuple(geom.plane_view(buf, 0, c).copy() for c in range(3))
