"""Parser-side emission of the packed formats.

The reference parses coefficients into one dense (N, N) int64 array per leaf
(`trans_coeff_level[c][x][y]`, tu.py:87-90, stored at tu.py:331) and reads them back through an
O(#leaves) tree walk PER COEFFICIENT (tu.py:667-684).  Round 1 replaced that with one walk over
the finished picture tree (packer.pack_pictures).  Here the coefficients leave the parser the
moment they are final: when a leaf CU's QP has been derived (`Cu.decode_leaf` -> `decode_qp`,
cu.py:483-488 -- qP needs cu_qp_delta, which is parsed inside the CU's first coded TU) its TBs
are written straight into the picture's *packed coefficient stream* (include/p265_b200.h:
significance bitmap + non-zero levels) and its TU descriptors are appended.  No second pass
over the picture, no dense arena, no transposition of whole pictures; the only end-of-picture
work is sorting the 16-byte descriptors by size.

    emit.hook_parser(cu_module)        # wraps Cu.decode_qp (the reference is not edited)
    ...decode...
    batch = emit.take(img, sps)        # PackedResidualBatch of that picture

INTEGRATION.md shows the equivalent one-line edit of cu.py for a maintainer.
"""
from __future__ import annotations

import numpy as np

from . import packer
from .picture import (TU_DESC, TU_LEVELS8, TU_ZC_SHIFT, TU_ZR_SHIFT, PackedResidualBatch, PicGeom, extent_code,
                      sort_by_size)


class PictureSink:
    """Accumulates one picture's coded TBs as packed records + descriptors."""

    def __init__(self, pic: int = 0):
        self.pic = pic
        self.stream = bytearray()
        self.recs: list = []
        self.area = 0

    def add_tb(self, c_idx: int, x: int, y: int, log2n: int, qp: int, flags: int, levels_yx) -> None:
        """One coded TB; `levels_yx` is the (N, N) row-major [y][x] view of its TransCoeffLevel."""
        n = 1 << log2n
        lv = np.ascontiguousarray(levels_yx).reshape(-1)
        if lv.size != n * n:
            raise ValueError("TB at (%d,%d) c_idx=%d: coefficient block has %d entries, expected %d"
                             % (x, y, c_idx, lv.size, n * n))
        nz = lv != 0
        vals = lv[nz]
        narrow = True
        if vals.size:
            lo, hi = int(vals.min()), int(vals.max())
            if lo < -32768 or hi > 32767:
                raise ValueError("TransCoeffLevel outside 16 bits at (%d,%d) c_idx=%d" % (x, y, c_idx))
            narrow = lo >= -128 and hi <= 127
        off = len(self.stream)
        self.stream += np.packbits(nz, bitorder="little").tobytes()
        self.stream += vals.astype(np.int8 if narrow else "<i2").tobytes()
        self.stream += b"\0" * (-len(self.stream) & 3)
        # rsvd: level count + the zero-extent codes of a 16x16 / 32x32 TB (the parser knows the last
        # significant position, tu.py:145-148; here: of the stored levels).  They are the ordering key of
        # the end-of-picture sort; the device derives the same codes from the bitmap.
        rsvd = int(vals.size)
        if log2n >= 4:
            pos = np.flatnonzero(nz)
            last_row = int(pos[-1]) >> log2n if pos.size else -1
            last_col = int((pos & (n - 1)).max()) if pos.size else -1
            rsvd |= int(extent_code(np.int64(last_row), n)) << TU_ZR_SHIFT
            rsvd |= int(extent_code(np.int64(last_col), n)) << TU_ZC_SHIFT
        self.recs.append((x, y, log2n, c_idx, qp, (flags & ~TU_LEVELS8) | (TU_LEVELS8 if narrow else 0),
                          off >> 2, self.pic, rsvd))
        self.area += n * n

    def add_cu(self, cu, sps) -> None:
        for c, x, y, l2, qp, fl, lv in packer.iter_cu_tbs(cu, sps):
            self.add_tb(c, x, y, l2, qp, fl, lv)

    def finish(self, geom: PicGeom, scaling_factor=None) -> PackedResidualBatch:
        tus = np.array(self.recs, dtype=TU_DESC) if self.recs else np.zeros(0, dtype=TU_DESC)
        stream = np.frombuffer(bytes(self.stream), dtype=np.uint8)
        covers = self.area == geom.width * geom.height + 2 * geom.width_c * geom.height_c and geom.n_pics == 1
        return PackedResidualBatch(geom, sort_by_size(tus, geom), stream, scaling_factor, covers_all=covers)


_ATTR = "_p265_b200_sink"


def on_cu_decoded(cu) -> None:
    """Call right after `Cu.decode_qp()` (cu.py:487): emits the CU's TBs into its picture's sink."""
    img = cu.ctx.img
    sink = getattr(img, _ATTR, None)
    if sink is None:
        sink = PictureSink()
        setattr(img, _ATTR, sink)
    sink.add_cu(cu, cu.ctx.sps)


def hook_parser(cu_module) -> None:
    """Wrap `cu_module.Cu.decode_qp` so that every leaf CU is emitted as soon as its QP is known.
    Idempotent; nothing of the reference is edited."""
    cls = cu_module.Cu
    if getattr(cls, "_p265_b200_emit", False):
        return
    orig = cls.decode_qp

    def decode_qp(self, *a, **k):
        r = orig(self, *a, **k)
        on_cu_decoded(self)
        return r

    cls.decode_qp = decode_qp
    cls._p265_b200_emit = True
    cls._p265_b200_decode_qp = orig


def unhook_parser(cu_module) -> None:
    cls = cu_module.Cu
    if getattr(cls, "_p265_b200_emit", False):
        cls.decode_qp = cls._p265_b200_decode_qp
        cls._p265_b200_emit = False


def take(img, sps, scaling_factor=None):
    """The picture's PackedResidualBatch (None when the parser was not hooked for it)."""
    sink = getattr(img, _ATTR, None)
    if sink is None:
        return None
    return sink.finish(packer.geom_from_sps(sps, 1), scaling_factor)
