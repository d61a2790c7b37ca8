"""Host-side intra sample prediction + reconstruction (H.265 8.4.4.2, 8.6.5).

The north star keeps intra prediction and reconstruction on the host: every intra TB
predicts from its already reconstructed neighbours, so the chain is sequential in
decoding order.  The reference's own `intra.py` is dead code with undefined names and an
inverted test (intra.py:155,184), so this module restates the standard
behind the reference's parser objects (`cu.intra_pred_mode_y[x][y]`, `cu.intra_pred_mode_c`,
cu.py:176-280; TU tree, tu.py:84-135).  It is the *caller* of the GPU path: residual
planes come in (from `Engine.residual`), reconstructed planes go out (to the deblocking
and SAO kernels).  It is what lets `sanity.bin` be decoded end to end and compared with
an independent conformant decoder (tests/test_decode_sanity.py).

This is host logic of the product (numpy), not a fallback for any kernel: nothing here
computes residuals, deblocking or SAO.
"""
from __future__ import annotations

import numpy as np

MODE_INTRA = 1            # cu.py:29
PLANAR, DC = 0, 1

#: intraPredAngle for predModeIntra 2..34 (Table 8-4)
_ANGLE = (32, 26, 21, 17, 13, 9, 5, 2, 0, -2, -5, -9, -13, -17, -21, -26, -32,
          -26, -21, -17, -13, -9, -5, -2, 0, 2, 5, 9, 13, 17, 21, 26, 32)
#: invAngle for predModeIntra 11..25 (Table 8-5)
_INV_ANGLE = (-4096, -1638, -910, -630, -482, -390, -315, -256, -315, -390, -482, -630,
              -910, -1638, -4096)


def _interleave(x: int, y: int, bits: int) -> int:
    z = 0
    for b in range(bits):
        z |= ((x >> b) & 1) << (2 * b) | ((y >> b) & 1) << (2 * b + 1)
    return z


class _Avail:
    """6.4.1 z-scan availability at min-TB (4x4 luma) granularity."""

    def __init__(self, img, sps, pps):
        self.w, self.h = int(sps.pic_width_in_luma_samples), int(sps.pic_height_in_luma_samples)
        self.ctb_log2 = int(sps.ctb_log2_size_y)
        self.wc = int(sps.pic_width_in_ctbs_y)
        hc = int(sps.pic_height_in_ctbs_y)
        bits = self.ctb_log2 - 2
        n = 1 << bits
        self.inner = np.array([[_interleave(x, y, bits) for x in range(n)] for y in range(n)])
        rs2ts = getattr(pps, "ctb_addr_rs2ts", None) if pps is not None and \
            getattr(pps, "tiles_enabled_flag", 0) else None
        self.rs2ts = list(rs2ts) if rs2ts is not None else list(range(self.wc * hc))
        self.tile = list(pps.tile_id) if rs2ts is not None else None
        self.slice_addr = {a: int(getattr(c, "slice_addr", 0)) for a, c in img.ctus.items()}
        self.constrained = bool(getattr(pps, "constrained_intra_pred_flag", 0)) if pps is not None else False
        self.bits = bits

    def zaddr(self, x: int, y: int):
        rs = (y >> self.ctb_log2) * self.wc + (x >> self.ctb_log2)
        m = (1 << self.bits) - 1
        return rs, (self.rs2ts[rs] << (2 * self.bits)) + int(self.inner[(y >> 2) & m, (x >> 2) & m])

    def available(self, x_cur: int, y_cur: int, x_nb: int, y_nb: int) -> bool:
        if x_nb < 0 or y_nb < 0 or x_nb >= self.w or y_nb >= self.h:
            return False
        rs_c, z_c = self.zaddr(x_cur, y_cur)
        rs_n, z_n = self.zaddr(x_nb, y_nb)
        if z_n > z_c:
            return False
        if self.slice_addr.get(rs_n) != self.slice_addr.get(rs_c):
            return False
        if self.tile is not None and self.tile[self.rs2ts[rs_n]] != self.tile[self.rs2ts[rs_c]]:
            return False
        return True


def _filter_neighbours(ref: np.ndarray, n: int, strong: bool, bit_depth: int) -> np.ndarray:
    """8.4.4.2.3 on the linear neighbour array (index 0 = p[-1][2n-1] ... 2n = corner ...
    4n = p[2n-1][-1])."""
    out = ref.copy()
    if strong and n == 32:
        c, bl, tr = int(ref[2 * n]), int(ref[0]), int(ref[4 * n])
        thr = 1 << (bit_depth - 5)
        if abs(c + tr - 2 * int(ref[3 * n])) < thr and abs(c + bl - 2 * int(ref[n])) < thr:
            i = np.arange(1, 64)
            # left column: p[-1][y], y = 0..62  <->  index 2n-1-y ; weight towards p[-1][63]
            out[2 * n - i] = ((64 - i) * c + i * bl + 32) >> 6
            out[2 * n + i] = ((64 - i) * c + i * tr + 32) >> 6
            return out
    out[1:-1] = (ref[:-2] + 2 * ref[1:-1] + ref[2:] + 2) >> 2
    return out


def predict_block(ref: np.ndarray, n: int, mode: int, c_idx: int, bit_depth: int,
                  strong_smoothing: bool = False, filtered_ok: bool = True) -> np.ndarray:
    """predSamples [y][x] (n x n) from the substituted neighbour array `ref` (4n+1 values,
    int64; layout in `_filter_neighbours`), 8.4.4.2.3 - 8.4.4.2.6."""
    ref = ref.astype(np.int64)
    if c_idx == 0 and mode != DC and n != 4 and filtered_ok:
        dist = min(abs(mode - 26), abs(mode - 10))
        if dist > {8: 7, 16: 1, 32: 0}[n]:
            ref = _filter_neighbours(ref, n, strong_smoothing, bit_depth)
    left = ref[2 * n - 1::-1][:2 * n]         # p[-1][y], y = 0..2n-1
    corner = ref[2 * n]
    top = ref[2 * n + 1:]                     # p[x][-1], x = 0..2n-1
    k = n.bit_length() - 1
    if mode == PLANAR:
        x = np.arange(n)[None, :]
        y = np.arange(n)[:, None]
        return ((n - 1 - x) * left[:n][:, None] + (x + 1) * top[n] +
                (n - 1 - y) * top[:n][None, :] + (y + 1) * left[n] + n) >> (k + 1)
    if mode == DC:
        dc = (int(top[:n].sum()) + int(left[:n].sum()) + n) >> (k + 1)
        pred = np.full((n, n), dc, dtype=np.int64)
        if c_idx == 0 and n < 32:
            pred[0, :] = (top[:n] + 3 * dc + 2) >> 2
            pred[:, 0] = (left[:n] + 3 * dc + 2) >> 2
            pred[0, 0] = (left[0] + 2 * dc + top[0] + 2) >> 2
        return pred
    angle = _ANGLE[mode - 2]
    vertical = mode >= 18
    main, side = (top, left) if vertical else (left, top)
    # ref_[i] for i = -n .. 2n, stored with offset n
    r = np.zeros(3 * n + 1, dtype=np.int64)
    r[n] = corner
    r[n + 1:2 * n + 1] = main[:n]
    last = (n * angle) >> 5
    if angle < 0:
        if last < -1:
            inv = _INV_ANGLE[mode - 11]
            for i in range(-1, last - 1, -1):
                j = -1 + ((i * inv + 128) >> 8)          # side index: -1 = corner
                r[n + i] = corner if j < 0 else side[j]
    else:
        r[2 * n + 1:3 * n + 1] = main[n:2 * n]
    pred = np.empty((n, n), dtype=np.int64)
    pos = np.arange(n)
    for a in range(n):                                   # a: y for vertical, x for horizontal
        idx = ((a + 1) * angle) >> 5
        fact = ((a + 1) * angle) & 31
        base = n + pos + idx + 1
        if fact:
            line = ((32 - fact) * r[base] + fact * r[base + 1] + 16) >> 5
        else:
            line = r[base]
        if vertical:
            pred[a, :] = line
        else:
            pred[:, a] = line
    if c_idx == 0 and n < 32:
        mx = (1 << bit_depth) - 1
        if mode == 26:
            pred[:, 0] = np.clip(top[0] + ((left[:n] - corner) >> 1), 0, mx)
        elif mode == 10:
            pred[0, :] = np.clip(left[0] + ((top[:n] - corner) >> 1), 0, mx)
    return pred


class IntraReconstructor:
    """Decoding order walk of one parsed picture: predict every TB from the planes being
    built, add the residual block, clip (8.6.5 / reconstruction.py:23-25)."""

    def __init__(self, img, sps, pps=None):
        self.img, self.sps, self.pps = img, sps, pps
        self.avail = _Avail(img, sps, pps)
        self.bd = (int(sps.bit_depth_y), int(sps.bit_depth_c), int(sps.bit_depth_c))
        self.strong = bool(getattr(sps, "strong_intra_smoothing_enabled_flag", 0))
        #: PcmBitDepthY, PcmBitDepthC (sps.py:97-98)
        self.pcm_bd = (int(getattr(sps, "pcm_sample_bit_depth_luma_minus1", self.bd[0] - 1)) + 1,
                       int(getattr(sps, "pcm_sample_bit_depth_chroma_minus1", self.bd[1] - 1)) + 1)
        w, h = self.avail.w, self.avail.h
        dt = np.uint8 if max(self.bd) <= 8 else np.uint16
        self.planes = [np.zeros((h, w), dt), np.zeros((h // 2, w // 2), dt), np.zeros((h // 2, w // 2), dt)]
        #: CuPredMode per min-TB (for constrained_intra_pred_flag), 1 = intra
        self.intra_map = np.zeros((h >> 2, w >> 2), dtype=bool)

    # -- neighbours -------------------------------------------------------------------
    def _neighbours(self, c_idx: int, xt: int, yt: int, n: int) -> np.ndarray:
        """Substituted neighbour array of the TB at component coords (xt, yt) (8.4.4.2.2)."""
        sh = 0 if c_idx == 0 else 1
        plane = self.planes[c_idx]
        x_cur, y_cur = xt << sh, yt << sh
        ref = np.zeros(4 * n + 1, dtype=np.int64)
        ok = np.zeros(4 * n + 1, dtype=bool)
        unit = 4 >> sh                      # samples of this component per min TB

        def usable(xn, yn):
            if not self.avail.available(x_cur, y_cur, xn << sh, yn << sh):
                return False
            if self.avail.constrained and not self.intra_map[(yn << sh) >> 2, (xn << sh) >> 2]:
                return False
            return True

        for y0 in range(0, 2 * n, unit):                # left column, top to bottom
            if usable(xt - 1, yt + y0):
                m = min(unit, 2 * n - y0)
                idx = 2 * n - 1 - y0 - np.arange(m)
                ref[idx] = plane[yt + y0:yt + y0 + m, xt - 1]
                ok[idx] = True
        if usable(xt - 1, yt - 1):
            ref[2 * n] = plane[yt - 1, xt - 1]
            ok[2 * n] = True
        for x0 in range(0, 2 * n, unit):
            if usable(xt + x0, yt - 1):
                m = min(unit, 2 * n - x0)
                ref[2 * n + 1 + x0:2 * n + 1 + x0 + m] = plane[yt - 1, xt + x0:xt + x0 + m]
                ok[2 * n + 1 + x0:2 * n + 1 + x0 + m] = True
        if not ok.any():
            ref[:] = 1 << (self.bd[c_idx] - 1)
            return ref
        if not ok.all():
            if not ok[0]:
                ref[0] = ref[int(np.argmax(ok))]
                ok[0] = True
            bad = np.flatnonzero(~ok)
            for i in bad:                                # ascending: copies the previous one
                ref[i] = ref[i - 1]
        return ref

    # -- one TB ------------------------------------------------------------------------
    def _tb(self, c_idx, xt, yt, log2n, mode, residual):
        n = 1 << log2n
        ref = self._neighbours(c_idx, xt, yt, n)
        pred = predict_block(ref, n, mode, c_idx, self.bd[c_idx], self.strong)
        res = residual[c_idx][yt:yt + n, xt:xt + n].astype(np.int64)
        mx = (1 << self.bd[c_idx]) - 1
        self.planes[c_idx][yt:yt + n, xt:xt + n] = np.clip(pred + res, 0, mx)

    def _tu(self, cu, tu, residual):
        if tu.children:
            for ch in tu.children:
                self._tu(cu, ch, residual)
            return
        l2 = tu.log2size
        pb = cu.intra_pb_size
        xp = cu.x + ((tu.x - cu.x) // pb) * pb
        yp = cu.y + ((tu.y - cu.y) // pb) * pb
        self._tb(0, tu.x, tu.y, l2, int(cu.intra_pred_mode_y[xp][yp]), residual)
        if l2 > 2:
            for c in (1, 2):
                self._tb(c, tu.x >> 1, tu.y >> 1, l2 - 1, int(cu.intra_pred_mode_c), residual)
        elif getattr(tu, "idx", 0) == 3 and tu.parent is not None:
            first = tu.parent.children[0]
            for c in (1, 2):
                self._tb(c, first.x >> 1, first.y >> 1, 2, int(cu.intra_pred_mode_c), residual)

    def _pcm(self, cu):
        """8.4.4.1 step for a pcm CU (cu.py:146-151): its samples are the pcm_sample_luma / pcm_sample_chroma
        values shifted up to the picture's bit depth, no prediction, no residual.  The parser has to have kept
        them ([row][col] arrays `pcm_sample_luma` (N, N), `pcm_sample_chroma` (2, N/2, N/2)); the reference
        has no pcm_sample() at all (oracle/refshim.enable_pcm supplies it for the tests)."""
        if not hasattr(cu, "pcm_sample_luma"):
            raise ValueError("pcm CU at (%d,%d) carries no pcm_sample_luma / pcm_sample_chroma" % (cu.x, cu.y))
        n = cu.size
        self.planes[0][cu.y:cu.y + n, cu.x:cu.x + n] = np.asarray(cu.pcm_sample_luma) << (self.bd[0] - self.pcm_bd[0])
        h = n >> 1
        for c in (1, 2):
            self.planes[c][cu.y >> 1:(cu.y >> 1) + h, cu.x >> 1:(cu.x >> 1) + h] = \
                np.asarray(cu.pcm_sample_chroma[c - 1]) << (self.bd[c] - self.pcm_bd[1])

    def run(self, residual):
        """residual: (Y, Cb, Cr) int16 [row][col] planes (zero where no TB is coded)."""
        for addr in sorted(self.img.ctus, key=lambda a: self.avail.rs2ts[a]):
            for cu in self.img.ctus[addr].get_leaves():
                if not hasattr(cu, "pred_mode"):
                    continue
                if cu.pred_mode != MODE_INTRA:
                    raise NotImplementedError("inter prediction is outside this decoder's scope "
                                              "(the reference parses no motion compensation)")
                self.intra_map[cu.y >> 2:(cu.y + cu.size) >> 2, cu.x >> 2:(cu.x + cu.size) >> 2] = True
                if getattr(cu, "pcm_flag", 0):
                    self._pcm(cu)
                    continue
                self._tu(cu, cu.tu, residual)
        return tuple(self.planes)


def reconstruct_intra_picture(img, sps, pps, residual):
    """Reconstructed (pre-deblocking) planes of an all-intra picture."""
    return IntraReconstructor(img, sps, pps).run(residual)
