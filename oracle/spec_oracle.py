"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the residual + SAO hot path (numpy int64).

This is a *restatement*, not the product: only tests/, bench.py's cpu_baseline /
`--impl reference` leg and __graft_entry__.smoke() may import it.  The product
package (p265_b200) never does and fails loudly when its CUDA library is missing.

What it restates, and what pins it:

* `inverse_scaling`      follows /root/reference/decoder/scaling.py:4-47 (== ITU-T
  H.265 (04/2013) 8.6.3).  PINNED against the reference's own function, run here
  through oracle/refshim.py on sanity.bin's real TBs and on random inputs
  (tests/golden/sanity_residual.npz, tests/test_oracle_vs_reference.py).
* `inverse_transform`    follows the *structure* of transform.py:89-106 (stage 1 over
  columns, clip16((e+64)>>7), stage 2) with the standard's orientation, stage order
  and final bdShift=20-BitDepth (8.6.2, 8.6.4.1-2).  The reference's function is
  defective as written, so the standard-conformant result is "parity
  unpinned" by the reference; it is pinned instead by (i) the tables being equal to
  transform.py:5-72, (ii) forward->inverse round trips, (iii) a naive triple-loop
  restatement, (iv) the independent C restatement oracle/spec_oracle.c.
* `ref_literal_transform` is the closed form of what transform.py:89-109 computes *as
  written* (G3 a-d).  PINNED against the reference's own function through the shim.
* transform-skip / bypass (8.6.2), ScalingFactor expansion (7.3.4 / 7.4.5; the
  reference's sld.py:118-153 is broken, G4) and the SAO filter (8.7.3; absent from the
  reference, G1): "parity unpinned" by the reference -- pinned by spec restatement,
  hand-derived known answers (tests/test_oracle_known_answers.py) and the C oracle.

Array conventions: the reference indexes every 2-D block as [x][y] (x horizontal,
first index; intra.py:34-37).  The device arena is row-major [y][x].  Functions here
whose name ends in `_xy` take/return the reference's [x][y] layout; all others are
row-major [y][x] ("yx") and batch over a leading axis.
"""
from __future__ import annotations

import numpy as np

I64 = np.int64

# --------------------------------------------------------------------------- tables
# 8.6.4.2: the 32x32 DCT basis.  Restated by structure, not copied: entry [j][i] is
# +/- mag[k] with k taken from the angle j*(2i+1)*pi/64 folded into [0, pi/2]; row 0 is
# the constant 64.  mag[k] are the standard's 31 integer magnitudes (k=1..31), mag[16]
# = 64.  tests/test_oracle_tables.py checks equality with transform.py:7-72.
_MAG = [64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67,
        64, 61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4, 0]


def _dct32() -> np.ndarray:
    m = np.zeros((32, 32), dtype=I64)
    for j in range(32):
        for i in range(32):
            if j == 0:
                m[j, i] = 64
                continue
            a = (j * (2 * i + 1)) % 128
            if a > 64:
                a = 128 - a
            m[j, i] = _MAG[a] if a <= 32 else -_MAG[64 - a]
    return m


#: trans_matrix_type0[j][i] -- same object layout as transform.py:7-72 (list of lists)
DCT32 = _dct32()
#: 4x4 DST-VII basis (8.6.4.2 eq. 8-XXX; transform.py:5): row j = basis function j
_a, _b, _c, _d = 29, 55, 74, 84
DST4 = np.array([[_a, _b, _c, _d],
                 [_c, _c, 0, -_c],
                 [_d, -_a, -_c, _b],
                 [_b, -_d, _c, -_a]], dtype=I64)

LEVEL_SCALE = np.array([40, 45, 51, 57, 64, 72], dtype=I64)  # scaling.py:28

COEFF_MIN, COEFF_MAX = -32768, 32767


def trans_matrix(log2size: int, tr_type: int) -> np.ndarray:
    """N x N matrix M with y[i] = sum_j M[j][i] * x[j]  (8.6.4.2)."""
    n = 1 << log2size
    if tr_type == 1:
        assert n == 4
        return DST4
    return DCT32[:: 32 // n, :n].copy()


# ------------------------------------------------------------------ scaling factors
def up_right_diagonal_scan(blk: int) -> np.ndarray:
    """6.5.3: returns (blk*blk, 2) array of (x, y) in up-right diagonal order."""
    out = []
    x = y = 0
    stop = False
    while not stop:
        while y >= 0:
            if x < blk and y < blk:
                out.append((x, y))
                if len(out) == blk * blk:
                    stop = True
                    break
            y -= 1
            x += 1
        y = x
        x = 0
    return np.array(out, dtype=I64)


# Table 7-5 / 7-6 default lists (values in diagonal-scan coefficient order; the same
# numbers as sld.py:4-31).
_DEF4 = [16] * 16
_DEF8_INTRA = [16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 17, 16, 17, 16, 17, 18,
               17, 18, 18, 17, 18, 21, 19, 20, 21, 20, 19, 21, 24, 22, 22, 24,
               24, 22, 22, 24, 25, 25, 27, 30, 27, 25, 25, 29, 31, 35, 35, 31,
               29, 36, 41, 44, 41, 36, 47, 54, 54, 47, 65, 70, 65, 88, 88, 115]
_DEF8_INTER = [16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 17, 17, 17, 17, 17, 18,
               18, 18, 18, 18, 18, 20, 20, 20, 20, 20, 20, 20, 24, 24, 24, 24,
               24, 24, 24, 24, 25, 25, 25, 25, 25, 25, 25, 28, 28, 28, 28, 28,
               28, 33, 33, 33, 33, 33, 41, 41, 41, 41, 54, 54, 54, 71, 71, 91]


def default_scaling_lists():
    """ScalingList[sizeId][matrixId] (coefficient lists) + dc[sizeId-2][matrixId]."""
    lists = {}
    for m in range(6):
        lists[(0, m)] = list(_DEF4)
        for s in (1, 2):
            lists[(s, m)] = list(_DEF8_INTRA if m < 3 else _DEF8_INTER)
    lists[(3, 0)] = list(_DEF8_INTRA)
    lists[(3, 1)] = list(_DEF8_INTER)
    dc = {(2, m): 16 for m in range(6)}
    dc.update({(3, 0): 16, (3, 1): 16})
    return lists, dc


def expand_scaling_factor(lists, dc):
    """7.4.5: ScalingFactor[sizeId][matrixId] as (N, N) arrays in [x][y] layout.

    Returns dict (sizeId, matrixId) -> int64 array indexed [x][y] like the reference's
    `sps.scaling_factor[size_id][matrix_id][x][y]` (scaling.py:44).
    """
    sf = {}
    scan4 = up_right_diagonal_scan(4)
    scan8 = up_right_diagonal_scan(8)
    for (s, m), lst in lists.items():
        n = 4 << s
        f = np.zeros((n, n), dtype=I64)
        if s == 0:
            for i in range(16):
                f[scan4[i, 0], scan4[i, 1]] = lst[i]
        else:
            rep = 1 << (s - 1)  # 1, 2, 4
            for i in range(64):
                x0, y0 = scan8[i, 0] * rep, scan8[i, 1] * rep
                f[x0:x0 + rep, y0:y0 + rep] = lst[i]
            if s >= 2:
                f[0, 0] = dc[(s, m)]
        sf[(s, m)] = f
    return sf


SF_OFFSETS = {0: 0, 1: 96, 2: 96 + 384, 3: 96 + 384 + 1536}
SF_BYTES = 4064


def pack_scaling_factor(sf) -> np.ndarray:
    """The 4064-byte device table: [sizeId][matrixId] row-major [y][x] uint8."""
    out = np.zeros(SF_BYTES, dtype=np.uint8)
    for (s, m), f in sf.items():
        n = 4 << s
        off = SF_OFFSETS[s] + m * n * n
        out[off:off + n * n] = f.T.reshape(-1).astype(np.uint8)  # [x][y] -> [y][x]
    return out


def matrix_id(log2size: int, c_idx: int, intra: bool) -> int:
    """scaling.py:33-42."""
    if log2size - 2 == 3:
        return 0 if intra else 1
    return c_idx if intra else c_idx + 3


# --------------------------------------------------------------------- dequantise
def inverse_scaling(levels, qp, bit_depth, log2size, m=None):
    """8.6.3 / scaling.py:23-47.  levels: (..., N, N) int; qp: scalar or (...,) array
    of qP *including* QpBdOffset (scaling.py:13-18); m: None (flat 16) or (N, N) /
    broadcastable scaling factors in the same layout as `levels`.  Returns int64."""
    lv = np.asarray(levels, dtype=I64)
    qp = np.asarray(qp, dtype=I64)
    bd_shift = bit_depth + log2size - 5
    scale = LEVEL_SCALE[qp % 6] << (qp // 6)
    scale = scale.reshape(scale.shape + (1, 1)) if scale.ndim else scale
    mm = I64(16) if m is None else np.asarray(m, dtype=I64)
    d = (lv * mm * scale + (1 << (bd_shift - 1))) >> bd_shift
    return np.clip(d, COEFF_MIN, COEFF_MAX)


# ----------------------------------------------------------------- inverse transform
def inverse_transform_1d(x, log2size, tr_type):
    """8.6.4.2 one-dimensional transform of a length-N vector (last axis)."""
    mat = trans_matrix(log2size, tr_type)
    return np.asarray(x, dtype=I64) @ mat  # y[i] = sum_j x[j] M[j][i]


def inverse_transform_yx(d, log2size, tr_type, bit_depth):
    """8.6.4.1 + the 8.6.2 rounding shift.  d: (..., N, N) row-major [y][x].

    Stage 1 transforms every *column* (over y), stage 2 every *row* (over x)."""
    mat = trans_matrix(log2size, tr_type)
    d = np.asarray(d, dtype=I64)
    e = np.einsum("ji,...jx->...ix", mat, d)             # vertical, per column x
    g = np.clip((e + 64) >> 7, COEFF_MIN, COEFF_MAX)
    r = np.einsum("...yj,ji->...yi", g, mat)              # horizontal, per row y
    bd_shift = 20 - bit_depth
    return (r + (1 << (bd_shift - 1))) >> bd_shift


def transform_skip_yx(d, bit_depth):
    """8.6.2 + 8.6.4.1 (04/2013): r = d << 7 then the common bdShift rounding."""
    r = np.asarray(d, dtype=I64) << 7
    bd_shift = 20 - bit_depth
    return (r + (1 << (bd_shift - 1))) >> bd_shift


def sat16(r):
    """Residual planes are int16: values are saturated.  For every bit depth <= 15
    clip1(pred + sat16(r)) == clip1(pred + r), so this loses nothing downstream
    (reconstruction.py:23-25)."""
    return np.clip(r, COEFF_MIN, COEFF_MAX)


def residual_block_yx(levels, qp, bit_depth, log2size, *, dst=False, ts=False,
                      bypass=False, m=None):
    """Full per-TB residual (8.6.2): levels (..., N, N) [y][x] -> residual [y][x]."""
    if bypass:
        return np.asarray(levels, dtype=I64).copy()
    d = inverse_scaling(levels, qp, bit_depth, log2size, m)
    if ts:
        return transform_skip_yx(d, bit_depth)
    return inverse_transform_yx(d, log2size, 1 if dst else 0, bit_depth)


# --------------------------------------------------- the reference "as written" (G3)
def ref_literal_transform_xy(d_xy, log2size, c_idx):
    """Closed form of transform.py:89-109 exactly as written ([x][y] in and out).

    (a) matrix used as C[i][j] = M[i][j*32/N] (transform.py:81,85), (b) stage 2 reads
    the stale loop variable `col` == N-1 for every row (transform.py:108-109), (c) no
    final shift, (d) DST for every 4x4 luma block (transform.py:97)."""
    n = 1 << log2size
    d = np.asarray(d_xy, dtype=I64)
    if n == 4 and c_idx == 0:
        c = DST4
    else:
        c = DCT32[:n, :: 32 // n]
    e = c @ d                                   # e[:, col] = C @ d[:, col]
    g = np.clip((e + 64) >> 7, COEFF_MIN, COEFF_MAX)
    row = c @ g[:, n - 1]
    return np.tile(row, (n, 1))                 # r[row, :] = C @ g[:, N-1]


# ---------------------------------------------------------------------------- SAO
def sao_offset_val(type_idx, offset_abs, offset_sign, bit_depth, log2_offset_scale=0):
    """7.4.9.3.2 SaoOffsetVal[1..4] for one CTB component (sao.py:43-77 fields).

    Edge offset: signs are fixed (+,+,-,-) whatever `offset_sign` holds (the reference
    leaves it 0 for non-merged edge CTBs, sao.py:111-116).  Scale: the 04/2013 edition
    shifts by bitDepth - Min(bitDepth, 10), which is 0 for every bit depth its profiles
    allow; from the 10/2014 edition on (the one that defines profiles above 10 bits) the
    shift is log2_sao_offset_scale_{luma,chroma} of the PPS range extension, 0 when absent.
    This follows the later text -- identical up to 10 bits, and what libavcodec decodes at
    12 bits (tests/golden/fuzz/rext12_lists_ctb32.bin)."""
    shift = int(log2_offset_scale)
    vals = []
    for i in range(4):
        if type_idx == 2:
            sign = 1 if i < 2 else -1
        else:
            sign = -1 if int(offset_sign[i]) else 1
        vals.append(sign * (int(offset_abs[i]) << shift))
    return vals


_HPOS = {0: (-1, 1), 1: (0, 0), 2: (-1, 1), 3: (1, -1)}
_VPOS = {0: (0, 0), 1: (-1, 1), 2: (-1, 1), 3: (-1, 1)}
_EDGE_REMAP = np.array([1, 2, 0, 3, 4], dtype=I64)


def sao_filter_plane(rec, bit_depth, ctb_size, type_idx, band_pos, eo_class, offset_val,
                     ctb_avail=None, no_filter=None, no_filter_log2=3):
    """8.7.3 on one component plane, out of place.

    rec: (H, W) samples (deblocked picture).  Per-CTB arrays are shaped
    (ctbs_h, ctbs_w[, 4]): type_idx {0 off, 1 band, 2 edge}, band_pos, eo_class,
    offset_val = SaoOffsetVal[1..4].  `ctb_size` is the CTB size *in this plane*.
    ctb_avail: optional (ctbs_h, ctbs_w) uint16 bit masks; bit (dy+1)*3+(dx+1) set
    means "samples of the neighbouring CTB in direction (dx, dy) may be used" (clear =
    other slice / tile with the respective loop_filter_across_* flag equal to 0; the
    host derives the masks, see p265_b200/sao.py:ctb_availability).  no_filter: optional bool array at (plane) granularity
    `1 << no_filter_log2` marking pcm+pcm_loop_filter_disabled / cu_transquant_bypass
    blocks whose samples stay unmodified."""
    rec = np.asarray(rec)
    h, w = rec.shape
    out = rec.astype(I64).copy()
    src = rec.astype(I64)
    max_val = (1 << bit_depth) - 1
    band_shift = bit_depth - 5
    ctbs_h = (h + ctb_size - 1) // ctb_size
    ctbs_w = (w + ctb_size - 1) // ctb_size
    ys, xs = np.mgrid[0:h, 0:w]
    cy, cx = ys // ctb_size, xs // ctb_size
    for ry in range(ctbs_h):
        for rx in range(ctbs_w):
            t = int(type_idx[ry][rx])
            if t == 0:
                continue
            y0, x0 = ry * ctb_size, rx * ctb_size
            y1, x1 = min(y0 + ctb_size, h), min(x0 + ctb_size, w)
            blk = src[y0:y1, x0:x1]
            off = np.array([0] + [int(v) for v in offset_val[ry][rx]], dtype=I64)
            if t == 1:
                table = np.zeros(32, dtype=I64)
                for k in range(4):
                    table[(k + int(band_pos[ry][rx])) & 31] = k + 1
                idx = table[blk >> band_shift]
                res = np.clip(blk + off[idx], 0, max_val)
            else:
                cls = int(eo_class[ry][rx])
                yy, xx = ys[y0:y1, x0:x1], xs[y0:y1, x0:x1]
                edge = np.full(blk.shape, 2, dtype=I64)
                valid = np.ones(blk.shape, dtype=bool)
                for k in range(2):
                    ny, nx = yy + _VPOS[cls][k], xx + _HPOS[cls][k]
                    inside = (ny >= 0) & (ny < h) & (nx >= 0) & (nx < w)
                    nyc, nxc = np.clip(ny, 0, h - 1), np.clip(nx, 0, w - 1)
                    if ctb_avail is not None:
                        bit = (nyc // ctb_size - ry + 1) * 3 + (nxc // ctb_size - rx + 1)
                        inside &= ((int(ctb_avail[ry][rx]) >> bit) & 1).astype(bool) | (bit == 4)
                    valid &= inside
                    edge += np.sign(blk - src[nyc, nxc])
                idx = np.where(valid, _EDGE_REMAP[edge], 0)
                res = np.clip(blk + off[idx], 0, max_val)
            if no_filter is not None:
                nf = np.asarray(no_filter)[(ys[y0:y1, x0:x1] >> no_filter_log2),
                                           (xs[y0:y1, x0:x1] >> no_filter_log2)]
                res = np.where(nf, blk, res)
            out[y0:y1, x0:x1] = res
    return out.astype(rec.dtype)


def sao_filter_plane_naive(rec, bit_depth, ctb_size, type_idx, band_pos, eo_class,
                           offset_val, ctb_avail=None):
    """Per-sample restatement of 8.7.3 (slow; cross-check for sao_filter_plane)."""
    h, w = len(rec), len(rec[0])
    out = [[int(v) for v in row] for row in rec]
    max_val = (1 << bit_depth) - 1
    for y in range(h):
        for x in range(w):
            ry, rx = y // ctb_size, x // ctb_size
            t = int(type_idx[ry][rx])
            if t == 0:
                continue
            c = int(rec[y][x])
            off = [0] + [int(v) for v in offset_val[ry][rx]]
            if t == 1:
                table = [0] * 32
                for k in range(4):
                    table[(k + int(band_pos[ry][rx])) & 31] = k + 1
                idx = table[c >> (bit_depth - 5)]
            else:
                cls = int(eo_class[ry][rx])
                idx = 2
                for k in range(2):
                    ny, nx = y + _VPOS[cls][k], x + _HPOS[cls][k]
                    if ny < 0 or ny >= h or nx < 0 or nx >= w:
                        idx = -1
                        break
                    bit = (ny // ctb_size - ry + 1) * 3 + (nx // ctb_size - rx + 1)
                    if ctb_avail is not None and bit != 4 and \
                            not (int(ctb_avail[ry][rx]) >> bit) & 1:
                        idx = -1
                        break
                    n = int(rec[ny][nx])
                    idx += (c > n) - (c < n)
                idx = 0 if idx < 0 else (1, 2, 0, 3, 4)[idx]
            out[y][x] = min(max(c + off[idx], 0), max_val)
    return out


# --------------------------------------------- reference-surface ([x][y]) wrappers
def inverse_scaling_xy(levels_xy, qp, bit_depth, log2size, m_xy=None):
    return inverse_scaling(levels_xy, qp, bit_depth, log2size, m_xy)


def inverse_transform_xy(d_xy, log2size, tr_type, bit_depth):
    d = np.asarray(d_xy, dtype=I64)
    return np.swapaxes(inverse_transform_yx(np.swapaxes(d, -1, -2), log2size, tr_type,
                                            bit_depth), -1, -2)


# ------------------------------------------------------------------ reconstruction
def reconstruct(pred, res, bit_depth):
    """8.6.7 picture construction prior to in-loop filtering == reconstruction.py:23-25:
    Clip1(predSamples + resSamples).  PINNED against the reference's own
    reconstruction.reconstruction (tests/test_oracle.py, through the shim)."""
    return np.clip(np.asarray(pred, dtype=I64) + np.asarray(res, dtype=I64), 0, (1 << bit_depth) - 1)


# --------------------------------------------------------------------------- deblocking
# H.265 (04/2013) 8.7.2.  The reference has no deblocking filter (only the control flags
# are parsed, pps.py:121-131, slice.py:170-179).  Inputs are the
# packed edge map the product's host side builds (picture.py DBK_*): per 8x8 luma block
# the Bs of the two edges starting in it, the CU's QpY and the no-filter bit; per CTB the
# slice's beta / tc offsets and the PPS chroma QP offsets.  Pinned end to end by the
# libavcodec decode of sanity.bin (tests/golden/sanity_ffmpeg.npz).
BETA_TABLE = np.array([0] * 16 + list(range(6, 19)) + list(range(20, 66, 2)), dtype=I64)     # Table 8-11
TC_TABLE = np.array([0] * 18 + [1] * 9 + [2] * 4 + [3] * 4 + [4] * 3 + [5, 5, 6, 6, 7, 8, 9, 10, 11, 13,
                    14, 16, 18, 20, 22, 24], dtype=I64)
assert BETA_TABLE.size == 52 and TC_TABLE.size == 54
_QPC_TABLE = [29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37, 37]                         # Table 8-10


def chroma_qp(qpi: int) -> int:
    """QpC as a function of qPi for ChromaArrayType == 1 (Table 8-10)."""
    if qpi < 30:
        return qpi
    if qpi >= 44:
        return qpi - 6
    return _QPC_TABLE[qpi - 30]


def _blk_bs(blk, vertical, x, y):
    """Bs of the edge segment whose first q sample is luma (x, y)."""
    e = int(blk[y >> 3, x >> 3])
    if vertical:
        return (e >> (2 if (y & 4) else 0)) & 3
    return (e >> (6 if (x & 4) else 4)) & 3


def _blk_qp(blk, x, y):
    q = (int(blk[y >> 3, x >> 3]) >> 8) & 0x7F
    return q - 128 if q >= 64 else q


def _blk_nofilter(blk, x, y):
    return bool(int(blk[y >> 3, x >> 3]) & 0x8000)


def _deblock_luma_segment(pix, blk, ctb, ctb_log2, bit_depth, vertical, x0, y0):
    """One 4-sample luma edge segment (8.7.2.5.3, .6, .7).  pix: int64 [row][col], in place.
    (x0, y0) = first q0 sample; p samples lie to the left (vertical edge) or above."""
    bs = _blk_bs(blk, vertical, x0, y0)
    if bs == 0:
        return
    dx, dy = (1, 0) if vertical else (0, 1)          # across the edge
    sx, sy = (0, 1) if vertical else (1, 0)          # along the edge

    def P(i, k):
        return int(pix[y0 - (i + 1) * dy + k * sy, x0 - (i + 1) * dx + k * sx])

    def Q(i, k):
        return int(pix[y0 + i * dy + k * sy, x0 + i * dx + k * sx])

    qp_q, qp_p = _blk_qp(blk, x0, y0), _blk_qp(blk, x0 - dx, y0 - dy)
    qpl = (qp_q + qp_p + 1) >> 1
    par = ctb[y0 >> ctb_log2, x0 >> ctb_log2]
    beta = int(BETA_TABLE[min(max(qpl + (int(par["beta_offset_div2"]) << 1), 0), 51)]) << (bit_depth - 8)
    tc = int(TC_TABLE[min(max(qpl + 2 * (bs - 1) + (int(par["tc_offset_div2"]) << 1), 0), 53)]) << (bit_depth - 8)
    dp0 = abs(P(2, 0) - 2 * P(1, 0) + P(0, 0))
    dp3 = abs(P(2, 3) - 2 * P(1, 3) + P(0, 3))
    dq0 = abs(Q(2, 0) - 2 * Q(1, 0) + Q(0, 0))
    dq3 = abs(Q(2, 3) - 2 * Q(1, 3) + Q(0, 3))
    dpq0, dpq3, dp, dq = dp0 + dq0, dp3 + dq3, dp0 + dp3, dq0 + dq3
    d = dpq0 + dpq3
    if d >= beta:
        return

    def d_sam(dpq, k):
        return (dpq < (beta >> 2) and abs(P(3, k) - P(0, k)) + abs(Q(0, k) - Q(3, k)) < (beta >> 3)
                and abs(P(0, k) - Q(0, k)) < ((5 * tc + 1) >> 1))

    strong = d_sam(2 * dpq0, 0) and d_sam(2 * dpq3, 3)
    d_ep = dp < ((beta + (beta >> 1)) >> 3)
    d_eq = dq < ((beta + (beta >> 1)) >> 3)
    no_p = _blk_nofilter(blk, x0 - dx, y0 - dy)
    no_q = _blk_nofilter(blk, x0, y0)
    mx = (1 << bit_depth) - 1

    def clip3(lo, hi, v):
        return lo if v < lo else (hi if v > hi else v)

    for k in range(4):
        p = [P(i, k) for i in range(4)]
        q = [Q(i, k) for i in range(4)]
        np_, nq_ = {}, {}
        if strong:
            np_[0] = clip3(p[0] - 2 * tc, p[0] + 2 * tc, (p[2] + 2 * p[1] + 2 * p[0] + 2 * q[0] + q[1] + 4) >> 3)
            np_[1] = clip3(p[1] - 2 * tc, p[1] + 2 * tc, (p[2] + p[1] + p[0] + q[0] + 2) >> 2)
            np_[2] = clip3(p[2] - 2 * tc, p[2] + 2 * tc, (2 * p[3] + 3 * p[2] + p[1] + p[0] + q[0] + 4) >> 3)
            nq_[0] = clip3(q[0] - 2 * tc, q[0] + 2 * tc, (p[1] + 2 * p[0] + 2 * q[0] + 2 * q[1] + q[2] + 4) >> 3)
            nq_[1] = clip3(q[1] - 2 * tc, q[1] + 2 * tc, (p[0] + q[0] + q[1] + q[2] + 2) >> 2)
            nq_[2] = clip3(q[2] - 2 * tc, q[2] + 2 * tc, (p[0] + q[0] + q[1] + 3 * q[2] + 2 * q[3] + 4) >> 3)
        else:
            delta = (9 * (q[0] - p[0]) - 3 * (q[1] - p[1]) + 8) >> 4
            if abs(delta) < tc * 10:
                delta = clip3(-tc, tc, delta)
                np_[0] = clip3(0, mx, p[0] + delta)
                nq_[0] = clip3(0, mx, q[0] - delta)
                if d_ep:
                    dl = clip3(-(tc >> 1), tc >> 1, (((p[2] + p[0] + 1) >> 1) - p[1] + delta) >> 1)
                    np_[1] = clip3(0, mx, p[1] + dl)
                if d_eq:
                    dl = clip3(-(tc >> 1), tc >> 1, (((q[2] + q[0] + 1) >> 1) - q[1] - delta) >> 1)
                    nq_[1] = clip3(0, mx, q[1] + dl)
        if not no_p:
            for i, v in np_.items():
                pix[y0 - (i + 1) * dy + k * sy, x0 - (i + 1) * dx + k * sx] = v
        if not no_q:
            for i, v in nq_.items():
                pix[y0 + i * dy + k * sy, x0 + i * dx + k * sx] = v


def _deblock_chroma_segment(pix, blk, ctb, ctb_log2, bit_depth, vertical, xc, yc, c_idx):
    """One 4-sample chroma edge segment (8.7.2.5.5, .8); (xc, yc) = first q0 sample in chroma
    coordinates (4:2:0), edges on the 8-sample chroma grid, only Bs == 2."""
    xl, yl = xc << 1, yc << 1
    if _blk_bs(blk, vertical, xl, yl) != 2:
        return
    dx, dy = (1, 0) if vertical else (0, 1)
    sx, sy = (0, 1) if vertical else (1, 0)
    qp_q, qp_p = _blk_qp(blk, xl, yl), _blk_qp(blk, xl - dx, yl - dy)
    par = ctb[yl >> ctb_log2, xl >> ctb_log2]
    off = int(par["cb_qp_offset"] if c_idx == 1 else par["cr_qp_offset"])
    qpc = chroma_qp(((qp_q + qp_p + 1) >> 1) + off)
    tc = int(TC_TABLE[min(max(qpc + 2 + (int(par["tc_offset_div2"]) << 1), 0), 53)]) << (bit_depth - 8)
    no_p = _blk_nofilter(blk, xl - dx, yl - dy)
    no_q = _blk_nofilter(blk, xl, yl)
    mx = (1 << bit_depth) - 1
    for k in range(4):
        yy, xx = yc + k * sy, xc + k * sx
        p0, p1 = int(pix[yy - dy, xx - dx]), int(pix[yy - 2 * dy, xx - 2 * dx])
        q0, q1 = int(pix[yy, xx]), int(pix[yy + dy, xx + dx])
        delta = min(max((((q0 - p0) << 2) + p1 - q1 + 4) >> 3, -tc), tc)
        if not no_p:
            pix[yy - dy, xx - dx] = min(max(p0 + delta, 0), mx)
        if not no_q:
            pix[yy, xx] = min(max(q0 - delta, 0), mx)


def deblock_picture(planes, blk, ctb, ctb_log2, bit_depth_y, bit_depth_c):
    """8.7.2 on one 4:2:0 picture: all vertical edges of the picture first, then all
    horizontal edges with the vertically filtered samples as input.  Returns new planes."""
    out = [np.asarray(p).astype(I64).copy() for p in planes]
    h, w = out[0].shape
    for vertical in (True, False):
        if vertical:
            for x in range(8, w, 8):
                for y in range(0, h, 4):
                    _deblock_luma_segment(out[0], blk, ctb, ctb_log2, bit_depth_y, True, x, y)
        else:
            for y in range(8, h, 8):
                for x in range(0, w, 4):
                    _deblock_luma_segment(out[0], blk, ctb, ctb_log2, bit_depth_y, False, x, y)
        hc, wc = out[1].shape
        for c in (1, 2):
            if vertical:
                for x in range(8, wc, 8):
                    for y in range(0, hc, 4):
                        _deblock_chroma_segment(out[c], blk, ctb, ctb_log2, bit_depth_c, True, x, y, c)
            else:
                for y in range(8, hc, 8):
                    for x in range(0, wc, 4):
                        _deblock_chroma_segment(out[c], blk, ctb, ctb_log2, bit_depth_c, False, x, y, c)
    return [o.astype(np.asarray(p).dtype) for o, p in zip(out, planes)]


# ------------------------------------------------------- packed coefficient stream (transport)
# Independent restatement of the record format documented in include/p265_b200.h (per TB:
# significance bitmap, N*N bits, bit y*N + x, least significant bit first; then the non-zero
# TransCoeffLevel values in raster order, int8 when the descriptor carries the LEVELS8 flag, else
# little-endian int16; records start at 4 * coeff_off).  It is how the parser's (position, level)
# stores (tu.py:331) travel; TB by TB, plain loops -- checker only.
_TU_LEVELS8 = 32


def unpack_stream(tus, stream):
    """(descriptors indexing a packed stream, stream bytes) -> (descriptors indexing a dense arena,
    int16 arena): TB i's N*N levels, row-major [y][x], at 16 * coeff_off of the result."""
    stream = np.asarray(stream, dtype=np.uint8)
    out_tus = np.array(tus, copy=True)
    sizes = 1 << (2 * out_tus["log2n"].astype(np.int64))
    offs = np.concatenate(([0], np.cumsum(sizes)[:-1])) if len(sizes) else np.zeros(0, np.int64)
    arena = np.zeros(int(sizes.sum()), dtype=np.int16)
    for i, t in enumerate(tus):
        nn = int(sizes[i])
        rec = int(t["coeff_off"]) * 4
        bits = np.unpackbits(stream[rec:rec + nn // 8], bitorder="little")[:nn].astype(bool)
        k = int(bits.sum())
        if k != int(t["rsvd"]) & 0x7FF:    # bits 0-10: level count; bits 11-14: zero-extent codes (include/p265_b200.h)
            raise ValueError("TB %d: descriptor announces %d levels, bitmap has %d bits set" % (i, int(t["rsvd"]) & 0x7FF, k))
        lv0 = rec + nn // 8
        if int(t["flags"]) & _TU_LEVELS8:
            lv = stream[lv0:lv0 + k].view(np.int8).astype(np.int16)
        else:
            lv = stream[lv0:lv0 + 2 * k].copy().view("<i2").astype(np.int16)
        arena[int(offs[i]):int(offs[i]) + nn][bits] = lv
    out_tus["coeff_off"] = (offs >> 4).astype(np.uint32)
    out_tus["flags"] = out_tus["flags"] & np.uint8(0xFF ^ _TU_LEVELS8)
    out_tus["rsvd"] = 0
    return out_tus, arena
