"""Packed per-picture formats at the host <-> C-ABI boundary.

Names follow the decoder's domain: a *TB* (transform block) is one component's block
of one leaf transform unit; a *TU descriptor* is its 16-byte record; the *coefficient
arena* holds every coded TB's TransCoeffLevel values, TB after TB, row-major [y][x]
int16 (the reference stores them [x][y] int64 per leaf, tu.py:87-90,331).

Everything here is plain numpy; nothing computes residuals or SAO on the CPU.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

#: mirrors `p265_tu_desc` in include/p265_b200.h (16 bytes, little endian)
TU_DESC = np.dtype([("x", "<u2"), ("y", "<u2"),        # component-plane sample coords
                    ("log2n", "u1"), ("c_idx", "u1"),
                    ("qp", "u1"),                        # qP incl. QpBdOffset (scaling.py:13-18)
                    ("flags", "u1"),
                    ("coeff_off", "<u4"),                # arena offset in units of 16 coeffs
                    ("pic", "<u2"), ("rsvd", "<u2")])
assert TU_DESC.itemsize == 16

TU_DST = 1        # trType = 1: 4x4 luma intra (8.6.4.2; transform.py:97)
TU_SKIP = 2       # transform_skip_flag (tu.py:142-143)
TU_BYPASS = 4     # cu_transquant_bypass_flag (cu.py:102-105)
TU_INTRA = 8      # CuPredMode == MODE_INTRA -> matrixId (scaling.py:33-42)
TU_PRESCALED = 16  # arena holds d[] already (pu.scaled_samples)
TU_LEVELS8 = 32   # packed coefficient stream: this TB's levels are int8

# `rsvd` of a descriptor: bits 0-10 number of levels (packed stream), bits 11-12 / 13-14 zero-extent
# codes of the rows / columns (16x16 and 32x32 TBs): code z = every coefficient in rows (columns)
# >= N >> z is zero; 0 = nothing known.  The parser knows the last significant position of a TB
# (tu.py:145-148); the kernels skip the products of the empty rows / columns.
TU_LEVELS_MASK = 0x07FF
TU_ZR_SHIFT, TU_ZC_SHIFT = 11, 13

#: mirrors `p265_sao_ctb` (24 bytes)
SAO_CTB = np.dtype([("type", "u1", 3), ("band_pos", "u1", 3), ("eo_class", "u1", 3),
                    ("offset_val", "i1", (3, 4)), ("pad", "u1"), ("avail", "<u2")])
assert SAO_CTB.itemsize == 24

#: mirrors `p265_dbk_ctb` (4 bytes): deblocking parameters of the slice / PPS a CTB belongs to
DBK_CTB = np.dtype([("beta_offset_div2", "i1"), ("tc_offset_div2", "i1"),
                    ("cb_qp_offset", "i1"), ("cr_qp_offset", "i1")])
assert DBK_CTB.itemsize == 4

# Deblocking edge map: one uint16 per 8x8 luma block (`p265_dbk_blk`, raster order)
DBK_BS_V0, DBK_BS_V1 = 0, 2      # bits 0-1 / 2-3: Bs of the vertical edge x = 8*bx, rows 0-3 / 4-7
DBK_BS_H0, DBK_BS_H1 = 4, 6      # bits 4-5 / 6-7: Bs of the horizontal edge y = 8*by, cols 0-3 / 4-7
DBK_QP_SHIFT = 8                 # bits 8-14: QpY of the CU, 7-bit two's complement
DBK_NO_FILTER = 0x8000           # bit 15: pcm + pcm_loop_filter_disabled / cu_transquant_bypass

#: all eight neighbouring CTBs usable (bit (dy+1)*3+(dx+1), centre bit unused)
AVAIL_ALL = 0x1FF

SF_BYTES = 4064
_SF_OFFSETS = (0, 96, 480, 2016)


def align_up(v: int, a: int) -> int:
    return (v + a - 1) // a * a


@dataclass
class PicGeom:
    """Plane geometry of a batch of 4:2:0 (or 4:0:0-like luma-only) pictures.

    Mirrors `p265_pic_geom`.  Planes of one picture are laid out Y, Cb, Cr inside one
    buffer; pictures follow each other `pic_stride` elements apart."""
    width: int                      # luma samples
    height: int
    n_pics: int = 1
    bit_depth_y: int = 8
    bit_depth_c: int = 8
    stride_align: int = 64          # elements; 128 B rows for int16 / uint16 planes
    width_c: int = field(init=False)
    height_c: int = field(init=False)
    stride_y: int = field(init=False)
    stride_c: int = field(init=False)
    plane_off: tuple = field(init=False)
    pic_stride: int = field(init=False)

    def __post_init__(self):
        if self.width <= 0 or self.height <= 0 or self.width % 2 or self.height % 2:
            raise ValueError("picture size must be positive and even (4:2:0)")
        self.width_c, self.height_c = self.width // 2, self.height // 2
        self.stride_y = align_up(self.width, self.stride_align)
        self.stride_c = align_up(self.width_c, self.stride_align)
        y = self.stride_y * self.height
        c = self.stride_c * self.height_c
        self.plane_off = (0, y, y + c)
        self.pic_stride = align_up(y + 2 * c, 64)

    # ---- helpers -------------------------------------------------------------
    def plane_shape(self, c_idx: int):
        return (self.height, self.width) if c_idx == 0 else (self.height_c, self.width_c)

    def stride(self, c_idx: int) -> int:
        return self.stride_y if c_idx == 0 else self.stride_c

    def total_elems(self) -> int:
        return self.pic_stride * self.n_pics

    def plane_view(self, buf: np.ndarray, pic: int, c_idx: int) -> np.ndarray:
        """(H, W) strided view of plane `c_idx` of picture `pic` inside flat `buf`."""
        h, w = self.plane_shape(c_idx)
        s = self.stride(c_idx)
        off = pic * self.pic_stride + self.plane_off[c_idx]
        return buf[off:off + h * s].reshape(h, s)[:, :w]

    def pixels(self) -> int:
        return self.width * self.height * self.n_pics


@dataclass
class ResidualBatch:
    """Everything one residual launch needs: descriptors binned by size (32,16,8,4),
    the coefficient arena and, optionally, the 4064-byte ScalingFactor table."""
    geom: PicGeom
    tus: np.ndarray                         # TU_DESC, sorted by log2n descending
    coeffs: np.ndarray                      # int16 arena
    scaling_factor: np.ndarray | None = None    # uint8[4064] or None (flat 16)
    covers_all: bool = False                # TBs tile every plane completely
    sf_replicated: bool | None = None       # None: detect from the table (see sf_is_replicated)

    bins: tuple | None = None               # TBs per size bin (32,16,8,4); None: counted once, on first use

    def __post_init__(self):
        if self.scaling_factor is not None and self.sf_replicated is None:
            self.sf_replicated = sf_is_replicated(self.scaling_factor)

    def bin_counts(self):
        """TBs per size bin (32, 16, 8, 4).  Part of the packed format: counted once per batch (the
        packer knows it when it sorts) -- `tus` must not change afterwards; the C-ABI re-validates
        the counts against the list on every call."""
        if self.bins is None:
            c = np.bincount(np.ascontiguousarray(self.tus["log2n"]), minlength=6)
            self.bins = (int(c[5]), int(c[4]), int(c[3]), int(c[2]))
        return self.bins

    def samples(self) -> int:
        return int((1 << (2 * self.tus["log2n"].astype(np.int64))).sum())

    def densified(self) -> "ResidualBatch":
        """The same batch with the arena re-laid in descriptor order (TB after TB, as sorted): the
        layout P265_RES_DENSE_ARENA describes.  A packer that emits coefficients while it sorts
        produces this directly; here it is one gather over 16-coefficient units."""
        units = (1 << (2 * self.tus["log2n"].astype(np.int64))) >> 4
        new_off = np.concatenate(([0], np.cumsum(units)[:-1])) if len(units) else np.zeros(0, np.int64)
        total = int(units.sum())
        src = np.repeat(self.tus["coeff_off"].astype(np.int64) - new_off, units) + np.arange(total, dtype=np.int64)
        n_units = self.coeffs.size // 16
        arena = np.ascontiguousarray(self.coeffs[:n_units * 16].reshape(n_units, 16)[src]).reshape(-1)
        tus = self.tus.copy()
        tus["coeff_off"] = new_off.astype(np.uint32)
        return ResidualBatch(self.geom, tus, arena, self.scaling_factor, self.covers_all, self.sf_replicated, self.bins)

    def with_extents(self) -> "ResidualBatch":
        """The same batch with the zero-extent codes of its 16x16 / 32x32 TBs in the descriptors
        (`extent_codes`) and the list re-ordered by the ordering rule (whole work items by code pair)."""
        tus = self.tus.copy()
        set_extents(tus, *extent_codes(tus, self.coeffs))
        return ResidualBatch(self.geom, sort_by_size(tus, self.geom), self.coeffs, self.scaling_factor,
                             self.covers_all, self.sf_replicated, self.bins)

    def has_extents(self) -> bool:
        return bool((self.tus["rsvd"] >> TU_ZR_SHIFT).any())

    def dense_small_bins(self) -> bool:
        """True when, inside the 8x8 bin and inside the 4x4 bin, every TB's coefficients directly
        follow the previous TB's in the arena (what P265_RES_DENSE_ARENA asserts for the device
        entry point; the host entry point checks it itself).  Computed once per batch."""
        if getattr(self, "_dense", None) is None:
            b = self.bin_counts()
            ok, k = True, b[0] + b[1]
            for n_tb, units in ((b[2], 4), (b[3], 1)):
                off = self.tus["coeff_off"][k:k + n_tb].astype(np.int64)
                if n_tb and not np.array_equal(off, off[0] + units * np.arange(n_tb, dtype=np.int64)):
                    ok = False
                k += n_tb
            self._dense = ok
        return self._dense


def extent_code(last: np.ndarray, n: int) -> np.ndarray:
    """Zero-extent code of a TB side from the index of its last non-zero row / column (-1: none)."""
    return np.where(last < n // 4, 2, np.where(last < n // 2, 1, 0)).astype(np.uint16)


def extent_codes(tus: np.ndarray, coeffs: np.ndarray):
    """(zr, zc) per descriptor of a dense arena: the zero-extent codes of the 16x16 / 32x32 TBs
    (0 for the smaller sizes, which do not use them).  A parser has them for free (emit.py); this is
    the vectorised form for synthetic / already packed data."""
    zr = np.zeros(len(tus), np.uint16)
    zc = np.zeros(len(tus), np.uint16)
    l2 = tus["log2n"]
    for k in (5, 4):
        idx = np.nonzero(l2 == k)[0]
        if not idx.size:
            continue
        n = 1 << k
        nz = coeffs[(tus["coeff_off"][idx].astype(np.int64) * 16)[:, None] +
                    np.arange(n * n, dtype=np.int64)[None, :]].reshape(-1, n, n) != 0
        ar = np.arange(1, n + 1)
        zr[idx] = extent_code((nz.any(axis=2) * ar).max(axis=1) - 1, n)
        zc[idx] = extent_code((nz.any(axis=1) * ar).max(axis=1) - 1, n)
    return zr, zc


def set_extents(tus: np.ndarray, zr: np.ndarray, zc: np.ndarray) -> None:
    keep = np.uint16(0xFFFF ^ (3 << TU_ZR_SHIFT) ^ (3 << TU_ZC_SHIFT))
    tus["rsvd"] = (tus["rsvd"] & keep) | (zr.astype(np.uint16) << TU_ZR_SHIFT) | (zc.astype(np.uint16) << TU_ZC_SHIFT)


def size_kind_order(tus: np.ndarray, geom=None) -> np.ndarray:
    """Permutation that sorts descriptors largest TBs first -- the order `p265_residual_*` requires -- and,
    inside a size, arranges TBs for the kernels.  The one ordering rule of the packed format: the packer, the
    parser-side emitter and the synthetic workloads share it; any order inside a size is CORRECT, this one is
    fast.  Inside a size:
      * kinds (normal, DST, transform-skip, bypass) are clustered so that a warp's 32 lanes (32 small TBs) agree
        (transform-skip only while rare, see below);
      * 16x16 / 32x32 TBs: when the bit depths are known (`geom`), TBs whose dequantisation is the left-shift form
        of 8.6.3 (qP / 6 >= bdShift) are clustered as well (the kernels pick that slower form per work item);
        otherwise the order they came in (decoding order: the 2 / 4 TBs of a work item are neighbours), whole
        work items ordered by zero-extent code pair when descriptors carry codes;
      * 8x8 / 4x4 TBs: raster order of their plane (picture, component, y, x)."""
    # Transform-skip TBs (4x4 only) get their own cluster only while they are rare: pulled out of the list they
    # break the raster runs of the 4x4 bin (below), left in they make the warps that hold one run both paths.
    # Measured on B200 (round 2): 10 % of the 4x4 TBs (config 3) -> unclustered, the 4x4 bin 12 % faster and the
    # chain 3.6 %; 1 % (config 2) -> clustered, 1.2 % faster than unclustered.
    kinds = TU_DST | TU_SKIP | TU_BYPASS
    n4 = int((tus["log2n"] == 2).sum())
    if n4 and int(((tus["flags"] & TU_SKIP) != 0).sum()) > 0.03 * n4:
        kinds = TU_DST | TU_BYPASS
    key = (-(tus["log2n"].astype(np.int32)) * 32 + (tus["flags"] & kinds).astype(np.int32) * 2)
    if geom is not None:
        bd = np.where(tus["c_idx"] == 0, geom.bit_depth_y, geom.bit_depth_c).astype(np.int32)
        left_shift = (tus["qp"].astype(np.int32) // 6 >= bd + tus["log2n"].astype(np.int32) - 5) & \
            ((tus["flags"] & (TU_PRESCALED | TU_BYPASS)) == 0)
        # (big TBs only: in the small bins the raster runs below are worth more than a warp-uniform dequantisation
        # path -- with this sub-key the runs break at every quadrant with another qP and the raster gain is gone:
        # BASELINE config 2 0.1183 ms with it, 0.1144 without, same box)
        key = key + (left_shift & (tus["log2n"] >= 4)).astype(np.int32)
    # Inside a cluster: big TBs stay in decoding order (a work item = 2 / 4 neighbours of one quadrant); the 8x8 and
    # 4x4 TBs go in RASTER order of their plane (picture, component, y, x).  A work item of the small bins is 32 TBs,
    # one per lane, and every store instruction writes one row of each: in decoding (z-scan) order those 32 rows lie
    # in 8-16 different plane rows, in raster order they are pieces of the same few rows -- fewer, longer runs for
    # the L2 to merge and for DRAM to write.  Measured on B200 (round 2, profiles/r2_small_bins.txt): BASELINE
    # config 2 -3.7 % (8x8 bin -6.7 %, 4x4 bin -5.8 %), config 3 -0.9 %; the 16x16 bin does not care.
    small = tus["log2n"] <= 3
    pos = (((tus["pic"].astype(np.int64) << 2) | tus["c_idx"].astype(np.int64)) << 32) | \
        (tus["y"].astype(np.int64) << 16) | tus["x"].astype(np.int64)
    order = np.lexsort((np.where(small, pos, np.arange(len(tus), dtype=np.int64)), key))
    # Second level, only when descriptors carry zero-extent codes: whole WORK ITEMS of the big bins (2
    # consecutive 32x32 TBs, 4 consecutive 16x16 TBs, counted from the start of the bin -- P265_ITEM_TBS in the
    # header) are ordered by the item's code pair, the weakest promise among its TBs.  Measured on B200 (round
    # 2): (a) TBs sorted individually by code lose their spatial neighbours inside an item (an item stores the
    # rows of its TBs together: 64 / 128 contiguous bytes per row in decoding order) and the 16x16 bin ran 57 %
    # slower; (b) items left in decoding order run a different pair of passes every few items, all copies of
    # the passes are hot at once and the instruction cache thrashes (32x32 bin 28 % slower than without codes).
    # Items as units keep (a) and give every copy of the passes a long uninterrupted run.
    if (tus["rsvd"] >> TU_ZR_SHIFT).any():
        t = tus[order]
        special = (t["flags"] & (TU_BYPASS | TU_SKIP)) != 0
        zr = np.where(special, 0, (t["rsvd"] >> TU_ZR_SHIFT) & 3).astype(np.int32)
        zc = np.where(special, 0, (t["rsvd"] >> TU_ZC_SHIFT) & 3).astype(np.int32)
        for log2n, per in ((5, 2), (4, 4)):
            idx = np.flatnonzero(t["log2n"] == log2n)
            g = idx.size // per
            if g < 2:
                continue
            lo = int(idx[0])
            code = zr[lo:lo + g * per].reshape(g, per).min(axis=1) * 4 + zc[lo:lo + g * per].reshape(g, per).min(axis=1)
            perm = np.argsort(code, kind="stable")
            order[lo:lo + g * per] = order[lo:lo + g * per].reshape(g, per)[perm].reshape(-1)
    return order


def sort_by_size(tus: np.ndarray, geom=None) -> np.ndarray:
    return np.ascontiguousarray(tus[size_kind_order(tus, geom)])


def sf_offset(size_id: int, matrix_id: int) -> int:
    n = 4 << size_id
    return _SF_OFFSETS[size_id] + matrix_id * n * n


def sf_is_replicated(table: np.ndarray) -> bool:
    """True when the 16x16 / 32x32 matrices of a packed table are an 8x8 list up-sampled
    2x / 4x with only [0][0] (the DC value) allowed to differ -- what 7.4.5 produces for
    every conformant stream.  Enables the kernel's per-column factor path."""
    table = np.asarray(table, dtype=np.uint8).reshape(-1)
    if table.size != SF_BYTES:
        return False
    for size_id, count, rep in ((2, 6, 2), (3, 2, 4)):
        n = 4 << size_id
        for m in range(count):
            off = sf_offset(size_id, m)
            f = table[off:off + n * n].reshape(n, n).copy()
            f[0, 0] = f[0, 1]
            if not np.array_equal(f, np.kron(f[::rep, ::rep], np.ones((rep, rep), np.uint8))):
                return False
    return True


def pack_scaling_factor(sf: dict) -> np.ndarray:
    """{(sizeId, matrixId): (N, N) [x][y] array} -> device table uint8[4064], each
    matrix row-major [y][x].  `sf` uses the reference's indexing
    sps.scaling_factor[size_id][matrix_id][x][y] (scaling.py:44)."""
    out = np.full(SF_BYTES, 16, dtype=np.uint8)
    for (s, m), f in sf.items():
        n = 4 << s
        f = np.asarray(f)
        if f.shape != (n, n):
            raise ValueError("ScalingFactor[%d][%d] must be %dx%d" % (s, m, n, n))
        if f.min() < 1 or f.max() > 255:
            raise ValueError("ScalingFactor entries must be in 1..255")
        off = sf_offset(s, m)
        out[off:off + n * n] = f.T.reshape(-1)
    return out


# ------------------------------------------------------------- packed coefficient stream
@dataclass
class PackedResidualBatch:
    """A residual batch whose coefficients travel as the packed stream of include/p265_b200.h
    (per TB: significance bitmap + non-zero levels, int8 when they fit) instead of a dense int16
    arena.  `tus[i].coeff_off` = byte offset of TB i's record / 4.  This is what a parser emits
    naturally -- it stores (position, level) pairs (tu.py:331) -- and what crosses PCIe."""
    geom: PicGeom
    tus: np.ndarray                         # TU_DESC sorted like ResidualBatch.tus; flags may carry TU_LEVELS8
    stream: np.ndarray                      # uint8
    scaling_factor: np.ndarray | None = None
    covers_all: bool = False
    sf_replicated: bool | None = None
    bins: tuple | None = None

    def __post_init__(self):
        if self.scaling_factor is not None and self.sf_replicated is None:
            self.sf_replicated = sf_is_replicated(self.scaling_factor)

    bin_counts = ResidualBatch.bin_counts
    samples = ResidualBatch.samples


def pack_coefficients(tus: np.ndarray, coeffs: np.ndarray):
    """(descriptors indexing a dense arena, arena) -> (descriptors indexing a packed stream, stream).
    Records are laid out in descriptor order.  Vectorised per TB size; the parser-side emitter
    (emit.PictureSink) writes the same format TB by TB."""
    tus = tus.copy()
    n_tb = len(tus)
    l2 = tus["log2n"].astype(np.int64)
    nnz = np.zeros(n_tb, np.int64)
    wide = np.zeros(n_tb, bool)
    per_size = {}
    for k in (5, 4, 3, 2):
        idx = np.nonzero(l2 == k)[0]
        if not idx.size:
            continue
        nn = 1 << (2 * k)
        blocks = coeffs[(tus["coeff_off"][idx].astype(np.int64) * 16)[:, None] + np.arange(nn, dtype=np.int64)[None, :]]
        nz = blocks != 0
        nnz[idx] = nz.sum(axis=1)
        wide[idx] = ((blocks > 127) | (blocks < -128)).any(axis=1)
        per_size[k] = (idx, blocks, nz)
    bm_bytes = (1 << (2 * l2)) >> 3
    rec = (bm_bytes + nnz * np.where(wide, 2, 1) + 3) & ~3
    rec_off = np.concatenate(([0], np.cumsum(rec)[:-1])) if n_tb else np.zeros(0, np.int64)
    stream = np.zeros(int(rec.sum()), np.uint8)
    for k, (idx, blocks, nz) in per_size.items():
        nn = 1 << (2 * k)
        bm = np.packbits(nz, axis=1, bitorder="little")
        stream[(rec_off[idx][:, None] + np.arange(nn // 8, dtype=np.int64)[None, :]).ravel()] = bm.ravel()
        vals = blocks[nz]                                    # TB after TB, raster order inside a TB
        cnt = nnz[idx]
        start = np.concatenate(([0], np.cumsum(cnt)[:-1]))
        rank = np.arange(int(cnt.sum()), dtype=np.int64) - np.repeat(start, cnt)
        w_rep = np.repeat(wide[idx], cnt)
        pos = np.repeat(rec_off[idx] + nn // 8, cnt) + rank * np.where(w_rep, 2, 1)
        u = vals.astype(np.int16).view(np.uint16)
        stream[pos] = (u & 0xFF).astype(np.uint8)
        stream[pos[w_rep] + 1] = (u[w_rep] >> 8).astype(np.uint8)
    if stream.size >> 2 > 0xFFFFFFFF:
        raise ValueError("packed stream too large for 32-bit record offsets")
    tus["coeff_off"] = (rec_off >> 2).astype(np.uint32)
    # number of levels in the record; the zero-extent codes (the host's ordering key) travel along,
    # the device derives its own from the bitmap
    tus["rsvd"] = (tus["rsvd"] & np.uint16(0xFFFF ^ TU_LEVELS_MASK)) | nnz.astype(np.uint16)
    tus["flags"] = (tus["flags"] & ~np.uint8(TU_LEVELS8)) | np.where(wide, 0, TU_LEVELS8).astype(np.uint8)
    return tus, stream


def unpack_coefficients(tus: np.ndarray, stream: np.ndarray):
    """Inverse of pack_coefficients: (descriptors indexing a dense arena laid out in descriptor order,
    int16 arena).  Host-side format conversion only (the per-TB drop-in keeps a d[] cache that is
    indexed like a dense arena); the device does its own expansion (csrc/coeffs.cu)."""
    tus = tus.copy()
    stream = np.asarray(stream, dtype=np.uint8)
    l2 = tus["log2n"].astype(np.int64)
    nn_all = 1 << (2 * l2)
    off = np.concatenate(([0], np.cumsum(nn_all)[:-1])) if len(tus) else np.zeros(0, np.int64)
    arena = np.zeros(int(nn_all.sum()), np.int16)
    rec = tus["coeff_off"].astype(np.int64) * 4
    narrow = (tus["flags"] & TU_LEVELS8) != 0
    for k in (5, 4, 3, 2):
        idx = np.nonzero(l2 == k)[0]
        if not idx.size:
            continue
        nn = 1 << (2 * k)
        bm = stream[rec[idx][:, None] + np.arange(nn // 8, dtype=np.int64)[None, :]]
        mask = np.unpackbits(bm, axis=1, bitorder="little").astype(bool)
        cnt = mask.sum(axis=1)
        start = np.concatenate(([0], np.cumsum(cnt)[:-1]))
        rank = np.arange(int(cnt.sum()), dtype=np.int64) - np.repeat(start, cnt)
        nar = np.repeat(narrow[idx], cnt)
        pos = np.repeat(rec[idx] + nn // 8, cnt) + rank * np.where(nar, 1, 2)
        lo = stream[pos].astype(np.uint16)
        hi = np.where(nar, np.where(lo & 0x80, 0xFF, 0), stream[np.minimum(pos + 1, stream.size - 1)]).astype(np.uint16)
        vals = (lo | (hi << 8)).astype(np.uint16).view(np.int16)
        blocks = np.zeros((idx.size, nn), np.int16)
        blocks[mask] = vals
        arena[(off[idx][:, None] + np.arange(nn, dtype=np.int64)[None, :]).ravel()] = blocks.ravel()
    tus["coeff_off"] = (off >> 4).astype(np.uint32)
    tus["rsvd"] = tus["rsvd"] & np.uint16(0xFFFF ^ TU_LEVELS_MASK)
    tus["flags"] = tus["flags"] & np.uint8(0xFF ^ TU_LEVELS8)
    return tus, arena


def _unpacked(self) -> ResidualBatch:
    tus, arena = unpack_coefficients(self.tus, self.stream)
    return ResidualBatch(self.geom, tus, arena, self.scaling_factor, self.covers_all, self.sf_replicated, self.bins)


PackedResidualBatch.unpacked = _unpacked


def _packed(self) -> PackedResidualBatch:
    """The same batch with its coefficients as a packed stream (host -> device transport)."""
    tus, stream = pack_coefficients(self.tus, self.coeffs)
    return PackedResidualBatch(self.geom, tus, stream, self.scaling_factor, self.covers_all, self.sf_replicated, self.bins)


ResidualBatch.packed = _packed
