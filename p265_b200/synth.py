"""Synthetic workloads of the BASELINE.json configs.

There is no network and the reference ships one 352x288 bitstream, so the 1080p / 4K
numbers are measured on synthetic TB / coefficient mixes of the documented shape:

  config 2  1920x1088 4:2:0  8-bit, luma area 4x4 23 % (DST) / 8x8 29 % / 16x16 31 % /
            32x32 17 %, flat lists, qP in {22,27,32,37}, TS on 1 % of 4x4   rng 26502
  config 3  3840x2176 4:2:0 10-bit, 5 / 10 / 25 / 60 %, default scaling lists, qP in
            34..49, TS on 10 % of 4x4, bypass on 1 % of CUs                  rng 26503
  config 4  3840x2160 10-bit reconstructed picture + per-CTB SAO parameters  rng 26504
  config 5  streams 26510.. = config 3 + config 4 per picture

Every TB of a picture is coded (the planes are covered completely).  Coefficients: a
non-zero mask of density ~0.20 biased to low frequencies, P(nz at (x,y)) ~
exp(-(x+y)/(N/4)), Laplacian magnitudes (scale 6, DC scale 40).
"""
from __future__ import annotations

import numpy as np

from .picture import (AVAIL_ALL, SAO_CTB, TU_BYPASS, TU_DESC, TU_DST, TU_INTRA, TU_SKIP,
                      PicGeom, ResidualBatch, pack_scaling_factor, sort_by_size)
from .scaling_list import default_scaling_factor

CONFIGS = {
    "1080p8": dict(width=1920, height=1088, bit_depth=8, shares=(0.23, 0.29, 0.31, 0.17),
                   qps=(22, 27, 32, 37), ts_frac=0.01, bypass_frac=0.0, scaling_lists=False,
                   seed=26502),
    "4k10": dict(width=3840, height=2176, bit_depth=10, shares=(0.05, 0.10, 0.25, 0.60),
                 qps=tuple(range(34, 50)), ts_frac=0.10, bypass_frac=0.01, scaling_lists=True,
                 seed=26503),
}


# "4k10_lowfreq": config 3 with the coefficients of every 16x16 / 32x32 TB confined to its first N >> zr
# rows and N >> zc columns, (zr, zc) drawn per 32x32 quadrant from the distribution measured on the 16x16 TBs of the
# reference's only real stream (sanity.bin, 368 TBs: tests/test_extents.py prints it).  The SURVEY 8(d)
# coefficient model itself leaves nothing to skip (a level in the last quarter of the rows of 99.85 % of its
# 32x32 TBs); real streams do, and this variant is what the zero-aware passes are measured on.
SANITY_EXTENT_MIX = (((0, 0), 0.41), ((2, 2), 0.29), ((1, 0), 0.07), ((2, 1), 0.055), ((0, 2), 0.05),
                     ((1, 1), 0.045), ((1, 2), 0.03), ((0, 1), 0.03), ((2, 0), 0.02))
CONFIGS["4k10_lowfreq"] = dict(CONFIGS["4k10"], extent_mix=SANITY_EXTENT_MIX)


def _nz_prob(n: int, density: float) -> np.ndarray:
    yy, xx = np.mgrid[0:n, 0:n]
    p = np.exp(-(xx + yy) / (n / 4.0))
    return np.minimum(1.0, p * (density * n * n / p.sum()))


def _coeff_blocks(rng, count: int, n: int, stress: bool, pick=None) -> np.ndarray:
    """`pick`: (count, 2) zero-extent codes (zr, zc) per TB or None."""
    if count == 0:
        return np.zeros((0, n, n), np.int16)
    box = None
    if pick is not None and n >= 16:   # per TB: coefficients only in rows < n >> zr, columns < n >> zc
        ar = np.arange(n)
        box = (ar[None, :, None] < (n >> pick[:, 0])[:, None, None]) & (ar[None, None, :] < (n >> pick[:, 1])[:, None, None])
    if stress:
        lv = rng.integers(-32768, 32768, (count, n, n), dtype=np.int64).astype(np.int16)
        return lv if box is None else np.where(box, lv, 0).astype(np.int16)
    mask = rng.random((count, n, n), dtype=np.float32) < _nz_prob(n, 0.20).astype(np.float32)
    if box is not None:
        mask &= box
    mag = rng.laplace(0.0, 6.0, (count, n, n)).astype(np.float32)
    mag[:, 0, 0] = rng.laplace(0.0, 40.0, count)
    lv = np.where(mask, np.rint(mag), 0.0)
    return np.clip(lv, -32767, 32767).astype(np.int16)


def _quadrant_tbs(qx, qy, cls):
    """Luma TB origins (x, y, log2n) for 32x32 quadrants split into class `cls`
    (0: 4x4, 1: 8x8, 2: 16x16, 3: 32x32)."""
    log2n = 2 + cls
    n = 1 << log2n
    per = 32 // n
    oy, ox = np.mgrid[0:per, 0:per]
    # z-order inside the quadrant is irrelevant for the result; keep raster order
    x = (qx[:, None] + (ox.reshape(-1) * n)[None, :]).reshape(-1)
    y = (qy[:, None] + (oy.reshape(-1) * n)[None, :]).reshape(-1)
    return x, y, log2n


def residual_picture(cfg: dict, rng, pic: int = 0, stress: bool = False):
    """TU descriptors (unsorted, coeff_off relative to 0) + coefficient arena of one
    synthetic picture.  Returns (tus, coeffs)."""
    w, h = cfg["width"], cfg["height"]
    bd = cfg["bit_depth"]
    qx, qy = np.meshgrid(np.arange(0, w, 32), np.arange(0, h, 32))
    qx, qy = qx.reshape(-1), qy.reshape(-1)
    nq = qx.size
    cls = rng.choice(4, size=nq, p=np.array(cfg["shares"]) / sum(cfg["shares"]))
    qp_q = rng.choice(np.array(cfg["qps"]), size=nq)
    byp_q = rng.random(nq) < cfg["bypass_frac"]
    recs, arenas, off = [], [], 0
    # extent_mix: one (zr, zc) pair per 32x32 quadrant -- how sparse a block is follows the local texture, so
    # the TBs of a quadrant (the four 16x16 TBs of one work item of the kernels) share it
    quad_codes = None
    if cfg.get("extent_mix") is not None:
        mix = cfg["extent_mix"]
        codes = np.array([c for c, _ in mix])
        quad_codes = codes[rng.choice(len(codes), size=nq, p=np.array([p for _, p in mix]) / sum(p for _, p in mix))]
    quads_per_row = (w + 31) // 32

    def emit(x, y, log2n, c_idx, qp, flags):
        nonlocal off
        cnt, n = x.size, 1 << log2n
        if cnt == 0:
            return
        r = np.zeros(cnt, dtype=TU_DESC)
        r["x"], r["y"], r["log2n"], r["c_idx"] = x, y, log2n, c_idx
        r["qp"], r["flags"], r["pic"] = qp, flags, pic
        r["coeff_off"] = (off + np.arange(cnt, dtype=np.int64) * n * n) >> 4
        pick = None
        if quad_codes is not None and n >= 16:
            sh = 1 if c_idx else 0
            pick = quad_codes[((y << sh) >> 5) * quads_per_row + ((x << sh) >> 5)]
        blocks = _coeff_blocks(rng, cnt, n, stress, pick)
        byp = (flags & TU_BYPASS) != 0
        if byp.any():      # bypass "coefficients" are residual samples: keep them small
            blocks[byp] = np.clip(blocks[byp], -(1 << bd), (1 << bd) - 1)
        recs.append(r)
        arenas.append(blocks.reshape(-1))
        off += cnt * n * n

    for c in range(4):
        sel = cls == c
        if not sel.any():
            continue
        x, y, log2n = _quadrant_tbs(qx[sel], qy[sel], c)
        per_q = (32 >> log2n) ** 2
        qp = np.repeat(qp_q[sel], per_q)
        byp = np.repeat(byp_q[sel], per_q)
        fl = np.full(x.size, TU_INTRA, np.uint8) | np.where(byp, TU_BYPASS, 0).astype(np.uint8)
        if log2n == 2:
            fl |= TU_DST
            ts = (rng.random(x.size) < cfg["ts_frac"]) & ~byp
            fl |= np.where(ts, TU_SKIP, 0).astype(np.uint8)
        emit(x, y, log2n, 0, qp, fl)
        # chroma: half-size TBs; four 4x4 luma TBs share one 4x4 Cb + Cr pair
        if log2n == 2:
            keep = ((x & 7) == 0) & ((y & 7) == 0)
            cx, cy, cl2 = x[keep] >> 1, y[keep] >> 1, 2
            cqp, cbyp = qp[keep], byp[keep]
        else:
            cx, cy, cl2 = x >> 1, y >> 1, log2n - 1
            cqp, cbyp = qp, byp
        for c_idx in (1, 2):
            cfl = np.full(cx.size, TU_INTRA, np.uint8) | np.where(cbyp, TU_BYPASS, 0).astype(np.uint8)
            if cl2 == 2:
                ts = (rng.random(cx.size) < cfg["ts_frac"]) & ~cbyp
                cfl |= np.where(ts, TU_SKIP, 0).astype(np.uint8)
            # chroma qP: one below luma as in the fixture stream (qp.log: 32 / 31)
            emit(cx, cy, cl2, c_idx, np.maximum(cqp - 1, 0), cfl)
    return np.concatenate(recs), np.concatenate(arenas)


def residual_batch(name: str, n_pics: int = 1, seed: int | None = None, stress: bool = False,
                   n_unique: int | None = None, extents: bool = False) -> ResidualBatch:
    """A batch of `n_pics` synthetic pictures of config `name`.  `n_unique` < n_pics
    generates that many distinct pictures and repeats them (distinct memory, same
    values) to keep host generation time down for large benchmark batches.  `extents`: descriptors
    carry the zero-extent codes of their coefficients and are ordered by them (what the parser-side
    emitter hands over, picture.ResidualBatch.with_extents)."""
    cfg = CONFIGS[name]
    rng = np.random.default_rng(cfg["seed"] if seed is None else seed)
    n_unique = n_pics if n_unique is None else min(n_unique, n_pics)
    uniq = [residual_picture(cfg, rng, 0, stress) for _ in range(n_unique)]
    tus, arenas, off = [], [], 0
    for p in range(n_pics):
        t, a = uniq[p % n_unique]
        t = t.copy()
        t["pic"] = p
        t["coeff_off"] = t["coeff_off"].astype(np.int64) + (off >> 4)
        tus.append(t)
        arenas.append(a)
        off += a.size
    geom = PicGeom(cfg["width"], cfg["height"], n_pics, cfg["bit_depth"], cfg["bit_depth"])
    sf = pack_scaling_factor(default_scaling_factor()) if cfg["scaling_lists"] else None
    # same ordering rule as the product packer (picture.sort_by_size: size, then kind)
    batch = ResidualBatch(geom=geom, tus=sort_by_size(np.concatenate(tus), geom),
                          coeffs=np.concatenate(arenas), scaling_factor=sf, covers_all=True)
    return batch.with_extents() if extents else batch


# ------------------------------------------------------------------------------ SAO
def sao_picture(width: int, height: int, bit_depth: int, rng, geom: PicGeom, buf: np.ndarray,
                pic: int) -> None:
    """Fill picture `pic` of `buf` with a smooth gradient + Gaussian noise (sigma 12 at
    10 bits) so that all five edge categories occur."""
    maxv = (1 << bit_depth) - 1
    scale = maxv / 1023.0
    for c in range(3):
        h, w = geom.plane_shape(c)
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        base = (0.15 + 0.7 * (xx / w * 0.6 + yy / h * 0.4)) * maxv
        if c:
            base = maxv - base if c == 2 else base * 0.9
        noise = rng.normal(0.0, 12.0 * scale, (h, w)).astype(np.float32)
        geom.plane_view(buf, pic, c)[:] = np.clip(np.rint(base + noise), 0, maxv).astype(buf.dtype)


def sao_params(ctbs_h: int, ctbs_w: int, bit_depth: int, rng, two_slices: bool = False) -> np.ndarray:
    """Per-CTB parameters of config 4: type uniform {off, band, edge} for luma and for
    the chroma pair (Cb/Cr share type and class, sao.py:58-59,76-77), class / band
    position uniform, offsets uniform 0..cMax, 25 % merge-left, 15 % merge-up."""
    c_max = (1 << (min(bit_depth, 10) - 5)) - 1
    shift = bit_depth - min(bit_depth, 10)
    tab = np.zeros((ctbs_h, ctbs_w), dtype=SAO_CTB)
    tab["avail"] = AVAIL_ALL
    split = (ctbs_h // 2) * ctbs_w + ctbs_w // 3 if two_slices else None   # 2nd slice start
    for ry in range(ctbs_h):
        for rx in range(ctbs_w):
            addr = ry * ctbs_w + rx
            e = tab[ry, rx]
            u = rng.random()
            first_in_slice = addr == 0 or addr == split
            left_ok = rx > 0 and not (split is not None and addr == split)
            up_ok = ry > 0 and not (split is not None and addr - ctbs_w < split <= addr)
            if u < 0.25 and left_ok and not first_in_slice:
                for f in ("type", "band_pos", "eo_class", "offset_val"):
                    e[f] = tab[ry, rx - 1][f]
                continue
            if u < 0.40 and up_ok:
                for f in ("type", "band_pos", "eo_class", "offset_val"):
                    e[f] = tab[ry - 1, rx][f]
                continue
            t_l, t_c = rng.integers(0, 3), rng.integers(0, 3)
            cls_l, cls_c = rng.integers(0, 4), rng.integers(0, 4)
            for c in range(3):
                t = t_l if c == 0 else t_c
                e["type"][c] = t
                if t == 0:
                    continue
                mag = rng.integers(0, c_max + 1, 4) << shift
                if t == 1:
                    e["band_pos"][c] = rng.integers(0, 32)
                    e["offset_val"][c] = mag * rng.choice((-1, 1), 4)
                else:
                    e["eo_class"][c] = cls_l if c == 0 else cls_c
                    e["offset_val"][c] = mag * np.array((1, 1, -1, -1))
    if two_slices:
        from .packer import ctb_availability
        addr = np.arange(ctbs_h * ctbs_w).reshape(ctbs_h, ctbs_w)
        slice_addr = np.where(addr >= split, split, 0)
        tab["avail"] = ctb_availability(slice_addr, {0: 1, int(split): 0},
                                        np.zeros_like(addr), True)
    return tab


def sao_batch(width: int = 3840, height: int = 2160, bit_depth: int = 10, n_pics: int = 1,
              ctb_log2: int = 6, seed: int = 26504, two_slices: bool = False, n_unique=None):
    """Returns (geom, rec_buffer, params[n_pics, ctbs_h, ctbs_w])."""
    rng = np.random.default_rng(seed)
    geom = PicGeom(width, height, n_pics, bit_depth, bit_depth)
    dtype = np.uint8 if bit_depth <= 8 else np.uint16
    buf = np.zeros(geom.total_elems(), dtype=dtype)
    ctb = 1 << ctb_log2
    ch, cw = (height + ctb - 1) // ctb, (width + ctb - 1) // ctb
    n_unique = n_pics if n_unique is None else min(n_unique, n_pics)
    params = np.zeros((n_pics, ch, cw), dtype=SAO_CTB)
    for p in range(n_pics):
        if p < n_unique:
            sao_picture(width, height, bit_depth, rng, geom, buf, p)
            params[p] = sao_params(ch, cw, bit_depth, rng, two_slices)
        else:
            q = p % n_unique
            buf[p * geom.pic_stride:(p + 1) * geom.pic_stride] = \
                buf[q * geom.pic_stride:(q + 1) * geom.pic_stride]
            params[p] = params[q]
    return geom, buf, params


# ------------------------------------------------------------------------ deblocking
def deblock_picture(bit_depth: int, rng, geom: PicGeom, buf: np.ndarray, pic: int) -> None:
    """Blocky reconstructed picture: smooth gradient, a DC step per 8x8 block (so edges
    exist), light noise -- strong, weak and "no filtering" decisions all occur."""
    maxv = (1 << bit_depth) - 1
    scale = maxv / 255.0
    for c in range(3):
        h, w = geom.plane_shape(c)
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        base = (0.2 + 0.6 * (xx / w * 0.5 + yy / h * 0.5)) * maxv
        step = rng.normal(0.0, 3.0 * scale, ((h + 7) // 8, (w + 7) // 8)).astype(np.float32)
        step = np.kron(step, np.ones((8, 8), np.float32))[:h, :w]
        noise = rng.normal(0.0, 0.8 * scale, (h, w)).astype(np.float32)
        geom.plane_view(buf, pic, c)[:] = np.clip(np.rint(base + step + noise), 0, maxv).astype(buf.dtype)


def deblock_edge_map(width: int, height: int, ctb_log2: int, rng, qp_lo: int = 22, qp_hi: int = 45,
                     no_filter_frac: float = 0.02, dense: bool = False):
    """Synthetic edge map + per-CTB parameters: an intra-like TU grid (every 8x8 edge is a
    transform edge with probability depending on a random 8/16/32 block structure; `dense`
    = every 8x8 edge), Bs 2 mostly, some Bs 1, QpY per 16x16 region, slice offsets per CTB."""
    from .picture import DBK_BS_H0, DBK_BS_H1, DBK_BS_V0, DBK_BS_V1, DBK_CTB, DBK_NO_FILTER, DBK_QP_SHIFT
    w8, h8 = width // 8, height // 8
    # block structure: each 32x32 region is one 32x32, four 16x16 or sixteen 8x8 blocks
    kind = rng.choice((8, 16, 32), size=((h8 + 3) // 4, (w8 + 3) // 4), p=(0.3, 0.3, 0.4))
    kind = np.kron(kind, np.ones((4, 4), np.int64))[:h8, :w8]
    if dense:
        kind[:] = 8
    by, bx = np.mgrid[0:h8, 0:w8]
    v_on = (bx * 8) % kind == 0
    h_on = (by * 8) % kind == 0
    v_on[:, 0] = False
    h_on[0, :] = False

    def bs(on):
        s = rng.choice((0, 1, 2), size=(2,) + on.shape, p=(0.05, 0.15, 0.8))
        return np.where(on[None], s, 0).astype(np.uint16)
    bv, bh = bs(v_on), bs(h_on)
    qp = rng.integers(qp_lo, qp_hi + 1, ((h8 + 1) // 2, (w8 + 1) // 2))
    qp = np.kron(qp, np.ones((2, 2), np.int64))[:h8, :w8]
    blk = (bv[0] << DBK_BS_V0) | (bv[1] << DBK_BS_V1) | (bh[0] << DBK_BS_H0) | (bh[1] << DBK_BS_H1) | \
        ((qp & 0x7F) << DBK_QP_SHIFT).astype(np.uint16)
    blk = blk.astype(np.uint16)
    blk[rng.random((h8, w8)) < no_filter_frac] |= DBK_NO_FILTER
    cs = 1 << ctb_log2
    ctb = np.zeros(((height + cs - 1) // cs, (width + cs - 1) // cs), dtype=DBK_CTB)
    ctb["beta_offset_div2"] = rng.integers(-6, 7, ctb.shape)
    ctb["tc_offset_div2"] = rng.integers(-6, 7, ctb.shape)
    ctb["cb_qp_offset"] = rng.integers(-4, 5)
    ctb["cr_qp_offset"] = rng.integers(-4, 5)
    return blk, ctb


def deblock_batch(width: int = 3840, height: int = 2160, bit_depth: int = 10, n_pics: int = 1,
                  ctb_log2: int = 6, seed: int = 26506, dense: bool = False, n_unique=None):
    """Returns (geom, rec_buffer, blk[n_pics, h/8, w/8], ctb[n_pics, ctbs_h, ctbs_w])."""
    from .picture import DBK_CTB
    rng = np.random.default_rng(seed)
    geom = PicGeom(width, height, n_pics, bit_depth, bit_depth)
    dtype = np.uint8 if bit_depth <= 8 else np.uint16
    buf = np.zeros(geom.total_elems(), dtype=dtype)
    cs = 1 << ctb_log2
    qp_off = 6 * (bit_depth - 8) * 0       # QpY itself does not include QpBdOffset
    blk = np.zeros((n_pics, height // 8, width // 8), np.uint16)
    ctb = np.zeros((n_pics, (height + cs - 1) // cs, (width + cs - 1) // cs), DBK_CTB)
    n_unique = n_pics if n_unique is None else min(n_unique, n_pics)
    for p in range(n_pics):
        if p < n_unique:
            deblock_picture(bit_depth, rng, geom, buf, p)
            blk[p], ctb[p] = deblock_edge_map(width, height, ctb_log2, rng, 22 + qp_off, 45 + qp_off, dense=dense)
        else:
            q = p % n_unique
            buf[p * geom.pic_stride:(p + 1) * geom.pic_stride] = buf[q * geom.pic_stride:(q + 1) * geom.pic_stride]
            blk[p], ctb[p] = blk[q], ctb[q]
    return geom, buf, blk, ctb
