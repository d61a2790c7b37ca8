"""Host side of the residual path with the reference's per-TB function surface.

    inverse_scaling(pu, x0, y0, log2size)      scaling.py:4-47
    inverse_transform(pu, x0, y0, log2size)    transform.py:89-109
    inverse_transform_1d(x, log2size, tr_type) transform.py:74-87

`pu` is the reference's IntraPu (intra.py:24-37) or anything with the same fields.
Both functions mutate the caller-owned `pu.scaled_samples` / `pu.transformed_samples`
views in place and return None, exactly like the reference.

Two ways to run:
  * per TB (works with the reference's decoder unchanged): each call packs one TB and
    runs it on the GPU -- correct, slow, useful for parity;
  * batched: `flush_picture(img, sps)` packs every coded TB of a parsed
    picture once, runs one residual launch and remembers the planes; the per-TB calls
    then only copy their block out of the planes.

Everything is computed by the sm_100a kernels through the C-ABI; there is no CPU path.
"""
from __future__ import annotations

import numpy as np

from . import emit, packer
from .engine import get_engine
from .picture import (TU_BYPASS, TU_DESC, TU_DST, TU_INTRA, TU_PRESCALED, TU_SKIP, PicGeom, ResidualBatch,
                      pack_scaling_factor)


#: "spec" = H.265 8.6.2-8.6.4 (default).  "ref_literal" = transform.py:89-109 exactly as
#: written -- parity tests only, never benchmarked.
MODE = "spec"


def set_mode(mode: str) -> None:
    global MODE
    if mode not in ("spec", "ref_literal"):
        raise ValueError("mode must be 'spec' or 'ref_literal'")
    MODE = mode


# ------------------------------------------------------------------ picture cache
class PictureResidual:
    """Residual planes + d[] arena of one flushed picture."""

    def __init__(self, batch: ResidualBatch, planes: np.ndarray, scaled: np.ndarray):
        self.batch, self.planes, self.scaled = batch, planes, scaled
        self.index = {(int(t["c_idx"]), int(t["x"]), int(t["y"])): i for i, t in enumerate(batch.tus)}

    def block(self, c_idx: int, x: int, y: int, n: int):
        """(d_yx, r_yx) of the TB at plane position (x, y) or None when it is not coded."""
        i = self.index.get((c_idx, x, y))
        if i is None:
            return None
        t = self.batch.tus[i]
        off = int(t["coeff_off"]) * 16
        d = self.scaled[off:off + n * n].reshape(n, n)
        r = self.batch.geom.plane_view(self.planes, int(t["pic"]), c_idx)[y:y + n, x:x + n]
        return d, r


def sps_scaling_table(sps, pps=None):
    """Packed 4064-byte table or None when scaling lists are off (scaling.py:32-33).

    `sps.scaling_factor[size_id][matrix_id][x][y]` (the attribute scaling.py:44 reads) wins
    when a caller has set it; the reference itself never does, so otherwise
    the table comes from the decoded scaling_list_data() of the PPS / SPS (p265_b200's `sld`
    drop-in) or the default lists (`scaling_list.active_table`)."""
    if not getattr(sps, "scaling_list_enabled_flag", 0):
        return None
    src = getattr(sps, "scaling_factor", None)
    if src is None:
        from . import scaling_list
        return scaling_list.active_table(sps, pps)
    sf = {}
    for s in range(4):
        for m in range(2 if s == 3 else 6):
            f = src[s][m]
            if f is not None:
                sf[(s, m)] = np.asarray(f)
    return pack_scaling_factor(sf)


def flush_picture(img, sps, device: int = 0, pps=None) -> PictureResidual:
    """Batched path: one residual launch (+ one dequant launch for `scaled_samples`) for
    all coded TBs of a parsed picture; the result is attached to `img`."""
    eng = get_engine(device)
    table = sps_scaling_table(sps, pps)
    packed = emit.take(img, sps, table)
    if packed is not None:
        # the parser was hooked (emit.hook_parser): the picture's packed stream is already there --
        # no walk over the finished tree; the stream is what crosses PCIe
        planes = eng.residual(packed)
        batch = packed.unpacked()          # the d[] cache of the per-TB drop-in is indexed like a dense arena
    else:
        batch = packer.pack_pictures([img], sps, table)
        planes = eng.residual(batch)
    scaled = eng.dequant(batch)
    img._p265_b200_residual = PictureResidual(batch, planes, scaled)
    return img._p265_b200_residual


# ------------------------------------------------------------------ per-TB helpers
def _qp(pu) -> int:
    sps = pu.cu.ctx.sps                                   # scaling.py:13-18
    if pu.c_idx == 0:
        return int(pu.cu.qp_y + sps.qp_bd_offset_y)
    if pu.c_idx == 1:
        return int(pu.cu.qp_cb + sps.qp_bd_offset_c)
    return int(pu.cu.qp_cr + sps.qp_bd_offset_c)


def _bit_depths(pu):
    sps = pu.cu.ctx.sps
    return int(sps.bit_depth_y), int(sps.bit_depth_c)


def _find_leaf(pu, x0, y0):
    tu = getattr(pu.cu, "tu", None)
    if tu is None or not hasattr(tu, "get_leaves"):
        return None
    for leaf in tu.get_leaves():
        if leaf.contain(x0, y0):
            return leaf
    return None


def _levels_xy(pu, x0, y0, log2size):
    """TransCoeffLevel block [x][y] of the TB.  Fast path: the leaf's own array
    (tu.py:87-90); otherwise the reference's per-coefficient accessor (tu.py:667-684)."""
    n = 1 << log2size
    leaf = _find_leaf(pu, x0, y0)
    if leaf is not None and hasattr(leaf, "trans_coeff_level"):
        arr = np.asarray(leaf.trans_coeff_level[pu.c_idx])
        if arr.shape == (n, n) and leaf.x == x0 and leaf.y == y0:
            return arr.astype(np.int64)
    tu = pu.cu.tu
    out = np.zeros((n, n), dtype=np.int64)
    for x in range(n):
        for y in range(n):
            out[x, y] = tu.get_trans_coeff_level(x0 + x, y0 + y, pu.c_idx)
    return out


def _flags(pu, x0, y0, log2size) -> int:
    fl = 0
    try:
        intra = bool(pu.cu.is_intra_mode())
    except Exception:
        intra = True
    if intra:
        fl |= TU_INTRA
    if log2size == 2 and pu.c_idx == 0 and (intra or MODE == "ref_literal"):
        fl |= TU_DST                                      # transform.py:97 (+ 8.6.4.2: intra only)
    if getattr(pu.cu, "cu_transquant_bypass_flag", 0):
        fl |= TU_BYPASS
    leaf = _find_leaf(pu, x0, y0)
    ts = getattr(leaf, "transform_skip_flag", None) if leaf is not None else None
    if ts is not None and log2size == 2 and ts[pu.c_idx]:
        fl |= TU_SKIP
    return fl


_GEOM = {}


def _tb_batch(pu, levels_yx, log2size, flags, qp, table):
    bdy, bdc = _bit_depths(pu)
    key = (bdy, bdc)
    geom = _GEOM.get(key)
    if geom is None:
        geom = _GEOM[key] = PicGeom(64, 64, 1, bdy, bdc)
    tus = np.zeros(1, dtype=TU_DESC)
    tus["log2n"], tus["c_idx"], tus["qp"], tus["flags"] = log2size, pu.c_idx, qp, flags
    lv = np.clip(levels_yx, -32768, 32767).astype(np.int16).reshape(-1)
    return ResidualBatch(geom, tus, lv, table, covers_all=False)


def _cached(pu, x0, y0, log2size):
    img = getattr(pu.cu.ctx, "img", None)
    for holder in (img, getattr(pu.cu, "_p265_b200_img", None)):
        cache = getattr(holder, "_p265_b200_residual", None) if holder is not None else None
        if cache is not None:
            sh = 0 if pu.c_idx == 0 else 1
            return cache.block(pu.c_idx, x0 >> sh, y0 >> sh, 1 << log2size)
    return None


# ------------------------------------------------------------------ reference surface
def inverse_scaling(pu, x0, y0, log2size):
    """Scaling process for transform coefficients (8.6.3), scaling.py:4-47.

    Differences from the reference, both required by the north star: scaling lists work
    (the reference never fills sps.scaling_factor) and cu_transquant_bypass stores the
    levels unchanged instead of raising ValueError("Unimplemented yet.")."""
    if not 2 <= int(log2size) <= 5:
        raise ValueError("log2size must be in 2..5")
    n = 1 << log2size
    sx, sy = x0 - pu.origin_x, y0 - pu.origin_y
    d = pu.scaled_samples[sx:sx + n, sy:sy + n]
    hit = _cached(pu, x0, y0, log2size)
    if hit is not None:
        d[...] = hit[0].T
        return
    levels_xy = _levels_xy(pu, x0, y0, log2size)
    if getattr(pu.cu, "cu_transquant_bypass_flag", 0):
        d[...] = levels_xy
        return
    batch = _tb_batch(pu, levels_xy.T, log2size, _flags(pu, x0, y0, log2size) & TU_INTRA, _qp(pu),
                      sps_scaling_table(pu.cu.ctx.sps, getattr(pu.cu.ctx, "pps", None)))
    out = get_engine().dequant(batch)
    d[...] = out[:n * n].reshape(n, n).T


def inverse_transform(pu, x0, y0, log2size):
    """Residual from pu.scaled_samples: 8.6.4 two-stage inverse transform with the
    16-bit clip between the stages + the 8.6.2 bdShift (MODE "spec"), or
    transform.py:89-109 as written (MODE "ref_literal")."""
    if not 2 <= int(log2size) <= 5:
        raise ValueError("log2size must be in 2..5")
    n = 1 << log2size
    sx, sy = x0 - pu.origin_x, y0 - pu.origin_y
    d = pu.scaled_samples[sx:sx + n, sy:sy + n]
    r = pu.transformed_samples[sx:sx + n, sy:sy + n]
    eng = get_engine()
    if MODE == "ref_literal":
        tus = np.zeros(1, dtype=TU_DESC)
        tus["log2n"], tus["c_idx"] = log2size, pu.c_idx
        sc = np.clip(np.asarray(d).T, -32768, 32767).astype(np.int16).reshape(-1)
        r[...] = eng.ref_literal(tus, sc).reshape(n, n)
        return
    hit = _cached(pu, x0, y0, log2size)
    if hit is not None:
        r[...] = hit[1].T
        return
    flags = _flags(pu, x0, y0, log2size) | TU_PRESCALED
    batch = _tb_batch(pu, np.asarray(d).T, log2size, flags, 0, None)
    planes = eng.residual(batch)
    r[...] = batch.geom.plane_view(planes, 0, pu.c_idx)[:n, :n].T


def inverse_transform_1d(x, log2size, tr_type):
    """One-dimensional transform (8.6.4.2): y[i] = sum_j transMatrix[j][i] * x[j]; in
    MODE "ref_literal" the reference's transposed indexing (transform.py:81,85)."""
    return get_engine().idct_1d(x, log2size, tr_type, as_written=(MODE == "ref_literal")).astype(np.int64)


def reconstruction(pu, x0, y0, log2size):
    """reconstruction.reconstruction(pu, x0, y0, log2size) (reconstruction.py:4-27):
    reconstructed = Clip1(predicted + transformed) on the TB window of the PU arrays, in
    place; returns the window like the reference.  One small GPU call per TB; whole
    pictures go through Engine.reconstruct (p265_reconstruct_batch)."""
    if not 2 <= int(log2size) <= 5:
        raise ValueError("log2size must be in 2..5")
    n = 1 << log2size
    sx, sy = x0 - pu.origin_x, y0 - pu.origin_y
    pred = pu.predicted_samples[sx:sx + n, sy:sy + n]
    res = pu.transformed_samples[sx:sx + n, sy:sy + n]
    rec = pu.reconstructed_samples[sx:sx + n, sy:sy + n]
    bdy, bdc = _bit_depths(pu)
    geom = _GEOM.get((bdy, bdc))
    if geom is None:
        geom = _GEOM[(bdy, bdc)] = PicGeom(64, 64, 1, bdy, bdc)
    dtype = np.uint8 if max(bdy, bdc) <= 8 else np.uint16
    bd = bdy if pu.c_idx == 0 else bdc
    pbuf = np.zeros(geom.total_elems(), dtype=dtype)
    rbuf = np.zeros(geom.total_elems(), dtype=np.int16)
    geom.plane_view(pbuf, 0, pu.c_idx)[:n, :n] = np.clip(np.asarray(pred).T, 0, (1 << bd) - 1)
    geom.plane_view(rbuf, 0, pu.c_idx)[:n, :n] = np.clip(np.asarray(res).T, -32768, 32767)
    out = get_engine().reconstruct(pbuf, rbuf, geom)
    rec[...] = geom.plane_view(out, 0, pu.c_idx)[:n, :n].T
    return rec
