"""The oracle is pinned here (CPU only): against the reference's own outputs committed
under tests/golden/, against the reference itself when /root/reference is mounted,
against hand-derived known answers, and numpy <-> C <-> naive loops against each other."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import refshim
from oracle import spec_oracle as so
from p265_b200 import scaling_list, synth
from p265_b200.picture import pack_scaling_factor

# sha256 of the 32x32 int8 basis as bytes (row-major) -- pins the table on boxes without
# the reference; test_tables_equal_reference checks it against transform.py:7-72 here.
DCT32_SHA = hashlib.sha256(so.DCT32.astype(np.int8).tobytes()).hexdigest()


def test_basis_structure():
    m = so.DCT32
    assert (m[0] == 64).all()
    assert list(m[1, :4]) == [90, 90, 88, 85] and m[1, 31] == -90
    assert list(m[16, :4]) == [64, -64, -64, 64]
    assert list(m[8, :4]) == [83, 36, -36, -83]
    # even rows of the 2N-point basis are the N-point basis (what the butterfly relies on)
    assert np.array_equal(m[::2, :16], so.trans_matrix(4, 0))
    # near-orthogonality of every size
    for l2 in (2, 3, 4, 5):
        t = so.trans_matrix(l2, 0)
        g = t @ t.T
        n = 1 << l2
        assert np.abs(g - np.diag(np.diag(g))).max() <= 64 * n * 0.02 * 64
    d = so.DST4 @ so.DST4.T
    assert np.abs(d - np.diag(np.diag(d))).max() <= 300


@pytest.mark.skipif(not refshim.reference_available(), reason="/root/reference not mounted")
def test_tables_equal_reference():
    ns = refshim.load()
    assert np.array_equal(np.array(ns.transform.trans_matrix_type0), so.DCT32)
    assert np.array_equal(np.array(ns.transform.trans_matrix_type1), so.DST4)
    assert list(ns.sld.ScalingListData.default_scaling_list_8x8_intra) == list(so._DEF8_INTRA)
    assert list(ns.sld.ScalingListData.default_scaling_list_8x8_inter) == list(so._DEF8_INTER)


def test_c_oracle_table_equals_numpy(c_oracle):
    assert np.array_equal(c_oracle.dct32(), so.DCT32)
    assert hashlib.sha256(c_oracle.dct32().astype(np.int8).tobytes()).hexdigest() == DCT32_SHA


def test_dequant_pinned_by_reference_outputs(sanity_batch, c_oracle):
    """scaling.py's own outputs for all 5,982 TBs of sanity.bin."""
    batch, _ = sanity_batch
    ref = np.load(os.path.join(GOLDEN, "sanity_residual.npz"))
    assert np.array_equal(c_oracle.dequant_batch(batch), ref["ref_scaled_yx"])
    for t in batch.tus[::37]:
        n = 1 << int(t["log2n"])
        off = int(t["coeff_off"]) * 16
        lv = batch.coeffs[off:off + n * n].reshape(n, n)
        d = so.inverse_scaling(lv, int(t["qp"]), 8, int(t["log2n"]))
        assert np.array_equal(d.reshape(-1), ref["ref_scaled_yx"][off:off + n * n])


def test_ref_literal_pinned_by_reference_outputs(sanity_batch, c_oracle):
    """transform.py as written, all TBs of sanity.bin."""
    batch, _ = sanity_batch
    ref = np.load(os.path.join(GOLDEN, "sanity_residual.npz"))
    got = c_oracle.ref_literal_batch(batch.tus, ref["ref_scaled_yx"])
    assert np.array_equal(got, ref["ref_literal_xy"])
    for t in batch.tus[::53]:
        n = 1 << int(t["log2n"])
        off = int(t["coeff_off"]) * 16
        d_xy = ref["ref_scaled_yx"][off:off + n * n].reshape(n, n).T
        lit = so.ref_literal_transform_xy(d_xy, int(t["log2n"]), int(t["c_idx"]))
        assert np.array_equal(lit.reshape(-1), ref["ref_literal_xy"][off:off + n * n])


def test_random_reference_vectors():
    """Seeded random TBs run through the reference (incl. full-range levels, 10-bit qP,
    scaling factors): dequantisation and the as-written transform."""
    rr = np.load(os.path.join(GOLDEN, "random_reference.npz"))
    sf = so.expand_scaling_factor(*so.default_scaling_lists())
    for i in range(int(rr["n"])):
        lv = rr["levels_xy_%d" % i]
        l2 = int(np.log2(lv.shape[0]))
        c, bd = int(rr["c_idx_%d" % i]), int(rr["bit_depth_%d" % i])
        m = sf[(l2 - 2, so.matrix_id(l2, c, True))] if bool(rr["use_sf_%d" % i]) else None
        d = so.inverse_scaling(lv, int(rr["qp_%d" % i]), bd, l2, m)
        assert np.array_equal(d, rr["scaled_xy_%d" % i])
        assert np.array_equal(so.ref_literal_transform_xy(d, l2, c), rr["literal_xy_%d" % i])


def _naive_residual(lv, qp, bd, l2, dst):
    n = 1 << l2
    ls = [40, 45, 51, 57, 64, 72]
    shift = bd + l2 - 5
    d = [[0] * n for _ in range(n)]
    for y in range(n):
        for x in range(n):
            v = (int(lv[y][x]) * 16 * (ls[qp % 6] << (qp // 6)) + (1 << (shift - 1))) >> shift
            d[y][x] = max(-32768, min(32767, v))
    mat = so.DST4 if dst else so.DCT32[:: 32 // n, :n]
    e = [[sum(int(mat[j][i]) * d[j][x] for j in range(n)) for x in range(n)] for i in range(n)]
    g = [[max(-32768, min(32767, (e[y][x] + 64) >> 7)) for x in range(n)] for y in range(n)]
    s2 = 20 - bd
    return [[(sum(int(mat[j][i]) * g[y][j] for j in range(n)) + (1 << (s2 - 1))) >> s2 for i in range(n)]
            for y in range(n)]


@pytest.mark.parametrize("l2", [2, 3, 4, 5])
def test_vectorised_equals_naive_loops(l2):
    rng = np.random.default_rng(l2)
    n = 1 << l2
    for k in range(3):
        lv = rng.integers(-32768, 32768, (n, n)) if k == 0 else rng.integers(-60, 61, (n, n))
        qp, bd = int(rng.integers(0, 52)), 8 + 2 * (k % 2)
        dst = l2 == 2 and k % 2 == 0
        want = np.array(_naive_residual(lv, qp, bd, l2, dst))
        got = so.residual_block_yx(lv, qp, bd, l2, dst=dst)
        assert np.array_equal(got, want)


def test_known_answers_by_hand():
    # DC only, 4x4 DCT, 8-bit, qP 28 (levelScale 64 << 4): d = (1*16*1024 + 16) >> 5 = 512
    lv = np.zeros((4, 4), np.int64)
    lv[0, 0] = 1
    assert so.inverse_scaling(lv, 28, 8, 2)[0, 0] == 512
    # stage 1: 64*512 = 32768 -> (32768+64)>>7 = 256; stage 2: 64*256 = 16384 -> (16384+2048)>>12 = 4
    assert (so.residual_block_yx(lv, 28, 8, 2) == 4).all()
    # transform skip: (512 << 7 + 2048) >> 12 = 16 at (0,0), 0 elsewhere
    r = so.residual_block_yx(lv, 28, 8, 2, ts=True)
    assert r[0, 0] == 16 and r.sum() == 16
    # bypass: residual = level
    assert np.array_equal(so.residual_block_yx(lv, 28, 8, 2, bypass=True), lv)
    # DST 4x4, coefficient at [y=0][x=0]: column pass gives 29,55,74,84 scaled; check symmetry
    r = so.residual_block_yx(lv, 28, 8, 2, dst=True)
    assert np.array_equal(r, r.T) and r[0, 0] < r[3, 3]
    # dequant clip: level 32767 at qP 51 saturates
    lv[0, 0] = 32767
    assert so.inverse_scaling(lv, 51, 8, 2)[0, 0] == 32767
    lv[0, 0] = -32768
    assert so.inverse_scaling(lv, 51, 8, 2)[0, 0] == -32768


@pytest.mark.parametrize("l2", [2, 3, 4, 5])
def test_forward_inverse_round_trip(l2):
    """Float forward DCT of a residual, then the integer inverse: error <= 3 LSB (the
    integer basis rows are only approximately of equal norm; the DC-row gain is used)."""
    rng = np.random.default_rng(40 + l2)
    n = 1 << l2
    m = so.trans_matrix(l2, 0).astype(np.float64)
    res = rng.integers(-200, 201, (6, n, n)).astype(np.float64)
    # forward with the same basis: coef = M res M^T / (norm), norm chosen so that the
    # inverse's two >>7 / >>12 shifts (8-bit) return the residual: total gain 64^2 * n... 
    coef = np.einsum("ji,byi->byj", m, res)
    coef = np.einsum("jy,byx->bjx", m, coef)
    gain = (m[0] @ m[0]) ** 2            # (64^2 N)^2 for the DC path
    d = np.rint(coef * (1 << (7 + 12)) / gain).astype(np.int64)
    r = so.inverse_transform_yx(d, l2, 0, 8)
    assert np.abs(r - res).max() <= 3


def test_sat16_is_lossless_for_reconstruction():
    rng = np.random.default_rng(5)
    r = rng.integers(-100000, 100000, 5000)
    for bd in (8, 10, 12):
        pred = rng.integers(0, 1 << bd, 5000)
        a = np.clip(pred + r, 0, (1 << bd) - 1)
        b = np.clip(pred + so.sat16(r), 0, (1 << bd) - 1)
        assert np.array_equal(a, b)


def test_densified_batch_is_the_same_batch():
    """ResidualBatch.densified(): arena re-laid in descriptor order, same TBs, same residual."""
    from oracle import c_oracle
    batch = synth.residual_batch("1080p8", n_pics=1)
    dense = batch.densified()
    assert dense.dense_small_bins() and not batch.dense_small_bins()
    assert dense.coeffs.size == batch.samples()
    assert np.array_equal(c_oracle.residual_batch(batch), c_oracle.residual_batch(dense))


@pytest.mark.parametrize("name", ["1080p8", "4k10"])
@pytest.mark.parametrize("stress", [False, True])
def test_c_oracle_equals_numpy_oracle(c_oracle, name, stress):
    from conftest import small_cfg
    batch = synth.residual_batch(small_cfg(name, 192, 128), n_pics=2, stress=stress)
    out = c_oracle.residual_batch(batch, zero_fill=False)
    sf = None
    if batch.scaling_factor is not None:
        sf = so.expand_scaling_factor(*so.default_scaling_lists())
        assert np.array_equal(so.pack_scaling_factor(sf), batch.scaling_factor)
    g = batch.geom
    for t in batch.tus:
        l2, c, fl = int(t["log2n"]), int(t["c_idx"]), int(t["flags"])
        n = 1 << l2
        off = int(t["coeff_off"]) * 16
        lv = batch.coeffs[off:off + n * n].reshape(n, n)
        m = None if sf is None else sf[(l2 - 2, so.matrix_id(l2, c, bool(fl & 8)))].T   # [x][y] -> [y][x]
        bd = g.bit_depth_c if c else g.bit_depth_y
        want = so.sat16(so.residual_block_yx(lv, int(t["qp"]), bd, l2, dst=bool(fl & 1), ts=bool(fl & 2),
                                            bypass=bool(fl & 4), m=m))
        got = g.plane_view(out, int(t["pic"]), c)[t["y"]:t["y"] + n, t["x"]:t["x"] + n]
        assert np.array_equal(got, want), (l2, c, fl)


def test_scaling_factor_expansion():
    sf = so.expand_scaling_factor(*so.default_scaling_lists())
    assert (sf[(0, 0)] == 16).all()
    f8 = sf[(1, 0)]
    assert f8[0, 0] == 16 and f8[7, 7] == 115 and f8[7, 6] == 88 and f8[6, 7] == 88
    assert np.array_equal(f8, f8.T)                       # default intra list is symmetric
    f16, f32 = sf[(2, 0)], sf[(3, 0)]
    assert np.array_equal(f16[::2, ::2], f8) and np.array_equal(f16[1::2, 1::2], f8)
    assert np.array_equal(f32[::4, ::4], f8) and np.array_equal(f32[3::4, 3::4], f8)
    assert sf[(3, 1)][31, 31] == 91
    # product-side construction agrees with the oracle's
    psf = scaling_list.default_scaling_factor()
    assert set(psf) == set(sf)
    for k in sf:
        assert np.array_equal(psf[k], sf[k])
    assert np.array_equal(pack_scaling_factor(psf), so.pack_scaling_factor(sf))
    # custom DC
    lists, dc = so.default_scaling_lists()
    dc[(2, 1)] = 40
    assert so.expand_scaling_factor(lists, dc)[(2, 1)][0, 0] == 40
    # scan order: 6.5.3 known prefix
    assert [tuple(v) for v in so.up_right_diagonal_scan(4)[:6]] == [(0, 0), (0, 1), (1, 0), (0, 2), (1, 1), (2, 0)]
    assert scaling_list.diag_scan(8) == [tuple(v) for v in so.up_right_diagonal_scan(8)]


# ------------------------------------------------------------------------------ SAO
def _rand_sao(rng, h, w, cs, bd):
    ch, cw = -(-h // cs), -(-w // cs)
    t = rng.integers(0, 3, (ch, cw))
    bp = rng.integers(0, 32, (ch, cw))
    cls = rng.integers(0, 4, (ch, cw))
    off = rng.integers(-7, 8, (ch, cw, 4))
    rec = rng.integers(0, 1 << bd, (h, w))
    return rec, t, bp, cls, off


@pytest.mark.parametrize("bd", [8, 10])
def test_sao_vectorised_equals_per_sample(bd):
    rng = np.random.default_rng(bd)
    rec, t, bp, cls, off = _rand_sao(rng, 40, 56, 16, bd)
    avail = rng.integers(0, 512, t.shape)
    a = so.sao_filter_plane(rec, bd, 16, t, bp, cls, off, ctb_avail=avail)
    b = np.array(so.sao_filter_plane_naive(rec, bd, 16, t, bp, cls, off, ctb_avail=avail))
    assert np.array_equal(a, b)


def test_sao_known_answers():
    # band offset, 8-bit: bandShift 3; band_position 4 -> bands 4..7 = samples 32..63
    rec = np.array([[31, 32, 39, 40, 56, 63, 64, 250]])
    out = so.sao_filter_plane(rec, 8, 8, [[1]], [[4]], [[0]], [[[1, -2, 3, -4]]])
    assert list(out[0]) == [31, 33, 40, 38, 52, 59, 64, 250]
    # clipping at both ends
    rec = np.array([[0, 255]])
    out = so.sao_filter_plane(rec, 8, 8, [[1]], [[0]], [[0]], [[[-5, 0, 0, 0]]])
    assert out[0, 0] == 0
    out = so.sao_filter_plane(rec, 8, 8, [[1]], [[31]], [[0]], [[[7, 0, 0, 0]]])
    assert out[0, 1] == 255
    # edge offset class 0 (horizontal): valley, concave corner, flat, convex corner, peak
    rec = np.array([[5, 3, 5, 5, 7, 7, 9, 7, 7]])
    off = [[[10, 20, -30, -40]]]
    out = so.sao_filter_plane(rec, 8, 16, [[2]], [[0]], [[0]], off)
    want = []
    for x in range(9):
        if x in (0, 8):
            want.append(rec[0, x]); continue
        c, a, b = rec[0, x], rec[0, x - 1], rec[0, x + 1]
        s = 2 + np.sign(c - a) + np.sign(c - b)
        idx = (1, 2, 0, 3, 4)[s]
        want.append(int(np.clip(c + ([0] + off[0][0])[idx], 0, 255)))
    assert list(out[0]) == want
    assert want[1] == 13 and want[6] == 0           # valley +10, peak -40 clipped to 0


def test_sao_offset_val_derivation():
    assert so.sao_offset_val(2, [1, 2, 3, 4], [0, 0, 0, 0], 8) == [1, 2, -3, -4]
    assert so.sao_offset_val(1, [1, 2, 3, 4], [1, 0, 1, 0], 10) == [-1, 2, -3, 4]
    assert so.sao_offset_val(1, [1, 2, 3, 4], [0, 0, 0, 0], 12) == [1, 2, 3, 4]          # 10/2014 edition on
    assert so.sao_offset_val(1, [1, 2, 3, 4], [0, 0, 0, 0], 12, 2) == [4, 8, 12, 16]   # log2_sao_offset_scale
    from p265_b200 import packer
    for args in ((2, [1, 2, 3, 4], [1, 1, 0, 0], 8), (1, [7, 0, 3, 4], [1, 0, 1, 0], 10),
                 (1, [31, 2, 3, 4], [0, 1, 0, 0], 12)):
        assert packer.sao_offset_val(*args) == so.sao_offset_val(*args)


@pytest.mark.parametrize("bit_depth", [8, 10])
@pytest.mark.parametrize("two_slices", [False, True])
def test_c_sao_equals_numpy_sao(c_oracle, bit_depth, two_slices):
    geom, rec, params = synth.sao_batch(200, 136, bit_depth, n_pics=1, ctb_log2=5, seed=3, two_slices=two_slices)
    out = c_oracle.sao_batch(rec, geom, 5, params)
    for c in range(3):
        cs = 32 >> (1 if c else 0)
        p = params[0]
        want = so.sao_filter_plane(geom.plane_view(rec, 0, c), bit_depth, cs, p["type"][..., c],
                                   p["band_pos"][..., c], p["eo_class"][..., c], p["offset_val"][..., c, :],
                                   ctb_avail=p["avail"])
        assert np.array_equal(geom.plane_view(out, 0, c), want), c


def test_c_sao_no_filter_equals_numpy(c_oracle):
    geom, rec, params = synth.sao_batch(128, 64, 8, n_pics=1, ctb_log2=6, seed=8)
    rng = np.random.default_rng(2)
    nf = (rng.random((1, 8, 16)) < 0.4).astype(np.uint8)
    out = c_oracle.sao_batch(rec, geom, 6, params, nf)
    for c in range(3):
        p = params[0]
        want = so.sao_filter_plane(geom.plane_view(rec, 0, c), 8, 64 >> (1 if c else 0), p["type"][..., c],
                                   p["band_pos"][..., c], p["eo_class"][..., c], p["offset_val"][..., c, :],
                                   ctb_avail=p["avail"], no_filter=nf[0], no_filter_log2=2 if c else 3)
        assert np.array_equal(geom.plane_view(out, 0, c), want), c


@pytest.mark.skipif(not refshim.reference_available(), reason="/root/reference not mounted")
def test_reconstruct_pinned_by_reference_function():
    """oracle.reconstruct == the reference's own reconstruction.reconstruction."""
    import types
    ns = refshim.load()
    rng = np.random.default_rng(17)
    for bd, c_idx in ((8, 0), (10, 1), (8, 2)):
        n = 8
        pred = rng.integers(0, 1 << bd, (n, n))
        res = rng.integers(-1500, 1500, (n, n))
        sps = types.SimpleNamespace(bit_depth_y=bd, bit_depth_c=bd)
        pu = types.SimpleNamespace(c_idx=c_idx, origin_x=0, origin_y=0,
                                   cu=types.SimpleNamespace(ctx=types.SimpleNamespace(sps=sps)),
                                   predicted_samples=pred.copy(), transformed_samples=res.copy(),
                                   reconstructed_samples=np.zeros((n, n), np.int64))
        ns.reconstruction.reconstruction(pu=pu, x0=0, y0=0, log2size=3)
        assert np.array_equal(pu.reconstructed_samples, so.reconstruct(pred, res, bd))


def test_residual_batch_bin_counts_are_counted_once():
    """bin counts are part of the packed batch (picture.ResidualBatch): counted on first use, then reused;
    explicit counts (the packer's) are taken as they are -- the C-ABI validates them against the list."""
    from p265_b200.picture import TU_DESC, PicGeom, ResidualBatch
    t = np.zeros(7, TU_DESC)
    t["log2n"] = (5, 4, 4, 3, 2, 2, 2)
    b = ResidualBatch(PicGeom(64, 64, 1, 8, 8), t, np.zeros(16, np.int16))
    assert b.bins is None and b.bin_counts() == (1, 2, 1, 3) and b.bins == (1, 2, 1, 3)
    assert ResidualBatch(PicGeom(64, 64, 1, 8, 8), t, np.zeros(16, np.int16), bins=(1, 2, 1, 3)).bin_counts() == (1, 2, 1, 3)
