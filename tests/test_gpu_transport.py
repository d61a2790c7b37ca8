"""GPU parity of the round-2 host transport: the packed coefficient stream (unpack_kernel +
residual kernels), SAO in place on page-locked memory (write-back of modified CTBs only), the fused
deblocking -> SAO entry point, the programmatic-launch chain of the residual bins, the device
entry points bench.py times, the config-5 stream seeds, and asynchronous contexts."""
import numpy as np
import pytest

from conftest import small_cfg
from p265_b200 import synth
from p265_b200.picture import (TU_DESC, TU_INTRA, TU_LEVELS8, PackedResidualBatch, PicGeom, ResidualBatch,
                               sort_by_size)

pytestmark = pytest.mark.gpu


def assert_planes_equal(geom, got, ref):
    for p in range(geom.n_pics):
        for c in range(3):
            a, b = geom.plane_view(got, p, c), geom.plane_view(ref, p, c)
            if not np.array_equal(a, b):
                ys, xs = np.nonzero(a != b)
                raise AssertionError("pic %d comp %d: %d mismatches, first at (x=%d,y=%d): got %d want %d"
                                     % (p, c, ys.size, xs[0], ys[0], a[ys[0], xs[0]], b[ys[0], xs[0]]))


def pinned_like(a):
    import torch
    t = torch.empty(a.nbytes, dtype=torch.uint8, pin_memory=True)
    v = t.numpy().view(a.dtype).reshape(a.shape)
    v[...] = a
    return t, v


def to_dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(torch.device("cuda", 0))


# ------------------------------------------------------------------ packed coefficient stream
@pytest.mark.parametrize("name", ["1080p8", "4k10"])
@pytest.mark.parametrize("stress", [False, True])
def test_packed_stream_small(engine, c_oracle, name, stress):
    batch = synth.residual_batch(small_cfg(name, 320, 192), n_pics=3, stress=stress)
    ref = c_oracle.residual_batch(batch, zero_fill=False)
    packed = batch.packed()
    if not stress:
        assert (packed.tus["flags"] & TU_LEVELS8).any() and packed.stream.nbytes < batch.coeffs.nbytes // 3
    assert_planes_equal(batch.geom, engine.residual(packed), ref)
    assert_planes_equal(batch.geom, engine.residual(batch), ref)


def test_packed_stream_full_4k_and_ragged(engine, c_oracle):
    batch = synth.residual_batch("4k10", n_pics=1)
    ref = c_oracle.residual_batch(batch, zero_fill=False)
    assert_planes_equal(batch.geom, engine.residual(batch.packed()), ref)
    # ragged bins (partial warp items in every bin) and planes that are not covered
    keep = np.ones(len(batch.tus), bool)
    for l2, drop in ((5, 1), (4, 3), (3, 13), (2, 37)):
        keep[np.flatnonzero(batch.tus["log2n"] == l2)[-drop:]] = False
    part = ResidualBatch(batch.geom, np.ascontiguousarray(batch.tus[keep]), batch.coeffs, batch.scaling_factor,
                         covers_all=False)
    assert_planes_equal(batch.geom, engine.residual(part.packed()), c_oracle.residual_batch(part, zero_fill=True))


def test_packed_stream_sanity_bin(engine, c_oracle, sanity_batch):
    batch, _ = sanity_batch
    assert_planes_equal(batch.geom, engine.residual(batch.packed()), c_oracle.residual_batch(batch))


def test_packed_stream_empty_and_single_tb(engine, c_oracle):
    geom = PicGeom(64, 64, 1, 8, 8)
    empty = ResidualBatch(geom, np.zeros(0, TU_DESC), np.zeros(0, np.int16))
    assert not engine.residual(empty.packed()).any()
    rng = np.random.default_rng(4)
    for l2 in (2, 3, 4, 5):
        tus = np.zeros(1, TU_DESC)
        tus["log2n"], tus["qp"], tus["flags"] = l2, 30, TU_INTRA
        tus["x"] = tus["y"] = 32 if l2 < 5 else 0
        co = np.zeros(1 << (2 * l2), np.int16)
        co[rng.integers(0, co.size, 3)] = rng.integers(-500, 500, 3)      # wide levels, very sparse
        b = ResidualBatch(geom, tus, co)
        assert_planes_equal(geom, engine.residual(b.packed()), c_oracle.residual_batch(b))
        b0 = ResidualBatch(geom, tus, np.zeros_like(co))                  # all-zero TB: empty level list
        assert not engine.residual(b0.packed()).any()


def test_packed_stream_device_entry(engine, c_oracle):
    import torch
    batch = synth.residual_batch(small_cfg("4k10", 512, 256), n_pics=2)
    pb = batch.packed()
    d_tus, d_st, d_sf = to_dev(pb.tus), to_dev(pb.stream), to_dev(pb.scaling_factor)
    dev = d_tus.device
    d_arena = torch.empty(batch.samples() * 2 + 64, dtype=torch.uint8, device=dev)
    d_tus2 = torch.empty_like(d_tus)
    d_out = torch.zeros(batch.geom.total_elems() * 2, dtype=torch.uint8, device=dev)
    engine.residual_packed_dev(d_tus.data_ptr(), pb.bin_counts(), d_st.data_ptr(), d_sf.data_ptr(), pb.geom,
                               d_arena.data_ptr(), d_tus2.data_ptr(), d_out.data_ptr(), zero_fill=False,
                               sf_replicated=bool(pb.sf_replicated))
    engine.sync()
    assert_planes_equal(batch.geom, d_out.cpu().numpy().view(np.int16), c_oracle.residual_batch(batch, zero_fill=False))
    # the expanded arena is the dense arena in descriptor order
    dense = batch.densified()
    assert np.array_equal(d_arena.cpu().numpy()[:dense.coeffs.nbytes].view(np.int16), dense.coeffs)
    # ... and the rewritten descriptors carry the zero-extent codes unpack_kernel derives from the bitmaps:
    # the same codes the host derives from the coefficients (picture.extent_codes)
    from p265_b200.picture import extent_codes, set_extents
    want = dense.tus.copy()
    set_extents(want, *extent_codes(dense.tus, dense.coeffs))
    assert np.array_equal(d_tus2.cpu().numpy().view(TU_DESC), want)
    low = synth.residual_batch(small_cfg("4k10_lowfreq", 512, 256), n_pics=2).densified()   # every code pair occurs
    pl = low.packed()
    d_tus, d_st = to_dev(pl.tus), to_dev(pl.stream)
    d_tus2 = torch.empty_like(d_tus)
    engine.residual_packed_dev(d_tus.data_ptr(), pl.bin_counts(), d_st.data_ptr(), d_sf.data_ptr(), pl.geom,
                               d_arena.data_ptr(), d_tus2.data_ptr(), d_out.data_ptr(), zero_fill=False,
                               sf_replicated=bool(pl.sf_replicated))
    engine.sync()
    want = low.tus.copy()
    zr, zc = extent_codes(low.tus, low.coeffs)
    assert len(set(zip(zr[low.tus["log2n"] >= 4].tolist(), zc[low.tus["log2n"] >= 4].tolist()))) >= 8
    set_extents(want, zr, zc)
    assert np.array_equal(d_tus2.cpu().numpy().view(TU_DESC), want)
    assert_planes_equal(low.geom, d_out.cpu().numpy().view(np.int16), c_oracle.residual_batch(low, zero_fill=False))


def test_packed_stream_rejects_malformed_input(engine):
    batch = synth.residual_batch(small_cfg("4k10", 128, 64), n_pics=1)
    pb = batch.packed()
    short = PackedResidualBatch(pb.geom, pb.tus, pb.stream[:-8].copy(), pb.scaling_factor, bins=pb.bins)
    with pytest.raises(ValueError):
        engine.residual(short)
    # a dense-arena call must not carry the packed-only flag
    bad = ResidualBatch(batch.geom, batch.tus.copy(), batch.coeffs, batch.scaling_factor)
    bad.tus["flags"][0] |= TU_LEVELS8
    with pytest.raises(ValueError):
        engine.residual(bad)
    # a descriptor that announces more levels than the stream holds behind its record
    k = int(np.argmax(pb.tus["coeff_off"]))
    liar = pb.tus.copy()
    liar["rsvd"][k] = 1 << (2 * int(liar["log2n"][k]))
    with pytest.raises(ValueError):
        engine.residual(PackedResidualBatch(pb.geom, liar, pb.stream, pb.scaling_factor, bins=pb.bins))
    liar["rsvd"][k] = 2000
    with pytest.raises(ValueError):
        engine.residual(PackedResidualBatch(pb.geom, liar, pb.stream, pb.scaling_factor, bins=pb.bins))
    # a bitmap with more bits than announced levels is memory-safe: the device reads `rsvd` levels, the
    # surplus positions stay zero (record at the very end of the stream, every bit set)
    evil = pb.stream.copy()
    last = int(pb.tus["coeff_off"][k]) * 4
    nn8 = (1 << (2 * int(pb.tus["log2n"][k]))) // 8
    evil[last:last + nn8] = 0xFF
    a = engine.residual(PackedResidualBatch(pb.geom, pb.tus, evil, pb.scaling_factor, bins=pb.bins))
    b = engine.residual(PackedResidualBatch(pb.geom, pb.tus, evil, pb.scaling_factor, bins=pb.bins))
    assert np.array_equal(a, b)


# ------------------------------------------------------------------ PDL chain of the bins
def test_bin_chain_is_transitive(engine, c_oracle):
    """A huge 32x32 bin followed by a tiny 4x4 bin, then the D2H copy straight away: the last bin's
    completion must imply every earlier bin's (griddepcontrol.wait in every chained bin)."""
    rng = np.random.default_rng(11)
    w, h = 4096, 2048
    geom = PicGeom(w, h, 1, 10, 10)
    n32 = (w // 32) * (h // 32) - 1
    tus = np.zeros(n32 + 3, TU_DESC)
    idx = np.arange(n32)
    tus["x"][:n32], tus["y"][:n32] = (idx % (w // 32)) * 32, (idx // (w // 32)) * 32
    tus["log2n"][:n32] = 5
    tus["coeff_off"][:n32] = idx * 64
    tus["log2n"][n32:] = 2
    tus["x"][n32:], tus["y"][n32:] = w - 32 + 4 * np.arange(3), h - 32
    tus["coeff_off"][n32:] = n32 * 64 + np.arange(3)
    tus["qp"], tus["flags"] = 40, TU_INTRA
    coeffs = rng.integers(-300, 300, n32 * 1024 + 3 * 16).astype(np.int16)
    batch = ResidualBatch(geom, sort_by_size(tus), coeffs)
    ref = c_oracle.residual_batch(batch)
    for _ in range(5):
        assert_planes_equal(geom, engine.residual(batch), ref)


# ------------------------------------------------------------------ SAO in place / write-back
@pytest.mark.parametrize("size,bit_depth,ctb_log2", [((3840, 2160), 10, 6), ((200, 136), 8, 4), ((264, 136), 10, 5),
                                                     ((72, 56), 8, 6), ((1920, 1080), 8, 5)])
def test_sao_in_place_on_pinned_memory(engine, c_oracle, size, bit_depth, ctb_log2):
    w, h = size
    geom, rec, params = synth.sao_batch(w, h, bit_depth, n_pics=2, ctb_log2=ctb_log2, seed=7 + w, n_unique=2)
    ref = c_oracle.sao_batch(rec, geom, ctb_log2, params)
    keep, buf = pinned_like(rec)
    n0 = engine.launch_count
    got = engine.sao(buf, geom, ctb_log2, params, inplace=True)
    assert got is buf or got.base is not None
    assert engine.launch_count - n0 == 2          # sao_kernel + sao_writeback_kernel: the zero-copy path ran
    assert_planes_equal(geom, buf, ref)
    # padding and gaps keep the caller's bytes
    mask = np.ones(rec.size, bool)
    for p in range(geom.n_pics):
        for c in range(3):
            hh, ww = geom.plane_shape(c)
            off = p * geom.pic_stride + geom.plane_off[c]
            mask[off:off + hh * geom.stride(c)].reshape(hh, geom.stride(c))[:, :ww] = False
    assert np.array_equal(buf[mask], rec[mask])
    # pageable memory: same result through the copy engine
    pag = rec.copy()
    engine.sao(pag, geom, ctb_log2, params, inplace=True)
    assert np.array_equal(pag, buf)
    # out of place: rec untouched, out complete
    out = engine.sao(rec, geom, ctb_log2, params)
    assert np.array_equal(out, buf)


# ------------------------------------------------------------------ fused loop filters
@pytest.mark.parametrize("size,bit_depth,ctb_log2", [((3840, 2160), 10, 6), ((200, 136), 8, 4), ((264, 136), 10, 5)])
def test_loop_filter_matches_deblock_then_sao(engine, c_oracle, size, bit_depth, ctb_log2):
    w, h = size
    geom, buf, blk, dctb = synth.deblock_batch(w, h, bit_depth, 2, ctb_log2, seed=w + 1)
    _, _, params = synth.sao_batch(w, h, bit_depth, n_pics=2, ctb_log2=ctb_log2, seed=w + 2)
    nf = (blk & 0x8000 != 0).astype(np.uint8)
    dbk = c_oracle.deblock_batch(buf, geom, ctb_log2, blk, dctb)
    ref = c_oracle.sao_batch(dbk, geom, ctb_log2, params, nf)
    work = buf.copy()
    engine.loop_filter(work, geom, ctb_log2, blk, dctb, params, nf)
    assert np.array_equal(work, ref)              # whole buffer: padding comes back as it went in
    # the two separate calls give the same
    two = engine.sao(engine.deblock(buf, geom, ctb_log2, blk, dctb), geom, ctb_log2, params, nf)
    assert np.array_equal(two, ref)
    # deblocking only / SAO only (pinned: write-back of the modified CTBs)
    only = buf.copy()
    engine.loop_filter(only, geom, ctb_log2, blk, dctb)
    assert np.array_equal(only, dbk)
    keep, pin = pinned_like(buf)
    engine.loop_filter(pin, geom, ctb_log2, sao_params=params)
    assert np.array_equal(pin, c_oracle.sao_batch(buf, geom, ctb_log2, params))
    with pytest.raises(ValueError):
        engine.loop_filter(buf.copy(), geom, ctb_log2)


# ------------------------------------------------------------------ device entry points bench.py times
def test_device_entry_points_match_oracle(engine, c_oracle):
    import torch
    geom, rec, params = synth.sao_batch(1920, 1080, 10, n_pics=2, ctb_log2=6, seed=31)
    d_rec, d_par = to_dev(rec), to_dev(params)
    d_out = d_rec.clone()
    engine.sao_dev(d_rec.data_ptr(), d_out.data_ptr(), geom, 6, d_par.data_ptr())
    engine.sync()
    assert_planes_equal(geom, d_out.cpu().numpy().view(rec.dtype), c_oracle.sao_batch(rec, geom, 6, params))
    with pytest.raises(ValueError):
        engine.sao_dev(d_rec.data_ptr(), d_rec.data_ptr(), geom, 6, d_par.data_ptr())
    dg, drec, dblk, dctb = synth.deblock_batch(1920, 1080 - 1080 % 8, 10, 2, 6, seed=32)
    d_pix, d_blk, d_ctb = to_dev(drec), to_dev(dblk), to_dev(dctb)
    engine.deblock_dev(d_pix.data_ptr(), dg, 6, d_blk.data_ptr(), d_ctb.data_ptr())
    engine.sync()
    assert np.array_equal(d_pix.cpu().numpy().view(drec.dtype), c_oracle.deblock_batch(drec, dg, 6, dblk, dctb))
    from oracle import spec_oracle as so
    rng = np.random.default_rng(33)
    pred = rng.integers(0, 1024, geom.total_elems()).astype(np.uint16)
    res = rng.integers(-1500, 1500, geom.total_elems()).astype(np.int16)
    d_pred, d_res = to_dev(pred), to_dev(res)
    d_r = torch.zeros_like(d_pred)
    engine.reconstruct_dev(d_pred.data_ptr(), d_res.data_ptr(), d_r.data_ptr(), geom)
    engine.sync()
    got = d_r.cpu().numpy().view(np.uint16)
    for p in range(geom.n_pics):
        for c in range(3):
            assert np.array_equal(geom.plane_view(got, p, c),
                                  so.reconstruct(geom.plane_view(pred, p, c), geom.plane_view(res, p, c), 10))


@pytest.mark.parametrize("seed", range(26510, 26518))
def test_config5_stream_seeds(engine, c_oracle, seed):
    """BASELINE config 5: the eight streams bench.py times (seeds 26510..26517), one 4K picture each,
    residual (dense and packed transport) + SAO against the oracle."""
    batch = synth.residual_batch("4k10", n_pics=1, seed=seed)
    ref = c_oracle.residual_batch(batch, zero_fill=False)
    assert_planes_equal(batch.geom, engine.residual(batch.packed()), ref)
    geom, rec, params = synth.sao_batch(3840, 2160, 10, n_pics=1, seed=seed + 1000)
    assert_planes_equal(geom, engine.sao(rec, geom, 6, params), c_oracle.sao_batch(rec, geom, 6, params))


# ------------------------------------------------------------------ asynchronous contexts
def test_async_context_interleaves_residual_and_sao(c_oracle):
    """One asynchronous context, residual and SAO calls interleaved (they share scratch slots; the
    context's single stream serialises them), results only read after sync()."""
    from p265_b200.engine import Engine
    eng = Engine(0)
    eng.set_async(True)
    work = []
    for i in range(3):
        batch = synth.residual_batch(small_cfg("4k10", 640 - 64 * i, 384), n_pics=2, seed=50 + i).packed()
        geom, rec, params = synth.sao_batch(520 + 8 * i, 264, 10, n_pics=2, ctb_log2=6, seed=60 + i)
        k1, h_res = pinned_like(np.zeros(batch.geom.total_elems(), np.int16))
        k2, h_rec = pinned_like(rec)
        k3, h_tus = pinned_like(batch.tus)
        k4, h_st = pinned_like(batch.stream)
        pb = PackedResidualBatch(batch.geom, h_tus, h_st, batch.scaling_factor, bins=batch.bins)
        eng.residual(pb, h_res)
        eng.sao(h_rec, geom, 6, params, inplace=True)
        work.append((pb, h_res, rec, h_rec, geom, params, (k1, k2, k3, k4)))
    eng.sync()
    sync_eng = Engine(0)
    for pb, h_res, rec, h_rec, geom, params, _ in work:
        assert np.array_equal(h_res, sync_eng.residual(pb))
        assert_planes_equal(geom, h_rec, c_oracle.sao_batch(rec, geom, 6, params))
    eng.close()


def test_pcie_probe_reports_both_directions(engine):
    h2d, d2h = engine.pcie_probe(32 << 20, reps=2)
    assert h2d > 1e9 and d2h > 1e9
    only, none = engine.pcie_probe(32 << 20, reps=2, d2h=False)
    assert only > 1e9 and none is None
    h2d, d2h = engine.pcie_probe(8 << 20, reps=8, n_buffers=4)
    assert h2d > 1e9 and d2h > 1e9
    with pytest.raises(ValueError):
        engine.pcie_probe(8 << 20, reps=1, n_buffers=0)
