"""GPU parity of the residual path (dequantisation + inverse transform) through the
C-ABI, bit-exact against the C oracle (oracle/spec_oracle.c)."""
import numpy as np
import pytest

from conftest import small_cfg
from p265_b200 import synth
from p265_b200.picture import (TU_BYPASS, TU_DESC, TU_DST, TU_INTRA, TU_SKIP, PicGeom, ResidualBatch,
                               pack_scaling_factor, sort_by_size)

pytestmark = pytest.mark.gpu


def assert_planes_equal(geom, got, ref):
    for p in range(geom.n_pics):
        for c in range(3):
            a, b = geom.plane_view(got, p, c), geom.plane_view(ref, p, c)
            if not np.array_equal(a, b):
                ys, xs = np.nonzero(a != b)
                raise AssertionError("pic %d comp %d: %d mismatches, first at (x=%d,y=%d): got %d want %d"
                                     % (p, c, ys.size, xs[0], ys[0], a[ys[0], xs[0]], b[ys[0], xs[0]]))


def test_sanity_bin_tbs(engine, c_oracle, sanity_batch):
    """BASELINE config 1: the real TB lists of sanity.bin (5,982 TBs, 41 transform-skip)."""
    batch, _ = sanity_batch
    got = engine.residual(batch)
    assert_planes_equal(batch.geom, got, c_oracle.residual_batch(batch))


@pytest.mark.parametrize("name", ["1080p8", "4k10"])
@pytest.mark.parametrize("stress", [False, True])
def test_synthetic_small(engine, c_oracle, name, stress):
    batch = synth.residual_batch(small_cfg(name, 320, 192), n_pics=3, stress=stress)
    got = engine.residual(batch)
    assert_planes_equal(batch.geom, got, c_oracle.residual_batch(batch, zero_fill=False))


def test_config2_full_1080p(engine, c_oracle):
    batch = synth.residual_batch("1080p8", n_pics=2)
    got = engine.residual(batch)
    assert_planes_equal(batch.geom, got, c_oracle.residual_batch(batch, zero_fill=False))


@pytest.mark.parametrize("name", ["1080p8", "4k10"])
@pytest.mark.parametrize("stress", [False, True])
def test_dense_arena_layout(engine, c_oracle, name, stress):
    """Arena re-laid in descriptor order: the host entry point detects it and the small bins
    address their tiles by index (P265_RES_DENSE_ARENA); results equal the scattered layout's."""
    batch = synth.residual_batch(name, n_pics=2, stress=stress)
    dense = batch.densified()
    assert dense.dense_small_bins() and not batch.dense_small_bins()
    ref = c_oracle.residual_batch(batch, zero_fill=False)
    assert_planes_equal(batch.geom, engine.residual(dense), ref)
    # ragged: drop TBs so that the last items of the small bins are partial, then re-lay
    keep = np.ones(len(batch.tus), bool)
    keep[-37:] = False
    keep[np.flatnonzero(batch.tus["log2n"] == 3)[-5:]] = False
    from p265_b200.picture import ResidualBatch
    part = ResidualBatch(batch.geom, np.ascontiguousarray(batch.tus[keep]), batch.coeffs, batch.scaling_factor,
                         covers_all=False).densified()
    assert part.dense_small_bins()
    assert_planes_equal(batch.geom, engine.residual(part), c_oracle.residual_batch(part, zero_fill=True))


def test_dense_arena_flag_on_device_entry(engine, c_oracle):
    import torch
    batch = synth.residual_batch("4k10", n_pics=1).densified()
    dev = torch.device("cuda", 0)
    to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)
    d_tus, d_co, d_sf = to_dev(batch.tus), to_dev(batch.coeffs), to_dev(batch.scaling_factor)
    ref = c_oracle.residual_batch(batch, zero_fill=False)
    for dense in (True, False):
        d_out = torch.zeros(batch.geom.total_elems() * 2, dtype=torch.uint8, device=dev)
        engine.residual_dev(d_tus.data_ptr(), batch.bin_counts(), d_co.data_ptr(), d_sf.data_ptr(), batch.geom,
                            d_out.data_ptr(), zero_fill=False, sf_replicated=bool(batch.sf_replicated),
                            dense_arena=dense)
        engine.sync()
        assert_planes_equal(batch.geom, d_out.cpu().numpy().view(np.int16), ref)


def test_config3_full_4k(engine, c_oracle):
    batch = synth.residual_batch("4k10", n_pics=1)
    got = engine.residual(batch)
    assert_planes_equal(batch.geom, got, c_oracle.residual_batch(batch, zero_fill=False))


def test_config3_full_4k_stress(engine, c_oracle):
    """Dense full-range levels, every qP: both 16-bit clips and the int16 saturation."""
    batch = synth.residual_batch("4k10", n_pics=1, stress=True, seed=99)
    got = engine.residual(batch)
    assert_planes_equal(batch.geom, got, c_oracle.residual_batch(batch, zero_fill=False))


def _one_size_batch(log2n, count, rng, bit_depth=8, qps=(0, 51), flags=TU_INTRA, c_idx=0, sf=None):
    n = 1 << log2n
    per_row = 256 // n
    geom = PicGeom(512, 64 * ((count * 2 + per_row - 1) // per_row + 1), 1, bit_depth, bit_depth)
    tus = np.zeros(count, TU_DESC)
    idx = np.arange(count)
    tus["x"] = (idx % per_row) * n
    tus["y"] = (idx // per_row) * n
    tus["log2n"], tus["c_idx"], tus["flags"] = log2n, c_idx, flags
    tus["qp"] = rng.choice(np.array(qps), count)
    tus["coeff_off"] = idx * (n * n // 16)
    coeffs = rng.integers(-32768, 32768, count * n * n).astype(np.int16)
    return ResidualBatch(geom, sort_by_size(tus), coeffs, sf, covers_all=False)


@pytest.mark.parametrize("log2n", [2, 3, 4, 5])
@pytest.mark.parametrize("count", [1, 3, 17, 33])
def test_ragged_bins_and_partial_cover(engine, c_oracle, log2n, count):
    """Counts that do not fill a warp item; planes mostly uncovered -> zero fill."""
    rng = np.random.default_rng(1000 + log2n * 100 + count)
    batch = _one_size_batch(log2n, count, rng, qps=tuple(range(0, 52)))
    got = engine.residual(batch)
    assert_planes_equal(batch.geom, got, c_oracle.residual_batch(batch))


@pytest.mark.parametrize("bit_depth", [8, 10, 12])
def test_every_qp_and_kind_4x4(engine, c_oracle, bit_depth):
    """4x4: DCT / DST / transform-skip / bypass at every legal qP (left-shift dequant)."""
    rng = np.random.default_rng(7 + bit_depth)
    qps = tuple(range(0, 52 + 6 * (bit_depth - 8)))
    for flags, c_idx in ((TU_INTRA | TU_DST, 0), (TU_INTRA, 1), (TU_INTRA | TU_SKIP, 0),
                         (TU_INTRA | TU_SKIP | TU_DST, 0), (TU_BYPASS | TU_INTRA, 2), (0, 0)):
        batch = _one_size_batch(2, 400, rng, bit_depth, qps, flags, c_idx)
        got = engine.residual(batch)
        assert_planes_equal(batch.geom, got, c_oracle.residual_batch(batch))


@pytest.mark.parametrize("log2n", [2, 3, 4, 5])
def test_scaling_lists_random_tables(engine, c_oracle, log2n):
    """Random ScalingFactor tables (1..255), intra and inter matrixId, all components."""
    rng = np.random.default_rng(500 + log2n)
    sf = {}
    for s in range(4):
        for m in range(2 if s == 3 else 6):
            sf[(s, m)] = rng.integers(1, 256, (4 << s, 4 << s))
    table = pack_scaling_factor(sf)
    for flags in (TU_INTRA, 0):
        for c_idx in ((0,) if log2n == 5 else (0, 1, 2)):
            batch = _one_size_batch(log2n, 40, rng, 10, tuple(range(0, 64)), flags, c_idx, table)
            got = engine.residual(batch)
            assert_planes_equal(batch.geom, got, c_oracle.residual_batch(batch))


def test_bypass_all_sizes(engine, c_oracle):
    rng = np.random.default_rng(3)
    for log2n in (2, 3, 4, 5):
        batch = _one_size_batch(log2n, 9, rng, 8, (30,), TU_BYPASS | TU_INTRA)
        got = engine.residual(batch)
        assert_planes_equal(batch.geom, got, c_oracle.residual_batch(batch))


def test_empty_batch(engine):
    geom = PicGeom(64, 64, 1, 8, 8)
    batch = ResidualBatch(geom, np.zeros(0, TU_DESC), np.zeros(0, np.int16))
    out = engine.residual(batch)
    assert not out.any()


def test_linearity_dc(engine):
    """A pure-DC block inverse-transforms to a constant block (size-independent property)."""
    rng = np.random.default_rng(11)
    for log2n in (2, 3, 4, 5):
        batch = _one_size_batch(log2n, 8, rng, 8, (28,), TU_INTRA, 1)
        n = 1 << log2n
        batch.coeffs[:] = 0
        batch.coeffs[::n * n] = rng.integers(-500, 500, 8)
        out = engine.residual(batch)
        pv = batch.geom.plane_view(out, 0, 1)
        for t in batch.tus:
            blk = pv[t["y"]:t["y"] + n, t["x"]:t["x"] + n]
            assert (blk == blk[0, 0]).all()


def test_bad_arguments_raise(engine):
    geom = PicGeom(64, 64, 1, 8, 8)
    tus = np.zeros(2, TU_DESC)
    tus["log2n"] = (2, 5)                      # not sorted largest-first
    with pytest.raises(ValueError):
        engine.residual(ResidualBatch(geom, tus, np.zeros(2048, np.int16)))
    tus = np.zeros(1, TU_DESC)
    tus["log2n"], tus["x"] = 3, 60             # leaves the plane
    with pytest.raises(ValueError):
        engine.residual(ResidualBatch(geom, tus, np.zeros(64, np.int16)))
    tus["x"], tus["coeff_off"] = 0, 100        # beyond the arena
    with pytest.raises(ValueError):
        engine.residual(ResidualBatch(geom, tus, np.zeros(64, np.int16)))


def test_descriptor_validation_cases(engine):
    """One-pass check of caller data at the C-ABI (api.cu: check_tus): every rule, first offender named."""
    geom = PicGeom(64, 64, 1, 8, 8)
    co = np.zeros(4096, np.int16)

    def one(**kw):
        t = np.zeros(1, TU_DESC)
        t["log2n"] = 2
        for k, v in kw.items():
            t[k] = v
        return t

    from p265_b200.residual_api import TU_PRESCALED
    bad = [one(c_idx=3), one(x=2), one(pic=1), one(log2n=3, flags=TU_SKIP), one(c_idx=1, flags=TU_DST),
           one(log2n=4, flags=TU_DST), one(qp=52), one(c_idx=1, x=32), one(log2n=5, y=48)]
    for t in bad:
        with pytest.raises(ValueError):
            engine.residual(ResidualBatch(geom, t, co))
    from p265_b200.scaling_list import default_scaling_factor
    sf = pack_scaling_factor(default_scaling_factor())
    with pytest.raises(ValueError):      # PRESCALED descriptors cannot be combined with a table
        engine.residual(ResidualBatch(geom, one(flags=TU_PRESCALED), co, scaling_factor=sf))
    # bin counts are caller data too: counts that do not match the list are rejected, never trusted
    t = np.zeros(3, TU_DESC)
    t["log2n"] = (3, 2, 2)
    t["x"] = (0, 8, 12)
    with pytest.raises(ValueError):
        engine.residual(ResidualBatch(geom, t, co, bins=(0, 0, 2, 1)))
    assert ResidualBatch(geom, t, co).bin_counts() == (0, 0, 1, 2)
    engine.residual(ResidualBatch(geom, t, co, bins=(0, 0, 1, 2)))


def test_dequant_matches_reference_outputs(engine, sanity_batch):
    """scaling.inverse_scaling: GPU d[] == the reference's own outputs on sanity.bin."""
    import os
    from conftest import GOLDEN
    batch, _ = sanity_batch
    ref = np.load(os.path.join(GOLDEN, "sanity_residual.npz"))["ref_scaled_yx"]
    assert np.array_equal(engine.dequant(batch), ref)


def test_ref_literal_matches_reference_outputs(engine, sanity_batch):
    """transform.py as written: GPU literal mode == the reference's own outputs."""
    import os
    from conftest import GOLDEN
    batch, _ = sanity_batch
    z = np.load(os.path.join(GOLDEN, "sanity_residual.npz"))
    got = engine.ref_literal(batch.tus, z["ref_scaled_yx"])
    assert np.array_equal(got, z["ref_literal_xy"])


def test_sf_general_and_replicated_paths_agree(engine, c_oracle):
    """Default (7.4.5-replicated) lists through both scaling-factor paths of the kernel."""
    batch = synth.residual_batch(small_cfg("4k10", 320, 192), n_pics=2)
    assert batch.sf_replicated is True
    ref = c_oracle.residual_batch(batch, zero_fill=False)
    assert_planes_equal(batch.geom, engine.residual(batch), ref)
    batch.sf_replicated = False
    assert_planes_equal(batch.geom, engine.residual(batch), ref)


def test_custom_dc_and_lists(engine, c_oracle):
    """Non-default lists with distinct DC values (still replicated -> fast path)."""
    from p265_b200 import scaling_list
    rng = np.random.default_rng(21)
    lists, dc = scaling_list.default_lists()
    for k in lists:
        lists[k] = [int(v) for v in rng.integers(1, 256, len(lists[k]))]
    for k in dc:
        dc[k] = int(rng.integers(1, 256))
    table = pack_scaling_factor(scaling_list.expand(lists, dc))
    batch = synth.residual_batch(small_cfg("4k10", 256, 128), n_pics=1, seed=5)
    batch = ResidualBatch(batch.geom, batch.tus, batch.coeffs, table, covers_all=True)
    assert batch.sf_replicated is True
    assert_planes_equal(batch.geom, engine.residual(batch), c_oracle.residual_batch(batch, zero_fill=False))


def test_many_items_per_warp(engine, c_oracle):
    """Enough TBs that every persistent warp loops over several work items of every size
    (exercises the double-buffered tile prefetch)."""
    batch = synth.residual_batch("1080p8", n_pics=6, n_unique=2)
    assert_planes_equal(batch.geom, engine.residual(batch), c_oracle.residual_batch(batch, zero_fill=False))


@pytest.mark.parametrize("log2n", [5, 4, 3, 2])
def test_each_size_bin_alone_at_scale(engine, c_oracle, log2n):
    """Four 4K pictures' worth of one TB size: every persistent warp loops over many items, every
    CTA slot of the SM is occupied (catches shared-memory layout overruns that small grids hide)."""
    from p265_b200.picture import ResidualBatch
    full = synth.residual_batch("4k10", n_pics=4, n_unique=2, seed=77)
    sel = np.ascontiguousarray(full.tus[full.tus["log2n"] == log2n])
    assert len(sel) > 15000
    b = ResidualBatch(full.geom, sel, full.coeffs, full.scaling_factor, covers_all=False)
    assert np.array_equal(engine.residual(b), c_oracle.residual_batch(b))
