"""GPU parity of the zero-aware passes (include/p265_b200.h: zero-extent codes) through the
C-ABI, bit-exact against the C oracle, which knows nothing about extents.  Dense arena: the codes are the
caller's promise in `rsvd`; packed stream: unpack_kernel derives them from the significance bitmap."""
import numpy as np
import pytest

from conftest import small_cfg
from p265_b200 import synth
from p265_b200.picture import (TU_DESC, TU_ZC_SHIFT, TU_ZR_SHIFT, PicGeom, ResidualBatch, set_extents)
from test_gpu_residual import assert_planes_equal

pytestmark = pytest.mark.gpu


def lowfreq(name):
    key = name + "_lowfreq"
    if key not in synth.CONFIGS:
        synth.CONFIGS[key] = dict(synth.CONFIGS[name], extent_mix=synth.SANITY_EXTENT_MIX)
    return key


@pytest.mark.parametrize("name", ["1080p8", "4k10"])
@pytest.mark.parametrize("stress", [False, True])
def test_dense_arena_with_codes(engine, c_oracle, name, stress):
    batch = synth.residual_batch(small_cfg(lowfreq(name), 512, 320), n_pics=3, stress=stress, extents=True)
    big = batch.tus["log2n"] >= 4
    pairs = set(zip(((batch.tus["rsvd"][big] >> TU_ZR_SHIFT) & 3).tolist(), ((batch.tus["rsvd"][big] >> TU_ZC_SHIFT) & 3).tolist()))
    assert len(pairs) >= 8
    ref = c_oracle.residual_batch(batch, zero_fill=False)
    assert_planes_equal(batch.geom, engine.residual(batch), ref)
    if batch.scaling_factor is not None:     # any table (SF_GENERAL) as well as the 7.4.5 replicated form
        general = ResidualBatch(batch.geom, batch.tus, batch.coeffs, batch.scaling_factor, True, sf_replicated=False)
        assert_planes_equal(batch.geom, engine.residual(general), ref)


def test_config3_lowfreq_full_size(engine, c_oracle):
    batch = synth.residual_batch("4k10_lowfreq", n_pics=1, extents=True)
    assert_planes_equal(batch.geom, engine.residual(batch), c_oracle.residual_batch(batch, zero_fill=False))


@pytest.mark.parametrize("stress", [False, True])
def test_packed_stream_derives_codes_on_the_device(engine, c_oracle, stress):
    batch = synth.residual_batch(small_cfg(lowfreq("4k10"), 512, 320), n_pics=2, stress=stress)   # no codes on the host
    ref = c_oracle.residual_batch(batch, zero_fill=False)
    packed = batch.packed()
    assert not (packed.tus["rsvd"] >> TU_ZR_SHIFT).any()
    assert_planes_equal(batch.geom, engine.residual(packed), ref)
    # codes a host wrote into a packed descriptor are not trusted: the bitmap decides.  Promise "first quarter
    # only" for every TB of a batch whose TBs are mostly full: the result must not change.
    liar = batch.packed()
    set_extents(liar.tus, np.full(len(liar.tus), 2), np.full(len(liar.tus), 2))
    assert_planes_equal(batch.geom, engine.residual(liar), ref)


def test_mixed_items_take_the_weakest_promise(engine, c_oracle):
    """Unsorted codes: the TBs of one work item (2 of 32x32, 4 of 16x16) disagree."""
    batch = synth.residual_batch(small_cfg(lowfreq("4k10"), 512, 320), n_pics=2, extents=True)
    rng = np.random.default_rng(3)
    tus = batch.tus.copy()
    b = batch.bin_counts()
    for lo, cnt in ((0, b[0]), (b[0], b[1])):       # shuffle inside the 32x32 bin and inside the 16x16 bin
        tus[lo:lo + cnt] = tus[lo:lo + cnt][rng.permutation(cnt)]
    mixed = ResidualBatch(batch.geom, tus, batch.coeffs, batch.scaling_factor, True, batch.sf_replicated)
    assert_planes_equal(batch.geom, engine.residual(mixed), c_oracle.residual_batch(batch, zero_fill=False))


def test_boundary_coefficients_and_every_pair(engine, c_oracle):
    """A level on the last row / column inside every promised extent, full-range levels, one code pair per picture."""
    rng = np.random.default_rng(11)
    geom = PicGeom(256, 128, 9, 10, 10)
    tus_all, co_all, off = [], [], 0
    for log2n in (5, 4):
        n = 1 << log2n
        per_row, cnt = 256 // n, (256 // n) * (64 // n)
        for pic in range(9):
            zr, zc = divmod(pic, 3)
            tus = np.zeros(cnt, TU_DESC)
            tus["log2n"], tus["qp"], tus["flags"], tus["pic"] = log2n, 34 + rng.integers(0, 16, cnt), 8, pic
            tus["x"] = (np.arange(cnt) % per_row) * n
            tus["y"] = (np.arange(cnt) // per_row) * n + (0 if log2n == 5 else 64)
            tus["coeff_off"] = (off + np.arange(cnt) * n * n) >> 4
            blk = np.zeros((cnt, n, n), np.int16)
            h, w = n >> zr, n >> zc
            blk[:, :h, :w] = rng.integers(-32768, 32768, (cnt, h, w))
            blk[:, h - 1, w - 1] |= 1
            set_extents(tus, np.full(cnt, zr), np.full(cnt, zc))
            tus_all.append(tus)
            co_all.append(blk.reshape(-1))
            off += cnt * n * n
    batch = ResidualBatch(geom, np.concatenate(tus_all), np.concatenate(co_all))
    ref = c_oracle.residual_batch(batch, zero_fill=True)
    assert_planes_equal(geom, engine.residual(batch), ref)
    assert_planes_equal(geom, engine.residual(batch.packed()), ref)


def test_code_3_is_rejected(engine):
    batch = synth.residual_batch(small_cfg("4k10", 256, 128), n_pics=1)
    tus = batch.tus.copy()
    tus["rsvd"][0] |= 3 << TU_ZC_SHIFT
    with pytest.raises(ValueError, match="zero-extent"):
        engine.residual(ResidualBatch(batch.geom, tus, batch.coeffs, batch.scaling_factor, True))
