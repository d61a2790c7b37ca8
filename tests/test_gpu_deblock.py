"""GPU parity of the deblocking kernel (through the C-ABI) against the plain-C oracle, which
tests/test_decode_sanity.py pins to libavcodec's output."""
import numpy as np
import pytest

from p265_b200 import synth
from p265_b200.picture import DBK_NO_FILTER

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("width,height,bit_depth,ctb_log2,dense", [
    (64, 64, 8, 6, True), (128, 72, 8, 5, False), (352, 288, 8, 6, False), (200, 120, 10, 4, True),
    (1920, 1088, 8, 6, False), (264, 136, 10, 6, False), (8, 8, 8, 4, True), (16, 8, 10, 4, True),
    (128, 64, 12, 6, True), (128, 64, 11, 6, True), (136, 72, 9, 5, False),
])
def test_deblock_matches_oracle(engine, c_oracle, width, height, bit_depth, ctb_log2, dense):
    geom, buf, blk, ctb = synth.deblock_batch(width, height, bit_depth, 2, ctb_log2, seed=width + height,
                                              dense=dense)
    got = engine.deblock(buf, geom, ctb_log2, blk, ctb)
    want = c_oracle.deblock_batch(buf, geom, ctb_log2, blk, ctb)
    assert np.array_equal(got, want)
    if width >= 64:
        assert not np.array_equal(got, buf)


def test_deblock_4k_10bit_batch(engine, c_oracle):
    geom, buf, blk, ctb = synth.deblock_batch(3840, 2160, 10, 3, 6, n_unique=2)
    got = engine.deblock(buf, geom, 6, blk, ctb)
    want = c_oracle.deblock_batch(buf, geom, 6, blk, ctb)
    assert np.array_equal(got, want)
    # idempotence does not hold for deblocking, but a map without edges is the identity
    none = engine.deblock(buf, geom, 6, blk & np.uint16(0xFF00), ctb)
    assert np.array_equal(none, buf)


def test_extreme_qp_offsets_and_no_filter_everywhere(engine, c_oracle):
    rng = np.random.default_rng(5)
    for qp_lo, qp_hi in ((0, 6), (46, 51), (-12, 0)):
        geom, buf, blk, ctb = synth.deblock_batch(256, 128, 10, 1, 6, seed=qp_hi + 20, dense=True)
        b, c = synth.deblock_edge_map(256, 128, 6, rng, qp_lo, qp_hi, no_filter_frac=0.3, dense=True)
        got = engine.deblock(buf, geom, 6, b, c)
        assert np.array_equal(got, c_oracle.deblock_batch(buf, geom, 6, b, c))
    allnf = blk | np.uint16(DBK_NO_FILTER)
    assert np.array_equal(engine.deblock(buf, geom, 6, allnf, ctb), buf)


def test_noise_picture_hits_every_decision(engine, c_oracle):
    """Full-range noise: most segments fail the beta test, the rest go every other way."""
    geom, buf, blk, ctb = synth.deblock_batch(512, 256, 8, 1, 6, seed=9, dense=True)
    rng = np.random.default_rng(10)
    buf[:] = rng.integers(0, 256, buf.size)
    smooth = (np.arange(buf.size) // 7 % 256).astype(np.uint8)
    buf[::2] = smooth[::2]
    got = engine.deblock(buf, geom, 6, blk, ctb)
    assert np.array_equal(got, c_oracle.deblock_batch(buf, geom, 6, blk, ctb))


def test_full_range_noise_at_every_bit_depth(engine, c_oracle):
    """Worst-case magnitudes for the packed arithmetic (bit depths <= 11) and the 32-bit path (12)."""
    for bd in (8, 9, 10, 11, 12):
        geom, buf, blk, ctb = synth.deblock_batch(256, 128, bd, 1, 6, seed=40 + bd, dense=True)
        rng = np.random.default_rng(bd)
        noise = rng.integers(0, 1 << bd, buf.size).astype(buf.dtype)
        flat = np.where(rng.random(buf.size) < 0.5, 0, (1 << bd) - 1).astype(buf.dtype)
        for data in (noise, flat, np.where(np.arange(buf.size) % 16 < 8, noise, buf)):
            got = engine.deblock(data, geom, 6, blk, ctb)
            assert np.array_equal(got, c_oracle.deblock_batch(data, geom, 6, blk, ctb)), bd


def test_bad_arguments(engine):
    geom, buf, blk, ctb = synth.deblock_batch(64, 64, 8, 1, 6)
    with pytest.raises(ValueError):
        engine.deblock(buf, geom, 6, blk[:, :4], ctb)
    with pytest.raises(ValueError):
        engine.deblock(buf, geom, 7, blk, ctb)
    bad = blk.copy()
    bad[0, 0, 0] |= 3
    with pytest.raises(ValueError):
        engine.deblock(buf, geom, 6, bad, ctb)
