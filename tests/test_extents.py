"""Zero-extent codes (include/p265_b200.h: P265_TU_ZR_SHIFT / ZC_SHIFT): the host side
that derives them (picture.extent_codes, emit.PictureSink) and the device code that uses them, run on
the CPU through the host emulation of residual_core.cuh -- the shortened column / row passes must give
the oracle's residuals bit for bit (the oracle knows nothing about extents)."""
import ctypes as C
from collections import Counter

import numpy as np
import pytest

from conftest import small_cfg
from p265_b200 import synth
from p265_b200.picture import (TU_DESC, TU_LEVELS_MASK, TU_ZC_SHIFT, TU_ZR_SHIFT, PicGeom, ResidualBatch, extent_code,
                               extent_codes, set_extents, sort_by_size)
from test_host_core import host_core, run  # noqa: F401  (fixture + driver of the emulation)


def codes(tus):
    return ((tus["rsvd"] >> TU_ZR_SHIFT) & 3).astype(int), ((tus["rsvd"] >> TU_ZC_SHIFT) & 3).astype(int)


def test_extent_code_thresholds():
    assert extent_code(np.array([-1, 0, 7, 8, 15, 16, 31]), 32).tolist() == [2, 2, 2, 1, 1, 0, 0]
    assert extent_code(np.array([-1, 3, 4, 7, 8, 15]), 16).tolist() == [2, 2, 1, 1, 0, 0]


def test_extent_codes_of_hand_made_blocks():
    tus = np.zeros(4, TU_DESC)
    tus["log2n"] = (5, 5, 4, 3)
    tus["coeff_off"] = (0, 64, 128, 144)
    co = np.zeros(128 * 16 + 256 + 64, np.int16)
    a = co[:1024].reshape(32, 32)
    a[7, 15] = 1                      # rows < 8, columns < 16
    b = co[1024:2048].reshape(32, 32)
    b[16, 0] = -3                     # row 16: nothing known for rows; columns < 8
    c = co[2048:2304].reshape(16, 16)
    c[3, 8] = 5                       # rows < 4, column 8: nothing known
    co[2304:] = 9                     # an 8x8 TB: sizes below 16 carry no codes
    zr, zc = extent_codes(tus, co)
    assert zr.tolist() == [2, 0, 2, 0] and zc.tolist() == [1, 2, 0, 0]
    tus["rsvd"] = 5                   # a level count in the low bits stays
    set_extents(tus, zr, zc)
    assert (tus["rsvd"] & TU_LEVELS_MASK).tolist() == [5, 5, 5, 5]
    assert codes(tus)[0].tolist() == [2, 0, 2, 0] and codes(tus)[1].tolist() == [1, 2, 0, 0]
    order = sort_by_size(tus)
    assert order["log2n"].tolist() == [5, 5, 4, 3]


def test_ordering_rule_moves_whole_items():
    """Items (4 consecutive 16x16 TBs, 2 consecutive 32x32 TBs) are ordered by their weakest promise and stay intact."""
    n = 4 * 6 + 3                                  # six items of 16x16 TBs and a partial one
    tus = np.zeros(n + 5, TU_DESC)
    tus["log2n"][:5] = 5
    tus["log2n"][5:] = 4
    tus["x"] = np.arange(n + 5)                    # identity tag
    zr = np.zeros(n + 5, int)
    zc = np.zeros(n + 5, int)
    zr[5:] = np.repeat([2, 0, 1, 2, 2, 0, 1], 4)[:n]
    zc[5:] = np.repeat([2, 0, 1, 2, 1, 0, 2], 4)[:n]
    zr[5 + 12] = 0                                 # one TB of the fourth item promises nothing -> item code (0, 2)
    zr[:5], zc[:5] = [2, 2, 0, 1, 2], [2, 2, 0, 1, 2]
    set_extents(tus, zr, zc)
    out = sort_by_size(tus)
    assert out["log2n"].tolist() == [5] * 5 + [4] * n
    items16 = out["x"][5:5 + 24].reshape(6, 4)
    assert all((row == row[0] + np.arange(4)).all() and (row[0] - 5) % 4 == 0 for row in items16)    # intact
    first = ((items16[:, 0] - 5) // 4).tolist()
    assert first == [1, 5, 3, 2, 4, 0]             # codes (0,0) (0,0) (0,2) (1,1) (2,1) (2,2)
    assert out["x"][5 + 24:].tolist() == list(range(5 + 24, 5 + 27))                                   # the partial item stays last
    assert out["x"][:5].tolist() == [2, 3, 0, 1, 4]  # 32x32: items (2,3) code (0,0), (0,1) code (2,2); the odd TB last


def test_config3_model_leaves_nothing_to_skip():
    """Why the zero-aware passes cannot move the BASELINE config-3 number: SURVEY 8(d)'s
    coefficient model puts a level into the last quarter of the rows of nearly every 32x32 TB."""
    b = synth.residual_batch(small_cfg("4k10", 1024, 512), n_pics=1, extents=True)
    zr, zc = codes(b.tus)
    big = b.tus["log2n"] == 5
    assert big.sum() > 100 and (zr[big] == 0).mean() > 0.99 and (zc[big] == 0).mean() > 0.99


def test_sanity_bin_extent_distribution(sanity_batch):
    """The only real stream of the reference: the distribution synth.SANITY_EXTENT_MIX restates."""
    batch, _ = sanity_batch
    zr, zc = extent_codes(batch.tus, batch.coeffs)
    m = batch.tus["log2n"] == 4
    dist = Counter(zip(zr[m].tolist(), zc[m].tolist()))
    print("16x16:", dist.most_common(), " 32x32:", Counter(zip(zr[batch.tus["log2n"] == 5].tolist(),
                                                              zc[batch.tus["log2n"] == 5].tolist())).most_common())
    for (code, p) in synth.SANITY_EXTENT_MIX:
        assert abs(dist[code] / m.sum() - p) < 0.01
    assert 0.55 < 1 - dist[(0, 0)] / m.sum() < 0.65     # 59 % of its 16x16 TBs promise something


@pytest.mark.parametrize("stress", [False, True])
@pytest.mark.parametrize("name", ["4k10_lowfreq", "1080p8_lowfreq"])
def test_shortened_passes_match_oracle(host_core, c_oracle, name, stress):  # noqa: F811
    if name == "1080p8_lowfreq":
        synth.CONFIGS[name] = dict(synth.CONFIGS["1080p8"], extent_mix=synth.SANITY_EXTENT_MIX)
    batch = synth.residual_batch(small_cfg(name, 512, 384), n_pics=2, stress=stress, extents=True)
    zr, zc = codes(batch.tus)
    big = batch.tus["log2n"] >= 4
    assert len(set(zip(zr[big].tolist(), zc[big].tolist()))) >= 8       # (nearly) every (row, column) code pair occurs
    ref = c_oracle.residual_batch(batch, zero_fill=True)
    for replicated in ((False, True) if batch.scaling_factor is not None else (False,)):
        assert np.array_equal(run(host_core, c_oracle, batch, replicated), ref)


def test_every_code_pair_alone_and_mixed_items(host_core, c_oracle):  # noqa: F811
    """One code pair per batch (every work item runs exactly that pair of passes), coefficients right up to
    the promised boundary (last row / column inside the extent is non-zero), full-range levels; then the
    unsorted list, where the TBs of an item disagree and the item must take the weakest promise."""
    rng = np.random.default_rng(7)
    geom = PicGeom(128, 64, 1, 10, 10)
    for log2n in (5, 4):
        n = 1 << log2n
        per_row = 128 // n
        cnt = per_row * (64 // n)
        all_tus, all_co = [], []
        for zr in range(3):
            for zc in range(3):
                tus = np.zeros(cnt, TU_DESC)
                tus["log2n"], tus["qp"], tus["flags"] = log2n, 34 + rng.integers(0, 16, cnt), 8
                tus["x"] = (np.arange(cnt) % per_row) * n
                tus["y"] = (np.arange(cnt) // per_row) * n
                tus["coeff_off"] = np.arange(cnt) * (n * n // 16)
                blk = np.zeros((cnt, n, n), np.int16)
                h, w = n >> zr, n >> zc
                blk[:, :h, :w] = rng.integers(-32768, 32768, (cnt, h, w))
                blk[:, h - 1, w - 1] |= 1
                set_extents(tus, np.full(cnt, zr), np.full(cnt, zc))
                batch = ResidualBatch(geom, tus, blk.reshape(-1), covers_all=True)
                assert np.array_equal(run(host_core, c_oracle, batch, False), c_oracle.residual_batch(batch))
                all_tus.append(tus)
                all_co.append(blk.reshape(-1))
        # mixed: TB i of every code pair at the same place is not allowed (overlap) -> one picture per pair
        geom9 = PicGeom(128, 64, 9, 10, 10)
        tus = np.concatenate(all_tus)
        tus["pic"] = np.repeat(np.arange(9), cnt)
        tus["coeff_off"] = np.arange(9 * cnt) * (n * n // 16)
        perm = rng.permutation(len(tus))          # items now mix code pairs
        batch = ResidualBatch(geom9, np.ascontiguousarray(tus[perm]), np.concatenate(all_co), covers_all=True)
        assert np.array_equal(run(host_core, c_oracle, batch, False), c_oracle.residual_batch(batch))
