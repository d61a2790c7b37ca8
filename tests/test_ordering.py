"""The ordering rule of the packed format (picture.size_kind_order): any order inside a size bin is correct, this
one is the fast one -- sizes 32, 16, 8, 4; kinds clustered inside a size; 16x16 / 32x32 TBs in the order they came
(decoding order: neighbours of one quadrant stay one work item); 8x8 / 4x4 TBs in raster order of their plane."""
import numpy as np

from p265_b200 import synth
from p265_b200.picture import TU_BYPASS, TU_DESC, TU_DST, TU_SKIP, PicGeom, size_kind_order, sort_by_size
from conftest import small_cfg


import pytest


@pytest.mark.parametrize("skip_share", [0.25, 0.02])
def test_sizes_kinds_and_raster_order_of_small_tbs(skip_share):
    """Transform-skip TBs are a cluster of their own only while they are rare (<= 3 % of the 4x4 TBs): pulled out
    of the list they break the raster runs, left in they make their warps run both paths."""
    rng = np.random.default_rng(1)
    n = 4000
    t = np.zeros(n, TU_DESC)
    t["log2n"] = rng.integers(2, 6, n)
    t["c_idx"] = rng.integers(0, 3, n)
    t["pic"] = rng.integers(0, 3, n)
    t["x"] = rng.integers(0, 64, n) * 8
    t["y"] = rng.integers(0, 32, n) * 8
    t["qp"] = 30
    rest = (1 - skip_share) / 3
    t["flags"] = np.where(t["log2n"] == 2, rng.choice([0, TU_DST, TU_SKIP, TU_BYPASS], n, p=[rest, rest, skip_share, rest]),
                          rng.choice([0, TU_BYPASS], n))
    clustered = TU_DST | TU_BYPASS | (TU_SKIP if skip_share < 0.03 else 0)
    t["coeff_off"] = np.arange(n)                     # identity tag
    out = sort_by_size(t, PicGeom(512, 256, 3, 8, 8))
    assert sorted(out["coeff_off"].tolist()) == list(range(n))                      # a permutation
    assert (np.diff(out["log2n"].astype(int)) <= 0).all()                           # 32, 16, 8, 4
    for l2 in (2, 3, 4, 5):
        b = out[out["log2n"] == l2]
        kind = (b["flags"] & clustered).astype(int)
        assert (np.diff(kind) >= 0).all()                                           # kinds clustered
        for k in np.unique(kind):
            c = b[kind == k]
            if l2 >= 4:
                assert (np.diff(c["coeff_off"].astype(int)) > 0).all()              # arrival (decoding) order kept
            else:
                pos = (c["pic"].astype(np.int64) * 4 + c["c_idx"]) * 2 ** 32 + c["y"].astype(np.int64) * 2 ** 16 + c["x"]
                assert (np.diff(pos) >= 0).all()                                    # raster order of the plane


def test_rule_is_idempotent_and_shared_by_the_synthetic_workloads():
    b = synth.residual_batch(small_cfg("4k10", 256, 192), n_pics=2)
    again = sort_by_size(b.tus, b.geom)
    assert np.array_equal(again, b.tus)
    order = size_kind_order(b.tus, b.geom)
    assert np.array_equal(order, np.arange(len(b.tus)))
