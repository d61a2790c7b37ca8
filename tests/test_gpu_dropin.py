"""GPU parity of the drop-in modules (`scaling`, `transform`, `sao` by bare name, the
reference's function surface) against the reference's OWN outputs (committed fixtures)
and against the spec oracle."""
import importlib
import os
import sys
import types

import numpy as np
import pytest

from conftest import GOLDEN, REPO
from oracle import spec_oracle as so

sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
from make_fixtures import fake_pu  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dropin(engine):
    import p265_b200
    sys.path.insert(0, p265_b200.dropin_path())
    for name in ("scaling", "transform"):
        sys.modules.pop(name, None)
    mods = types.SimpleNamespace(scaling=importlib.import_module("scaling"),
                                 transform=importlib.import_module("transform"))
    assert "p265_b200" in mods.scaling.__file__ and "p265_b200" in mods.transform.__file__
    yield mods
    mods.transform.set_mode("spec")
    sys.path.remove(p265_b200.dropin_path())
    for name in ("scaling", "transform"):
        sys.modules.pop(name, None)


def _cases():
    rr = np.load(os.path.join(GOLDEN, "random_reference.npz"))
    sf = so.expand_scaling_factor(*so.default_scaling_lists())
    sf_ref = [[sf.get((s, m)) for m in range(6)] for s in range(4)]
    for i in range(int(rr["n"])):
        yield (rr["levels_xy_%d" % i], int(rr["c_idx_%d" % i]), int(rr["qp_%d" % i]), int(rr["bit_depth_%d" % i]),
               sf_ref if bool(rr["use_sf_%d" % i]) else None, rr["scaled_xy_%d" % i], rr["literal_xy_%d" % i])


def test_per_tb_calls_match_the_reference_outputs(dropin):
    """inverse_scaling == scaling.py's outputs; inverse_transform in ref_literal mode ==
    transform.py's outputs as written; in spec mode == the H.265 oracle."""
    n_checked = 0
    for lv, c_idx, qp, bd, sf, ref_scaled, ref_literal in _cases():
        l2 = int(np.log2(lv.shape[0]))
        pu = fake_pu(lv, c_idx, qp, bd, sf=sf)
        assert dropin.scaling.inverse_scaling(pu=pu, x0=0, y0=0, log2size=l2) is None
        assert np.array_equal(pu.scaled_samples, ref_scaled)
        dropin.transform.set_mode("ref_literal")
        dropin.transform.inverse_transform(pu=pu, x0=0, y0=0, log2size=l2)
        assert np.array_equal(pu.transformed_samples, ref_literal)
        dropin.transform.set_mode("spec")
        dropin.transform.inverse_transform(pu=pu, x0=0, y0=0, log2size=l2)
        want = so.sat16(so.inverse_transform_xy(ref_scaled, l2, 1 if (l2 == 2 and c_idx == 0) else 0, bd))
        assert np.array_equal(pu.transformed_samples, want)
        n_checked += 1
    assert n_checked == 52


def test_sub_block_of_a_larger_pu(dropin):
    """The reference calls the functions on TB-sized windows of a larger PU array
    (start_x = x0 - origin_x, intra.py:34-37, scaling.py:7-9)."""
    rng = np.random.default_rng(4)
    lv = rng.integers(-200, 200, (8, 8))
    pu = fake_pu(lv, 0, 30, 8)
    pu.origin_x, pu.origin_y = 16, 32
    pu.scaled_samples = np.zeros((16, 16), np.int64)
    pu.transformed_samples = np.zeros((16, 16), np.int64)
    pu.cu.tu.get_trans_coeff_level = lambda x, y, c: int(lv[x - 24][y - 40])
    dropin.scaling.inverse_scaling(pu=pu, x0=24, y0=40, log2size=3)
    dropin.transform.inverse_transform(pu=pu, x0=24, y0=40, log2size=3)
    d = so.inverse_scaling(lv, 30, 8, 3)
    assert np.array_equal(pu.scaled_samples[8:16, 8:16], d)
    assert np.array_equal(pu.transformed_samples[8:16, 8:16], so.inverse_transform_xy(d, 3, 0, 8))
    assert not pu.scaled_samples[:8].any() and not pu.transformed_samples[:, :8].any()


def test_bypass_and_bad_arguments(dropin):
    lv = np.arange(16).reshape(4, 4) - 8
    pu = fake_pu(lv, 1, 30, 8)
    pu.cu.cu_transquant_bypass_flag = 1
    dropin.scaling.inverse_scaling(pu=pu, x0=0, y0=0, log2size=2)       # reference: ValueError
    dropin.transform.inverse_transform(pu=pu, x0=0, y0=0, log2size=2)
    assert np.array_equal(pu.scaled_samples, lv) and np.array_equal(pu.transformed_samples, lv)
    with pytest.raises(ValueError):
        dropin.scaling.inverse_scaling(pu=pu, x0=0, y0=0, log2size=6)
    with pytest.raises(ValueError):
        dropin.transform.inverse_transform_1d(np.zeros(8), 2, 0)


def test_inverse_transform_1d_and_tables(dropin):
    assert np.array_equal(np.array(dropin.transform.trans_matrix_type0), so.DCT32)
    assert np.array_equal(np.array(dropin.transform.trans_matrix_type1), so.DST4)
    rng = np.random.default_rng(8)
    for l2 in (2, 3, 4, 5):
        x = rng.integers(-32768, 32768, 1 << l2)
        dropin.transform.set_mode("spec")
        assert np.array_equal(dropin.transform.inverse_transform_1d(x, l2, 0), so.inverse_transform_1d(x, l2, 0))
        dropin.transform.set_mode("ref_literal")
        n = 1 << l2
        c = so.DCT32[:n, :: 32 // n]
        assert np.array_equal(dropin.transform.inverse_transform_1d(x, l2, 0), c @ x)
    x = rng.integers(-1000, 1000, 4)
    dropin.transform.set_mode("spec")
    assert np.array_equal(dropin.transform.inverse_transform_1d(x, 2, 1), x @ so.DST4)
    dropin.transform.set_mode("ref_literal")
    assert np.array_equal(dropin.transform.inverse_transform_1d(x, 2, 1), so.DST4 @ x)
    dropin.transform.set_mode("spec")


def test_batched_flush_of_parsed_pictures(engine, c_oracle, dropin, parsed_sanity, sanity_batch):
    """Two-pass driver: parse everything with the reference, then one residual launch per
    picture; the per-TB functions afterwards only copy out of the planes."""
    from p265_b200 import packer, residual_api
    imgs, sps, _ = parsed_sanity
    assert len(imgs) == 3
    fixture, _ = sanity_batch
    n_tbs = 0
    for p, img in enumerate(imgs):
        cache = residual_api.flush_picture(img, sps)
        assert np.array_equal(cache.planes, c_oracle.residual_batch(cache.batch))
        assert np.array_equal(cache.scaled, c_oracle.dequant_batch(cache.batch))
        n_tbs += len(cache.batch.tus)
        # a per-TB call on a real CU/TU of this picture is served from the cache
        t = cache.batch.tus[cache.batch.tus["c_idx"] == 0][5]
        n = 1 << int(t["log2n"])
        cu = next(c for a in sorted(img.ctus) for c in packer._leaf_cus(img.ctus[a])
                  if c.contain(int(t["x"]), int(t["y"])))
        cu._p265_b200_img = img
        pu = types.SimpleNamespace(c_idx=0, origin_x=cu.x, origin_y=cu.y, cu=cu,
                                   scaled_samples=np.zeros((cu.size, cu.size), np.int64),
                                   transformed_samples=np.zeros((cu.size, cu.size), np.int64))
        x0, y0 = int(t["x"]), int(t["y"])
        dropin.scaling.inverse_scaling(pu=pu, x0=x0, y0=y0, log2size=int(t["log2n"]))
        dropin.transform.inverse_transform(pu=pu, x0=x0, y0=y0, log2size=int(t["log2n"]))
        d, r = cache.block(0, x0, y0, n)
        sx, sy = x0 - cu.x, y0 - cu.y
        assert np.array_equal(pu.scaled_samples[sx:sx + n, sy:sy + n], d.T)
        assert np.array_equal(pu.transformed_samples[sx:sx + n, sy:sy + n], r.T)
    assert n_tbs == len(fixture.tus) == 5982


def test_filter_picture_with_parsed_sao_params(engine, c_oracle, parsed_sanity):
    from p265_b200 import packer, sao_api, synth
    from p265_b200.picture import PicGeom
    imgs, sps, pps = parsed_sanity
    rng = np.random.default_rng(12)
    geom = PicGeom(352, 288, 1, 8, 8)
    for img in imgs:
        buf = np.zeros(geom.total_elems(), np.uint8)
        synth.sao_picture(352, 288, 8, rng, geom, buf, 0)
        planes = tuple(geom.plane_view(buf, 0, c).copy() for c in range(3))
        got = sao_api.filter_picture(planes, img, sps, pps)
        params = packer.sao_params_from_picture(img, sps)
        want = c_oracle.sao_batch(buf, geom, 6, params)
        assert (params["type"] != 0).any()
        for c in range(3):
            assert np.array_equal(got[c], geom.plane_view(want, 0, c))
