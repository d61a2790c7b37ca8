"""GPU parity of the reconstruction step: Clip1(pred + residual)."""
import importlib
import sys
import types

import numpy as np
import pytest

from oracle import spec_oracle as so
from p265_b200.picture import PicGeom

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bit_depth", [8, 10, 12])
@pytest.mark.parametrize("size", [(64, 64), (200, 136), (3840, 2160)])
def test_reconstruct_planes(engine, bit_depth, size):
    w, h = size
    geom = PicGeom(w, h, 2 if w < 1000 else 1, bit_depth, bit_depth)
    rng = np.random.default_rng(w + bit_depth)
    dtype = np.uint8 if bit_depth <= 8 else np.uint16
    pred = rng.integers(0, 1 << bit_depth, geom.total_elems()).astype(dtype)
    res = rng.integers(-32768, 32768, geom.total_elems()).astype(np.int16)
    res[::3] = rng.integers(-40, 41, res[::3].size)          # mostly small residuals + full-range ones
    got = engine.reconstruct(pred, res, geom)
    for p in range(geom.n_pics):
        for c in range(3):
            want = so.reconstruct(geom.plane_view(pred, p, c), geom.plane_view(res, p, c), bit_depth)
            assert np.array_equal(geom.plane_view(got, p, c), want), (p, c)


def test_residual_then_reconstruct_then_sao_chain(engine, c_oracle):
    """The three launches back to back on one small picture batch, each vs its oracle."""
    from conftest import small_cfg
    from p265_b200 import synth
    batch = synth.residual_batch(small_cfg("4k10", 256, 128), n_pics=2)
    geom = batch.geom
    res = engine.residual(batch)
    assert np.array_equal(res, c_oracle.residual_batch(batch, zero_fill=False))
    rng = np.random.default_rng(3)
    pred = rng.integers(0, 1024, geom.total_elems()).astype(np.uint16)
    rec = engine.reconstruct(pred, res, geom)
    _, _, params = synth.sao_batch(256, 128, 10, n_pics=2, ctb_log2=6, seed=4)
    out = engine.sao(rec, geom, 6, params)
    want_rec = pred.copy()
    for p in range(2):
        for c in range(3):
            geom.plane_view(want_rec, p, c)[:] = so.reconstruct(geom.plane_view(pred, p, c),
                                                                geom.plane_view(res, p, c), 10)
    want = c_oracle.sao_batch(want_rec, geom, 6, params)
    for p in range(2):
        for c in range(3):
            assert np.array_equal(geom.plane_view(rec, p, c), geom.plane_view(want_rec, p, c))
            assert np.array_equal(geom.plane_view(out, p, c), geom.plane_view(want, p, c))


def test_dropin_reconstruction_module(engine):
    import p265_b200
    sys.path.insert(0, p265_b200.dropin_path())
    sys.modules.pop("reconstruction", None)
    try:
        mod = importlib.import_module("reconstruction")
        assert "p265_b200" in mod.__file__
        rng = np.random.default_rng(9)
        for bd, c_idx, l2 in ((8, 0, 2), (10, 1, 3), (8, 2, 4), (10, 0, 5)):
            n = 1 << l2
            pred = rng.integers(0, 1 << bd, (n, n))
            res = rng.integers(-2000, 2000, (n, n))
            sps = types.SimpleNamespace(bit_depth_y=bd, bit_depth_c=bd)
            pu = types.SimpleNamespace(c_idx=c_idx, origin_x=0, origin_y=0,
                                       cu=types.SimpleNamespace(ctx=types.SimpleNamespace(sps=sps)),
                                       predicted_samples=pred.copy(), transformed_samples=res.copy(),
                                       reconstructed_samples=np.zeros((n, n), np.int64))
            ret = mod.reconstruction(pu=pu, x0=0, y0=0, log2size=l2)
            assert np.array_equal(pu.reconstructed_samples, so.reconstruct(pred, res, bd))
            assert np.array_equal(ret, pu.reconstructed_samples)
    finally:
        sys.path.remove(p265_b200.dropin_path())
        sys.modules.pop("reconstruction", None)
