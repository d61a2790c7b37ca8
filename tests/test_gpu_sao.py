"""GPU parity of the SAO kernel through the C-ABI, bit-exact against the C oracle."""
import numpy as np
import pytest

from p265_b200 import synth
from p265_b200.picture import AVAIL_ALL, SAO_CTB, PicGeom

pytestmark = pytest.mark.gpu


def assert_planes_equal(geom, got, ref):
    for p in range(geom.n_pics):
        for c in range(3):
            a, b = geom.plane_view(got, p, c), geom.plane_view(ref, p, c)
            if not np.array_equal(a, b):
                ys, xs = np.nonzero(a != b)
                raise AssertionError("pic %d comp %d: %d mismatches, first at (x=%d,y=%d): got %d want %d"
                                     % (p, c, ys.size, xs[0], ys[0], a[ys[0], xs[0]], b[ys[0], xs[0]]))


@pytest.mark.parametrize("bit_depth", [8, 10])
@pytest.mark.parametrize("ctb_log2", [4, 5, 6])
@pytest.mark.parametrize("size", [(256, 128), (200, 136), (72, 56), (8, 8)])
def test_small_pictures(engine, c_oracle, bit_depth, ctb_log2, size):
    """Pictures that end inside a CTB, chroma widths of 4 mod 8, every CTB size."""
    w, h = size
    geom, rec, params = synth.sao_batch(w, h, bit_depth, n_pics=2, ctb_log2=ctb_log2,
                                        seed=100 + bit_depth + ctb_log2 + w)
    got = engine.sao(rec, geom, ctb_log2, params)
    assert_planes_equal(geom, got, c_oracle.sao_batch(rec, geom, ctb_log2, params))


def test_config4_full_4k(engine, c_oracle):
    geom, rec, params = synth.sao_batch(3840, 2160, 10, n_pics=1)
    got = engine.sao(rec, geom, 6, params)
    ref = c_oracle.sao_batch(rec, geom, 6, params)
    assert_planes_equal(geom, got, ref)
    assert not np.array_equal(geom.plane_view(ref, 0, 0), geom.plane_view(rec, 0, 0))


def test_config4_two_slices_no_filter_across(engine, c_oracle):
    geom, rec, params = synth.sao_batch(3840, 2160, 10, n_pics=1, two_slices=True, seed=26505)
    assert (params["avail"] != AVAIL_ALL).any()
    got = engine.sao(rec, geom, 6, params)
    assert_planes_equal(geom, got, c_oracle.sao_batch(rec, geom, 6, params))


@pytest.mark.parametrize("bit_depth", [8, 10, 12])
@pytest.mark.parametrize("ctb_log2", [4, 5, 6])
def test_all_edge_interior_ctbs(engine, c_oracle, bit_depth, ctb_log2):
    """Every CTB edge-filtered, all classes, every neighbour available: the CTBs that do not touch
    the picture border take the kernel's mask-free path, the border CTBs the masked one; one CTB in
    the middle has a neighbour switched off and must fall back to the masked path."""
    geom, rec, params = synth.sao_batch(448, 320, bit_depth, n_pics=2, ctb_log2=ctb_log2, seed=300 + ctb_log2)
    rng = np.random.default_rng(301 + bit_depth)
    params["type"][:] = 2
    params["eo_class"][:] = rng.integers(0, 4, params["eo_class"].shape)
    params["offset_val"][:] = (7, 2, -2, -7)
    params["avail"][:] = AVAIL_ALL
    params["avail"][0, 2, 3] = AVAIL_ALL & ~(1 << 5)   # right neighbour of one interior CTB unavailable
    got = engine.sao(rec, geom, ctb_log2, params)
    ref = c_oracle.sao_batch(rec, geom, ctb_log2, params)
    assert_planes_equal(geom, got, ref)
    assert not np.array_equal(geom.plane_view(ref, 0, 0), geom.plane_view(rec, 0, 0))


@pytest.mark.parametrize("bit_depth", [8, 10])
def test_random_availability_masks(engine, c_oracle, bit_depth):
    geom, rec, params = synth.sao_batch(320, 256, bit_depth, n_pics=2, ctb_log2=5, seed=77)
    rng = np.random.default_rng(78)
    params["avail"] = rng.integers(0, 512, params["avail"].shape)
    params["type"][..., 0] = 2                       # all luma CTBs edge offset
    params["eo_class"][..., 0] = rng.integers(0, 4, params["eo_class"][..., 0].shape)
    params["offset_val"][..., 0, :] = (5, 3, -3, -5)
    got = engine.sao(rec, geom, 5, params)
    assert_planes_equal(geom, got, c_oracle.sao_batch(rec, geom, 5, params))


@pytest.mark.parametrize("bit_depth", [8, 10])
@pytest.mark.parametrize("width,height,ctb_log2", [(256, 192, 6), (136, 72, 6), (72, 40, 5), (24, 24, 4)])
def test_no_filter_blocks(engine, c_oracle, bit_depth, width, height, ctb_log2):
    """pcm + pcm_loop_filter_disabled / cu_transquant_bypass blocks keep their samples.  Odd
    numbers of 8x8 block columns matter: a chroma strip at the right edge then owns a single
    flag (found by the libavcodec fuzz stream main10_lists_bypass)."""
    geom, rec, params = synth.sao_batch(width, height, bit_depth, n_pics=2, ctb_log2=ctb_log2, seed=5)
    rng = np.random.default_rng(6)
    nf = (rng.random((2, height // 8, width // 8)) < 0.4).astype(np.uint8)
    got = engine.sao(rec, geom, ctb_log2, params, no_filter=nf)
    assert_planes_equal(geom, got, c_oracle.sao_batch(rec, geom, ctb_log2, params, nf))


def test_extreme_offsets_clip(engine, c_oracle):
    """Offsets +-31 on samples at 0 and max: both sides of Clip1."""
    geom, rec, params = synth.sao_batch(128, 128, 10, n_pics=1, ctb_log2=6, seed=9)
    rec[:] = np.where(np.random.default_rng(1).random(rec.size) < 0.5, 0, 1023).astype(rec.dtype)
    params["type"][:] = 1
    params["band_pos"][..., 0] = 0
    params["band_pos"][..., 1] = 28
    params["band_pos"][..., 2] = 30                # wraps around band 31 -> 0
    params["offset_val"][:] = (-31, 31, -31, 31)
    got = engine.sao(rec, geom, 6, params)
    assert_planes_equal(geom, got, c_oracle.sao_batch(rec, geom, 6, params))


def test_sanity_bin_sao_params(engine, c_oracle, sanity_batch):
    """BASELINE config 1: sanity.bin's real per-CTB SAO syntax (90 CTBs) on a synthetic
    352x288 8-bit picture."""
    _, z = sanity_batch
    params = z["sao"].view(SAO_CTB).reshape(z["sao"].shape[:3]) if z["sao"].dtype != SAO_CTB else z["sao"]
    geom = PicGeom(352, 288, params.shape[0], 8, 8)
    rng = np.random.default_rng(26501)
    rec = np.zeros(geom.total_elems(), np.uint8)
    for p in range(geom.n_pics):
        synth.sao_picture(352, 288, 8, rng, geom, rec, p)
    assert (params["type"] != 0).any()
    got = engine.sao(rec, geom, 6, params)
    assert_planes_equal(geom, got, c_oracle.sao_batch(rec, geom, 6, params))


def test_idempotent_when_off(engine):
    geom, rec, params = synth.sao_batch(192, 128, 10, n_pics=1, ctb_log2=6, seed=3)
    params["type"][:] = 0
    got = engine.sao(rec, geom, 6, params)
    assert_planes_equal(geom, got, rec)


def test_bad_arguments_raise(engine):
    geom, rec, params = synth.sao_batch(64, 64, 8, n_pics=1, ctb_log2=6, seed=3)
    with pytest.raises(ValueError):
        engine.sao(rec, geom, 7, params)
    bad = params.copy()
    bad["eo_class"][..., 0] = 4
    with pytest.raises(ValueError):
        engine.sao(rec, geom, 6, bad)
    with pytest.raises(ValueError):
        engine.sao(rec, geom, 6, params[:, :0])
