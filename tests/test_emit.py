"""Parser-side emission (p265_b200/emit.py): the packed stream written CU by
CU while the reference's parser runs equals the round-1 packer's walk over the finished tree."""
import time

import numpy as np

from oracle import spec_oracle
from p265_b200 import emit, packer
from p265_b200.picture import TU_LEVELS8


def _tb_set(tus, arena):
    out = {}
    for t in tus:
        n2 = 1 << (2 * int(t["log2n"]))
        o = int(t["coeff_off"]) * 16
        key = (int(t["pic"]), int(t["c_idx"]), int(t["x"]), int(t["y"]), int(t["log2n"]), int(t["qp"]),
               int(t["flags"]) & ~TU_LEVELS8)
        assert key not in out
        out[key] = arena[o:o + n2].tobytes()
    return out


def test_emitted_stream_equals_the_tree_walk(parsed_sanity, c_oracle):
    imgs, sps, pps = parsed_sanity
    assert len(imgs) == 3
    t_walk = t_take = 0.0
    for img in imgs:
        t0 = time.perf_counter()
        dense = packer.pack_pictures([img], sps)
        t_walk += time.perf_counter() - t0
        t0 = time.perf_counter()
        packed = emit.take(img, sps)
        t_take += time.perf_counter() - t0
        assert packed is not None and len(packed.tus) == len(dense.tus) > 1000
        assert packed.stream.nbytes < dense.coeffs.nbytes // 4
        tus, arena = spec_oracle.unpack_stream(packed.tus, packed.stream)      # the oracle's record reader
        assert _tb_set(tus, arena) == _tb_set(dense.tus, dense.coeffs)
        # same bins, same order rule; and the product's own vectorised unpacker agrees with the oracle's
        assert packed.bin_counts() == dense.bin_counts()
        assert np.array_equal(packed.unpacked().coeffs, arena)
        # residual of the emitted picture == residual of the walked picture (oracle arithmetic)
        from p265_b200.picture import ResidualBatch
        a = c_oracle.residual_batch(ResidualBatch(dense.geom, tus, arena))
        assert np.array_equal(a, c_oracle.residual_batch(dense))
    # the end-of-picture work left is a descriptor sort: far below the tree walk it replaces
    assert t_take < t_walk


def test_sink_rejects_bad_blocks():
    s = emit.PictureSink()
    import pytest
    with pytest.raises(ValueError):
        s.add_tb(0, 0, 0, 2, 30, 0, np.zeros((4, 3)))
    with pytest.raises(ValueError):
        s.add_tb(0, 0, 0, 2, 30, 0, np.full((4, 4), 40000))
    s.add_tb(0, 0, 0, 2, 30, 0, np.zeros((4, 4), np.int64))
    assert len(s.stream) == 4 and s.recs[0][5] & TU_LEVELS8


def _fuzz_streams():
    import fuzz_common as fz
    return sorted(fz.manifest().items())


import pytest  # noqa: E402


@pytest.mark.parametrize("name,cfg", _fuzz_streams(), ids=[n for n, _ in _fuzz_streams()])
def test_emission_on_the_fuzz_streams(name, cfg, c_oracle):
    """The same equality on the libavcodec-pinned fuzz streams: bypass CUs, several slices, tiles (CUs arrive in
    tile-scan order), PPS scaling lists, 9 .. 12 bit, and pcm CUs (a leaf CU without any TB)."""
    import sys
    import fuzz_common as fz
    from p265_b200 import scaling_list
    from oracle import refshim
    if not refshim.shim_available() and not refshim.reference_available():
        pytest.skip("baseline/_ref shim not present")
    refshim.load()
    emit.hook_parser(sys.modules["cu"])
    imgs, sps, pps = fz.parse(name, cfg)
    for img in imgs:
        sf = scaling_list.active_table(sps, pps)
        dense = packer.pack_pictures([img], sps, sf)
        packed = emit.take(img, sps, sf)
        assert packed is not None and len(packed.tus) == len(dense.tus)
        tus, arena = spec_oracle.unpack_stream(packed.tus, packed.stream)
        assert _tb_set(tus, arena) == _tb_set(dense.tus, dense.coeffs)
        assert packed.bin_counts() == dense.bin_counts()
        assert np.array_equal(c_oracle.residual_batch(packed.unpacked()), c_oracle.residual_batch(dense))
