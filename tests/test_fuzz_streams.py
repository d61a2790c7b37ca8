"""The ORACLE against libavcodec on the fuzz streams (tests/fuzz_common.py): 10-bit, default and
explicit (PPS) scaling lists, cu_transquant_bypass, chroma QP offsets, deblocking offsets / disable,
several slices per picture with slice_loop_filter_across_slices_enabled_flag = 0, CTB sizes 16 / 32 /
64, coefficient magnitudes large enough to hit every clip of 8.6.3 / 8.6.4.  Bit-exact for Y, Cb and
Cr before and after the loop filters."""
import pytest

import fuzz_common as fz

STREAMS = sorted(fz.manifest().items())


@pytest.mark.parametrize("name,cfg", STREAMS, ids=[n for n, _ in STREAMS])
def test_oracle_chain_equals_libavcodec(name, cfg, c_oracle):
    diffs = fz.check_stream(name, cfg, fz.OracleBackend(c_oracle), fz.answers())
    total = [sum(d[c] for d in diffs) for c in range(3)]
    restore = cfg["bypass"] or bool(cfg.get("pcm") and cfg["pcm"]["lf_disabled"])
    if cfg.get("lav_chroma_tc_dev"):
        assert total[1] > 0 and total[2] > 0      # the deviation is there (chroma SAO is off in that stream; luma is
        #                                           compared exactly inside check_stream)
    if cfg["slices"] == 1 and not restore:
        assert total == [0, 0, 0]            # no rule on which the standard and libavcodec differ is in play
    if restore:
        assert total[0] == 0                 # libavcodec's restore deviation is chroma-only
    if cfg.get("tiles"):
        assert total == [0, 0, 0]            # tiles: one slice each, every slice flag 1 -> only the tile rule acts


def test_the_streams_cover_what_sanity_bin_does_not():
    m = fz.manifest()
    assert {c["bit_depth"] for c in m.values()} == {8, 9, 10, 12}
    assert {c["ctb_log2"] for c in m.values()} == {4, 5, 6}
    assert {c["scaling_lists"] for c in m.values()} == {"off", "default", "pps", "sps", "sps+pps"}
    assert any(c["bypass"] for c in m.values()) and any(c["slices"] > 1 for c in m.values())
    assert any(c["dbk_disable"] for c in m.values()) and any(c["tc_offset_div2"] for c in m.values())
    assert any(c["cb_qp_offset"] for c in m.values()) and sum(c["tbs"] for c in m.values()) > 3000
    pcm = [c["pcm"] for c in m.values() if c.get("pcm")]
    assert {c["lf_disabled"] for c in pcm} == {0, 1} and any(c["bits_y"] < 8 for c in pcm)
    tiles = [c for c in m.values() if c.get("tiles")]
    assert any(not c["lf_across_tiles"] for c in tiles) and any(c["lf_across_tiles"] for c in tiles)
