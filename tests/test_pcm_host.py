"""Host reconstruction of pcm coding units (p265_b200/intra_host.py: `_pcm`, H.265 8.4.4.1) on hand-made parser objects:
samples are shifted up from PcmBitDepth to the picture's bit depth, pcm CUs count as intra neighbours, and a parser
that did not keep the samples (the reference as shipped, cu.py:146-151) is reported, not guessed around.  The
libavcodec-pinned streams with pcm CUs are in tests/test_fuzz_streams.py."""
import types

import numpy as np
import pytest

from p265_b200 import intra_host


def _picture(cu):
    ctu = types.SimpleNamespace(get_leaves=lambda: [cu], slice_addr=0)
    img = types.SimpleNamespace(ctus={0: ctu})
    sps = types.SimpleNamespace(pic_width_in_luma_samples=16, pic_height_in_luma_samples=16, ctb_log2_size_y=4,
                                pic_width_in_ctbs_y=1, pic_height_in_ctbs_y=1, bit_depth_y=10, bit_depth_c=9,
                                strong_intra_smoothing_enabled_flag=0, pcm_sample_bit_depth_luma_minus1=6,
                                pcm_sample_bit_depth_chroma_minus1=4)
    return img, sps


def _cu(**kw):
    return types.SimpleNamespace(x=0, y=0, size=16, log2size=4, pred_mode=intra_host.MODE_INTRA, pcm_flag=1, **kw)


def test_pcm_samples_are_shifted_up_to_the_bit_depth():
    rng = np.random.default_rng(3)
    luma = rng.integers(0, 1 << 7, (16, 16))
    chroma = rng.integers(0, 1 << 5, (2, 8, 8))
    img, sps = _picture(_cu(pcm_sample_luma=luma, pcm_sample_chroma=chroma))
    zero = [np.zeros((16, 16), np.int16), np.zeros((8, 8), np.int16), np.zeros((8, 8), np.int16)]
    y, cb, cr = intra_host.reconstruct_intra_picture(img, sps, None, zero)
    assert y.dtype == np.uint16
    assert np.array_equal(y, luma << 3)                  # 10 - 7
    assert np.array_equal(cb, chroma[0] << 4) and np.array_equal(cr, chroma[1] << 4)   # 9 - 5


def test_a_parser_that_dropped_the_samples_is_reported():
    img, sps = _picture(_cu())
    zero = [np.zeros((16, 16), np.int16), np.zeros((8, 8), np.int16), np.zeros((8, 8), np.int16)]
    with pytest.raises(ValueError, match="pcm_sample_luma"):
        intra_host.reconstruct_intra_picture(img, sps, None, zero)
