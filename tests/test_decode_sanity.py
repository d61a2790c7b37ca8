"""End-to-end pin of the oracle against an INDEPENDENT conformant decoder.

The reference's tests hold no output vectors for the residual / SAO path,
and its own transform is not the standard's (G3).  libavcodec's decode of the reference's
sanity.bin (tests/golden/sanity_ffmpeg.npz, make_ffmpeg_fixture.py) is the known answer:

  reference parser (shim) -> packed TB list -> ORACLE residual -> host intra prediction +
  reconstruction  == libavcodec with the loop filters skipped        (Y, Cb, Cr, 3 pictures)
  ... -> ORACLE deblocking -> ORACLE SAO (parameters from the parsed sao() syntax)
                          == libavcodec's final pictures

This pins, bit for bit: dequantisation, the inverse DCT's orientation / stage order /
intermediate clip / final shift, the 4x4 DST, transform-skip (41 TBs), the SAO filter and
`SaoOffsetVal` derivation, the deblocking oracle and the host edge map.  Not exercised by
this stream: scaling lists, cu_transquant_bypass, 10-bit, multiple slices / tiles.
The GPU twin of this test is tests/test_gpu_decode_sanity.py.
"""
import numpy as np

from oracle import spec_oracle as so
from p265_b200 import deblock_api, intra_host, packer
from p265_b200.picture import PicGeom

COMPS = ("y", "cb", "cr")


def _planes(geom, buf, pic=0):
    return [geom.plane_view(buf, pic, c) for c in range(3)]


def test_oracle_residual_plus_host_intra_equals_libavcodec_reconstruction(parsed_sanity, ffmpeg_sanity,
                                                                             c_oracle):
    imgs, sps, pps = parsed_sanity
    assert len(imgs) == 3
    n_ts = 0
    for p, img in enumerate(imgs):
        batch = packer.pack_pictures([img], sps)
        n_ts += int(((batch.tus["flags"] & 2) != 0).sum())
        res = c_oracle.residual_batch(batch)
        rec = intra_host.reconstruct_intra_picture(img, sps, pps, _planes(batch.geom, res))
        for c, n in enumerate(COMPS):
            assert np.array_equal(rec[c], ffmpeg_sanity["rec%d_%s" % (p, n)]), (p, n)
    assert n_ts == 41


def test_numpy_oracle_blocks_equal_libavcodec_residual_on_one_picture(parsed_sanity, ffmpeg_sanity):
    """Same through the numpy restatement (spec_oracle.py), TB by TB, on picture 0."""
    imgs, sps, pps = parsed_sanity
    img = imgs[0]
    geom = packer.geom_from_sps(sps)
    res = [np.zeros(geom.plane_shape(c), np.int16) for c in range(3)]
    for c, x, y, l2, qp, fl, lv in packer.iter_tbs(img, sps):
        n = 1 << l2
        r = so.residual_block_yx(np.asarray(lv), qp, 8, l2, dst=bool(fl & 1), ts=bool(fl & 2))
        res[c][y:y + n, x:x + n] = so.sat16(r)
    rec = intra_host.reconstruct_intra_picture(img, sps, pps, res)
    for c, n in enumerate(COMPS):
        assert np.array_equal(rec[c], ffmpeg_sanity["rec0_%s" % n])


def test_oracle_deblocking_and_sao_equal_libavcodec_output(parsed_sanity, ffmpeg_sanity, c_oracle):
    imgs, sps, pps = parsed_sanity
    geom = PicGeom(352, 288, 1, 8, 8)
    for p, img in enumerate(imgs):
        rec = [ffmpeg_sanity["rec%d_%s" % (p, n)] for n in COMPS]
        blk, ctb = deblock_api.edge_map_from_picture(img, sps, pps)
        dbk = so.deblock_picture(rec, blk, ctb, int(sps.ctb_log2_size_y), 8, 8)
        assert any(not np.array_equal(a, b) for a, b in zip(dbk, rec))
        # the plain-C deblocking restatement agrees with the numpy one
        buf = np.zeros(geom.total_elems(), np.uint8)
        for c in range(3):
            geom.plane_view(buf, 0, c)[:] = rec[c]
        dbk_c = c_oracle.deblock_batch(buf, geom, int(sps.ctb_log2_size_y), blk, ctb)
        for c in range(3):
            assert np.array_equal(geom.plane_view(dbk_c, 0, c), dbk[c])
        params = packer.sao_params_from_picture(img, sps)
        assert (params["type"] != 0).any()
        out = c_oracle.sao_batch(dbk_c, geom, int(sps.ctb_log2_size_y), params)
        for c, n in enumerate(COMPS):
            assert np.array_equal(geom.plane_view(out, 0, c), ffmpeg_sanity["out%d_%s" % (p, n)]), (p, n)
        # numpy SAO restatement on the luma plane as well
        y = so.sao_filter_plane(dbk[0], 8, 64, params["type"][:, :, 0], params["band_pos"][:, :, 0],
                                params["eo_class"][:, :, 0], params["offset_val"][:, :, 0, :])
        assert np.array_equal(y, ffmpeg_sanity["out%d_y" % p])
