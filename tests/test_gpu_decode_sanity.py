"""sanity.bin decoded end to end with the GPU path and compared with libavcodec (the
independent known answer, tests/golden/sanity_ffmpeg.npz):

  reference parser (py3 shim, host) -> packer / parser-side emitter -> residual kernels (GPU) -> host intra
  prediction + reconstruction -> deblocking kernel (GPU) -> SAO kernel (GPU)

Y, Cb and Cr of all three pictures, before the loop filters and after them, bit for bit.
The CPU twin (oracle instead of GPU) is tests/test_decode_sanity.py."""
import numpy as np
import pytest

from p265_b200 import deblock_api, emit, intra_host, loop_filter_api, packer, sao_api

pytestmark = pytest.mark.gpu
COMPS = ("y", "cb", "cr")


def test_gpu_decode_of_sanity_bin_equals_libavcodec(engine, parsed_sanity, ffmpeg_sanity):
    imgs, sps, pps = parsed_sanity
    assert len(imgs) == 3
    launches = engine.launch_count
    for p, img in enumerate(imgs):
        batch = packer.pack_pictures([img], sps)
        res = engine.residual(batch)
        # the same picture as the parser emitted it (packed coefficient stream, emit.hook_parser)
        assert np.array_equal(engine.residual(emit.take(img, sps)), res)
        planes = [batch.geom.plane_view(res, 0, c) for c in range(3)]
        rec = intra_host.reconstruct_intra_picture(img, sps, pps, planes)
        for c, n in enumerate(COMPS):
            assert np.array_equal(rec[c], ffmpeg_sanity["rec%d_%s" % (p, n)]), ("rec", p, n)
        dbk = deblock_api.filter_picture(rec, img, sps, pps)
        out = sao_api.filter_picture(dbk, img, sps, pps)
        fused = loop_filter_api.filter_picture(rec, img, sps, pps)      # both filters, one round trip
        for c, n in enumerate(COMPS):
            assert np.array_equal(out[c], ffmpeg_sanity["out%d_%s" % (p, n)]), ("out", p, n)
            assert np.array_equal(fused[c], out[c]), ("fused", p, n)
    assert engine.launch_count > launches
