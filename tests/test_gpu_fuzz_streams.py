"""The GPU kernels against libavcodec on the fuzz streams (see tests/fuzz_common.py and the CPU twin
tests/test_fuzz_streams.py): residual kernels, deblocking kernel and SAO kernel through the C-ABI,
host intra prediction in between; Y, Cb, Cr bit-exact before and after the loop filters."""
import pytest

import fuzz_common as fz

pytestmark = pytest.mark.gpu
STREAMS = sorted(fz.manifest().items())


@pytest.mark.parametrize("name,cfg", STREAMS, ids=[n for n, _ in STREAMS])
def test_gpu_chain_equals_libavcodec(name, cfg, engine):
    launches = engine.launch_count
    fz.check_stream(name, cfg, fz.GpuBackend(engine), fz.answers())
    assert engine.launch_count > launches
