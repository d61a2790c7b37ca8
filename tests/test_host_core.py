"""CPU check of the device code's indexing and arithmetic: tests/host/host_core.cpp
compiles p265_b200/csrc/residual_core.cuh for the host and runs the SAME per-lane
phase functions the sm_100a kernel runs, lane after lane.  Bit-exact vs the C oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import REPO, small_cfg
from p265_b200 import synth


@pytest.fixture(scope="module")
def host_core(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hostcore") / "libhost_core.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(REPO, "include"),
                           "-o", out, os.path.join(REPO, "tests", "host", "host_core.cpp")])
    return C.CDLL(out)


def run(lib, c_oracle, batch, replicated):
    out = np.zeros(batch.geom.total_elems(), np.int16)
    gs = c_oracle.geom_struct(batch.geom)
    sf = batch.scaling_factor
    lib.host_residual_batch(batch.tus.ctypes.data_as(C.c_void_p), (C.c_int32 * 4)(*batch.bin_counts()),
                            batch.coeffs.ctypes.data_as(C.c_void_p),
                            sf.ctypes.data_as(C.c_void_p) if sf is not None else None,
                            C.c_int(1 if replicated else 0), C.byref(gs), out.ctypes.data_as(C.c_void_p))
    return out


def test_slot_permutation(host_core):
    for n in (4, 8, 16, 32):
        idx = [host_core.host_slot_index(n, s, h) for s in range(n // 2) for h in (0, 1)]
        assert sorted(idx) == list(range(n))
    assert [host_core.host_slot_index(32, s, 0) for s in range(4)] == [0, 8, 4, 20]
    assert [host_core.host_slot_index(32, s, 1) for s in range(4)] == [16, 24, 12, 28]


@pytest.mark.parametrize("name", ["1080p8", "4k10"])
@pytest.mark.parametrize("stress", [False, True])
def test_host_emulation_matches_oracle(host_core, c_oracle, name, stress):
    batch = synth.residual_batch(small_cfg(name, 256, 192), n_pics=2, stress=stress)
    ref = c_oracle.residual_batch(batch, zero_fill=True)
    for replicated in ((False, True) if batch.scaling_factor is not None else (False,)):
        assert np.array_equal(run(host_core, c_oracle, batch, replicated), ref)


def test_host_emulation_sanity_bin(host_core, c_oracle, sanity_batch):
    batch, _ = sanity_batch
    assert np.array_equal(run(host_core, c_oracle, batch, False), c_oracle.residual_batch(batch))
