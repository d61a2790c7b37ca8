"""Packed coefficient stream (include/p265_b200.h): the product's vectorised packer
(picture.pack_coefficients) against the oracle's independent record reader
(oracle/spec_oracle.py: unpack_stream) and a hand-written known answer.  CPU only."""
import numpy as np

from conftest import small_cfg
from oracle import spec_oracle
from p265_b200 import synth
from p265_b200.picture import TU_DESC, TU_LEVELS8, PicGeom, ResidualBatch, pack_coefficients


def test_known_answer_record():
    # one 4x4 TB: levels 3 at (x=0,y=0), -2 at (x=1,y=0), 300 at (x=3,y=2)  -> wide (int16) record
    blk = np.zeros((4, 4), np.int16)
    blk[0, 0], blk[0, 1], blk[2, 3] = 3, -2, 300
    tus = np.zeros(1, TU_DESC)
    tus["log2n"] = 2
    t, s = pack_coefficients(tus, blk.reshape(-1))
    assert not (t["flags"][0] & TU_LEVELS8)
    # bitmap bits 0, 1 and 2*4+3 = 11 -> bytes 0x03, 0x08; then 3, -2, 300 as little-endian int16; padded to 4
    assert s.tolist() == [0x03, 0x08, 3, 0, 0xFE, 0xFF, 0x2C, 0x01]
    assert t["rsvd"][0] == 3                                  # number of levels travels in the descriptor
    # the same block without the 300: int8 levels
    blk[2, 3] = -128
    t, s = pack_coefficients(tus, blk.reshape(-1))
    assert t["flags"][0] & TU_LEVELS8
    assert s.tolist() == [0x03, 0x08, 3, 0xFE, 0x80, 0, 0, 0]


def test_round_trip_matches_oracle_reader():
    for name, stress in (("4k10", False), ("1080p8", False), ("4k10", True)):
        b = synth.residual_batch(small_cfg(name, 192, 128), n_pics=2, stress=stress)
        pb = b.packed()
        assert pb.stream.size % 4 == 0 and pb.stream.size < b.coeffs.nbytes * (1.1 if stress else 0.3)
        tus, arena = spec_oracle.unpack_stream(pb.tus, pb.stream)
        for t_in, t_out in zip(b.tus, tus):
            n2 = 1 << (2 * int(t_in["log2n"]))
            a = b.coeffs[int(t_in["coeff_off"]) * 16:int(t_in["coeff_off"]) * 16 + n2]
            o = arena[int(t_out["coeff_off"]) * 16:int(t_out["coeff_off"]) * 16 + n2]
            assert np.array_equal(a, o)
        for f in ("x", "y", "log2n", "c_idx", "qp", "pic"):
            assert np.array_equal(tus[f], b.tus[f])
        assert np.array_equal(tus["flags"], b.tus["flags"])


def test_empty_batch_packs():
    b = ResidualBatch(PicGeom(64, 64), np.zeros(0, TU_DESC), np.zeros(0, np.int16))
    pb = b.packed()
    assert pb.stream.size == 0 and len(pb.tus) == 0 and pb.bin_counts() == (0, 0, 0, 0)
