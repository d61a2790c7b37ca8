"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/libspec_oracle.so (plain-C
restatement, see spec_oracle.c).  Used by tests/ (checker at full 4K sizes) and by
bench.py's cpu_baseline / --impl reference leg ("port").  Never imported by p265_b200."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libspec_oracle.so")


class Geom(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("n_pics", C.c_int32),
                ("bit_depth_y", C.c_int32), ("bit_depth_c", C.c_int32),
                ("stride_y", C.c_int32), ("stride_c", C.c_int32), ("rsvd", C.c_int32),
                ("plane_off", C.c_int64 * 3), ("pic_stride", C.c_int64)]


def geom_struct(g) -> Geom:
    return Geom(g.width, g.height, g.n_pics, g.bit_depth_y, g.bit_depth_c, g.stride_y,
                g.stride_c, 0, (C.c_int64 * 3)(*g.plane_off), g.pic_stride)


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "spec_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "-B", "libspec_oracle.so"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def dct32() -> np.ndarray:
    out = np.zeros((32, 32), np.int32)
    lib().oracle_dct32(_p(out))
    return out


def residual_batch(batch, zero_fill=True) -> np.ndarray:
    g = batch.geom
    out = np.zeros(g.total_elems(), np.int16)
    tus = np.ascontiguousarray(batch.tus)
    co = np.ascontiguousarray(batch.coeffs, dtype=np.int16)
    sf = None if batch.scaling_factor is None else np.ascontiguousarray(batch.scaling_factor, np.uint8)
    gs = geom_struct(g)
    lib().oracle_residual_batch(_p(tus), C.c_int32(len(tus)), _p(co), _p(sf), C.byref(gs),
                                _p(out), C.c_int(1 if zero_fill else 0))
    return out


def dequant_batch(batch) -> np.ndarray:
    g = batch.geom
    out = np.zeros(batch.coeffs.size, np.int16)
    tus = np.ascontiguousarray(batch.tus)
    co = np.ascontiguousarray(batch.coeffs, dtype=np.int16)
    sf = None if batch.scaling_factor is None else np.ascontiguousarray(batch.scaling_factor, np.uint8)
    lib().oracle_dequant_batch(_p(tus), C.c_int32(len(tus)), _p(co), _p(sf),
                               C.c_int(g.bit_depth_y), C.c_int(g.bit_depth_c), _p(out))
    return out


def ref_literal_batch(tus, scaled) -> np.ndarray:
    out = np.zeros(scaled.size, np.int32)
    tus = np.ascontiguousarray(tus)
    sc = np.ascontiguousarray(scaled, dtype=np.int16)
    lib().oracle_ref_literal_batch(_p(tus), C.c_int32(len(tus)), _p(sc), _p(out))
    return out


def sao_batch(rec, geom, ctb_log2, params, no_filter=None) -> np.ndarray:
    rec = np.ascontiguousarray(rec)
    out = rec.copy()
    params = np.ascontiguousarray(params)
    nf = None if no_filter is None else np.ascontiguousarray(no_filter, np.uint8)
    gs = geom_struct(geom)
    lib().oracle_sao_batch(_p(rec), _p(out), C.byref(gs), C.c_int(ctb_log2), _p(params), _p(nf))
    return out


def deblock_batch(planes, geom, ctb_log2, blk, ctb) -> np.ndarray:
    """8.7.2 on a batch (returns a filtered copy of the flat plane buffer)."""
    out = np.ascontiguousarray(planes).copy()
    blk = np.ascontiguousarray(blk, dtype=np.uint16)
    ctb = np.ascontiguousarray(ctb)
    gs = geom_struct(geom)
    if lib().oracle_deblock_batch(_p(out), C.byref(gs), C.c_int(ctb_log2), _p(blk), _p(ctb)) != 0:
        raise ValueError("oracle_deblock_batch: bad geometry")
    return out
