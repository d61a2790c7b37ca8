"""ctypes binding of libp265b200.so (include/p265_b200.h).

There is deliberately no fallback: if the CUDA library is missing or no B200 is
visible, importing works (so CPU-only tooling can inspect the ABI) but every
compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# P265_LIB: another build of the same library (A/B tuning runs, tools/kbench.py); never a fallback
LIB_PATH = os.environ.get("P265_LIB") or os.path.join(HERE, "libp265b200.so")

P265_OK, P265_EINVAL, P265_ECUDA, P265_ENOMEM = 0, -1, -2, -3
RES_ZERO_FILL = 1
RES_SF_REPLICATED = 2
RES_DENSE_ARENA = 4
RES_ZERO_EXTENTS = 8

#: every symbol include/p265_b200.h declares (tests check the .so exports them all)
SYMBOLS = (
    "p265_abi_version", "p265_last_error", "p265_device_count", "p265_ctx_create",
    "p265_ctx_destroy", "p265_sync", "p265_ctx_set_async", "p265_sm_count", "p265_launch_count",
    "p265_residual_batch", "p265_residual_batch_dev", "p265_dequant_batch",
    "p265_ref_literal_batch", "p265_idct_1d", "p265_sao_batch", "p265_sao_batch_dev",
    "p265_reconstruct_batch", "p265_reconstruct_batch_dev", "p265_deblock_batch",
    "p265_deblock_batch_dev", "p265_int_peak",
    "p265_residual_batch_packed", "p265_residual_batch_packed_dev", "p265_loop_filter_batch",
    "p265_pcie_probe", "p265_ctx_set_trace", "p265_trace_read",
)
ABI_VERSION = 3


class Geom(C.Structure):
    """p265_pic_geom"""
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("n_pics", C.c_int32),
                ("bit_depth_y", C.c_int32), ("bit_depth_c", C.c_int32),
                ("stride_y", C.c_int32), ("stride_c", C.c_int32), ("rsvd", C.c_int32),
                ("plane_off", C.c_int64 * 3), ("pic_stride", C.c_int64)]


def geom_struct(g) -> Geom:
    return Geom(g.width, g.height, g.n_pics, g.bit_depth_y, g.bit_depth_c, g.stride_y,
                g.stride_c, 0, (C.c_int64 * 3)(*g.plane_off), g.pic_stride)


_lib = None


def load():
    """Load the shared library or raise RuntimeError (never falls back)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "p265_b200: %s is missing -- build it with `python -m p265_b200.build` "
            "(needs nvcc); there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64p = C.c_void_p, C.c_int32, C.POINTER(C.c_int32)
    lib.p265_abi_version.restype = C.c_int
    lib.p265_last_error.restype = C.c_char_p
    lib.p265_device_count.restype = C.c_int
    lib.p265_ctx_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    lib.p265_ctx_destroy.argtypes = [vp]
    lib.p265_sync.argtypes = [vp]
    lib.p265_ctx_set_async.argtypes = [vp, C.c_int]
    lib.p265_sm_count.argtypes = [vp]
    lib.p265_launch_count.argtypes = [vp]
    lib.p265_launch_count.restype = C.c_uint64
    lib.p265_residual_batch.argtypes = [vp, vp, i64p, vp, C.c_size_t, vp, C.POINTER(Geom), vp, C.c_int]
    lib.p265_residual_batch_dev.argtypes = [vp, vp, i64p, vp, vp, C.POINTER(Geom), vp, C.c_int]
    lib.p265_dequant_batch.argtypes = [vp, vp, i32, vp, C.c_size_t, vp, C.c_int, C.c_int, vp]
    lib.p265_ref_literal_batch.argtypes = [vp, vp, i32, vp, C.c_size_t, vp]
    lib.p265_idct_1d.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp]
    lib.p265_sao_batch.argtypes = [vp, vp, vp, C.POINTER(Geom), C.c_int, vp, vp]
    lib.p265_sao_batch_dev.argtypes = [vp, vp, vp, C.POINTER(Geom), C.c_int, vp, vp]
    lib.p265_reconstruct_batch.argtypes = [vp, vp, vp, vp, C.POINTER(Geom)]
    lib.p265_reconstruct_batch_dev.argtypes = [vp, vp, vp, vp, C.POINTER(Geom)]
    lib.p265_deblock_batch.argtypes = [vp, vp, C.POINTER(Geom), C.c_int, vp, vp]
    lib.p265_deblock_batch_dev.argtypes = [vp, vp, C.POINTER(Geom), C.c_int, vp, vp]
    lib.p265_int_peak.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.p265_residual_batch_packed.argtypes = [vp, vp, i64p, vp, C.c_size_t, vp, C.POINTER(Geom), vp, C.c_int]
    lib.p265_residual_batch_packed_dev.argtypes = [vp, vp, i64p, vp, vp, C.POINTER(Geom), vp, vp, vp, C.c_int]
    lib.p265_loop_filter_batch.argtypes = [vp, vp, C.POINTER(Geom), C.c_int, vp, vp, vp, vp]
    lib.p265_ctx_set_trace.argtypes = [vp, C.c_int]
    lib.p265_trace_read.argtypes = [vp, C.POINTER(C.c_double), C.c_int]
    lib.p265_pcie_probe.argtypes = [vp, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    _lib = lib
    return lib


def check(rc: int) -> None:
    """0 -> ok; EINVAL -> ValueError; anything else -> RuntimeError (the reference's
    error style: ValueError for bad arguments, scaling.py:21)."""
    if rc == P265_OK:
        return
    msg = load().p265_last_error().decode("utf-8", "replace")
    if rc == P265_EINVAL:
        raise ValueError(msg)
    raise RuntimeError(msg)


def ptr(a):
    """Host pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("array must be C-contiguous")
    return a.ctypes.data_as(C.c_void_p)


def bins(counts):
    return (C.c_int32 * 4)(*[int(c) for c in counts])


def as_array(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)
