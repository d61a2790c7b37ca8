"""Engine: one C-ABI context (one GPU, one stream) behind a small Python surface.

Host-array methods copy in, run the sm_100a kernels and copy out (what the drop-in
modules and the end-to-end benchmark use); `*_dev` methods take raw device pointers of
buffers that are already resident (what the roofline benchmark times).  torch is not
needed here; bench.py uses it only to own device memory and the stream.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _lib
from .picture import DBK_CTB, SAO_CTB, SF_BYTES, TU_DESC, PackedResidualBatch, PicGeom, ResidualBatch


class Engine:
    def __init__(self, device: int = 0, stream: int | None = None):
        lib = _lib.load()
        if lib.p265_abi_version() != _lib.ABI_VERSION:
            raise RuntimeError("libp265b200.so ABI mismatch")
        handle = C.c_void_p()
        _lib.check(lib.p265_ctx_create(int(device), C.c_void_p(stream) if stream else None,
                                       C.byref(handle)))
        self._lib, self._ctx, self.device = lib, handle, int(device)
        self._async, self._inflight = False, []

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.p265_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _lib.check(self._lib.p265_sync(self._ctx))
        self._inflight.clear()

    def _hold(self, *arrays):
        """Asynchronous mode: keep the host arrays of queued copies alive until sync()."""
        if self._async:
            self._inflight.extend(a for a in arrays if a is not None)

    def set_async(self, enable: bool = True):
        """Host-array methods return once their copies and kernels are queued; call sync()
        before reading the outputs (keep inputs alive, ideally page-locked, until then)."""
        _lib.check(self._lib.p265_ctx_set_async(self._ctx, 1 if enable else 0))
        self._async = bool(enable)

    def set_trace(self, enable: bool = True):
        """Record a timeline of the host calls on this context (see trace())."""
        _lib.check(self._lib.p265_ctx_set_trace(self._ctx, 1 if enable else 0))

    def trace(self, max_marks: int = 4096):
        """[(kind, phase, ms)] since the last read: kind 1 residual / 2 SAO; phase 0 call start, 1 inputs
        copied, 2 kernels done, 3 outputs back; ms on the device clock since tracing was first enabled."""
        buf = (C.c_double * (3 * max_marks))()
        n = self._lib.p265_trace_read(self._ctx, buf, max_marks)
        if n < 0:
            _lib.check(n)
        return [(int(buf[3 * i]), int(buf[3 * i + 1]), float(buf[3 * i + 2])) for i in range(n)]

    @property
    def sm_count(self) -> int:
        return int(self._lib.p265_sm_count(self._ctx))

    @property
    def launch_count(self) -> int:
        return int(self._lib.p265_launch_count(self._ctx))

    # ------------------------------------------------------------------- residual
    def residual(self, batch, out: np.ndarray | None = None) -> np.ndarray:
        """Residual planes (flat int16 buffer laid out by `batch.geom`) of a batch: a
        `ResidualBatch` (dense coefficient arena) or a `PackedResidualBatch` (packed stream)."""
        g = batch.geom
        if out is None:
            out = np.empty(g.total_elems(), dtype=np.int16)
        elif out.dtype != np.int16 or out.size < g.total_elems():
            raise ValueError("out must be int16 with at least geom.total_elems() elements")
        if isinstance(batch, PackedResidualBatch):
            return self._residual_packed(batch, out)
        tus = _lib.as_array(batch.tus, TU_DESC)
        co = _lib.as_array(batch.coeffs, np.int16)
        sf = None
        if batch.scaling_factor is not None:
            sf = _lib.as_array(batch.scaling_factor, np.uint8)
            if sf.size != SF_BYTES:
                raise ValueError("scaling_factor table must have %d bytes" % SF_BYTES)
        gs = _lib.geom_struct(g)
        flags = 0 if batch.covers_all else _lib.RES_ZERO_FILL
        if sf is not None and batch.sf_replicated:
            flags |= _lib.RES_SF_REPLICATED
        _lib.check(self._lib.p265_residual_batch(
            self._ctx, _lib.ptr(tus), _lib.bins(batch.bin_counts()), _lib.ptr(co), co.size,
            _lib.ptr(sf), C.byref(gs), _lib.ptr(out), flags))
        self._hold(tus, co, sf, out)
        return out

    def _residual_packed(self, batch: PackedResidualBatch, out: np.ndarray) -> np.ndarray:
        tus = _lib.as_array(batch.tus, TU_DESC)
        st = _lib.as_array(batch.stream, np.uint8)
        sf = None
        if batch.scaling_factor is not None:
            sf = _lib.as_array(batch.scaling_factor, np.uint8)
            if sf.size != SF_BYTES:
                raise ValueError("scaling_factor table must have %d bytes" % SF_BYTES)
        gs = _lib.geom_struct(batch.geom)
        flags = _lib.RES_SF_REPLICATED if (sf is not None and batch.sf_replicated) else 0
        _lib.check(self._lib.p265_residual_batch_packed(
            self._ctx, _lib.ptr(tus), _lib.bins(batch.bin_counts()), _lib.ptr(st), st.size,
            _lib.ptr(sf), C.byref(gs), _lib.ptr(out), flags))
        self._hold(tus, st, sf, out)
        return out

    def residual_packed_dev(self, d_tus: int, bin_counts, d_stream: int, d_sf: int | None, geom: PicGeom,
                            d_arena: int, d_tus_out: int, d_out: int, zero_fill: bool = False,
                            sf_replicated: bool = False):
        """Device-resident packed batch: unpack_kernel into `d_arena` / `d_tus_out`, then the residual kernels."""
        gs = _lib.geom_struct(geom)
        flags = (_lib.RES_ZERO_FILL if zero_fill else 0) | (_lib.RES_SF_REPLICATED if (d_sf and sf_replicated) else 0)
        _lib.check(self._lib.p265_residual_batch_packed_dev(
            self._ctx, C.c_void_p(d_tus), _lib.bins(bin_counts), C.c_void_p(d_stream),
            C.c_void_p(d_sf) if d_sf else None, C.byref(gs), C.c_void_p(d_arena), C.c_void_p(d_tus_out),
            C.c_void_p(d_out), flags))

    def residual_dev(self, d_tus: int, bin_counts, d_coeffs: int, d_sf: int | None, geom: PicGeom,
                     d_out: int, zero_fill: bool = False, sf_replicated: bool = False, dense_arena: bool = False,
                     zero_extents: bool = False):
        """Device-resident batch.  `dense_arena` = ResidualBatch.dense_small_bins() of the batch the
        device buffers were filled from (P265_RES_DENSE_ARENA: asserted by the caller here);
        `zero_extents` = its descriptors carry zero-extent codes (P265_RES_ZERO_EXTENTS)."""
        gs = _lib.geom_struct(geom)
        flags = (_lib.RES_ZERO_FILL if zero_fill else 0) | \
            (_lib.RES_SF_REPLICATED if (d_sf and sf_replicated) else 0) | \
            (_lib.RES_DENSE_ARENA if dense_arena else 0) | (_lib.RES_ZERO_EXTENTS if zero_extents else 0)
        _lib.check(self._lib.p265_residual_batch_dev(
            self._ctx, C.c_void_p(d_tus), _lib.bins(bin_counts), C.c_void_p(d_coeffs),
            C.c_void_p(d_sf) if d_sf else None, C.byref(gs), C.c_void_p(d_out), flags))

    def dequant(self, batch: ResidualBatch) -> np.ndarray:
        """scaling.inverse_scaling for every TB: d[] in arena layout ([y][x] per TB)."""
        tus = _lib.as_array(batch.tus, TU_DESC)
        co = _lib.as_array(batch.coeffs, np.int16)
        sf = None if batch.scaling_factor is None else _lib.as_array(batch.scaling_factor, np.uint8)
        out = np.zeros(co.size, dtype=np.int16)
        _lib.check(self._lib.p265_dequant_batch(
            self._ctx, _lib.ptr(tus), len(tus), _lib.ptr(co), co.size, _lib.ptr(sf),
            batch.geom.bit_depth_y, batch.geom.bit_depth_c, _lib.ptr(out)))
        return out

    def ref_literal(self, tus: np.ndarray, scaled: np.ndarray) -> np.ndarray:
        """transform.py:89-109 as written (parity tests only): int32 arena, [x][y] per TB."""
        tus = _lib.as_array(tus, TU_DESC)
        sc = _lib.as_array(scaled, np.int16)
        out = np.zeros(sc.size, dtype=np.int32)
        _lib.check(self._lib.p265_ref_literal_batch(self._ctx, _lib.ptr(tus), len(tus), _lib.ptr(sc),
                                                    sc.size, _lib.ptr(out)))
        return out

    def idct_1d(self, x, log2size: int, tr_type: int, as_written: bool = False) -> np.ndarray:
        n = 1 << int(log2size)
        xv = _lib.as_array(np.asarray(x).reshape(-1), np.int32)
        if xv.size != n:
            raise ValueError("vector length %d does not match log2size %d" % (xv.size, log2size))
        out = np.zeros(n, dtype=np.int32)
        _lib.check(self._lib.p265_idct_1d(self._ctx, _lib.ptr(xv), int(log2size), int(tr_type),
                                          1 if as_written else 0, _lib.ptr(out)))
        return out

    # ------------------------------------------------------------------------ SAO
    def sao(self, rec: np.ndarray, geom: PicGeom, ctb_log2: int, params: np.ndarray,
            no_filter: np.ndarray | None = None, out: np.ndarray | None = None,
            inplace: bool = False) -> np.ndarray:
        """SAO (8.7.3) of reconstructed planes.  `inplace` (or out is rec): the host buffer `rec` itself
        receives the result -- with page-locked memory only the CTBs SAO modifies are written back.
        Otherwise `out` (a copy of `rec` when not given, so that row padding carries over) gets the
        plane rows."""
        dtype = np.uint8 if max(geom.bit_depth_y, geom.bit_depth_c) <= 8 else np.uint16
        rec = _lib.as_array(rec, dtype).reshape(-1)
        if rec.size < geom.total_elems():
            raise ValueError("rec buffer smaller than the geometry")
        if inplace:
            if not rec.flags.writeable:
                raise ValueError("inplace SAO needs a writable rec buffer")
            out = rec
        elif out is None:
            out = rec.copy()
        par = _lib.as_array(params, SAO_CTB)
        ctb = 1 << ctb_log2
        ctbs = ((geom.width + ctb - 1) // ctb) * ((geom.height + ctb - 1) // ctb) * geom.n_pics
        if par.size != ctbs:
            raise ValueError("expected %d SAO CTB records, got %d" % (ctbs, par.size))
        nf = None
        if no_filter is not None:
            nf = _lib.as_array(no_filter, np.uint8)
            need = ((geom.width + 7) // 8) * ((geom.height + 7) // 8) * geom.n_pics
            if nf.size != need:
                raise ValueError("no_filter needs %d entries" % need)
        gs = _lib.geom_struct(geom)
        _lib.check(self._lib.p265_sao_batch(self._ctx, _lib.ptr(rec), _lib.ptr(out), C.byref(gs),
                                            int(ctb_log2), _lib.ptr(par), _lib.ptr(nf)))
        self._hold(rec, par, nf, out)
        return out

    def sao_dev(self, d_rec: int, d_out: int, geom: PicGeom, ctb_log2: int, d_params: int,
                d_no_filter: int | None = None):
        gs = _lib.geom_struct(geom)
        _lib.check(self._lib.p265_sao_batch_dev(
            self._ctx, C.c_void_p(d_rec), C.c_void_p(d_out), C.byref(gs), int(ctb_log2),
            C.c_void_p(d_params), C.c_void_p(d_no_filter) if d_no_filter else None))

    # ------------------------------------------------------------- reconstruction
    def reconstruct(self, pred: np.ndarray, residual: np.ndarray, geom: PicGeom,
                    out: np.ndarray | None = None) -> np.ndarray:
        """rec = Clip1(pred + residual) over whole planes (reconstruction.py:4-27)."""
        dtype = np.uint8 if max(geom.bit_depth_y, geom.bit_depth_c) <= 8 else np.uint16
        pred = _lib.as_array(pred, dtype).reshape(-1)
        res = _lib.as_array(residual, np.int16).reshape(-1)
        if pred.size < geom.total_elems() or res.size < geom.total_elems():
            raise ValueError("pred / residual buffers smaller than the geometry")
        if out is None:
            out = np.empty_like(pred)
        gs = _lib.geom_struct(geom)
        _lib.check(self._lib.p265_reconstruct_batch(self._ctx, _lib.ptr(pred), _lib.ptr(res), _lib.ptr(out),
                                                    C.byref(gs)))
        self._hold(pred, res, out)
        return out

    def reconstruct_dev(self, d_pred: int, d_residual: int, d_rec: int, geom: PicGeom):
        gs = _lib.geom_struct(geom)
        _lib.check(self._lib.p265_reconstruct_batch_dev(self._ctx, C.c_void_p(d_pred), C.c_void_p(d_residual),
                                                        C.c_void_p(d_rec), C.byref(gs)))

    # ------------------------------------------------------------------ deblocking
    def deblock(self, rec: np.ndarray, geom: PicGeom, ctb_log2: int, blk: np.ndarray,
                ctb: np.ndarray) -> np.ndarray:
        """Deblocked copy (8.7.2) of reconstructed planes; blk / ctb: the edge map and per-CTB
        parameters of `deblock_api.edge_map_from_picture`, stacked per picture."""
        dtype = np.uint8 if max(geom.bit_depth_y, geom.bit_depth_c) <= 8 else np.uint16
        out = np.array(_lib.as_array(rec, dtype).reshape(-1), copy=True)
        if out.size < geom.total_elems():
            raise ValueError("rec buffer smaller than the geometry")
        if geom.width % 8 or geom.height % 8:
            raise ValueError("picture size must be a multiple of 8")
        blk = _lib.as_array(blk, np.uint16).reshape(-1)
        par = _lib.as_array(ctb, DBK_CTB).reshape(-1)
        cs = 1 << ctb_log2
        n_ctb = ((geom.width + cs - 1) // cs) * ((geom.height + cs - 1) // cs) * geom.n_pics
        if blk.size != (geom.width // 8) * (geom.height // 8) * geom.n_pics or par.size != n_ctb:
            raise ValueError("edge map / CTB parameter tables do not match the geometry")
        gs = _lib.geom_struct(geom)
        _lib.check(self._lib.p265_deblock_batch(self._ctx, _lib.ptr(out), C.byref(gs), int(ctb_log2),
                                                _lib.ptr(blk), _lib.ptr(par)))
        self._hold(blk, par, out)
        return out

    def deblock_dev(self, d_planes: int, geom: PicGeom, ctb_log2: int, d_blk: int, d_ctb: int):
        gs = _lib.geom_struct(geom)
        _lib.check(self._lib.p265_deblock_batch_dev(self._ctx, C.c_void_p(d_planes), C.byref(gs), int(ctb_log2),
                                                    C.c_void_p(d_blk), C.c_void_p(d_ctb)))

    # ---------------------------------------------------------- loop filters, fused
    def loop_filter(self, planes: np.ndarray, geom: PicGeom, ctb_log2: int, blk=None, dbk_ctb=None,
                    sao_params=None, no_filter=None) -> np.ndarray:
        """Deblocking (8.7.2, when `blk` / `dbk_ctb` are given) then SAO (8.7.3, when `sao_params` is
        given) in ONE host round trip, in place on `planes` (one H2D + one D2H of the planes instead of
        two of each through deblock() + sao())."""
        dtype = np.uint8 if max(geom.bit_depth_y, geom.bit_depth_c) <= 8 else np.uint16
        if not (isinstance(planes, np.ndarray) and planes.dtype == dtype and planes.flags["C_CONTIGUOUS"]
                and planes.flags.writeable):
            raise ValueError("planes must be a writable C-contiguous %s array (filtered in place)" % np.dtype(dtype).name)
        buf = planes.reshape(-1)
        if buf.size < geom.total_elems():
            raise ValueError("planes buffer smaller than the geometry")
        cs = 1 << ctb_log2
        n_ctb = ((geom.width + cs - 1) // cs) * ((geom.height + cs - 1) // cs) * geom.n_pics
        b = c = par = nf = None
        if (blk is None) != (dbk_ctb is None):
            raise ValueError("edge map and per-CTB deblocking parameters go together")
        if blk is not None:
            b = _lib.as_array(blk, np.uint16).reshape(-1)
            c = _lib.as_array(dbk_ctb, DBK_CTB).reshape(-1)
            if geom.width % 8 or geom.height % 8:
                raise ValueError("picture size must be a multiple of 8")
            if b.size != (geom.width // 8) * (geom.height // 8) * geom.n_pics or c.size != n_ctb:
                raise ValueError("edge map / CTB parameter tables do not match the geometry")
        if sao_params is not None:
            par = _lib.as_array(sao_params, SAO_CTB).reshape(-1)
            if par.size != n_ctb:
                raise ValueError("expected %d SAO CTB records, got %d" % (n_ctb, par.size))
            if no_filter is not None:
                nf = _lib.as_array(no_filter, np.uint8).reshape(-1)
                if nf.size != ((geom.width + 7) // 8) * ((geom.height + 7) // 8) * geom.n_pics:
                    raise ValueError("no_filter map does not match the geometry")
        gs = _lib.geom_struct(geom)
        _lib.check(self._lib.p265_loop_filter_batch(self._ctx, _lib.ptr(buf), C.byref(gs), int(ctb_log2), _lib.ptr(b),
                                                    _lib.ptr(c), _lib.ptr(par), _lib.ptr(nf)))
        self._hold(buf, b, c, par, nf)
        return planes

    # ---------------------------------------------------------------- measurement
    def pcie_probe(self, n_bytes: int = 256 << 20, reps: int = 8, h2d: bool = True, d2h: bool = True,
                   n_buffers: int = 1):
        """(H2D, D2H) bytes per second of plain page-locked copies of `n_bytes`, both directions at once
        when both are requested (None for a direction that was not), cycling through `n_buffers` distinct
        host buffers per direction (1 = one cache-resident buffer, the usual copy benchmark)."""
        a, b = C.c_double(), C.c_double()
        _lib.check(self._lib.p265_pcie_probe(self._ctx, int(n_bytes), int(n_buffers), int(reps),
                                             C.byref(a) if h2d else None, C.byref(b) if d2h else None))
        return (a.value if h2d else None), (b.value if d2h else None)

    def int_peak(self, kind: int):
        ops, ms = C.c_double(), C.c_double()
        _lib.check(self._lib.p265_int_peak(self._ctx, int(kind), C.byref(ops), C.byref(ms)))
        return ops.value, ms.value


_engines: dict = {}
_lock = threading.Lock()


def get_engine(device: int = 0) -> Engine:
    """Process-wide default engine per device (what the drop-in modules use)."""
    with _lock:
        e = _engines.get(device)
        if e is None:
            e = _engines[device] = Engine(device)
        return e
