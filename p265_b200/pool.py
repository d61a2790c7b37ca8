"""EnginePool: the in-process multi-GPU dispatcher of the path (SURVEY.md 8(b) "Threading", 8(e)).

The reference decodes in one thread, one NAL loop (dec.py:18-82).  Independent pictures / streams
share nothing, so a decoder process that owns several GPUs hands picture p to GPU p mod G
(`partition.pictures_of_rank`) -- no collective, no data exchange between devices.  One host thread
per GPU drives that GPU's asynchronous contexts (ctypes releases the GIL during every C-ABI call, so
the G threads run concurrently); a picture's call is queued on one context, and its future resolves
once that context's stream has delivered the outputs into the caller's host buffers.

    pool = EnginePool()                       # every visible GPU
    futs = [pool.residual(batch_p, out_p, picture=p) for p, ... ]
    planes = [f.result() for f in futs]
    pool.close()
"""
from __future__ import annotations

import queue
import threading
from concurrent.futures import Future

from . import _lib
from .engine import Engine


class _Worker(threading.Thread):
    """One host thread = one GPU = `n_ctx` asynchronous contexts used round robin; at most one
    call in flight per context (its future is resolved when the context is reused or idle)."""

    def __init__(self, device: int, n_ctx: int):
        super().__init__(daemon=True, name="p265-gpu%d" % device)
        self.device, self.n_ctx = device, n_ctx
        self.jobs: "queue.Queue" = queue.Queue()
        self.engines: list[Engine] = []
        self.ready = threading.Event()
        self.error: BaseException | None = None
        self.calls = 0

    def run(self):
        try:
            self.engines = [Engine(self.device) for _ in range(self.n_ctx)]
            for e in self.engines:
                e.set_async(True)
        except BaseException as exc:           # no device / no library: surface it to the creator
            self.error = exc
            self.ready.set()
            return
        self.ready.set()
        pending: list = [None] * self.n_ctx    # (future, value) per context
        k = 0

        def settle(i):
            if pending[i] is None:
                return
            fut, value = pending[i]
            pending[i] = None
            try:
                self.engines[i].sync()
                fut.set_result(value)
            except BaseException as exc:
                fut.set_exception(exc)

        while True:
            try:
                job = self.jobs.get(timeout=0.0 if any(p is not None for p in pending) else None)
            except queue.Empty:                # nothing queued: deliver what is in flight
                for i in range(self.n_ctx):
                    settle((k + i) % self.n_ctx)
                continue
            if job is None:
                for i in range(self.n_ctx):
                    settle((k + i) % self.n_ctx)
                for e in self.engines:
                    e.close()
                return
            fut, method, args, kwargs = job
            if not fut.set_running_or_notify_cancel():
                continue
            settle(k)                          # the context is free again once its previous call is delivered
            try:
                value = getattr(self.engines[k], method)(*args, **kwargs)
                pending[k] = (fut, value)
                self.calls += 1
            except BaseException as exc:
                fut.set_exception(exc)
            k = (k + 1) % self.n_ctx


class EnginePool:
    def __init__(self, devices=None, contexts_per_device: int = 3):
        n = _lib.load().p265_device_count()
        if n <= 0:
            raise RuntimeError("p265_b200: no CUDA device -- the pool has no CPU fallback")
        self.devices = list(range(n)) if devices is None else [int(d) for d in devices]
        if not self.devices or any(d < 0 or d >= n for d in self.devices):
            raise ValueError("devices %r not within the %d visible GPUs" % (self.devices, n))
        self._workers = [_Worker(d, max(1, int(contexts_per_device))) for d in self.devices]
        for w in self._workers:
            w.start()
        for w in self._workers:
            w.ready.wait()
            if w.error is not None:
                self.close()
                raise w.error
        self._next = 0
        self._lock = threading.Lock()

    # ---------------------------------------------------------------- dispatch
    def device_of(self, picture: int) -> int:
        """Picture p -> GPU p mod G (the partition of SURVEY.md 8(e))."""
        return self.devices[picture % len(self.devices)]   # == partition.pictures_of_rank(n, rank, G) inverted

    def submit(self, method: str, *args, picture: int | None = None, **kwargs) -> Future:
        """Queue `Engine.<method>(*args, **kwargs)` on the GPU that owns `picture` (round robin over
        the GPUs when no picture index is given).  Host buffers must stay alive and untouched until
        the future resolves; the result is the method's return value (the output array)."""
        with self._lock:
            if picture is None:
                picture = self._next
            self._next = picture + 1
        w = self._workers[picture % len(self._workers)]
        fut: Future = Future()
        w.jobs.put((fut, method, args, kwargs))
        return fut

    def residual(self, batch, out=None, picture=None) -> Future:
        return self.submit("residual", batch, out, picture=picture)

    def sao(self, rec, geom, ctb_log2, params, no_filter=None, out=None, inplace=False, picture=None) -> Future:
        return self.submit("sao", rec, geom, ctb_log2, params, no_filter, out, inplace, picture=picture)

    def loop_filter(self, planes, geom, ctb_log2, blk=None, dbk_ctb=None, sao_params=None, no_filter=None,
                    picture=None) -> Future:
        return self.submit("loop_filter", planes, geom, ctb_log2, blk, dbk_ctb, sao_params, no_filter, picture=picture)

    def reconstruct(self, pred, residual, geom, out=None, picture=None) -> Future:
        return self.submit("reconstruct", pred, residual, geom, out, picture=picture)

    # ---------------------------------------------------------------- bookkeeping
    def launch_counts(self) -> dict:
        """Kernels launched so far per device (proof that every GPU worked)."""
        return {w.device: sum(e.launch_count for e in w.engines) for w in self._workers}

    def calls(self) -> dict:
        return {w.device: w.calls for w in self._workers}

    def close(self):
        for w in getattr(self, "_workers", []):
            if w.is_alive():
                w.jobs.put(None)
        for w in getattr(self, "_workers", []):
            w.join(timeout=30)
        self._workers = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
