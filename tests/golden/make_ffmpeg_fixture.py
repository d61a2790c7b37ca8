#!/usr/bin/env python
"""Decode the reference's sanity.bin with an INDEPENDENT conformant HEVC decoder and keep
its output as a known answer (run in the dev container; needs /root/reference).

    python tests/golden/make_ffmpeg_fixture.py        # -> tests/golden/sanity_ffmpeg.npz

The decoder is libavcodec 62.11 (FFmpeg), the shared library bundled with the
opencv-python-headless wheel of this image, driven through ctypes (no ffmpeg binary and
no PyAV here).  Two decodes of the three pictures (all I slices) are stored:

  rec{p}_{y,cb,cr}    skip_loop_filter=all  -> reconstructed samples BEFORE deblocking/SAO
                      (pins dequantisation + inverse transform/DST/transform-skip + intra
                      prediction + reconstruction, bit for bit)
  out{p}_{y,cb,cr}    normal decode          -> final pictures (adds deblocking + SAO)

The reference itself holds no output vectors for this path: these planes
are the independent pin for everything the reference's golden logs cannot pin.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("P265_REFERENCE_ROOT", "/root/reference")
AV_CODEC_ID_HEVC = 173


class _FrameHead(C.Structure):          # first members of AVFrame (libavutil 60)
    _fields_ = [("data", C.c_void_p * 8), ("linesize", C.c_int * 8), ("extended_data", C.c_void_p),
                ("width", C.c_int), ("height", C.c_int), ("nb_samples", C.c_int), ("format", C.c_int)]


def _libs():
    import cv2
    libdir = os.path.join(os.path.dirname(cv2.__file__), "..", "opencv_python_headless.libs")

    def load(pat):
        return C.CDLL(glob.glob(os.path.join(libdir, pat))[0], mode=C.RTLD_GLOBAL)
    avutil = load("libavutil-*")
    load("libswresample-*")
    return avutil, load("libavcodec-*")


def access_units(data: bytes):
    """Annex-B byte stream -> one packet per picture: a new packet starts at the first non-VCL
    NAL unit after a VCL one, or at a VCL NAL unit with first_slice_segment_in_pic_flag = 1."""
    pos = [m.start() for m in re.finditer(b"\x00\x00\x01", data)]
    cur, have_vcl = b"", False
    for i, p in enumerate(pos):
        e = pos[i + 1] if i + 1 < len(pos) else len(data)
        if i + 1 < len(pos) and data[e - 1] == 0:
            e -= 1
        nal = data[p:e]
        vcl = ((nal[3] >> 1) & 0x3F) < 32
        first = vcl and len(nal) > 5 and (nal[5] & 0x80) != 0
        if have_vcl and (not vcl or first):
            yield cur
            cur, have_vcl = b"", False
        cur += b"\x00" + nal
        have_vcl = have_vcl or vcl
    if have_vcl:
        yield cur


def decode(data: bytes, skip_loop_filter: bool):
    avutil, avcodec = _libs()
    avcodec.avcodec_find_decoder.restype = C.c_void_p
    avcodec.avcodec_find_decoder.argtypes = [C.c_int]
    avcodec.avcodec_alloc_context3.restype = C.c_void_p
    avcodec.avcodec_alloc_context3.argtypes = [C.c_void_p]
    avcodec.avcodec_open2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    avcodec.av_packet_alloc.restype = C.c_void_p
    avutil.av_frame_alloc.restype = C.c_void_p
    avcodec.av_new_packet.argtypes = [C.c_void_p, C.c_int]
    avcodec.avcodec_send_packet.argtypes = [C.c_void_p, C.c_void_p]
    avcodec.avcodec_receive_frame.argtypes = [C.c_void_p, C.c_void_p]
    avcodec.av_packet_unref.argtypes = [C.c_void_p]
    avutil.av_opt_set.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
    avutil.av_get_pix_fmt_name.restype = C.c_char_p
    avutil.av_get_pix_fmt_name.argtypes = [C.c_int]
    dec = avcodec.avcodec_find_decoder(AV_CODEC_ID_HEVC)
    ctx = avcodec.avcodec_alloc_context3(dec)
    if skip_loop_filter and avutil.av_opt_set(ctx, b"skip_loop_filter", b"all", 0) != 0:
        raise RuntimeError("av_opt_set(skip_loop_filter) failed")
    if avcodec.avcodec_open2(ctx, dec, None) != 0:
        raise RuntimeError("avcodec_open2 failed")
    pkt, frm = avcodec.av_packet_alloc(), avutil.av_frame_alloc()
    frames = []

    def drain():
        while avcodec.avcodec_receive_frame(ctx, frm) >= 0:
            h = _FrameHead.from_address(frm)
            name = avutil.av_get_pix_fmt_name(h.format).decode()
            if name not in ("yuv420p", "yuv420p9le", "yuv420p10le", "yuv420p12le"):
                raise RuntimeError("unexpected pixel format %s" % name)
            dt = np.uint8 if name == "yuv420p" else np.dtype("<u2")
            planes = []
            for c in range(3):
                hh, ww = (h.height, h.width) if c == 0 else (h.height // 2, h.width // 2)
                buf = (C.c_uint8 * (h.linesize[c] * hh)).from_address(h.data[c])
                rows = np.frombuffer(buf, np.uint8).reshape(hh, h.linesize[c])
                planes.append(rows[:, :ww * dt.itemsize].copy().view(dt) if dt != np.uint8 else rows[:, :ww].copy())
            frames.append(planes)

    for au in access_units(data):
        avcodec.av_new_packet(pkt, len(au))
        C.memmove(C.c_void_p.from_address(pkt + 24).value, au, len(au))   # AVPacket.data
        if avcodec.avcodec_send_packet(ctx, pkt) != 0:
            raise RuntimeError("avcodec_send_packet failed")
        avcodec.av_packet_unref(pkt)
        drain()
    avcodec.avcodec_send_packet(ctx, None)
    drain()
    return frames


def main():
    data = open(os.path.join(REF_ROOT, "sanity.bin"), "rb").read()
    out = {}
    for key, skip in (("rec", True), ("out", False)):
        frames = decode(data, skip)
        assert len(frames) == 3
        for p, planes in enumerate(frames):
            for c, n in enumerate(("y", "cb", "cr")):
                out["%s%d_%s" % (key, p, n)] = planes[c]
    path = os.path.join(HERE, "sanity_ffmpeg.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")
    return 0


if __name__ == "__main__":
    sys.exit(main())
