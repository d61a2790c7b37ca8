#!/usr/bin/env python
"""Regenerate tests/golden/*.npz / *.json from the reference (run in the dev container).

    python tests/golden/make_fixtures.py

Needs /root/reference (read-only).  It
  1. builds the py3 shim (oracle/refshim.py -> baseline/_ref/p265ref/, git-ignored),
  2. decodes sanity.bin with the reference's own parser, splits the log with the
     reference's tools/gen_logs.py and byte-compares all 95 files of test/golden
     (BASELINE config 1) -> golden_manifest.json (sha256 per file, no content),
  3. packs the hot path's *inputs* the parser produced (TB descriptors, coefficient
     arena, per-CTB SAO syntax) -> sanity_inputs.npz,
  4. runs the reference's OWN scaling.inverse_scaling / transform.inverse_transform
     (through the shim, on the duck-typed `pu` of SURVEY 8(b)) on every coded TB of
     sanity.bin and on seeded random TBs -> sanity_residual.npz / random_reference.npz.
     These pin the oracle's dequantisation and its `ref_literal` transform.

Nothing in the `-m gpu` tests, smoke() or bench.py reads /root/reference; they use the
committed fixtures only.
"""
from __future__ import annotations

import hashlib
import json
import os
import runpy
import shutil
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import refshim  # noqa: E402
from p265_b200 import packer  # noqa: E402

REF_GOLDEN = os.path.join(refshim.REF_ROOT, "test", "golden")


def decode_sanity(workdir):
    ns = refshim.load(workdir)
    args = types.SimpleNamespace(bitstream=os.path.join(refshim.SHIM_DIR, "sanity.bin"),
                                 skip_syntax_dump=0, output=None, plot=None)
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        d = ns.dec.Decoder(args)
        try:
            d.decode()
        except SystemExit:          # bsb.py:65-67 calls exit() at end of stream
            pass
        import logging
        for name in ("p265", "p265.location", "p265.syntax_element", "p265.cabac",
                     "p265.qp", "p265.intra"):
            for h in logging.getLogger(name).handlers:
                h.flush()
        os.chdir(os.path.join(workdir, "logs"))
        runpy.run_path(os.path.join(refshim.SHIM_DIR, "gen_logs.py"))
    finally:
        os.chdir(cwd)
    return ns, d


def golden_manifest(workdir):
    logs = os.path.join(workdir, "logs")
    names = sorted(f for f in os.listdir(REF_GOLDEN) if f.endswith(".log"))
    man, bad = {}, []
    for n in names:
        with open(os.path.join(REF_GOLDEN, n), "rb") as fh:
            ref = fh.read()
        p = os.path.join(logs, n)
        got = open(p, "rb").read() if os.path.exists(p) else b""
        man[n] = {"sha256": hashlib.sha256(ref).hexdigest(), "bytes": len(ref)}
        if got != ref:
            bad.append(n)
    return man, bad


class _FakeTu:
    def __init__(self, lv):
        self.lv = lv

    def get_trans_coeff_level(self, x, y, c_idx):
        return int(self.lv[x][y])


def fake_pu(levels_xy, c_idx, qp, bit_depth, intra=True, sf=None):
    """The duck-typed `pu` of SURVEY 8(b) (intra.py:24-37 fields)."""
    n = levels_xy.shape[0]
    sps = types.SimpleNamespace(bit_depth_y=bit_depth, bit_depth_c=bit_depth,
                                qp_bd_offset_y=0, qp_bd_offset_c=0,
                                scaling_list_enabled_flag=0 if sf is None else 1,
                                scaling_factor=sf)
    cu = types.SimpleNamespace(qp_y=qp, qp_cb=qp, qp_cr=qp, cu_transquant_bypass_flag=0,
                               ctx=types.SimpleNamespace(sps=sps),
                               tu=_FakeTu(levels_xy), is_intra_mode=lambda: intra)
    return types.SimpleNamespace(c_idx=c_idx, origin_x=0, origin_y=0, cu=cu,
                                 scaled_samples=np.zeros((n, n), np.int64),
                                 transformed_samples=np.zeros((n, n), np.int64))


def run_reference_tb(ns, levels_xy, c_idx, qp, bit_depth, sf=None):
    l2 = int(np.log2(levels_xy.shape[0]))
    pu = fake_pu(levels_xy, c_idx, qp, bit_depth, sf=sf)
    ns.scaling.inverse_scaling(pu=pu, x0=0, y0=0, log2size=l2)
    ns.transform.inverse_transform(pu=pu, x0=0, y0=0, log2size=l2)
    return pu.scaled_samples.copy(), pu.transformed_samples.copy()


def main():
    workdir = tempfile.mkdtemp(prefix="p265_fixtures_")
    refshim.build(force=True)
    ns, dec = decode_sanity(workdir)
    man, bad = golden_manifest(workdir)
    print("golden files: %d, mismatching: %d %s" % (len(man), len(bad), bad[:5]))
    with open(os.path.join(HERE, "golden_manifest.json"), "w") as fh:
        json.dump({"source": "/root/reference/test/golden", "files": man,
                   "regenerated_identical": len(man) - len(bad), "mismatch": bad}, fh,
                  indent=1, sort_keys=True)

    imgs = dec.ctx.dpb.images
    sps = dec.ctx.sps
    batch = packer.pack_pictures(imgs, sps)
    sao = np.stack([packer.sao_params_from_picture(img, sps) for img in imgs])
    raw_sao = np.zeros((len(imgs), sao.shape[1] * sao.shape[2], 3, 11), np.int16)
    for p, img in enumerate(imgs):
        for addr, ctu in img.ctus.items():
            s = ctu.sao
            for c in range(3):
                raw_sao[p, addr, c] = ([s.sao_type_idx[c], s.sao_band_position[c],
                                        s.sao_eo_class[c]] + list(s.sao_offset_abs[c]) +
                                       list(s.sao_offset_sign[c]))
    np.savez_compressed(
        os.path.join(HERE, "sanity_inputs.npz"),
        tus=batch.tus, coeffs=batch.coeffs, sao=sao, sao_raw=raw_sao,
        geom=np.array([batch.geom.width, batch.geom.height, batch.geom.n_pics,
                       batch.geom.bit_depth_y, batch.geom.bit_depth_c,
                       sps.ctb_log2_size_y]))
    print("TBs:", len(batch.tus), "coeffs:", batch.coeffs.size, "bins:", batch.bin_counts())

    # the reference's own functions on every coded TB of sanity.bin
    scaled = np.zeros(batch.coeffs.size, np.int16)          # [y][x] per TB, arena layout
    literal = np.zeros(batch.coeffs.size, np.int32)         # [x][y] per TB (as written)
    for t in batch.tus:
        n = 1 << int(t["log2n"])
        off = int(t["coeff_off"]) * 16
        lv_yx = batch.coeffs[off:off + n * n].reshape(n, n).astype(np.int64)
        d_xy, r_xy = run_reference_tb(ns, lv_yx.T, int(t["c_idx"]), int(t["qp"]),
                                      batch.geom.bit_depth_y)
        scaled[off:off + n * n] = d_xy.T.reshape(-1)
        literal[off:off + n * n] = r_xy.reshape(-1)
    np.savez_compressed(os.path.join(HERE, "sanity_residual.npz"),
                        ref_scaled_yx=scaled, ref_literal_xy=literal)

    # seeded random TBs incl. both clips, 8- and 10-bit qP range, and a scaling table
    rng = np.random.default_rng(26501)
    recs = []
    from oracle import spec_oracle as so
    lists, dc = so.default_scaling_lists()
    sf_xy = so.expand_scaling_factor(lists, dc)
    sf_ref = [[sf_xy.get((s, m)) for m in range(6)] for s in range(4)]
    for l2, count in ((2, 24), (3, 16), (4, 8), (5, 4)):
        n = 1 << l2
        for i in range(count):
            kind = i % 4
            if kind == 0:
                lv = rng.integers(-32768, 32768, (n, n))
            elif kind == 1:
                lv = (rng.laplace(0, 6, (n, n)) * (rng.random((n, n)) < 0.3)).astype(np.int64)
            elif kind == 2:
                lv = rng.integers(-300, 301, (n, n))
            else:
                lv = np.zeros((n, n), np.int64)
                lv[0, 0] = rng.integers(-2000, 2000)
            bd = 8 if i % 2 == 0 else 10
            qp = int(rng.integers(0, 52 + (12 if bd == 10 else 0)))
            c_idx = int(rng.integers(0, 3))
            use_sf = (i % 3 == 0) and not (l2 == 5 and c_idx > 0)
            d_xy, r_xy = run_reference_tb(ns, lv, c_idx, qp, bd, sf=sf_ref if use_sf else None)
            recs.append(dict(levels_xy=lv, c_idx=c_idx, qp=qp, bit_depth=bd, use_sf=use_sf,
                             scaled_xy=d_xy, literal_xy=r_xy))
    np.savez_compressed(
        os.path.join(HERE, "random_reference.npz"),
        n=len(recs),
        **{"%s_%d" % (k, i): np.asarray(r[k]) for i, r in enumerate(recs) for k in r})
    print("random reference TBs:", len(recs))
    shutil.rmtree(workdir, ignore_errors=True)
    if bad:
        sys.exit("golden regression FAILED for %d files" % len(bad))


if __name__ == "__main__":
    main()
