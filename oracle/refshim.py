"""TEST INFRASTRUCTURE ONLY -- load-time Python-3 shim of the reference decoder.

The reference (jacke121/p265, /root/reference) is Python 2 and "all rights reserved"
(LICENSE:1), so nothing of it is committed here.  `build()` mechanically transforms
the reference's own files into the git-ignored directory `baseline/_ref/p265ref/`
(which still travels to the GPU box with gpurun) and `load()` imports them.

Transform, nothing else is touched:
  1. `print X[,]`            -> `print(X[, end=' '])`
  2. every `/` operator token -> `//` except on lines containing `float(`
     (sps.py:151,155, bsb.py:155 really want true division)
  3. dec.py: `import decoder.X as X` -> `import X` (so `log` is one module)
  4. stub `matplotlib`, `matplotlib.pyplot`, `matplotlib.patches` (imported at
     image.py:1-2, tree.py:2-3, slice.py:1; never used while decoding)
  5. pps.py:64-65: the misspelt `num_tile_colums_minus1` (tile branch only; see transform_source)
  6. slice.py:174-175: the unimplemented deblocking_filter_override branch (see transform_source)
  7. cu.py:557 / tu.py:100: two one-token fixes on the cu_qp_delta path (see transform_source)
  8. sps.py:90: `scaling_list_data.decode()` -> `self.scaling_list_data.decode()` (see transform_source)

Only tests/, bench.py's cpu_baseline / --impl reference leg and
__graft_entry__ may import this module; the product package never does.
"""
from __future__ import annotations

import io
import os
import re
import sys
import tokenize
import types

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = os.environ.get("P265_REFERENCE_ROOT", "/root/reference")
SHIM_DIR = os.path.join(REPO, "baseline", "_ref", "p265ref")

_PRINT_RE = re.compile(r"^(\s*)print\s+(?!\()(.*?)(,?)\s*$")
_PRINT_PAREN_RE = re.compile(r"^(\s*)print\s+(\(.*\))\s*$")


def _fix_print(line: str) -> str:
    body = line.rstrip("\n")
    m = _PRINT_RE.match(body)
    if m:
        indent, expr, comma = m.groups()
        end = ", end=' '" if comma else ""
        return "%sprint(%s%s)\n" % (indent, expr, end)
    return line


def _fix_division(src: str) -> str:
    """Replace the `/` OP token by `//` (py2 int division) except on float( lines."""
    lines = src.splitlines(keepends=True)
    edits = []  # (row, col)
    for tok in tokenize.generate_tokens(io.StringIO(src).readline):
        if tok.type == tokenize.OP and tok.string == "/":
            row, col = tok.start
            if "float(" in lines[row - 1]:
                continue
            edits.append((row, col))
    for row, col in sorted(edits, reverse=True):
        ln = lines[row - 1]
        lines[row - 1] = ln[:col] + "//" + ln[col + 1:]
    return "".join(lines)


#: bump when the transform changes: build() regenerates a shim directory with another stamp
SHIM_VERSION = "8"


def transform_source(src: str, name: str) -> str:
    src = src.replace("\t", "        ")
    src = "".join(_fix_print(l) for l in src.splitlines(keepends=True))
    src = _fix_division(src)
    if name == "dec.py":
        src = re.sub(r"^import decoder\.(\w+) as \1$", r"import \1", src, flags=re.M)
    if name == "pps.py":
        # 5. harness patch for the tile fuzz stream (tests/golden/make_fuzz_streams.py): the PPS tile
        #    branch cannot run as written -- pps.py:64-65 read the misspelt attribute
        #    `num_tile_colums_minus1` (AttributeError) and derive the ROW count from the COLUMN syntax
        #    element.  Only reached when tiles_enabled_flag = 1 (sanity.bin: 0; the 95 golden logs are
        #    unaffected).
        src = src.replace("self.num_tile_rows = self.num_tile_colums_minus1 + 1", "self.num_tile_rows = self.num_tile_rows_minus1 + 1")
        src = src.replace("self.num_tile_columns = self.num_tile_colums_minus1 + 1",
                          "self.num_tile_columns = self.num_tile_columns_minus1 + 1")
        #    ... and pps.py:178 reads `self.pic_width_in_ctbs_y` (an Sps attribute) for every tile row but the first
        src = src.replace("+= self.pic_width_in_ctbs_y * self.row_height[j]",
                          "+= self.sps.pic_width_in_ctbs_y * self.row_height[j]")
    if name == "cu.py":
        #    ... and cu.py:518 (tile branch of decode_qp) reads the misspelt `sps.ctb_log2size_y`
        src = src.replace("self.ctx.sps.ctb_log2size_y", "self.ctx.sps.ctb_log2_size_y")
        # 7. harness patches for the cu_qp_delta fuzz stream (only reached when cu_qp_delta_enabled_flag = 1;
        #    sanity.bin: 0).  cu.py:557: qPY_B falls back to qPY_PREV when the upper neighbour is unavailable OR lies in
        #    another CTB (8.6.1, as cu.py:545 has it for qPY_A); the reference wrote `and`
        old = "if available_b == False and self.get_root().addr_ts != derived_ctb_addr_b:"
        assert old in src, "cu.py: qPY_B condition not found"
        src = src.replace(old, old.replace(" and ", " or "))
    if name == "sps.py":
        # 8. harness patch for the SPS scaling-list fuzz streams: sps.py:90 calls `scaling_list_data.decode()` on an
        #    undefined local name (the object is `self.scaling_list_data`, sps.py:18)
        old = "                scaling_list_data.decode()"
        assert old in src, "sps.py: scaling_list_data.decode() not found"
        src = src.replace(old, "                self.scaling_list_data.decode()")
    if name == "tu.py":
        #    ... and tu.py:100 reads the misspelt `self.cu_qp_data_abs`
        assert "self.cu_qp_data_abs" in src
        src = src.replace("self.cu_qp_data_abs", "self.cu_qp_delta_abs")
    if name == "slice.py":
        #    ... and slice.py:182 calls `self.ue(...)` (no such method) for num_entry_point_offsets, a syntax
        #    element that only exists when tiles or wavefronts are enabled
        src = src.replace('self.ue("num_entry_point_offsets")', 'bs.ue("num_entry_point_offsets")')
        # 6. harness patch for the deblocking-override fuzz stream: slice.py:174-175 raises "Unimplemented yet" when
        #    deblocking_filter_override_flag = 1 (and :178 reads slice_deblocking_filter_disabled_flag, which nothing
        #    assigns).  7.3.6.1: the flag, then the two offsets unless the slice disables deblocking.
        old = ('            if self.deblocking_filter_override_flag:\n'
               '                raise "Unimplemented yet"\n')
        new = ('            self.slice_deblocking_filter_disabled_flag = getattr(self.pps, "pps_deblocking_filter_disabled_flag", 0)\n'
               '            if self.deblocking_filter_override_flag:\n'
               '                self.slice_deblocking_filter_disabled_flag = bs.u(1, "slice_deblocking_filter_disabled_flag")\n'
               '                if not self.slice_deblocking_filter_disabled_flag:\n'
               '                    self.slice_beta_offset_div2 = bs.se("slice_beta_offset_div2")\n'
               '                    self.slice_tc_offset_div2 = bs.se("slice_tc_offset_div2")\n')
        assert old in src, "slice.py: deblocking override branch not found"
        src = src.replace(old, new)
    return src


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "decoder", "transform.py"))


def _stamp() -> str:
    try:
        with open(os.path.join(SHIM_DIR, ".shim_version")) as fh:
            return fh.read().strip()
    except OSError:
        return ""


def shim_available() -> bool:
    return os.path.isfile(os.path.join(SHIM_DIR, "transform.py"))


def build(force: bool = False) -> str:
    """Generate baseline/_ref/p265ref/ from REF_ROOT.  Returns the shim dir."""
    if shim_available() and not force and (_stamp() == SHIM_VERSION or not reference_available()):
        return SHIM_DIR
    if not reference_available():
        raise FileNotFoundError("reference tree not found at %s" % REF_ROOT)
    os.makedirs(SHIM_DIR, exist_ok=True)
    dec_dir = os.path.join(REF_ROOT, "decoder")
    files = [(os.path.join(dec_dir, f), f) for f in sorted(os.listdir(dec_dir))
             if f.endswith(".py") and f != "__init__.py"]
    files.append((os.path.join(REF_ROOT, "dec.py"), "dec.py"))
    files.append((os.path.join(REF_ROOT, "tools", "gen_logs.py"), "gen_logs.py"))
    for path, name in files:
        with open(path, "r") as fh:
            src = fh.read()
        out = transform_source(src, name)
        with open(os.path.join(SHIM_DIR, name), "w") as fh:
            fh.write(out)
    # the bitstream fixture travels too (28 KB) so the GPU box can decode it
    with open(os.path.join(REF_ROOT, "sanity.bin"), "rb") as fi, \
            open(os.path.join(SHIM_DIR, "sanity.bin"), "wb") as fo:
        fo.write(fi.read())
    with open(os.path.join(SHIM_DIR, ".shim_version"), "w") as fh:
        fh.write(SHIM_VERSION)
    return SHIM_DIR


def _stub_matplotlib() -> None:
    if "matplotlib" in sys.modules:
        return
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]


def load(workdir: str | None = None):
    """Import the shimmed reference; returns a namespace of its modules.

    The reference's log.py opens logs/*.log relative to the cwd at import time
    (log.py:33-83), so the first call chdir()s into `workdir` (a scratch dir with a
    logs/ sub-directory) for the import.
    """
    if not shim_available() or (_stamp() != SHIM_VERSION and reference_available()):
        build()
    _stub_matplotlib()
    if SHIM_DIR not in sys.path:
        sys.path.insert(0, SHIM_DIR)
    import tempfile
    if "log" not in sys.modules or not hasattr(sys.modules["log"], "syntax"):
        wd = workdir or tempfile.mkdtemp(prefix="p265ref_")
        os.makedirs(os.path.join(wd, "logs"), exist_ok=True)
        cwd = os.getcwd()
        os.chdir(wd)
        try:
            import log  # noqa: F401
        finally:
            os.chdir(cwd)
    import importlib
    ns = types.SimpleNamespace()
    for name in ("utils", "scaling", "transform", "reconstruction", "sao", "tu", "cu",
                 "ctu", "nalu", "context", "log", "dec", "sld", "scan", "image", "intra"):
        setattr(ns, name, importlib.import_module(name))
    return ns


if __name__ == "__main__":
    print(build(force=True))


def enable_multi_slice(ns) -> None:
    """Test-harness patch: the reference creates a CTU object only for the first slice
    segment of a picture (slice.py:237-238), so a second slice re-parses the previous CTU and
    trips `assert self.ctu.addr_rs not in self.ctus` (image.py:15).  Wrap SliceSegmentData.parse
    so that a non-first, independent slice segment starts at its slice_segment_address --
    needed only for the multi-slice fuzz streams (tests/golden/make_fuzz_streams.py)."""
    cls = ns.dec.nalu.slice.SliceSegmentData if hasattr(ns.dec, "nalu") else sys.modules["slice"].SliceSegmentData
    if getattr(cls, "_p265_multi_slice", False):
        return
    ctu_mod = sys.modules["ctu"]
    orig = cls.parse

    def parse(self):
        hdr = self.ctx.img.slice_hdrs[-1]
        if not hdr.first_slice_segment_in_pic_flag and not hdr.dependent_slice_segment_flag:
            self.ctx.img.ctu = ctu_mod.Ctu(self.ctx, addr_rs=int(hdr.slice_segment_address))
        return orig(self)

    cls.parse = parse
    cls._p265_multi_slice = True
    # ... and gives `slice_addr` only to the first CTU of a slice segment (slice.py:252; every
    # other Ctu keeps the constructor's 0, ctu.py:13): propagate it when the next CTU is created
    img_cls = sys.modules["image"].Image
    orig_next = img_cls.next_ctu

    def next_ctu(self, end_of_slice_segment_flag):
        orig_next(self, end_of_slice_segment_flag)
        if not end_of_slice_segment_flag:
            self.ctu.slice_addr = self.slice_hdr.slice_segment_address

    img_cls.next_ctu = next_ctu
    # the reference's Sao.parse leaves sao_merge_left_flag / sao_merge_up_flag unassigned when
    # the neighbouring CTB belongs to another slice (sao.py:26-41 -> AttributeError at :50);
    # the product's drop-in `sao` module (golden-log identical, tests/test_dropin_golden.py)
    # takes its place, exactly as INTEGRATION.md describes
    from p265_b200 import sao_api
    sys.modules["ctu"].sao = sao_api


def enable_transquant_bypass(ns) -> None:
    """Test-harness patch: cu.py:102-103 calls `self.parse__cu_transquant_bypass_flag()`, which
    the reference never defines.  The syntax element is one context-coded bin, ctxInc 0
    (9.3.4.2, Table 9-8; the reference's own init table has the entry, cabac.py:16)."""
    cls = sys.modules["cu"].Cu
    if hasattr(cls, "parse__cu_transquant_bypass_flag"):
        return

    def parse__cu_transquant_bypass_flag(self):
        init_type = int(getattr(self.ctx.img.slice_hdr, "init_type", 0) or 0)
        bit = self.ctx.cabac.decode_decision("cu_transquant_bypass_flag", init_type)
        sys.modules["log"].syntax.info("cu_transquant_bypass_flag = %d" % bit)
        return bit

    cls.parse__cu_transquant_bypass_flag = parse__cu_transquant_bypass_flag


def enable_pcm(ns) -> None:
    """Test-harness patch: the reference's pcm branch cannot run.  cu.py:139-151 reads
    `sps.log2_min_pcm_luma_coding_block_size` / `log2_max_...` (sps.py:95-101 stores only the
    `_minus3` / `_diff_` syntax elements) and calls `parse__pcm_flag`, `parse__pcm_sample` and
    (cu.py:484-485) `decode_pcm`, none of which exist.  Added here, for the pcm fuzz streams only:

      * the two derived SPS variables (7.4.3.2.1: Log2MinIpcmCbSizeY, Log2MaxIpcmCbSizeY);
      * pcm_flag: one terminate bin (9.3.4.3.5);
      * pcm_sample(): 7.3.8.7 -- (1 << (log2CbSize << 1)) luma samples of PcmBitDepthY bits, then
        Cb and Cr of PcmBitDepthC bits each, raster order inside the coding block; kept on the CU as
        `pcm_sample_luma` (N, N) and `pcm_sample_chroma` (2, N/2, N/2), [row][col]; afterwards the
        arithmetic decoding engine is initialised again (9.3.2.5);
      * decode_pcm: only the CU's QpY (8.6.1; the in-loop filters read it), like every other CU."""
    cls = sys.modules["cu"].Cu
    if hasattr(cls, "parse__pcm_sample"):
        return
    import numpy as np
    sps_cls = sys.modules["sps"].Sps
    sps_cls.log2_min_pcm_luma_coding_block_size = property(
        lambda self: self.log2_min_pcm_luma_coding_block_size_minus3 + 3)
    sps_cls.log2_max_pcm_luma_coding_block_size = property(
        lambda self: self.log2_min_pcm_luma_coding_block_size_minus3 + 3 + self.log2_diff_max_min_pcm_luma_coding_block_size)

    def parse__pcm_flag(self):
        bit = self.ctx.cabac.decode_terminate()
        sys.modules["log"].syntax.info("pcm_flag = %d" % bit)
        return bit

    def parse__pcm_sample(self):
        bs, sps = self.ctx.bs, self.ctx.sps
        n = 1 << self.log2size
        dy, dc = sps.pcm_sample_bit_depth_luma_minus1 + 1, sps.pcm_sample_bit_depth_chroma_minus1 + 1
        self.pcm_sample_luma = np.array([bs.read_bits(dy) for _ in range(n * n)], np.int64).reshape(n, n)
        h = n // 2
        self.pcm_sample_chroma = np.array([bs.read_bits(dc) for _ in range(2 * h * h)], np.int64).reshape(2, h, h)
        self.ctx.cabac.initialization_process_arithmetic_decoding_engine()

    def decode_pcm(self):
        self.decode_qp()

    cls.parse__pcm_flag, cls.parse__pcm_sample, cls.decode_pcm = parse__pcm_flag, parse__pcm_sample, decode_pcm


def enable_cu_qp_delta(ns) -> None:
    """Test-harness patch: tu.py:94,96 call `parse__cu_qp_delta_abs` / `parse__cu_qp_delta_sign_flag`, which the
    reference never defines.  9.3.3.10 / Table 9-4x: cu_qp_delta_abs = prefix TR(cMax 5) with context-coded bins
    (ctxInc 0 for the first, 1 for the others; two contexts per initType, cabac.py:32) + suffix EG0 in bypass bins
    when the prefix is 5; the sign is one bypass bin."""
    cls = sys.modules["tu"].Tu
    if hasattr(cls, "parse__cu_qp_delta_abs"):
        return

    def parse__cu_qp_delta_abs(self):
        cab = self.ctx.cabac
        base = 2 * int(getattr(self.ctx.img.slice_hdr, "init_type", 0) or 0)
        prefix = 0
        while prefix < 5 and cab.decode_decision("cu_qp_delta_abs", base + (1 if prefix else 0)):
            prefix += 1
        value = prefix
        if prefix == 5:
            k = 0
            while cab.decode_bypass():
                value += 1 << k
                k += 1
            for i in range(k - 1, -1, -1):
                value += cab.decode_bypass() << i
        sys.modules["log"].syntax.info("cu_qp_delta_abs = %d" % value)
        return value

    def parse__cu_qp_delta_sign_flag(self):
        bit = self.ctx.cabac.decode_bypass()
        sys.modules["log"].syntax.info("cu_qp_delta_sign_flag = %d" % bit)
        return bit

    cls.parse__cu_qp_delta_abs, cls.parse__cu_qp_delta_sign_flag = parse__cu_qp_delta_abs, parse__cu_qp_delta_sign_flag
