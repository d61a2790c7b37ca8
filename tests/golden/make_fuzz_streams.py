#!/usr/bin/env python
"""Generate small all-intra HEVC test streams that exercise what sanity.bin does not, and
their known answers from libavcodec (run in the dev container; needs /root/reference).

    python tests/golden/make_fuzz_streams.py        # -> tests/golden/fuzz/*.bin + fuzz_ffmpeg.npz

How a stream is made ("bin-level fuzzing through the reference's own parser"):

  1. VPS / SPS / PPS / slice-segment headers are written by the little bit writer below
     (H.265 7.3.1.1 - 7.3.6.1) for the feature set of the stream: bit depth 8 or 10, CTB size,
     scaling_list_enabled_flag (default lists, or explicit lists in the PPS), cu_transquant_bypass,
     transform_skip, sign data hiding, chroma QP offsets, deblocking offsets / disable, several
     slices per picture with slice_loop_filter_across_slices_enabled_flag = 0, ...
  2. The reference's parser (through the py3 shim) is run over the headers with its CABAC
     *decoder hooked*: every decode_decision / decode_bypass / decode_terminate call returns a
     bin drawn from a seeded policy and, at the same time, that bin is ENCODED by the CABAC
     encoder below (9.3.4.x mirrored) with the context state the parser itself selected.  The
     parser's control flow therefore writes a random but syntactically valid slice_segment_data().
  3. Headers + encoded payloads are assembled into an Annex-B stream (emulation prevention
     included).  The unhooked reference parser must read back exactly what the hooked run
     produced (this checks the encoder), and libavcodec decodes the stream twice (loop filters
     skipped / normal) -> the known answers.

The streams are this repository's own data (nothing of the reference is stored in them).
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

from make_ffmpeg_fixture import decode as ffmpeg_decode  # noqa: E402
from oracle import refshim  # noqa: E402

OUT_DIR = os.path.join(HERE, "fuzz")


# ------------------------------------------------------------------------ bit writing
class BitWriter:
    def __init__(self):
        self.bits = []

    def u(self, n, v):
        assert 0 <= v < (1 << n), (n, v)
        self.bits += [(v >> (n - 1 - i)) & 1 for i in range(n)]

    def ue(self, v):
        v += 1
        n = v.bit_length()
        self.bits += [0] * (n - 1)
        self.u(n, v)

    def se(self, v):
        self.ue(2 * v - 1 if v > 0 else -2 * v)

    def align_one(self):                       # rbsp_trailing_bits() / byte_alignment()
        self.bits.append(1)
        while len(self.bits) % 8:
            self.bits.append(0)

    def to_bytes(self) -> bytes:
        assert len(self.bits) % 8 == 0
        return bytes(int("".join(map(str, self.bits[i:i + 8])), 2) for i in range(0, len(self.bits), 8))


def escape(rbsp: bytes) -> bytes:
    """emulation_prevention_three_byte insertion (7.4.2)."""
    out, zeros = bytearray(), 0
    for b in rbsp:
        if zeros >= 2 and b <= 3:
            out.append(3)
            zeros = 0
        out.append(b)
        zeros = zeros + 1 if b == 0 else 0
    if out and out[-1] == 0:                   # cabac_zero_words are not used; keep a NAL from ending in 0x00
        out.append(3)
    return bytes(out)


def nal_unit(nal_type: int, rbsp: bytes) -> bytes:
    hdr = BitWriter()
    hdr.u(1, 0); hdr.u(6, nal_type); hdr.u(6, 0); hdr.u(3, 1)
    return b"\x00\x00\x00\x01" + hdr.to_bytes() + escape(rbsp)


def profile_tier_level(w: BitWriter, profile_idc: int):
    w.u(2, 0); w.u(1, 0); w.u(5, profile_idc)
    for i in range(32):
        w.u(1, 1 if i == profile_idc or (profile_idc == 1 and i == 2) else 0)
    w.u(1, 1); w.u(1, 0); w.u(1, 0); w.u(1, 1)                 # progressive, frame only
    w.u(32, 0); w.u(12, 0)                                     # general_reserved_zero_44bits
    w.u(8, 93)                                                 # level 3.1


def vps_rbsp(cfg) -> bytes:
    w = BitWriter()
    w.u(4, 0); w.u(2, 3); w.u(6, 0); w.u(3, 0); w.u(1, 1); w.u(16, 0xFFFF)
    profile_tier_level(w, cfg["profile"])
    w.u(1, 1); w.ue(1); w.ue(0); w.ue(0)                       # ordering info: dpb 2, no reorder
    w.u(6, 0); w.ue(0)                                         # vps_max_layer_id, num_layer_sets_minus1
    w.u(1, 0); w.u(1, 0)                                       # timing info, extension
    w.align_one()
    return w.to_bytes()


def sps_rbsp(cfg, sl=None) -> bytes:
    w = BitWriter()
    w.u(4, 0); w.u(3, 0); w.u(1, 1)
    profile_tier_level(w, cfg["profile"])
    w.ue(0); w.ue(1)                                           # sps id, chroma_format_idc 4:2:0
    w.ue(cfg["width"]); w.ue(cfg["height"])
    w.u(1, 0)                                                  # conformance_window_flag
    w.ue(cfg["bit_depth"] - 8); w.ue(cfg["bit_depth"] - 8)
    w.ue(4)                                                    # log2_max_pic_order_cnt_lsb_minus4
    w.u(1, 1); w.ue(1); w.ue(0); w.ue(0)
    w.ue(0); w.ue(cfg["ctb_log2"] - 3)                         # min CB 8
    w.ue(0); w.ue(min(cfg["ctb_log2"], 5) - 2)                 # TB 4 .. min(CTB, 32)
    w.ue(cfg["tu_depth"]); w.ue(cfg["tu_depth"])
    w.u(1, cfg["scaling_lists"] != "off")
    if cfg["scaling_lists"] != "off":
        w.u(1, 1 if sl else 0)                                 # sps_scaling_list_data_present_flag
        if sl:
            write_scaling_list_data(w, *sl)
    w.u(1, 0)                                                  # amp
    w.u(1, 1)                                                  # sample_adaptive_offset_enabled_flag
    pcm = cfg.get("pcm")
    w.u(1, 1 if pcm else 0)                                    # pcm_enabled_flag
    if pcm:
        w.u(4, pcm["bits_y"] - 1); w.u(4, pcm["bits_c"] - 1)   # pcm_sample_bit_depth_{luma,chroma}_minus1
        w.ue(pcm["log2_min"] - 3); w.ue(pcm["log2_max"] - pcm["log2_min"])
        w.u(1, pcm["lf_disabled"])                             # pcm_loop_filter_disabled_flag
    w.ue(0)                                                    # num_short_term_ref_pic_sets
    w.u(1, 0); w.u(1, 0)                                       # long-term refs, temporal mvp
    w.u(1, cfg["strong_smoothing"])
    w.u(1, 0); w.u(1, 0)                                       # vui, extension
    w.align_one()
    return w.to_bytes()


def write_scaling_list_data(w: BitWriter, lists, dc, modes):
    """scaling_list_data() (7.3.4); modes[(s, m)] = 'explicit' | 'default' | ('ref', delta)."""
    for s in range(4):
        for m in range(2 if s == 3 else 6):
            how = modes[(s, m)]
            if how == "explicit":
                w.u(1, 1)
                nxt = 8
                if s >= 2:
                    w.se(dc[(s, m)] - 8)
                    nxt = dc[(s, m)]
                for v in lists[(s, m)]:
                    w.se((v - nxt + 128) % 256 - 128)
                    nxt = v
            else:
                w.u(1, 0)
                w.ue(0 if how == "default" else how[1])


def pps_rbsp(cfg, sl=None) -> bytes:
    w = BitWriter()
    w.ue(0); w.ue(0)
    w.u(1, 0); w.u(1, 0); w.u(3, 0)                            # dependent slices, output flag, extra bits
    w.u(1, cfg["sdh"]); w.u(1, 0)                              # sign data hiding, cabac_init_present
    w.ue(0); w.ue(0)
    w.se(0)                                                    # init_qp_minus26
    w.u(1, 0)                                                  # constrained_intra_pred_flag
    w.u(1, cfg["transform_skip"])
    w.u(1, 1 if cfg.get("cu_qp_delta") is not None else 0)     # cu_qp_delta_enabled_flag
    if cfg.get("cu_qp_delta") is not None:
        w.ue(cfg["cu_qp_delta"])                               # diff_cu_qp_delta_depth
    w.se(cfg["cb_qp_offset"]); w.se(cfg["cr_qp_offset"])
    w.u(1, 1 if cfg.get("slice_chroma_offsets") else 0)        # pps_slice_chroma_qp_offsets_present_flag
    w.u(1, 0); w.u(1, 0)                                       # weighted pred / bipred
    w.u(1, cfg["bypass"])                                      # transquant_bypass_enabled_flag
    tiles = cfg.get("tiles")
    w.u(1, 1 if tiles else 0); w.u(1, 0)                       # tiles_enabled_flag, wavefront
    if tiles:
        w.ue(tiles[0] - 1); w.ue(tiles[1] - 1)                 # num_tile_columns_minus1, num_tile_rows_minus1
        w.u(1, 1)                                              # uniform_spacing_flag
        w.u(1, cfg["lf_across_tiles"])                         # loop_filter_across_tiles_enabled_flag
    w.u(1, 1)                                                  # pps_loop_filter_across_slices_enabled_flag
    ctl = cfg["dbk_disable"] or cfg["beta_offset_div2"] or cfg["tc_offset_div2"] or cfg.get("dbk_override")
    w.u(1, 1 if ctl else 0)                                    # deblocking_filter_control_present_flag
    if ctl:
        w.u(1, 1 if cfg.get("dbk_override") else 0)            # deblocking_filter_override_enabled_flag
        w.u(1, cfg["dbk_disable"])
        if not cfg["dbk_disable"]:
            w.se(cfg["beta_offset_div2"]); w.se(cfg["tc_offset_div2"])
    w.u(1, 1 if sl else 0)                                     # pps_scaling_list_data_present_flag
    if sl:
        write_scaling_list_data(w, *sl)
    w.u(1, 0)                                                  # lists_modification_present_flag
    w.ue(0)                                                    # log2_parallel_merge_level_minus2
    w.u(1, 0); w.u(1, 0)                                       # slice header extension, pps extension
    w.align_one()
    return w.to_bytes()


def slice_header_bits(cfg, first: bool, address: int, qp: int, across: int, index: int = 0) -> BitWriter:
    w = BitWriter()
    w.u(1, 1 if first else 0)
    w.u(1, 0)                                                  # no_output_of_prior_pics_flag (IDR)
    w.ue(0)                                                    # slice_pic_parameter_set_id
    if not first:
        n_ctb = cfg["ctbs_w"] * cfg["ctbs_h"]
        w.u(max(1, (n_ctb - 1).bit_length()), address)
    w.ue(2)                                                    # slice_type I
    w.u(1, 1); w.u(1, cfg["sao_chroma"])                       # slice_sao_luma_flag, slice_sao_chroma_flag
    w.se(qp - 26)                                              # slice_qp_delta (init_qp_minus26 = 0)
    if cfg.get("slice_chroma_offsets"):                        # slice_cb_qp_offset, slice_cr_qp_offset: dequantisation
        cb, cr = cfg["slice_chroma_offsets"][index % len(cfg["slice_chroma_offsets"])]   # only (8.6.1); deblocking's
        w.se(cb); w.se(cr)                                     # cQpPicOffset is the PPS offset alone (8.7.2.5.5)
    if cfg.get("dbk_override"):                                # None: PPS values | "off" | (beta_offset_div2, tc_offset_div2)
        o = cfg["dbk_override"][index % len(cfg["dbk_override"])]
        w.u(1, 0 if o is None else 1)                          # deblocking_filter_override_flag
        if o is not None:
            w.u(1, 1 if o == "off" else 0)                     # slice_deblocking_filter_disabled_flag
            if o != "off":
                w.se(o[0]); w.se(o[1])                         # slice_beta_offset_div2, slice_tc_offset_div2
    w.u(1, across)                                             # slice_loop_filter_across_slices_enabled_flag
    if cfg.get("tiles"):
        w.ue(0)                                                # num_entry_point_offsets: one tile per slice
    w.align_one()                                              # byte_alignment()
    return w


# ------------------------------------------------------------------------ CABAC encoder
class CabacEncoder:
    """9.3.4.x arithmetic *encoding* (the mirror of cabac.py:214-294), context state supplied
    by the caller for every decision."""

    def __init__(self, lps_range_table):
        self.lps = lps_range_table
        self.reset()

    def reset(self):
        self.low, self.range, self.first, self.outstanding, self.bits = 0, 510, True, 0, []

    def _put(self, b):
        if self.first:
            self.first = False
        else:
            self.bits.append(b)
        while self.outstanding:
            self.bits.append(1 - b)
            self.outstanding -= 1

    def _renorm(self):
        while self.range < 256:
            if self.low < 256:
                self._put(0)
            elif self.low >= 512:
                self.low -= 512
                self._put(1)
            else:
                self.low -= 256
                self.outstanding += 1
            self.range <<= 1
            self.low <<= 1

    def decision(self, p_state_idx, val_mps, b):
        r_lps = self.lps[p_state_idx][(self.range >> 6) & 3]
        self.range -= r_lps
        if b != val_mps:
            self.low += self.range
            self.range = r_lps
        self._renorm()

    def bypass(self, b):
        self.low <<= 1
        if b:
            self.low += self.range
        if self.low >= 1024:
            self._put(1)
            self.low -= 1024
        elif self.low < 512:
            self._put(0)
        else:
            self.low -= 512
            self.outstanding += 1

    def terminate(self, b):
        self.range -= 2
        if b:
            self.low += self.range
            self.range = 2                       # EncodeFlush
            self._renorm()
            self._put((self.low >> 9) & 1)
            self.bits += [(self.low >> 8) & 1, 1]   # WriteBits(((low >> 7) & 3) | 1, 2): last bit = rbsp_stop_one_bit
        else:
            self._renorm()

    def pcm_samples(self, values, n_bits):
        """After terminate(1) for pcm_flag: pcm_alignment_zero_bits, the raw samples (7.3.8.7), and the
        arithmetic encoder starts again (9.3.2.5) -- the bits written so far stay."""
        while len(self.bits) % 8:
            self.bits.append(0)
        for v, n in zip(values, n_bits):
            self.bits += [(v >> (n - 1 - i)) & 1 for i in range(n)]
        assert len(self.bits) % 8 == 0
        self.low, self.range, self.first, self.outstanding = 0, 510, True, 0

    def payload(self) -> bytes:
        bits = list(self.bits)
        while len(bits) % 8:
            bits.append(0)
        return bytes(int("".join(map(str, bits[i:i + 8])), 2) for i in range(0, len(bits), 8))


# ------------------------------------------------------------------------ bin policy
class Policy:
    def __init__(self, seed, dense, big=False):
        self.rng = np.random.default_rng(seed)
        self.max_ones = 13 if big else 9
        d = dense
        self.p = {"split_cu_flag": 0.7, "part_mode": 0.5, "prev_intra_luma_pred_flag": 0.5,
                  "intra_chroma_pred_mode": 0.5, "split_transform_flag": 0.5,
                  "cbf_luma": 0.75 if d else 0.35, "cbf_chroma": 0.5 if d else 0.2,
                  "transform_skip_flag": 0.3, "cu_transquant_bypass_flag": 0.15,
                  "last_sig_coeff_x_prefix": 0.55 if d else 0.3, "last_sig_coeff_y_prefix": 0.55 if d else 0.3,
                  "coded_sub_block_flag": 0.5 if d else 0.3, "sig_coeff_flag": 0.45 if d else 0.25,
                  "coeff_abs_level_greater1_flag": 0.4 if d else 0.15,
                  "coeff_abs_level_greater2_flag": 0.4 if d else 0.15,
                  "sao_merge_leftup_flag": 0.3, "sao_type_idx_lumachroma_flag": 0.75, "pcm_flag": 0.35}
        self.p_bypass = 0.62 if big else (0.4 if d else 0.3)    # big: long remaining-level prefixes
        self.ones = 0
        self.qpd = 0

    def decision(self, name):
        b = int(self.rng.random() < self.p.get(name, 0.5))
        if name == "cu_qp_delta_abs":                      # prefix of at most 4 ones: |CuQpDeltaVal| <= 4, no suffix
            if self.qpd >= 4:
                b = 0
            self.qpd = self.qpd + 1 if b else 0
        else:
            self.qpd = 0
        return b

    def bypass(self):
        b = int(self.rng.random() < self.p_bypass)
        if self.ones >= self.max_ones:                     # bounds coeff_abs_level_remaining prefixes (|level| stays in int16)
            b = 0
        self.ones = self.ones + 1 if b else 0
        return b


# ------------------------------------------------------------------------ generation
def run_parser(ns, path, hook=None):
    """Run the reference's Decoder over `path`; returns (images, sps, pps)."""
    import logging
    wd = tempfile.mkdtemp(prefix="p265_fuzz_")
    os.makedirs(os.path.join(wd, "logs"), exist_ok=True)
    args = types.SimpleNamespace(bitstream=path, skip_syntax_dump=1000000, output=None, plot=None)
    cwd = os.getcwd()
    os.chdir(wd)
    prev, logging.raiseExceptions = logging.raiseExceptions, False
    try:
        d = ns.dec.Decoder(args)
        if hook:
            hook(d)
        try:
            d.decode()
        except SystemExit:
            pass
    finally:
        logging.raiseExceptions = prev
        os.chdir(cwd)
    return d.ctx.dpb.images, d.ctx.sps, d.ctx.pps


def make_stream(ns, cfg):
    ctb = 1 << cfg["ctb_log2"]
    cfg["ctbs_w"], cfg["ctbs_h"] = -(-cfg["width"] // ctb), -(-cfg["height"] // ctb)
    n_ctb = cfg["ctbs_w"] * cfg["ctbs_h"]
    sl = sl_sps = None
    if "pps" in cfg["scaling_lists"]:
        sl = random_scaling_lists(cfg["seed"])
    if "sps" in cfg["scaling_lists"]:
        sl_sps = random_scaling_lists(cfg["seed"] + 100)
    head = nal_unit(32, vps_rbsp(cfg)) + nal_unit(33, sps_rbsp(cfg, sl_sps)) + nal_unit(34, pps_rbsp(cfg, sl))
    # slices of every picture: list of (first CTB address, qp, across flag)
    rng = np.random.default_rng(cfg["seed"] + 1)
    pictures = []
    tile_ends = None
    if cfg.get("tiles"):
        # one slice per tile, in tile-scan order (6.5.1): slice_segment_address = raster address of the
        # tile's first CTB.  Every slice has slice_loop_filter_across_slices_enabled_flag = 1, so the only
        # thing that stops the in-loop filters at these boundaries is loop_filter_across_tiles_enabled_flag
        # (uniform spacing with CTB counts divisible by the tile counts, see pps.py:87-91)
        tc, tr = cfg["tiles"]
        assert cfg["ctbs_w"] % tc == 0 and cfg["ctbs_h"] % tr == 0 and cfg["slices"] == tc * tr
        cw, rh = cfg["ctbs_w"] // tc, cfg["ctbs_h"] // tr
        starts = [j * rh * cfg["ctbs_w"] + i * cw for j in range(tr) for i in range(tc)]
        tile_ends = [((j + 1) * rh - 1) * cfg["ctbs_w"] + (i + 1) * cw - 1 for j in range(tr) for i in range(tc)]
    for _ in range(cfg["pictures"]):
        if tile_ends is not None:
            pictures.append([(a, int(rng.choice(cfg["qps"])), 1) for a in starts])
            continue
        starts = [0] + sorted(rng.choice(np.arange(1, n_ctb), size=min(cfg["slices"] - 1, n_ctb - 1),
                                         replace=False).tolist()) if cfg["slices"] > 1 else [0]
        # slice_loop_filter_across_slices_enabled_flag: 1, 0, 1, ... (the first slice has no left / upper slice)
        pictures.append([(a, int(rng.choice(cfg["qps"])), 1 - (i & 1)) for i, a in enumerate(starts)])
    headers = [[slice_header_bits(cfg, a == 0, a, qp, across, i).to_bytes() for i, (a, qp, across) in enumerate(pic)]
               for pic in pictures]
    # pass 1: headers only; the hooked parser writes the slice data
    skeleton = head + b"".join(nal_unit(19, h) for pic in headers for h in pic) + b"\x00" * 16
    tmp = tempfile.NamedTemporaryFile(suffix=".bin", delete=False)
    tmp.write(skeleton)
    tmp.close()
    policy = Policy(cfg["seed"], cfg["dense"], cfg["big"])
    payloads = []
    ends = [[(pic[i + 1][0] if i + 1 < len(pic) else n_ctb) - 1 for i in range(len(pic))] for pic in pictures]
    if tile_ends is not None:
        ends = [list(tile_ends) for _ in pictures]             # a tile's last CTB in raster addressing
    state = {"pic": 0, "slice": 0}

    def hook(d):
        cab = d.ctx.cabac
        enc = CabacEncoder(cab.tables.lps_range_table)
        cab.initialization_process_arithmetic_decoding_engine = enc.reset

        def decision(ctx_table, ctx_idx):
            m = cab.context_models[ctx_table][ctx_idx]
            b = policy.decision(ctx_table)
            enc.decision(m.p_state_idx, m.val_mps, b)
            cab.state_transition_process(ctx_table, ctx_idx, b)
            return b

        def bypass():
            b = policy.bypass()
            enc.bypass(b)
            return b

        def terminate():
            last = ends[state["pic"]][state["slice"]]
            b = int(d.ctx.img.ctu.addr_rs == last)
            enc.terminate(b)
            if b:
                payloads.append(enc.payload())
                state["slice"] += 1
                if state["slice"] == len(ends[state["pic"]]):
                    state["pic"], state["slice"] = state["pic"] + 1, 0
            return b
        cab.decode_decision, cab.decode_bypass, cab.decode_terminate = decision, bypass, terminate
        state["enc"] = enc

    prepare(ns, cfg)
    cu_cls, saved = sys.modules["cu"].Cu, None
    if cfg.get("pcm"):
        # pcm_flag is a terminate bin of its own (not the end of the slice segment) and pcm_sample() is raw
        # bits between two arithmetic codewords: the hooked run draws both from the policy
        saved = (cu_cls.parse__pcm_flag, cu_cls.parse__pcm_sample)
        pcm = cfg["pcm"]

        def parse__pcm_flag(self):
            b = policy.decision("pcm_flag")
            state["enc"].terminate(b)
            return b

        def parse__pcm_sample(self):
            n = 1 << self.log2size
            h = n // 2
            # smooth-ish content: a random level per CU plus noise, so that the loop filters have decisions to make
            def block(count, bits):
                base = int(policy.rng.integers(0, 1 << bits))
                amp = int(policy.rng.choice([1, 3, 1 << max(bits - 2, 1)]))
                return np.clip(base + policy.rng.integers(-amp, amp + 1, count), 0, (1 << bits) - 1).astype(np.int64)
            y = block(n * n, pcm["bits_y"])
            c = np.concatenate([block(h * h, pcm["bits_c"]), block(h * h, pcm["bits_c"])])
            self.pcm_sample_luma = y.reshape(n, n)
            self.pcm_sample_chroma = c.reshape(2, h, h)
            state["enc"].pcm_samples([int(v) for v in y] + [int(v) for v in c],
                                     [pcm["bits_y"]] * (n * n) + [pcm["bits_c"]] * (2 * h * h))
        cu_cls.parse__pcm_flag, cu_cls.parse__pcm_sample = parse__pcm_flag, parse__pcm_sample
    try:
        imgs, sps, pps = run_parser(ns, tmp.name, hook)
    finally:
        if saved:
            cu_cls.parse__pcm_flag, cu_cls.parse__pcm_sample = saved
    os.unlink(tmp.name)
    assert len(imgs) == cfg["pictures"] and len(payloads) == sum(len(p) for p in pictures)
    k, body = 0, b""
    for pic in headers:
        for h in pic:
            body += nal_unit(19, h + payloads[k])
            k += 1
    return head + body + b"\x00" * 8, (imgs, sps, pps)


def random_scaling_lists(seed):
    rng = np.random.default_rng(seed + 7)
    lists, dc, modes = {}, {}, {}
    for s in range(4):
        for m in range(2 if s == 3 else 6):
            r = rng.random()
            if m > 0 and r < 0.2:
                delta = int(rng.integers(1, m + 1))
                modes[(s, m)] = ("ref", delta)
                lists[(s, m)] = list(lists[(s, m - delta)])
                if s >= 2:
                    dc[(s, m)] = dc[(s, m - delta)]
            elif r < 0.35:
                from p265_b200 import scaling_list
                modes[(s, m)] = "default"
                lists[(s, m)] = scaling_list.default_list(s, m)
                if s >= 2:
                    dc[(s, m)] = 16
            else:
                modes[(s, m)] = "explicit"
                base = rng.integers(8, 40)
                lists[(s, m)] = [int(np.clip(base + i // 3 + rng.integers(-4, 5), 1, 255))
                                 for i in range(16 if s == 0 else 64)]
                if s >= 2:
                    dc[(s, m)] = int(rng.integers(4, 64))
    return lists, dc, modes


def use_sld_dropin(ns):
    """PPS-level scaling_list_data(): the reference's own sld.ScalingListData.decode cannot run
    (SURVEY.md G4); swap in p265_b200's drop-in class for pps.py:11."""
    from p265_b200.scaling_list import ScalingListData
    sys.modules["sld"].ScalingListData = ScalingListData
    sys.modules["pps"].sld.ScalingListData = ScalingListData


def summarize(imgs, sps):
    """Everything the hot path consumes, for the hooked-vs-clean comparison."""
    from p265_b200 import packer
    out = []
    for img in imgs:
        b = packer.pack_pictures([img], sps)
        modes = []
        for a in sorted(img.ctus):
            for cu in packer._leaf_cus(img.ctus[a]):
                if getattr(cu, "pcm_flag", 0):
                    modes.append((cu.x, cu.y, cu.log2size, "pcm", int(cu.qp_y), cu.pcm_sample_luma.tobytes(),
                                  cu.pcm_sample_chroma.tobytes()))
                    continue
                modes.append((cu.x, cu.y, cu.log2size, cu.part_mode, cu.intra_pred_mode_c,
                              tuple(sorted((x, y, int(v)) for x, col in cu.intra_pred_mode_y.items()
                                           for y, v in col.items())), int(cu.cu_transquant_bypass_flag)))
        out.append((b.tus.tobytes(), b.coeffs.tobytes(), packer.sao_params_from_picture(img, sps).tobytes(), modes))
    return out


STREAMS = [
    # name, overrides of BASE
    ("main8_dense", dict()),
    ("main8_smooth_slices", dict(dense=False, slices=3, qps=(27, 34, 40), ctb_log2=5, seed=11,
                                 beta_offset_div2=2, tc_offset_div2=-3)),
    ("main10_lists_bypass", dict(bit_depth=10, profile=2, scaling_lists="default", bypass=1, dense=False, seed=12,
                                 qps=(22, 30, 37), cb_qp_offset=3, cr_qp_offset=-4, width=136, height=72)),
    # libavcodec's SAO lags deblocking by one CTB, which is not enough for chroma with 16x16 CTBs
    # (observed: a chroma edge-offset neighbour read before its horizontal-edge deblocking): luma SAO only
    ("main10_pps_lists_ctb16", dict(bit_depth=10, profile=2, scaling_lists="pps", ctb_log2=4, dense=True, seed=13,
                                    qps=(17, 26, 45), width=64, height=48, tu_depth=1, sdh=0, sao_chroma=0)),
    ("main8_dbk_off_bypass", dict(dbk_disable=1, bypass=1, dense=False, seed=14, qps=(30, 51), sao_chroma=0,
                                  width=72, height=40, ctb_log2=5)),
    ("main10_big_levels_ctb32", dict(bit_depth=10, profile=2, ctb_log2=5, dense=True, big=True, seed=15,
                                     qps=(40, 46, 51), width=96, height=64, strong_smoothing=0)),
    # 12 bit: libavcodec's qP % 6 / qP / 6 tables stop at qP 73, so QpY 50 / 51 (qP 74 / 75) dequantise as
    # qP 0 / 1 there (observed); the stream stays at QpY <= 49
    ("rext12_lists_ctb32", dict(bit_depth=12, profile=4, scaling_lists="default", ctb_log2=5, dense=True, seed=17,
                                qps=(20, 33, 49), width=96, height=64, cb_qp_offset=2, cr_qp_offset=-2,
                                beta_offset_div2=-2, tc_offset_div2=3)),
    ("main8_big_levels_slices", dict(dense=True, big=True, slices=2, seed=16, qps=(8, 51), width=192, height=128,
                                     tu_depth=3, cb_qp_offset=-5, cr_qp_offset=6, tc_offset_div2=4)),
    # 2x2 tiles (one slice each), loop_filter_across_tiles_enabled_flag = 0: tile-scan CTU order, intra
    # availability, deblocking filterEdgeFlag and SAO neighbour availability at tile boundaries
    ("main8_tiles_2x2_no_lf_across", dict(tiles=(2, 2), lf_across_tiles=0, slices=4, ctb_log2=5, width=128, height=128,
                                          dense=True, seed=18, qps=(26, 34, 41), tc_offset_div2=2)),
    ("main10_tiles_3x2_lf_across", dict(tiles=(3, 2), lf_across_tiles=1, slices=6, ctb_log2=4, width=96, height=64,
                                        bit_depth=10, profile=2, dense=False, seed=19, qps=(24, 36), pictures=1)),
    # 9 bit (round 2): the odd depth libavcodec can decode (9-bit packed arithmetic of deblocking and SAO, bdShift 11 / 6)
    ("rext9_sparse_ctb32", dict(bit_depth=9, profile=4, ctb_log2=5, dense=False, seed=21, qps=(20, 31, 42), width=96,
                                height=64, cb_qp_offset=-2, cr_qp_offset=3, beta_offset_div2=1, tc_offset_div2=-1)),
    # slice-level chroma QP offsets (round 2): they move the chroma qP of dequantisation, not deblocking's QpC
    ("main8_slice_chroma_offsets", dict(slices=3, slice_chroma_offsets=((5, -6), (-7, 4), (0, 8)), cb_qp_offset=4,
                                        cr_qp_offset=-3, ctb_log2=5, width=128, height=64, dense=True, seed=24,
                                        qps=(23, 31, 38), tc_offset_div2=1)),
    # per-slice deblocking override (round 2): PPS offsets, own offsets, deblocking off, in neighbouring slices; the
    # parameters of an edge are those of the slice that holds its q0 sample (8.7.2.5.3), an edge belongs to the CU on
    # its right / lower side (8.7.2.3) -- slices 2 and 4 filter across their upper / left boundaries INTO slices with
    # deblocking off (slice_loop_filter_across_slices_enabled_flag: 1, 0, 1, 0, 1)
    ("main8_dbk_override_slices", dict(slices=5, dbk_override=("off", (3, 2), None, "off", (-4, 2)), beta_offset_div2=-1,
                                       tc_offset_div2=2, ctb_log2=4, width=96, height=64, dense=True, seed=25,
                                       qps=(28, 35, 42), sao_chroma=0)),
    # ... the same with slice_tc_offset_div2 differing between slices as well.  Luma is bit-exact; for CHROMA libavcodec
    # picks the tc offset of an edge segment from the current or the left CTB by loop position, not from the slice
    # of the q0 sample (observed; 8.7.2.5.5 says the latter): the test confines the chroma differences to the edge
    # samples next to CTBs whose neighbours carry another tc offset (lav_chroma_tc_dev)
    ("main8_dbk_override_tc", dict(slices=5, dbk_override=("off", (3, -2), None, "off", (-4, 5)), beta_offset_div2=-1,
                                   tc_offset_div2=2, ctb_log2=4, width=96, height=64, dense=True, seed=25,
                                   qps=(28, 35, 42), sao_chroma=0, lav_chroma_tc_dev=1)),
    # cu_qp_delta (round 2): quantisation groups of 8x8, a QP of its own for every CU with a coded block -- qP differs
    # from TB to TB inside a warp's 32 TBs, deblocking averages QpP / QpQ across every CU edge, QpC over its whole table
    ("main10_cu_qp_delta", dict(cu_qp_delta=2, bit_depth=10, profile=2, ctb_log2=5, width=128, height=64, dense=True,
                                seed=26, qps=(24, 37), slices=2, cb_qp_offset=5, cr_qp_offset=-4, tc_offset_div2=1)),
    # scaling lists in the SPS (round 2), and in both parameter sets: the PPS lists win (7.4.3.3.1)
    ("main8_sps_lists", dict(scaling_lists="sps", ctb_log2=5, dense=True, seed=27, qps=(21, 29, 36), width=96, height=64,
                             sdh=0)),
    ("main10_sps_and_pps_lists", dict(scaling_lists="sps+pps", bit_depth=10, profile=2, ctb_log2=5, dense=True, seed=28,
                                      qps=(25, 33), width=64, height=64, pictures=1)),
    # pcm coding units (round 2): raw samples at a lower PcmBitDepth between two arithmetic codewords, 8x8 .. 32x32;
    # pcm_loop_filter_disabled_flag = 1 (deblocking and SAO leave them alone) / 0 (filtered like any intra CU)
    ("main8_pcm_lf_disabled", dict(pcm=dict(bits_y=7, bits_c=5, log2_min=3, log2_max=5, lf_disabled=1), ctb_log2=5,
                                   width=96, height=64, dense=True, seed=22, qps=(25, 33, 39))),
    ("main10_pcm_filtered", dict(pcm=dict(bits_y=8, bits_c=10, log2_min=3, log2_max=4, lf_disabled=0), ctb_log2=4,
                                 bit_depth=10, profile=2, width=64, height=48, dense=False, seed=23, qps=(22, 30, 38),
                                 tu_depth=1, sao_chroma=0)),
]
BASE = dict(width=128, height=96, bit_depth=8, profile=1, ctb_log2=6, tu_depth=2, scaling_lists="off",
            strong_smoothing=1, sdh=1, transform_skip=1, bypass=0, cb_qp_offset=0, cr_qp_offset=0,
            dbk_disable=0, beta_offset_div2=0, tc_offset_div2=0, sao_chroma=1, pictures=2, slices=1,
            qps=(24, 32), dense=True, big=False, seed=10, tiles=None, lf_across_tiles=1, pcm=None, slice_chroma_offsets=None, dbk_override=None, lav_chroma_tc_dev=0, cu_qp_delta=None)


def prepare(ns, cfg):
    """Harness patches a stream needs before the reference's parser can read it (also used by
    the tests): see oracle/refshim.py for what each one works around."""
    if cfg["scaling_lists"] in ("pps", "sps", "sps+pps"):
        use_sld_dropin(ns)
    if cfg["slices"] > 1:
        refshim.enable_multi_slice(ns)
    if cfg["bypass"]:
        refshim.enable_transquant_bypass(ns)
    if cfg.get("pcm"):
        refshim.enable_pcm(ns)
    if cfg.get("cu_qp_delta") is not None:
        refshim.enable_cu_qp_delta(ns)


def main():
    import json
    ns = refshim.load(tempfile.mkdtemp(prefix="p265ref_"))
    os.makedirs(OUT_DIR, exist_ok=True)
    answers, manifest = {}, {}
    only = set(sys.argv[1:])          # stream names: (re)generate just these, keep every other answer
    if only:
        answers = dict(np.load(os.path.join(HERE, "fuzz_ffmpeg.npz")))
        with open(os.path.join(OUT_DIR, "manifest.json")) as fh:
            manifest = json.load(fh)
    for name, over in STREAMS:
        if only and name not in only:
            continue
        cfg = dict(BASE, **over)
        stream, (imgs, sps, pps) = make_stream(ns, cfg)
        path = os.path.join(OUT_DIR, name + ".bin")
        with open(path, "wb") as fh:
            fh.write(stream)
        want = summarize(imgs, sps)
        imgs2, sps2, _ = run_parser(ns, path)                  # the unhooked parser reads back the same syntax
        got = summarize(imgs2, sps2)
        assert got == want, "%s: clean parse differs from the generating run" % name
        n_tb = sum(len(np.frombuffer(w[0], dtype=np.uint8)) // 16 for w in want)
        for key, skip in (("rec", True), ("out", False)):
            frames = ffmpeg_decode(stream, skip)
            assert len(frames) == cfg["pictures"], (name, len(frames))
            for p, planes in enumerate(frames):
                for c, n in enumerate(("y", "cb", "cr")):
                    answers["%s/%s%d_%s" % (name, key, p, n)] = planes[c]
        manifest[name] = {k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()}
        manifest[name]["tbs"] = n_tb
        print("%-28s %6d bytes  %5d TBs  %dx%d %d-bit" % (name, len(stream), n_tb, cfg["width"], cfg["height"],
                                                         cfg["bit_depth"]))
    np.savez_compressed(os.path.join(HERE, "fuzz_ffmpeg.npz"), **answers)
    with open(os.path.join(OUT_DIR, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
