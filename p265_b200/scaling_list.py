"""ScalingFactor construction on the host (H.265 7.3.4 / 7.4.5).

The reference's own version (decoder/sld.py:118-153) cannot run (undefined names,
`range(0,63)` dropping coefficient 63, `for k in [0,3]`, DC written from index [0];
SURVEY.md G4) and nothing ever assigns `sps.scaling_factor`, which
scaling.inverse_scaling reads (scaling.py:44).  This module restates the standard's
derivation and produces the object scaling.py expects:

    sps.scaling_factor[size_id][matrix_id][x][y]

plus the packed 4064-byte device table (picture.pack_scaling_factor).  It is tiny,
sequential, host-only work.
"""
from __future__ import annotations

import numpy as np

# Table 7-5 (4x4: flat) and Table 7-6 (8x8 intra / inter), coefficient order = up-right
# diagonal scan.  The values are the standard's (identical to sld.py:4-31).
DEFAULT_4x4 = (16,) * 16
DEFAULT_8x8_INTRA = (
    16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 17, 16, 17, 16, 17, 18,
    17, 18, 18, 17, 18, 21, 19, 20, 21, 20, 19, 21, 24, 22, 22, 24,
    24, 22, 22, 24, 25, 25, 27, 30, 27, 25, 25, 29, 31, 35, 35, 31,
    29, 36, 41, 44, 41, 36, 47, 54, 54, 47, 65, 70, 65, 88, 88, 115)
DEFAULT_8x8_INTER = (
    16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 17, 17, 17, 17, 17, 18,
    18, 18, 18, 18, 18, 20, 20, 20, 20, 20, 20, 20, 24, 24, 24, 24,
    24, 24, 24, 24, 25, 25, 25, 25, 25, 25, 25, 28, 28, 28, 28, 28,
    28, 33, 33, 33, 33, 33, 41, 41, 41, 41, 54, 54, 54, 71, 71, 91)


def num_matrices(size_id: int) -> int:
    """matrixId range: 6 for 4x4..16x16, 2 for 32x32 (the reference's numbering,
    scaling.py:35-36: 0 intra, 1 inter)."""
    return 2 if size_id == 3 else 6


def diag_scan(blk: int):
    """6.5.3 up-right diagonal scan: list of (x, y)."""
    out, x, y = [], 0, 0
    while len(out) < blk * blk:
        while y >= 0 and len(out) < blk * blk:
            if x < blk and y < blk:
                out.append((x, y))
            y -= 1
            x += 1
        y, x = x, 0
    return out


def default_list(size_id: int, matrix_id: int):
    if size_id == 0:
        return list(DEFAULT_4x4)
    intra = matrix_id < 3 if size_id < 3 else matrix_id == 0
    return list(DEFAULT_8x8_INTRA if intra else DEFAULT_8x8_INTER)


def default_lists():
    lists = {(s, m): default_list(s, m) for s in range(4) for m in range(num_matrices(s))}
    dc = {(s, m): 16 for s in (2, 3) for m in range(num_matrices(s))}
    return lists, dc


def expand(lists: dict, dc: dict) -> dict:
    """7.4.5 equations 7-XX: {(sizeId, matrixId): (N, N) int array indexed [x][y]}.

    4x4 and 8x8 map the list through the diagonal scan; 16x16 / 32x32 replicate each
    8x8 entry 2x2 / 4x4 times and overwrite [0][0] with the DC value."""
    sf = {}
    scan4, scan8 = diag_scan(4), diag_scan(8)
    for (s, m), lst in lists.items():
        n = 4 << s
        f = np.zeros((n, n), dtype=np.int64)
        if s == 0:
            if len(lst) != 16:
                raise ValueError("4x4 scaling list needs 16 entries")
            for i, (x, y) in enumerate(scan4):
                f[x, y] = lst[i]
        else:
            if len(lst) != 64:
                raise ValueError("8x8-based scaling list needs 64 entries")
            rep = 1 << (s - 1)
            for i, (x, y) in enumerate(scan8):
                f[x * rep:(x + 1) * rep, y * rep:(y + 1) * rep] = lst[i]
            if s >= 2:
                f[0, 0] = dc[(s, m)]
        sf[(s, m)] = f
    return sf


def default_scaling_factor() -> dict:
    return expand(*default_lists())


def as_reference_object(sf: dict):
    """Nested list form `scaling_factor[size_id][matrix_id]` -> (N, N) [x][y] array, the
    attribute scaling.py:44 reads from `sps`."""
    return [[sf.get((s, m)) for m in range(6)] for s in range(4)]


def parse_scaling_list_data(read_flag, read_ue, read_se):
    """scaling_list_data() syntax (7.3.4) driven by three bit-reader callables
    (u(1), ue(v), se(v) -- e.g. bsb.BitStreamBuffer.u / ue / se, bsb.py:132-168).
    Returns (lists, dc) ready for `expand`.  Restates what sld.py:63-116 attempts."""
    lists, dc = {}, {}
    for s in range(4):
        for m in range(num_matrices(s)):
            lists[(s, m)], one_dc = _parse_one(s, m, lists, dc, read_flag, read_ue, read_se)
            if s >= 2:
                dc[(s, m)] = one_dc
    return lists, dc

class ScalingListData:
    """Drop-in for the reference's `sld.ScalingListData(bs)` (sld.py:3, used by sps.py:18,90
    and pps.py:11,135).  `decode()` reads scaling_list_data() with the reference's bit reader
    (`bs.u(1, name)`, `bs.ue(name)`, `bs.se(name)`, bsb.py:143-168, logging the syntax
    elements under the names sld.py:78-111 uses) and leaves

        .scaling_list[size_id][matrix_id]          the coefficient lists (scan order)
        .scaling_list_dc_coef_minus8[size_id-2][matrix_id]
        .scaling_factor[size_id][matrix_id][x][y]  what scaling.inverse_scaling reads (scaling.py:44)
        .table                                      the packed 4064-byte device table

    The reference's own decode() cannot run (undefined names, sld.py:82-94)."""
    default_scaling_list_4x4 = list(DEFAULT_4x4)
    default_scaling_list_8x8_intra = list(DEFAULT_8x8_INTRA)
    default_scaling_list_8x8_inter = list(DEFAULT_8x8_INTER)

    def __init__(self, bs=None):
        self.bs = bs
        self.present = False
        self._set(*default_lists())

    def _set(self, lists, dc):
        from .picture import pack_scaling_factor
        self.lists, self.dc = lists, dc
        self.scaling_list = [[lists.get((s, m)) for m in range(num_matrices(s))] for s in range(4)]
        self.scaling_list_dc_coef_minus8 = [[dc[(s, m)] - 8 for m in range(num_matrices(s))] for s in (2, 3)]
        sf = expand(lists, dc)
        self.scaling_factor = as_reference_object(sf)
        self.table = pack_scaling_factor(sf)

    def decode(self):
        bs = self.bs
        pos = {"s": 0, "m": 0, "i": 0}

        # the syntax element names carry their indices (sld.py:78-106); track them by call order
        def read_flag():
            return bs.u(1, "scaling_list_pred_mode_flag[%d][%d]" % (pos["s"], pos["m"]))

        def read_ue():
            return bs.ue("scaling_list_pred_matrix_id_delta[%d][%d]" % (pos["s"], pos["m"]))

        def read_se():
            if pos["i"] < 0:
                pos["i"] = 0
                return bs.se("scaling_list_dc_coef_minus8[%d][%d]" % (pos["s"] - 2, pos["m"]))
            name = "scaling_list_delta_coef[%d][%d][%d]" % (pos["s"], pos["m"], pos["i"])
            pos["i"] += 1
            return bs.se(name)

        lists, dc = {}, {}
        for s in range(4):
            for m in range(num_matrices(s)):
                pos.update(s=s, m=m, i=-1 if s >= 2 else 0)
                one, one_dc = _parse_one(s, m, lists, dc, read_flag, read_ue, read_se)
                lists[(s, m)] = one
                if s >= 2:
                    dc[(s, m)] = one_dc
        self.present = True
        self._set(lists, dc)


def _parse_one(s, m, lists, dc, read_flag, read_ue, read_se):
    """One (sizeId, matrixId) entry of scaling_list_data() (7.3.4)."""
    coef_num = min(64, 1 << (4 + (s << 1)))
    if not read_flag():                               # scaling_list_pred_mode_flag == 0
        delta = read_ue()                             # scaling_list_pred_matrix_id_delta
        if delta == 0:
            return default_list(s, m), 16
        ref = m - delta
        if ref < 0:
            raise ValueError("scaling_list_pred_matrix_id_delta out of range")
        return list(lists[(s, ref)]), dc.get((s, ref), 16)
    next_coef, dc_val = 8, 16
    if s >= 2:
        next_coef = read_se() + 8                     # scaling_list_dc_coef_minus8
        if not 1 <= next_coef <= 255:
            raise ValueError("scaling_list_dc_coef_minus8 out of range")
        dc_val = next_coef
    lst = []
    for _ in range(coef_num):
        next_coef = (next_coef + read_se() + 256) % 256   # scaling_list_delta_coef
        if next_coef == 0:
            raise ValueError("ScalingList entries must be positive (7.4.5)")
        lst.append(next_coef)
    return lst, dc_val


def active_table(sps, pps=None):
    """The packed ScalingFactor table a picture uses (7.4.3.2.1 / 7.4.3.3.1) or None when
    scaling lists are off: PPS lists when pps_scaling_list_data_present_flag, else SPS lists
    when sps_scaling_list_data_present_flag, else the default lists (Tables 7-5 / 7-6)."""
    if not getattr(sps, "scaling_list_enabled_flag", 0):
        return None
    for owner, flag in ((pps, "pps_scaling_list_data_present_flag"), (sps, "sps_scaling_list_data_present_flag")):
        if owner is not None and getattr(owner, flag, 0):
            data = getattr(owner, "scaling_list_data", None)
            if isinstance(data, ScalingListData) and data.present:
                return data.table
            raise ValueError("%s is set but the scaling list data were not decoded by p265_b200's sld module" % flag)
    return ScalingListData().table
