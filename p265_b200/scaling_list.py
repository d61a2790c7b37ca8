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
            coef_num = min(64, 1 << (4 + (s << 1)))
            if not read_flag():                       # scaling_list_pred_mode_flag == 0
                delta = read_ue()                     # scaling_list_pred_matrix_id_delta
                if delta == 0:
                    lists[(s, m)] = default_list(s, m)
                    if s >= 2:
                        dc[(s, m)] = 16
                else:
                    ref = m - delta
                    if ref < 0:
                        raise ValueError("scaling_list_pred_matrix_id_delta out of range")
                    lists[(s, m)] = list(lists[(s, ref)])
                    if s >= 2:
                        dc[(s, m)] = dc[(s, ref)]
            else:
                next_coef = 8
                if s >= 2:
                    next_coef = read_se() + 8          # scaling_list_dc_coef_minus8
                    if not 1 <= next_coef <= 255:
                        raise ValueError("scaling_list_dc_coef_minus8 out of range")
                    dc[(s, m)] = next_coef
                lst = []
                for _ in range(coef_num):
                    next_coef = (next_coef + read_se() + 256) % 256   # scaling_list_delta_coef
                    lst.append(next_coef)
                lists[(s, m)] = lst
    return lists, dc
