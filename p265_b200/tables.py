"""H.265 8.6.4.2 basis tables in the shape the reference exposes them
(`transform.trans_matrix_type0` = 32x32 DCT, `trans_matrix_type1` = 4x4 DST-VII,
transform.py:5-72), rebuilt from their structure rather than copied: entry [j][i] of the
DCT basis is +/- mag[k], k being the angle j*(2i+1)*pi/64 folded into [0, pi/2]."""
from __future__ import annotations

_MAG = (64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64,
        61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4, 0)


def _dct32():
    rows = []
    for j in range(32):
        row = []
        for i in range(32):
            if j == 0:
                row.append(64)
                continue
            a = (j * (2 * i + 1)) % 128
            if a > 64:
                a = 128 - a
            row.append(_MAG[a] if a <= 32 else -_MAG[64 - a])
        rows.append(row)
    return rows


def _dst4():
    a, b, c, d = 29, 55, 74, 84
    return [[a, b, c, d], [c, c, 0, -c], [d, -a, -c, b], [b, -d, c, -a]]


trans_matrix_type0 = _dct32()
trans_matrix_type1 = _dst4()
