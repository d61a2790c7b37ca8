"""scaling_list_data() syntax (7.3.4) + ScalingFactor derivation (7.4.5) behind the reference's
`sld.ScalingListData(bs)` surface (sld.py:58-153, which cannot run: SURVEY.md G4)."""
import importlib
import sys

import numpy as np
import pytest

from oracle import spec_oracle as so
from p265_b200 import scaling_list
from p265_b200.picture import pack_scaling_factor, sf_is_replicated


class BitWriter:
    def __init__(self):
        self.bits = []

    def u(self, n, v):
        self.bits += [(v >> (n - 1 - i)) & 1 for i in range(n)]

    def ue(self, v):
        v += 1
        n = v.bit_length()
        self.bits += [0] * (n - 1)
        self.u(n, v)

    def se(self, v):
        self.ue(2 * v - 1 if v > 0 else -2 * v)


class BitReader:
    """Same method surface as the reference's bsb.BitStreamBuffer (bsb.py:143-168)."""

    def __init__(self, bits):
        self.bits, self.pos, self.names = bits, 0, []

    def read_bits(self, n):
        v = 0
        for _ in range(n):
            v = (v << 1) | self.bits[self.pos]
            self.pos += 1
        return v

    def u(self, n, name):
        self.names.append(name)
        return self.read_bits(n)

    def ue(self, name, no_print=False):
        z = 0
        while self.read_bits(1) == 0:
            z += 1
        if not no_print:
            self.names.append(name)
        return (1 << z) - 1 + self.read_bits(z)

    def se(self, name):
        k = self.ue(name, True)
        self.names.append(name)
        return (k + 1) // 2 if k & 1 else -(k // 2)


def encode(lists, dc, mode):
    """mode[(s, m)]: 'explicit', 'default' or ('ref', delta)."""
    w = BitWriter()
    for s in range(4):
        for m in range(scaling_list.num_matrices(s)):
            how = mode[(s, m)]
            if how == "explicit":
                w.u(1, 1)
                nxt = 8
                if s >= 2:
                    w.se(dc[(s, m)] - 8)
                    nxt = dc[(s, m)]
                for v in lists[(s, m)]:
                    d = (v - nxt + 128) % 256 - 128
                    w.se(d)
                    nxt = v
            else:
                w.u(1, 0)
                w.ue(0 if how == "default" else how[1])
    return w.bits


def test_round_trip_of_random_lists_through_the_reference_surface():
    rng = np.random.default_rng(745)
    lists, dc, mode = {}, {}, {}
    for s in range(4):
        for m in range(scaling_list.num_matrices(s)):
            r = rng.random()
            if m > 0 and r < 0.25:
                delta = int(rng.integers(1, m + 1))
                mode[(s, m)] = ("ref", delta)
                lists[(s, m)] = list(lists[(s, m - delta)])
                if s >= 2:
                    dc[(s, m)] = dc[(s, m - delta)]
            elif r < 0.4:
                mode[(s, m)] = "default"
                lists[(s, m)] = scaling_list.default_list(s, m)
                if s >= 2:
                    dc[(s, m)] = 16
            else:
                mode[(s, m)] = "explicit"
                lists[(s, m)] = [int(v) for v in rng.integers(1, 256, 16 if s == 0 else 64)]
                if s >= 2:
                    dc[(s, m)] = int(rng.integers(1, 256))
    bs = BitReader(encode(lists, dc, mode))
    sys.path.insert(0, __import__("p265_b200").dropin_path())
    try:
        sys.modules.pop("sld", None)
        sld = importlib.import_module("sld")                  # bare name, like sps.py:3
        assert "p265_b200" in sld.__file__
        data = sld.ScalingListData(bs)
        data.decode()
    finally:
        sys.path.remove(__import__("p265_b200").dropin_path())
        sys.modules.pop("sld", None)
    assert bs.pos == len(bs.bits)
    assert bs.names[0] == "scaling_list_pred_mode_flag[0][0]"
    for (s, m), lst in lists.items():
        assert data.scaling_list[s][m] == lst
        if s >= 2:
            assert data.scaling_list_dc_coef_minus8[s - 2][m] == dc[(s, m)] - 8
    # ScalingFactor == the oracle's 7.4.5 expansion; device table == the packed form
    want = so.expand_scaling_factor(lists, dc)
    for (s, m), f in want.items():
        assert np.array_equal(np.asarray(data.scaling_factor[s][m]), f)
    assert np.array_equal(data.table, so.pack_scaling_factor(want))
    assert np.array_equal(data.table, pack_scaling_factor(scaling_list.expand(lists, dc)))
    assert sf_is_replicated(data.table)
    # and the function form agrees
    bs2 = BitReader(encode(lists, dc, mode))
    l2, d2 = scaling_list.parse_scaling_list_data(lambda: bs2.u(1, ""), lambda: bs2.ue(""), lambda: bs2.se(""))
    assert l2 == lists and d2 == dc


def test_active_table_resolution():
    import types
    off = types.SimpleNamespace(scaling_list_enabled_flag=0)
    assert scaling_list.active_table(off) is None
    sps = types.SimpleNamespace(scaling_list_enabled_flag=1, sps_scaling_list_data_present_flag=0)
    default = scaling_list.active_table(sps)
    assert np.array_equal(default, so.pack_scaling_factor(so.expand_scaling_factor(*so.default_scaling_lists())))
    lists, dc = scaling_list.default_lists()
    lists[(1, 0)] = [17] * 64
    bits = encode(lists, dc, {k: "explicit" for k in lists})
    data = scaling_list.ScalingListData(BitReader(bits))
    data.decode()
    pps = types.SimpleNamespace(pps_scaling_list_data_present_flag=1, scaling_list_data=data)
    tab = scaling_list.active_table(sps, pps)
    assert not np.array_equal(tab, default) and (tab[96:96 + 64] == 17).all()
    # a flag that is set without decoded data (the reference's own sld.py) must not be ignored
    bad = types.SimpleNamespace(pps_scaling_list_data_present_flag=1, scaling_list_data=object())
    with pytest.raises(ValueError):
        scaling_list.active_table(sps, bad)


def test_malformed_syntax_is_rejected():
    w = BitWriter()
    w.u(1, 0)
    w.ue(1)                                    # matrix 0 cannot refer to matrix -1
    with pytest.raises(ValueError):
        scaling_list.ScalingListData(BitReader(w.bits + [0] * 64)).decode()
