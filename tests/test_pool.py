"""EnginePool (p265_b200/pool.py): dispatch logic on CPU with stand-in engines, and on the GPU box
bit-exact results through the pool (one device always; two devices when the box has them)."""
import numpy as np
import pytest

from p265_b200 import partition, pool as pool_mod


class _FakeEngine:
    def __init__(self, device):
        self.device, self.launch_count, self.synced = device, 0, 0

    def set_async(self, enable):
        assert enable

    def sync(self):
        self.synced += 1

    def close(self):
        pass

    def residual(self, batch, out):
        self.launch_count += 5
        if batch == "boom":
            raise ValueError("bad batch")
        return ("residual", self.device, batch)


class _FakeLib:
    def p265_device_count(self):
        return 4


def test_pool_partitions_pictures_round_robin(monkeypatch):
    monkeypatch.setattr(pool_mod, "Engine", _FakeEngine)
    monkeypatch.setattr(pool_mod._lib, "load", lambda: _FakeLib())
    with pool_mod.EnginePool(contexts_per_device=2) as pool:
        futs = [pool.residual(p, None, picture=p) for p in range(11)]
        got = [f.result(timeout=10) for f in futs]
        assert got == [("residual", p % 4, p) for p in range(11)]
        for rank in range(4):          # the same partition the multi-process benchmark uses
            assert [p for p in range(11) if pool.device_of(p) == rank] == partition.pictures_of_rank(11, rank, 4)
        assert pool.calls() == {0: 3, 1: 3, 2: 3, 3: 2}
        assert all(v > 0 for v in pool.launch_counts().values())
        bad = pool.residual("boom", None)
        with pytest.raises(ValueError):
            bad.result(timeout=10)
        assert pool.residual(99, None).result(timeout=10)[2] == 99      # the worker survives a failed call
    with pytest.raises(ValueError):
        pool_mod.EnginePool(devices=[7])


@pytest.mark.gpu
def test_pool_on_the_gpus_of_this_box(c_oracle):
    from conftest import small_cfg
    from p265_b200 import _lib, synth
    n_dev = _lib.load().p265_device_count()
    devices = list(range(min(n_dev, 2)))
    with pool_mod.EnginePool(devices, contexts_per_device=2) as pool:
        jobs = []
        for p in range(6):
            batch = synth.residual_batch(small_cfg("4k10", 512, 256), n_pics=1, seed=700 + p)
            geom, rec, params = synth.sao_batch(512, 264, 10, n_pics=1, ctb_log2=6, seed=800 + p)
            jobs.append((batch, pool.residual(batch.packed(), picture=p), geom, rec, params,
                         pool.sao(rec, geom, 6, params, picture=p)))
        for batch, f_res, geom, rec, params, f_sao in jobs:
            assert np.array_equal(f_res.result(timeout=120), c_oracle.residual_batch(batch, zero_fill=False))
            got, ref = f_sao.result(timeout=120), c_oracle.sao_batch(rec, geom, 6, params)
            for c in range(3):
                assert np.array_equal(geom.plane_view(got, 0, c), geom.plane_view(ref, 0, c))
        counts = pool.launch_counts()
        assert set(counts) == set(devices) and all(v > 0 for v in counts.values())


@pytest.mark.gpu
def test_pool_uses_two_devices():
    from p265_b200 import _lib
    if _lib.load().p265_device_count() < 2:
        pytest.skip("needs two GPUs")
    from conftest import small_cfg
    from p265_b200 import synth
    from p265_b200.engine import Engine
    ref_eng = Engine(0)
    with pool_mod.EnginePool([0, 1]) as pool:
        batches = [synth.residual_batch(small_cfg("4k10", 512, 256), n_pics=1, seed=900 + p) for p in range(8)]
        futs = [pool.residual(b, picture=p) for p, b in enumerate(batches)]
        for b, f in zip(batches, futs):
            assert np.array_equal(f.result(timeout=120), ref_eng.residual(b))
        counts = pool.launch_counts()
        assert counts[0] > 0 and counts[1] > 0
